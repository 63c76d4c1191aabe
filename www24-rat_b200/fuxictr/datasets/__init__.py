"""Dataset entry points used by run_expid.py (reference: fuxictr/datasets/data_utils.py:1189-1280).
csv -> h5 building (build_dataset, split_train_test, save_hdf5) and the kkbox / tmall FeatureEncoders are the offline
preparation in front of the hot path (reference data_utils.py:37-54,1067-1186, datasets/kkbox.py,
datasets/tmall.py)."""
import glob
import logging
import os
import re


def _blocks(pattern):
    found = glob.glob(pattern)
    if not found:      # accept npz / npy mirrors of the h5 blocks
        stem = os.path.splitext(pattern)[0]
        found = glob.glob(stem + ".npz") + glob.glob(stem + ".npy")
    found = [f for f in found if not os.path.basename(f).startswith("retrieval_")]
    if len(found) > 1:
        found.sort(key=lambda x: int(x.split("_")[-1].split(".")[0]))
    return found


def _log(tag, gen):
    logging.info("{} samples: total/{:d}, pos/{:.0f}, neg/{:.0f}, ratio/{:.2f}%, blocks/{:.0f}".format(
        tag, gen.num_samples, gen.num_positives, gen.num_negatives, 100. * gen.num_positives / gen.num_samples,
        gen.num_blocks))


def h5_generator(feature_map, stage="both", train_data=None, valid_data=None, test_data=None, batch_size=32,
                 shuffle=True, retrieval_configs=None, retrieval_augmented=False, **kwargs):
    from ..pytorch.data_generator import get_data_generator
    logging.info("Loading data...")
    xfold = retrieval_configs is not None and re.match(r"\d+-fold", str(retrieval_configs["split_type"])) is not None
    if retrieval_configs is not None and "used_cols" in retrieval_configs:      # reference data_utils.py:1194-1205
        retrieval_configs["used_col_indices"] = [feature_map.feature_specs[c]["index"] for c in retrieval_configs["used_cols"]]
        exact = retrieval_configs.get("exact_match_cols") or []
        retrieval_configs["exact_match_col_indices"] = [retrieval_configs["used_cols"].index(c) for c in exact] if exact else None

    def pool_for(first_train_block, own_is_train):
        if retrieval_configs is None:
            return None
        if xfold:                                   # X-fold: train retrieves from itself, others from train block 0
            return "self" if own_is_train else first_train_block
        return retrieval_configs["retrieval_pool_data"]

    common = dict(batch_size=batch_size, feature_map=feature_map, retrieval_configs=retrieval_configs,
                  retrieval_augmented=retrieval_augmented)
    train_gen = valid_gen = test_gen = None
    if stage in ("both", "train"):
        train_blocks, valid_blocks = _blocks(train_data), _blocks(valid_data)
        assert len(train_blocks) > 0 and len(valid_blocks) > 0, "invalid data files or paths."
        train_gen = get_data_generator(train_blocks, shuffle=shuffle, retrieval_pool_fname=pool_for(train_blocks[0], True),
                                       **common, **kwargs)
        valid_gen = get_data_generator(valid_blocks, shuffle=False, retrieval_pool_fname=pool_for(train_blocks[0], False),
                                       **common, **kwargs)
        _log("Train", train_gen)
        _log("Validation", valid_gen)
        if stage == "train":
            logging.info("Loading train data done.")
            return train_gen, valid_gen
    if stage in ("both", "test"):
        test_blocks = _blocks(test_data)
        if test_blocks:
            first_train = _blocks(train_data)[0] if xfold else None
            test_gen = get_data_generator(test_blocks, shuffle=False, retrieval_pool_fname=pool_for(first_train, False),
                                          **common, **kwargs)
            _log("Test", test_gen)
        if stage == "test":
            logging.info("Loading test data done.")
            return test_gen
    logging.info("Loading data done.")
    return train_gen, valid_gen, test_gen


def save_hdf5(data_array, data_path, key="data"):
    """one data block: `.h5` with the reference's key (data_utils.py:37-44) when h5py is importable, else the `.npz` mirror
    that pytorch/data_generator.py reads"""
    import numpy as np
    logging.info("Saving data to h5: " + data_path)
    dir_name = os.path.dirname(data_path)
    if dir_name:
        os.makedirs(dir_name, exist_ok=True)
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:
        with h5py.File(data_path, "a") as hf:
            hf.create_dataset(key, data=data_array)
    else:
        np.savez(os.path.splitext(data_path)[0] + ".npz", **{key: data_array})


def load_hdf5(data_path, key=None, verbose=True):
    """one data block back (reference data_utils.py:46-54): `.h5` through h5py, or the `.npz` / `.npy` mirror of the same stem"""
    from ..pytorch.data_generator import _load_array
    if verbose:
        logging.info("Loading data from h5: " + data_path)
    return _load_array(data_path, key)


def split_train_test(train_ddf=None, valid_ddf=None, test_ddf=None, valid_size=0, test_size=0, split_type="sequential"):
    """tail splits of the training frame (reference data_utils.py:1067-1088): the LAST test_size rows become the test set, the
    valid_size rows before them the validation set; sizes < 1 are fractions of the original frame; "random" shuffles the row
    order first (numpy global RNG, as the reference)."""
    import numpy as np
    n = len(train_ddf)
    order = np.arange(n)
    if split_type == "random":
        np.random.shuffle(order)
    train_size = n
    if test_size > 0:
        if test_size < 1:
            test_size = int(n * test_size)
        train_size -= test_size
        test_ddf = train_ddf.loc[order[train_size:], :].reset_index()
        order = order[:train_size]
    if valid_size > 0:
        if valid_size < 1:
            valid_size = int(n * valid_size)
        train_size -= valid_size
        valid_ddf = train_ddf.loc[order[train_size:], :].reset_index()
        order = order[:train_size]
    if valid_size > 0 or test_size > 0:
        train_ddf = train_ddf.loc[order, :].reset_index()
    return train_ddf, valid_ddf, test_ddf


def _save_blocks(array, data_dir, stem, block_size):
    if block_size > 0:
        for block_id, idx in enumerate(range(0, len(array), block_size)):
            save_hdf5(array[idx:idx + block_size, :], os.path.join(data_dir, "{}_part_{}.h5".format(stem, block_id)))
    else:
        save_hdf5(array, os.path.join(data_dir, stem + ".h5"))


def build_dataset(feature_encoder, train_data=None, valid_data=None, test_data=None, valid_size=0, test_size=0,
                  split_type="sequential", retrieval_configs=None, **kwargs):
    """csv -> feature_map.json + train / valid / test (/ retrieval_pool) id blocks (reference data_utils.py:1091-1186):
    the vocabulary is fitted on the training frame (plus the retrieval pool when it is a file of its own); without an
    X-fold split a `pool_ratio` head of the training frame becomes the retrieval pool."""
    import pandas as pd
    enc = feature_encoder
    train_ddf = enc.preprocess(enc.read_csv(train_data))
    valid_ddf = enc.preprocess(enc.read_csv(valid_data)) if valid_data else None
    test_ddf = enc.preprocess(enc.read_csv(test_data)) if test_data else None
    if valid_size > 0 or test_size > 0:
        train_ddf, valid_ddf, test_ddf = split_train_test(train_ddf, valid_ddf, test_ddf, valid_size, test_size, split_type)
    pool_ddf = None
    xfold = retrieval_configs is not None and re.match(r"\d+-fold", str(retrieval_configs.get("split_type"))) is not None
    if retrieval_configs is not None:
        if "retrieval_pool_data" in retrieval_configs:
            pool_ddf = enc.preprocess(enc.read_csv(retrieval_configs["retrieval_pool_data"]))
            enc.fit(pd.concat([train_ddf, pool_ddf]), **kwargs)
        else:
            assert "pool_ratio" in retrieval_configs and "split_type" in retrieval_configs
            enc.fit(train_ddf, **kwargs)
            if not xfold:       # the pool is the HEAD of the training frame, the rest stays the training set
                pool_ddf, train_ddf, _ = split_train_test(train_ddf=train_ddf, valid_size=(1 - retrieval_configs["pool_ratio"]),
                                                          split_type=retrieval_configs["split_type"])
    else:
        enc.fit(train_ddf, **kwargs)
    block_size = int(kwargs.get("data_block_size", 0))
    _save_blocks(enc.transform(train_ddf), enc.data_dir, "train", block_size)
    if retrieval_configs is not None and not xfold:
        _save_blocks(enc.transform(pool_ddf), enc.data_dir, "retrieval_pool", block_size)
    if valid_ddf is not None:
        _save_blocks(enc.transform(valid_ddf), enc.data_dir, "valid", block_size)
    if test_ddf is not None:
        _save_blocks(enc.transform(test_ddf), enc.data_dir, "test", block_size)
    logging.info("Transform csv data to h5 done.")


from . import kkbox, tmall  # noqa: E402  (run_expid.py: getattr(datasets, <dataset>).FeatureEncoder)
from .data_utils import BM25_topk_retrieval_v4, BM25_topk_retrieval_v4 as BM25_topk_retrieval  # noqa: E402,F401
