"""Dataset entry points used by run_expid.py (reference: fuxictr/datasets/data_utils.py:1189-1280).
csv -> h5 building (build_dataset), BM25 retrieval and the kkbox/tmall FeatureEncoders are offline preparation
outside the hot path (SURVEY.md 2 #5,#6)."""
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
    if retrieval_configs is not None and "used_cols" in retrieval_configs:
        retrieval_configs["used_col_indices"] = [feature_map.feature_specs[c]["index"] for c in retrieval_configs["used_cols"]]

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


def build_dataset(feature_encoder, **kwargs):
    raise NotImplementedError("csv -> h5 dataset building is offline preparation outside the B200 hot path; "
                              "run the reference's build_dataset once and point data_root at the result")
