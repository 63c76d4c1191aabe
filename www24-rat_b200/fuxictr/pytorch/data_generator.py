"""Batch producers for the RAT hot path.

Two producers with the same `len()` / re-iterable contract the training loop needs (base_model.py:197,220):

* DataGenerator       -- host producer with the reference's ctor signature and WIRE FORMAT: a torch DataLoader
                         over `Dataset`, yielding (X f64 [B,1+K,L], y f64 [B,1+K], retrieved_values f64 [B,K],
                         retrieved_lens i64 [B]) exactly like fuxictr/pytorch/data_generator.py:66-78,84-251.
* DeviceDataGenerator -- B200-native producer: the int32 id matrix, uint8 labels, the retrieval pool and the
                         [Q,K] neighbour index live in HBM; a batch is just a row list, and the model assembles the
                         retrieval set on the device with rat_assemble_ids.

Data blocks are read from `.h5` when h5py is importable, else from `.npz` / `.npy` mirrors with the same stem
(key "data"; retrieval files hold "indices", "values", "lens" like data_generator.py:107-113,213-215).
BM25 pre-retrieval (pre_retrieval: true is asserted by the reference, :101): the cached retrieval file is read when it
exists, otherwise it is computed on the GPU (compute_retrieval -> rat_bm25_topk) and saved, like data_generator.py:114-215.
"""
import logging
import os

import numpy as np
import torch
from torch.utils import data


def _load_array(path, key=None):
    """h5 / npz / npy loader (reference: load_hdf5 fuxictr/datasets/data_utils.py:46-54)."""
    stem, ext = os.path.splitext(path)
    if ext == ".h5" and os.path.exists(path):
        try:
            import h5py
        except ImportError:
            h5py = None
        if h5py is not None:
            with h5py.File(path, "r") as hf:
                return hf[key][()] if key is not None else hf[list(hf.keys())[0]][()]
    for cand in (path, stem + ".npz", stem + ".npy"):
        if os.path.exists(cand) and cand.endswith(".npz"):
            with np.load(cand) as z:
                return z[key] if key is not None else z[z.files[0]]
        if os.path.exists(cand) and cand.endswith(".npy"):
            if key is not None:
                raise KeyError("{} has no key {}".format(cand, key))
            return np.load(cand)
    raise FileNotFoundError("no readable data block for {} (h5 needs h5py; npz/npy mirrors are accepted)".format(path))


def _exists(path):
    stem = os.path.splitext(path)[0]
    return any(os.path.exists(p) for p in (path, stem + ".npz", stem + ".npy"))


class Dataset(data.Dataset):
    """host retrieval-set assembly (numpy fancy indexing: index -1 wraps to the last pool row)."""

    def __init__(self, darray, feature_map=None, graph_processor=None, retr_pool_darray=None, retr_indices=None,
                 retr_values=None, retr_lens=None):
        if graph_processor:
            raise NotImplementedError("graph processors (DGL) are not part of the RAT path")
        self.darray = darray
        have = [a is not None for a in (retr_pool_darray, retr_indices, retr_values, retr_lens)]
        self.retrieval_augmented = all(have)
        self.retr_pool_darray, self.retr_indices = retr_pool_darray, retr_indices
        self.retr_values, self.retr_lens = retr_values, retr_lens
        if self.retrieval_augmented:
            n = len(darray)
            assert n == len(retr_indices) == len(retr_values) == len(retr_lens), "retrieval arrays must align with data"
            assert retr_indices.shape[-1] == retr_values.shape[-1]

    def __getitem__(self, index):
        row = self.darray[index]
        if not self.retrieval_augmented:
            return row[..., :-1], row[..., -1]
        block = np.concatenate([row[None], self.retr_pool_darray[self.retr_indices[index]]])
        return block[..., :-1], block[..., -1], self.retr_values[index], self.retr_lens[index]

    def __len__(self):
        return len(self.darray)


def _save_retrieval(path, indices, values, lens):
    """h5 with the reference's keys (data_generator.py:213-215) when h5py exists, else the npz mirror _load_array reads"""
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:
        with h5py.File(path, "a") as hf:
            for k, v in (("indices", indices), ("values", values), ("lens", lens)):
                hf.create_dataset(k, data=v)
    else:
        np.savez(os.path.splitext(path)[0] + ".npz", indices=indices, values=values, lens=lens)


def compute_retrieval(data_array, retrieval_configs, retrieval_pool_fname, db_array=None):
    """BM25 pre-retrieval of one data block on the GPU: the driver logic of the reference (data_generator.py:115-212 --
    X-fold self retrieval, label-wise positive / negative pools, external pool) around the device kernel
    (fuxictr.datasets.data_utils.BM25_topk_retrieval -> rat_bm25_topk).  Returns (indices, values, lens)."""
    import re
    from ..datasets.data_utils import BM25_topk_retrieval
    cfg = retrieval_configs
    cols = cfg["used_col_indices"]

    def split_by_label(db, labels, qry, remap):
        """label-wise: K from the positives, K from the negatives (data_generator.py:133-166,182-206)"""
        outs = []
        for sel in (np.nonzero(labels)[0], np.nonzero(1 - labels)[0]):
            r = BM25_topk_retrieval(db_np_data=db[sel], qry_np_data=qry, **cfg)
            outs.append((remap(sel[r.indices]), r.values, r.lens))       # like the reference, -1 wraps to the last row
        return (np.concatenate([outs[0][0], outs[1][0]], axis=-1), np.concatenate([outs[0][1], outs[1][1]], axis=-1),
                np.stack([outs[0][2], outs[1][2]], axis=-1))

    if retrieval_pool_fname == "self":
        arr = data_array[:, cols].astype(int)
        labels = data_array[:, -1].astype(int) if cfg["label_wise"] else None
        fold_num = int(re.match(r"\d+-fold", cfg["split_type"]).group().split("-")[0])
        fold_size = int(np.ceil(len(arr) / fold_num))
        ind, val, lens = [], [], []
        for fi in range(fold_num):
            lo, hi = fi * fold_size, (fi + 1) * fold_size
            qry = arr[lo:hi]
            db = np.concatenate([arr[:lo], arr[hi:]], axis=0)
            db_idx = np.concatenate([np.arange(lo), np.arange(hi, len(arr))], axis=0)
            if cfg["label_wise"]:
                i, v, n = split_by_label(db, np.concatenate([labels[:lo], labels[hi:]]), qry, lambda x: db_idx[x])
            else:
                r = BM25_topk_retrieval(db_np_data=db, qry_np_data=qry, **cfg)
                i, v, n = db_idx[r.indices], r.values, r.lens
            ind.append(i); val.append(v); lens.append(n)
        return np.concatenate(ind), np.concatenate(val), np.concatenate(lens)
    db = db_array[:, cols].astype(int)
    qry = data_array[:, cols].astype(int)
    if cfg["label_wise"]:
        return split_by_label(db, db_array[:, -1].astype(int), qry, lambda x: x)
    r = BM25_topk_retrieval(db_np_data=db, qry_np_data=qry, **cfg)
    return r.indices, r.values, r.lens


def _load_retrieval(data_path, retrieval_configs, data_array=None, retrieval_pool_fname=None, db_array=None):
    """the cached retrieval file of the reference (data_generator.py:107-113) or, when it does not exist yet, the BM25
    pre-retrieval itself on the GPU followed by the same save (data_generator.py:114-215)"""
    root, fname = os.path.split(data_path)
    path = os.path.join(root, "retrieval_{}_".format(retrieval_configs["topK"]) + fname)
    if _exists(path):
        return (_load_array(path, "indices"), _load_array(path, "values"), _load_array(path, "lens"))
    if data_array is None or "used_col_indices" not in retrieval_configs:
        raise FileNotFoundError("pre-computed retrieval file {} not found and no retrieval columns configured".format(path))
    logging.info("BM25 pre-retrieval on the GPU -> %s", path)
    ind, val, lens = compute_retrieval(data_array, retrieval_configs, retrieval_pool_fname, db_array)
    _save_retrieval(path, ind, val, lens)
    return ind, val, lens


def _stats(obj, data_array, batch_size):
    obj.num_blocks = 1
    obj.num_samples = len(data_array)
    obj.num_batches = int(np.ceil(obj.num_samples * 1.0 / batch_size))
    obj.num_positives = float(data_array[:, -1].sum())
    obj.num_negatives = obj.num_samples - obj.num_positives


class DataGenerator(data.DataLoader):
    def __init__(self, data_path, batch_size=32, shuffle=False, num_workers=1, feature_map=None, graph_processor=None,
                 retrieval_configs=None, retrieval_pool_fname=None, retrieval_augmented=False, **kwargs):
        if isinstance(data_path, list):
            data_path = data_path[0]
        data_array = _load_array(data_path)
        if retrieval_configs is not None:
            assert retrieval_configs["pre_retrieval"], "only the pre-retrieval strategy exists (as in the reference)"
            pool = data_array if retrieval_pool_fname == "self" else _load_array(retrieval_pool_fname)
            idx, vals, lens = _load_retrieval(data_path, retrieval_configs, data_array, retrieval_pool_fname, pool)
            if retrieval_augmented:
                self.dataset = Dataset(data_array, feature_map, None, pool, idx, vals, lens)
            else:
                self.dataset = Dataset(data_array, feature_map)
        else:
            assert not retrieval_augmented, "retrieval-augmented mode requires retrieval_configs"
            self.dataset = Dataset(data_array, feature_map)
        super().__init__(dataset=self.dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers,
                         pin_memory=torch.cuda.is_available())
        _stats(self, data_array, batch_size)


class DeviceBatch(object):
    """A batch that is already resident on the GPU: the model gathers it with rat_assemble_ids."""
    __slots__ = ("gen", "rows", "row0", "size")

    def __init__(self, gen, rows, row0, size):
        self.gen, self.rows, self.row0, self.size = gen, rows, row0, size


class DeviceDataGenerator(object):
    """HBM-resident dataset + retrieval pool.  rank/world shard every batch for data-parallel training."""

    def __init__(self, data_array, pool_array, retr_indices, batch_size=32, shuffle=False, device="cuda:0", seed=2021,
                 rank=0, world=1, drop_last=False):
        dev = torch.device(device)
        self.device = dev
        self.batch_size, self.shuffle, self.rank, self.world = int(batch_size), shuffle, rank, world
        self.q_ids = torch.from_numpy(np.ascontiguousarray(data_array[:, :-1]).astype(np.int32)).to(dev)
        self.q_labels = torch.from_numpy(np.ascontiguousarray(data_array[:, -1]).astype(np.uint8)).to(dev)
        if pool_array is data_array:
            self.pool_ids, self.pool_labels = self.q_ids, self.q_labels
        else:
            self.pool_ids = torch.from_numpy(np.ascontiguousarray(pool_array[:, :-1]).astype(np.int32)).to(dev)
            self.pool_labels = torch.from_numpy(np.ascontiguousarray(pool_array[:, -1]).astype(np.uint8)).to(dev)
        self.nbr = torch.from_numpy(np.ascontiguousarray(retr_indices).astype(np.int64)).to(dev)
        self.K = int(self.nbr.shape[1])
        self.L = int(self.q_ids.shape[1])
        self.n_pool = int(self.pool_ids.shape[0])
        self.drop_last = drop_last
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(seed)
        _stats(self, data_array, batch_size)
        if drop_last:
            self.num_batches = self.num_samples // self.batch_size

    def __len__(self):
        return self.num_batches

    def __iter__(self):
        Q, bs = self.num_samples, self.batch_size
        perm = torch.randperm(Q, device=self.device, generator=self._gen) if self.shuffle else None
        for b in range(self.num_batches):
            lo, hi = b * bs, min(Q, (b + 1) * bs)
            n = hi - lo
            # contiguous per-rank slice of the global batch (data-parallel batch sharding).  Every rank must get the SAME
            # number of samples (the engine scales the loss by 1 / (B_local * world) and BatchNorm counts B_local * world):
            # a ragged last batch is truncated to a multiple of the world size (at most world - 1 samples are dropped)
            per = n // self.world
            if per == 0:
                continue
            s, e = lo + self.rank * per, lo + (self.rank + 1) * per
            if perm is not None:
                yield DeviceBatch(self, perm[s:e], 0, e - s)
            else:
                yield DeviceBatch(self, None, s, e - s)


class DeviceDataFileGenerator(DeviceDataGenerator):
    """DeviceDataGenerator built from the same files / arguments as DataGenerator."""

    def __init__(self, data_path, batch_size=32, shuffle=False, num_workers=1, feature_map=None, graph_processor=None,
                 retrieval_configs=None, retrieval_pool_fname=None, retrieval_augmented=False, device="cuda:0",
                 seed=2021, **kwargs):
        if isinstance(data_path, list):
            data_path = data_path[0]
        assert retrieval_configs is not None and retrieval_augmented, "device generator serves retrieval-augmented data"
        data_array = _load_array(data_path)
        pool = data_array if retrieval_pool_fname == "self" else _load_array(retrieval_pool_fname)
        idx, _, _ = _load_retrieval(data_path, retrieval_configs, data_array, retrieval_pool_fname, pool)
        # only the (shuffled) training generator is sharded over the ranks: validation / test generators serve the whole set on
        # every rank, so all ranks compute identical metrics and take identical early-stop / lr-decay decisions
        dp = bool(kwargs.get("data_parallel")) and bool(shuffle)
        rank = int(os.environ.get("RANK", 0)) if dp else 0
        world = int(os.environ.get("WORLD_SIZE", 1)) if dp else 1
        super().__init__(data_array, pool, idx, batch_size, shuffle, device, seed, rank, world)


def get_data_generator(data_path_list, batch_size=32, shuffle=False, num_workers=1, feature_map=None,
                       retrieval_configs=None, retrieval_pool_fname=None, retrieval_augmented=False, **kwargs):
    """reference: data_generator.py:479-508.  `device_resident: true` (or gpu>=0 with `device_resident` unset and
    retrieval-augmented data) selects the HBM-resident producer."""
    assert len(data_path_list) > 0, "invalid data files or paths."
    if len(data_path_list) > 1:
        raise NotImplementedError("multi-block h5 generators (DataBlockGenerator) are broken in the reference "
                                  "(data_generator.py:292,419) and not part of the RAT path; merge the blocks")
    device_resident = kwargs.pop("device_resident", None)
    gpu = kwargs.pop("gpu", -1)
    seed = kwargs.pop("seed", 2021)
    if device_resident is None:
        device_resident = bool(retrieval_augmented and retrieval_configs is not None and gpu is not None and gpu >= 0
                               and torch.cuda.is_available())
    if device_resident:
        logging.info("HBM-resident data generator: " + str(data_path_list[0]))
        return DeviceDataFileGenerator(data_path_list[0], batch_size=batch_size, shuffle=shuffle,
                                       feature_map=feature_map, retrieval_configs=retrieval_configs,
                                       retrieval_pool_fname=retrieval_pool_fname, retrieval_augmented=retrieval_augmented,
                                       device="cuda:{}".format(gpu), seed=seed, **kwargs)
    return DataGenerator(data_path=data_path_list[0], batch_size=batch_size, shuffle=shuffle, num_workers=num_workers,
                         feature_map=feature_map, retrieval_configs=retrieval_configs,
                         retrieval_pool_fname=retrieval_pool_fname, retrieval_augmented=retrieval_augmented, **kwargs)
