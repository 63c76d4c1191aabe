"""FeatureMap: the schema object consumed by the gather (reference: fuxictr/features.py:36-90; same JSON layout), and
FeatureEncoder: the offline csv -> id preparation in front of it (reference :93-337)."""
import io
import json
import logging
import os
from collections import OrderedDict


class FeatureMap(object):
    def __init__(self, dataset_id, data_dir, version="pytorch"):
        self.data_dir = data_dir
        self.dataset_id = dataset_id
        self.version = version
        self.num_fields = 0
        self.num_features = 0
        self.input_length = 0
        self.feature_specs = OrderedDict()

    def set_feature_index(self):
        """sequence features take max_len consecutive columns (reference features.py:46-57)."""
        logging.info("Set feature index...")
        idx = 0
        for feature, spec in self.feature_specs.items():
            if spec["type"] != "sequence":
                spec["index"] = idx
                idx += 1
            else:
                spec["index"] = [i + idx for i in range(spec["max_len"])]
                idx += spec["max_len"]
        self.input_length = idx

    def get_feature_index(self, feature_type=None):
        if feature_type is None:
            return []
        if not isinstance(feature_type, list):
            feature_type = [feature_type]
        return [spec["index"] for _, spec in self.feature_specs.items() if spec["type"] in feature_type]

    def load(self, json_file):
        logging.info("Load feature_map from json: " + json_file)
        with io.open(json_file, "r", encoding="utf-8") as fd:
            feature_map = json.load(fd, object_pairs_hook=OrderedDict)
        if feature_map["dataset_id"] != self.dataset_id:
            raise RuntimeError("dataset_id={} does not match to feature_map!".format(self.dataset_id))
        self.num_fields = feature_map["num_fields"]
        self.num_features = feature_map.get("num_features", None)
        self.input_length = feature_map.get("input_length", None)
        self.feature_specs = OrderedDict(feature_map["feature_specs"])

    def save(self, json_file):
        logging.info("Save feature_map to json: " + json_file)
        os.makedirs(os.path.dirname(json_file), exist_ok=True)
        feature_map = OrderedDict()
        feature_map["dataset_id"] = self.dataset_id
        feature_map["num_fields"] = self.num_fields
        feature_map["num_features"] = self.num_features
        feature_map["input_length"] = self.input_length
        feature_map["feature_specs"] = self.feature_specs
        with open(json_file, "w") as fd:
            json.dump(feature_map, fd, indent=4)


class FeatureEncoder(object):
    """csv -> id arrays + feature_map.json (reference fuxictr/features.py:93-337, same constructor, same
    methods, same feature_map / vocabulary / array contents -- pinned by tests/golden/encoder_*.json, which the reference's
    own FeatureEncoder produced).  Offline host-side preparation; the per-column work is done by the array-based Tokenizer of
    fuxictr/preprocess.py.  Not supported, as everywhere else in this package: pretrained embeddings."""

    def __init__(self, feature_cols=[], label_col={}, dataset_id=None, data_root="../data/", version="pytorch",
                 **kwargs):
        logging.info("Set up feature encoder...")
        self.data_dir = os.path.join(data_root, dataset_id)
        self.pickle_file = os.path.join(self.data_dir, "feature_encoder.pkl")
        self.json_file = os.path.join(self.data_dir, "feature_map.json")
        self.feature_cols = self._complete_feature_cols(feature_cols)
        self.label_col = label_col
        self.version = version
        self.feature_map = FeatureMap(dataset_id, self.data_dir, version)
        self.encoders = dict()

    @staticmethod
    def _complete_feature_cols(feature_cols):
        """a column spec whose `name` is a list stands for one spec per name (reference :111-122)"""
        full = []
        for col in feature_cols:
            names = col["name"]
            if isinstance(names, list):
                for n in names:
                    c = dict(col)
                    c["name"] = n
                    full.append(c)
            else:
                full.append(col)
        return full

    # ------------------------------------------------------------------ csv
    @staticmethod
    def _dtype_of(spec):
        d = spec["dtype"]
        return {"str": str, "float": float, "int": int}.get(d, d) if isinstance(d, str) else d

    def read_csv(self, data_path):
        import pandas as pd
        assert isinstance(data_path, (list, str)), "require a string filename or a list of string filenames"
        dtypes = dict((c["name"], self._dtype_of(c)) for c in self.feature_cols + [self.label_col])
        paths = data_path if isinstance(data_path, list) else [data_path]
        logging.info("Reading file: " + ", ".join(paths))
        frames = [pd.read_csv(p, dtype=dtypes, memory_map=True) for p in paths]
        return frames[0] if len(frames) == 1 else pd.concat(frames, ignore_index=True)

    def preprocess(self, ddf, fill_na=True):
        """fill missing values, run the per-column `preprocess` hooks (label first, then the feature columns in REVERSE
        order, as the reference does: a hook may read a column that a later spec rewrites), keep the active columns."""
        logging.info("Preprocess feature columns...")
        for col in [self.label_col] + self.feature_cols[::-1]:
            name = col["name"]
            if fill_na and name in ddf.columns and ddf[name].isnull().values.any():
                ddf[name] = self._fill_na(col, ddf[name])
            if col.get("preprocess", "") != "":
                ddf[name] = getattr(self, col["preprocess"])(ddf, name)
        active = [self.label_col["name"]] + [c["name"] for c in self.feature_cols if c["active"]]
        return ddf.loc[:, active]

    @staticmethod
    def _fill_na(col, series):
        na_value = col.get("na_value")
        if na_value is not None:
            return series.fillna(na_value)
        if col["dtype"] in ["str", str]:
            return series.fillna("")
        raise RuntimeError("Feature column={} requires to assign na_value!".format(col["name"]))

    # ------------------------------------------------------------------ fit
    def fit_transform(self, ddf, min_categr_count=1, num_buckets=10, **kwargs):
        self.fit(ddf, min_categr_count=min_categr_count, num_buckets=num_buckets, **kwargs)
        return self.transform(ddf)

    def fit(self, ddf, min_categr_count=1, num_buckets=10, **kwargs):
        logging.info("Fit feature encoder...")
        self.feature_map.num_fields = 0
        for col in self.feature_cols:
            if col["active"]:
                logging.info("Processing column: {}".format(col))
                self.fit_feature_col(col, ddf[col["name"]].values, min_categr_count=min_categr_count, num_buckets=num_buckets)
                self.feature_map.num_fields += 1
        self.feature_map.set_feature_index()
        self.save_pickle(self.pickle_file)
        self.feature_map.save(self.json_file)
        logging.info("Set feature encoder done.")

    def fit_feature_col(self, feature_column, feature_values, min_categr_count=1, num_buckets=10):
        from .preprocess import Normalizer, Tokenizer
        name, ftype = feature_column["name"], feature_column["type"]
        spec = self.feature_map.feature_specs[name] = {"source": feature_column.get("source", ""), "type": ftype}
        if "min_categr_count" in feature_column:
            min_categr_count = feature_column["min_categr_count"]
            spec["min_categr_count"] = min_categr_count
        if "embedding_dim" in feature_column:
            spec["embedding_dim"] = feature_column["embedding_dim"]
        if "pretrained_emb" in feature_column:
            raise NotImplementedError("pretrained embeddings (feature column {}) are not supported".format(name))
        if ftype == "numeric":
            if feature_column.get("normalizer") is not None:
                normalizer = Normalizer(feature_column["normalizer"])
                normalizer.fit(feature_values)
                self.encoders[name + "_normalizer"] = normalizer
            self.feature_map.num_features += 1
        elif ftype == "categorical":
            encoder = feature_column.get("encoder", "")
            if encoder != "":
                spec["encoder"] = encoder
            if encoder == "":
                tokenizer = Tokenizer(min_freq=min_categr_count, na_value=feature_column.get("na_value", ""))
                if "share_embedding" in feature_column:
                    spec["share_embedding"] = feature_column["share_embedding"]
                    tokenizer.set_vocab(self.encoders["{}_tokenizer".format(feature_column["share_embedding"])].vocab)
                else:       # a table shared with a sequence feature needs the padding row
                    tokenizer.fit_on_texts(feature_values, use_padding=self.is_share_embedding_with_sequence(name))
                if tokenizer.use_padding:
                    spec["padding_idx"] = tokenizer.vocab_size - 1
                self.encoders[name + "_tokenizer"] = tokenizer
                self.feature_map.num_features += tokenizer.vocab_size
                spec["vocab_size"] = tokenizer.vocab_size
            else:
                # "numeric_bucket" / "hash_bucket": the reference fits them but raises NotImplementedError when it has to
                # transform them (features.py:283-286), so no dataset can be built with them there either
                raise NotImplementedError("categorical encoder={}".format(encoder))
        elif ftype == "sequence":
            tokenizer = Tokenizer(min_freq=min_categr_count, splitter=feature_column.get("splitter", " "),
                                  na_value=feature_column.get("na_value", ""), max_len=feature_column.get("max_len", 0),
                                  padding=feature_column.get("padding", "post"))
            if "share_embedding" in feature_column:
                if feature_column.get("max_len") is None:
                    tokenizer.fit_on_texts(feature_values, use_padding=True)      # only to learn max_len
                spec["share_embedding"] = feature_column["share_embedding"]
                tokenizer.set_vocab(self.encoders["{}_tokenizer".format(feature_column["share_embedding"])].vocab)
            else:
                tokenizer.fit_on_texts(feature_values, use_padding=True)
            self.encoders[name + "_tokenizer"] = tokenizer
            self.feature_map.num_features += tokenizer.vocab_size
            spec.update({"encoder": feature_column.get("encoder", "MaskedAveragePooling"),
                         "padding_idx": tokenizer.vocab_size - 1, "vocab_size": tokenizer.vocab_size,
                         "max_len": tokenizer.max_len})
        else:
            raise NotImplementedError("feature_col={}".format(feature_column))

    # ------------------------------------------------------------------ transform
    def transform(self, ddf):
        """[n, input_length + 1] float64: id columns in feature_map order (a sequence feature = max_len columns), label last"""
        import numpy as np
        logging.info("Transform feature columns...")
        arrays = []
        for feature, spec in self.feature_map.feature_specs.items():
            ftype = spec["type"]
            if ftype == "numeric":
                arr = ddf.loc[:, feature].fillna(0).astype(float).values
                normalizer = self.encoders.get(feature + "_normalizer")
                arrays.append(normalizer.normalize(arr) if normalizer else arr)
            elif ftype == "categorical":
                if spec.get("encoder", "") != "":
                    raise NotImplementedError("transform of categorical encoder={} (not implemented by the reference either, "
                                              "features.py:283-286)".format(spec["encoder"]))
                arrays.append(self.encoders[feature + "_tokenizer"].encode_category(ddf.loc[:, feature].values))
            elif ftype == "sequence":
                arrays.append(self.encoders[feature + "_tokenizer"].encode_sequence(ddf.loc[:, feature].values))
        arrays.append(ddf.loc[:, self.label_col["name"]].astype(np.float64).values)          # the label column last
        return np.hstack([a.reshape(-1, 1) if a.ndim == 1 else a for a in arrays]).astype(np.float64, copy=False)

    def is_share_embedding_with_sequence(self, feature):
        return any(c.get("share_embedding") == feature and c["type"] == "sequence" for c in self.feature_cols)

    # ------------------------------------------------------------------ persistence
    def load_pickle(self, pickle_file=None):
        import pickle
        pickle_file = self.pickle_file if pickle_file is None else pickle_file
        logging.info("Load feature_encoder from pickle: " + pickle_file)
        if os.path.exists(pickle_file):
            with open(pickle_file, "rb") as fd:
                enc = pickle.load(fd)
            if enc.feature_map.dataset_id == self.feature_map.dataset_id:
                enc.version = self.version
                return enc
        raise IOError("pickle_file={} not valid.".format(pickle_file))

    def save_pickle(self, pickle_file):
        import pickle
        logging.info("Pickle feature_encoder: " + pickle_file)
        os.makedirs(os.path.dirname(pickle_file), exist_ok=True)
        with open(pickle_file, "wb") as fd:
            pickle.dump(self, fd)

    def load_json(self, json_file):
        self.feature_map.load(json_file)
