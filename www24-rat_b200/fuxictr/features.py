"""FeatureMap: the schema object consumed by the gather (reference: fuxictr/features.py:36-90; same JSON layout).
FeatureEncoder (csv -> ids, pandas-bound offline preparation) is out of the hot-path scope."""
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
    """csv -> id encoding is offline data preparation and outside the B200 hot path.
    The constructor keeps the attributes run_expid.py reads so that a prepared dataset directory
    (feature_map.json + h5/npz blocks) can be used through the reference's csv branch as well."""

    def __init__(self, feature_cols=[], label_col={}, dataset_id=None, data_root="../data/", version="pytorch",
                 **kwargs):
        self.data_dir = os.path.join(data_root, dataset_id)
        self.pickle_file = os.path.join(self.data_dir, "feature_encoder.pkl")
        self.json_file = os.path.join(self.data_dir, "feature_map.json")
        self.feature_cols = feature_cols
        self.label_col = label_col
        self.version = version
        self.feature_map = FeatureMap(dataset_id, self.data_dir, version)

    def fit(self, *a, **k):
        raise NotImplementedError("csv preprocessing is out of scope of the B200 hot path; prepare the dataset "
                                  "(feature_map.json + h5/npz) with the reference's FeatureEncoder")

    transform = read_csv = preprocess = fit
