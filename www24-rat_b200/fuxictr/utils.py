"""Config / logging helpers with the reference's semantics (reference: fuxictr/utils.py:26-104)."""
import glob
import json
import logging
import logging.config
import os
from collections import OrderedDict

import yaml


def load_config(config_dir, experiment_id):
    """model_config.yaml (Base U expid) U dataset config; dataset keys win (reference utils.py:26-52)."""
    model_configs = glob.glob(os.path.join(config_dir, "model_config.yaml"))
    if not model_configs:
        model_configs = glob.glob(os.path.join(config_dir, "model_config/*.yaml"))
    if not model_configs:
        raise RuntimeError("config_dir={} is not valid!".format(config_dir))
    found = dict()
    for config in model_configs:
        with open(config, "r") as cfg:
            config_dict = yaml.load(cfg, Loader=yaml.FullLoader)
        if "Base" in config_dict:
            found["Base"] = config_dict["Base"]
        if experiment_id in config_dict:
            found[experiment_id] = config_dict[experiment_id]
        if len(found) == 2:
            break
    if experiment_id not in found:
        raise ValueError("expid={} not found in config".format(experiment_id))
    params = dict()
    params.update(found.get("Base", {}))
    params.update(found.get(experiment_id))
    params["model_id"] = experiment_id
    params.update(load_dataset_config(config_dir, params["dataset_id"]))
    return params


def load_dataset_config(config_dir, dataset_id):
    """reference utils.py:55-64; additionally searches <config_dir>/../datasets/*.yaml, where the reference ships
    its dataset YAMLs (the shipped tree never finds them, SURVEY.md 5 'Config')."""
    candidates = glob.glob(os.path.join(config_dir, "dataset_config.yaml"))
    if not candidates:
        candidates = glob.glob(os.path.join(config_dir, "dataset_config/*.yaml"))
    candidates += glob.glob(os.path.join(config_dir, "..", "datasets", "*.yaml"))
    candidates += glob.glob(os.path.join(config_dir, "..", "..", "datasets", "*.yaml"))
    for config in candidates:
        with open(config, "r") as cfg:
            config_dict = yaml.load(cfg, Loader=yaml.FullLoader)
        if config_dict and dataset_id in config_dict:
            return config_dict[dataset_id]
    raise RuntimeError("dataset_id={} is not found in config.".format(dataset_id))


def set_logger(params, log_file=None):
    if log_file is None:
        log_dir = os.path.join(params["model_root"], params["dataset_id"])
        log_file = os.path.join(log_dir, params["model_id"] + ".log")
    os.makedirs(os.path.dirname(log_file), exist_ok=True)
    for handler in logging.root.handlers[:]:
        logging.root.removeHandler(handler)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s P%(process)d %(levelname)s %(message)s",
                        handlers=[logging.FileHandler(log_file, mode="w"), logging.StreamHandler()])


def print_to_json(data, sort_keys=True):
    new_data = dict((k, str(v)) for k, v in data.items())
    if sort_keys:
        new_data = OrderedDict(sorted(new_data.items(), key=lambda x: x[0]))
    return json.dumps(new_data, indent=4)


def print_to_list(data):
    return " - ".join("{}: {:.6f}".format(k, v) for k, v in data.items())


class Monitor(object):
    """weighted sum of validation metrics (reference utils.py:94-104)."""

    def __init__(self, kv):
        if isinstance(kv, str):
            kv = {kv: 1}
        self.kv_pairs = kv

    def get_value(self, logs):
        value = 0
        for k, v in self.kv_pairs.items():
            value += logs.get(k, 0) * v
        return value
