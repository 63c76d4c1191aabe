"""Config / logging helpers behind run_expid.py (same call signatures and results as the reference's fuxictr/utils.py:
load_config :26-52, load_dataset_config :55-64, set_logger :67-80, print_to_json / print_to_list :83-92, Monitor :94-104)."""
import glob
import json
import logging
import os

import yaml


def _read_yaml(path):
    with open(path, "r") as fh:
        return yaml.load(fh, Loader=yaml.FullLoader) or {}


def _yaml_files(config_dir, stem):
    """`<config_dir>/<stem>.yaml`, else every yaml under `<config_dir>/<stem>/`."""
    single = os.path.join(config_dir, stem + ".yaml")
    return [single] if os.path.isfile(single) else sorted(glob.glob(os.path.join(config_dir, stem, "*.yaml")))


def load_config(config_dir, experiment_id):
    """params = Base section, overridden by the expid section, overridden by the dataset section (dataset keys win)."""
    files = _yaml_files(config_dir, "model_config")
    if not files:
        raise RuntimeError("config_dir={} is not valid!".format(config_dir))
    base, expid = None, None
    for path in files:
        sections = _read_yaml(path)
        base = sections.get("Base", base)
        expid = sections.get(experiment_id, expid)
        if base is not None and expid is not None:
            break
    if expid is None:
        raise ValueError("expid={} not found in config".format(experiment_id))
    params = dict(base or {})
    params.update(expid)
    params["model_id"] = experiment_id
    params.update(load_dataset_config(config_dir, params["dataset_id"]))
    return params


def load_dataset_config(config_dir, dataset_id):
    """Looks where the reference looks (dataset_config.yaml / dataset_config/*.yaml) and, in addition, in
    `<config_dir>/../datasets` and `<config_dir>/../../datasets`, where the reference actually ships its dataset
    YAMLs (its own loader never finds them, SURVEY.md 5 'Config')."""
    search = _yaml_files(config_dir, "dataset_config")
    for up in ("..", os.path.join("..", "..")):
        search += sorted(glob.glob(os.path.join(config_dir, up, "datasets", "*.yaml")))
    for path in search:
        sections = _read_yaml(path)
        if dataset_id in sections:
            return sections[dataset_id]
    raise RuntimeError("dataset_id={} is not found in config.".format(dataset_id))


def set_logger(params, log_file=None):
    """INFO log to `<model_root>/<dataset_id>/<model_id>.log` (truncated) and to the console."""
    if log_file is None:
        log_file = os.path.join(params["model_root"], params["dataset_id"], params["model_id"] + ".log")
    os.makedirs(os.path.dirname(log_file), exist_ok=True)
    root = logging.getLogger()
    for old in list(root.handlers):
        root.removeHandler(old)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s P%(process)d %(levelname)s %(message)s",
                        handlers=[logging.FileHandler(log_file, mode="w"), logging.StreamHandler()])


def print_to_json(data, sort_keys=True):
    """every value stringified, 4-space indented JSON (keys sorted by default)."""
    keys = sorted(data) if sort_keys else list(data)
    return json.dumps({k: str(data[k]) for k in keys}, indent=4)


def print_to_list(data):
    return " - ".join("{}: {:.6f}".format(name, value) for name, value in data.items())


class Monitor(object):
    """Early-stopping / checkpoint monitor: a weighted sum of validation metrics, e.g. {"AUC": 1, "logloss": -1};
    a bare metric name means weight 1; missing metrics count as 0."""

    def __init__(self, kv):
        self.kv_pairs = {kv: 1} if isinstance(kv, str) else kv

    def get_value(self, logs):
        return sum(weight * logs.get(name, 0) for name, weight in self.kv_pairs.items())
