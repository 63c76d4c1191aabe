"""Shared test helpers: load golden fixtures into oracle structures."""
import json
import os
from collections import OrderedDict

import numpy as np
import torch

from oracle import rat_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES_M2 = ["ml_small", "kkbox_small", "tmall_small"]
CASES_VAR = ["rat_m0_small", "rat_m1_small", "rat_m3_small"]
# the variants on the kkbox schema (BASELINE configs[3]: S = 84 flat / 14 + 6 / head width 20)
CASES_VAR_KKBOX = ["rat_m0_kkbox", "rat_m1_kkbox", "rat_m3_kkbox"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    feats = [O.Feature(n, t, v, max_len=ml) for n, t, v, ml in meta["feats"]]
    hp = dict(meta["hp"])
    hp["dnn_hidden_units"] = tuple(hp["dnn_hidden_units"])
    hp.pop("net_regularizer", None)
    spec = O.ModelSpec(features=feats, model=meta["model"], **hp)
    sd0 = OrderedDict((k[4:], torch.from_numpy(z[k].copy())) for k in z.files if k.startswith("sd0/"))
    sd2 = OrderedDict((k[4:], torch.from_numpy(z[k].copy())) for k in z.files if k.startswith("sd2/"))
    grad1 = OrderedDict((k[6:], torch.from_numpy(z[k].copy())) for k in z.files if k.startswith("grad1/"))
    return dict(name=name, spec=spec, meta=meta, z=z, sd0=sd0, sd2=sd2, grad1=grad1,
                X=torch.from_numpy(z["X"].copy()), y=torch.from_numpy(z["y"].copy()))


def split_state(sd):
    """state_dict -> (params, buffers) in oracle form; m3 alias keys are dropped."""
    params, bufs = OrderedDict(), OrderedDict()
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
            bufs[k] = v.clone()
        elif ".fn.W_" in k:          # RAT_m3 aliases of encoder.encoder.l.W_* 
            continue
        else:
            params[k] = v.clone()
    return params, bufs


def add_dead_params(params, spec):
    """query_proj is excluded from fixtures (dead parameter); restore zeros for counting."""
    FD = spec.num_fields * spec.embedding_dim
    params["query_proj.weight"] = torch.zeros(FD, FD)
    params["query_proj.bias"] = torch.zeros(FD)
    return params


def noise_grad_param(name, spec):
    """Linear biases that feed a train-mode BatchNorm have an exactly-zero true gradient; what autograd
    returns is rounding noise, which Adam normalises to +-lr steps.  Such parameters cannot be compared
    tightly after an optimizer step (neither between torch versions nor between CPU and GPU)."""
    if not spec.batch_norm:
        return False
    layers, _ = O.dnn_layout(spec)
    return any(name == f"dnn.dnn.{lin}.bias" for lin, bn in layers if bn is not None)
