"""reference: fuxictr/pytorch/torch_utils.py:26-94 (seed / device / regulariser factories)."""
import os
import random

import numpy as np
import torch


def seed_everything(seed=1029):
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    torch.backends.cudnn.deterministic = True


def get_device(gpu=-1):
    if gpu >= 0 and torch.cuda.is_available():
        return torch.device("cuda:" + str(gpu))
    return torch.device("cpu")


def get_regularizer(reg):
    """float -> [(2, lambda)]; 'l2(x)' / 'l1(x)' / 'l1_l2(a,b)' strings (torch_utils.py:65-81)."""
    reg_pair = []
    if isinstance(reg, (float, int)) and not isinstance(reg, bool):
        if reg:
            reg_pair.append((2, float(reg)))
    elif isinstance(reg, str):
        try:
            if reg.startswith("l1(") or reg.startswith("l2("):
                reg_pair.append((int(reg[1]), float(reg.rstrip(")").split("(")[-1])))
            elif reg.startswith("l1_l2"):
                l1_reg, l2_reg = reg.rstrip(")").split("(")[-1].split(",")
                reg_pair.append((1, float(l1_reg)))
                reg_pair.append((2, float(l2_reg)))
            else:
                raise NotImplementedError
        except Exception:
            raise NotImplementedError("regularizer={} is not supported.".format(reg))
    return reg_pair


def l2_lambda(reg):
    """the fused Adam kernel supports the L2 form only (all shipped configs use a float)."""
    pairs = get_regularizer(reg)
    for p, _ in pairs:
        if p != 2:
            raise NotImplementedError("only L2 regularisation is implemented in the fused B200 optimizer")
    return float(sum(l for _, l in pairs))
