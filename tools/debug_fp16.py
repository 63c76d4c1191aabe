import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
import numpy as np, torch
from fuxictr.pytorch import models
from rat_native import shapes
from rat_native.engine import set_precision
mode = sys.argv[1] if len(sys.argv) > 1 else "fp16"
drop = float(sys.argv[2]) if len(sys.argv) > 2 else -1
set_precision(mode)
fm = shapes.make_feature_map("tmall", vocab_scale=0.001, data_dir="/tmp/dbg")
kw = {}
if drop >= 0: kw = dict(emb_dropout=drop, net_dropout=drop)
params = shapes.model_params("tmall", K=5, gpu=0, model_root="/tmp/dbg/exps", dnn_hidden_units=[64, 32], embedding_regularizer=1e-6, **kw)
os.makedirs(os.path.join(params["model_root"], fm.dataset_id), exist_ok=True)
model = models.RAT_m2(fm, **params)
pool = shapes.synthetic_array(fm.feature_specs, 3000, seed=3)
pool[:, -1] = (pool[:, 2] % 2 == 0).astype(np.float64)
nbr = shapes.synthetic_neighbours(3000, 3000, 5, seed=3)
model.train()
eng = model._engine
for i in range(30):
    rows = np.arange(i * 100, (i + 1) * 100) % 3000
    batch = tuple(torch.from_numpy(t) for t in shapes.host_wire_batch(pool, pool, nbr, rows))
    loss = float(model.train_step(batch))
    ws = list(eng._ws.values())[0]
    bad = {k: int((~torch.isfinite(v)).sum()) for k, v in [("W", eng.store.W), ("G", eng.store.G), ("dact", ws["dact"]), ("dxemb", ws["dxemb"]), ("y_pred", ws["y_pred"])] }
    acts_bad = [int((~torch.isfinite(a)).sum()) for a in ws["acts"]]
    dh_bad = [int((~torch.isfinite(a)).sum()) for a in ws["dh"]]
    print(i, loss, "gradnorm", float(eng.opt_state[0]), bad, "acts", acts_bad, "dh", dh_bad, "max|dact|", float(ws["dact"].abs().max()), "max|dh|", [float(a.abs().max()) for a in ws["dh"]])
    if not np.isfinite(loss): break
