"""Launch each hot kernel once at the kkbox BASELINE shape (for ncu captures)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
import numpy as np, torch
import rat_native as rn
from rat_native import shapes
from fuxictr.pytorch import models
from fuxictr.pytorch.data_generator import DeviceDataGenerator

shape = sys.argv[1] if len(sys.argv) > 1 else "kkbox"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
K = 5
from rat_native.engine import set_precision
set_precision(sys.argv[4] if len(sys.argv) > 4 else "fp16")
vocab_scale = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
fm = shapes.make_feature_map(shape, vocab_scale=vocab_scale)
params = shapes.model_params(shape, K=K, gpu=0)
os.makedirs(os.path.join(params["model_root"], fm.dataset_id), exist_ok=True)
model = models.RAT_m2(fm, **params)
pool = shapes.synthetic_array(fm.feature_specs, 200000, seed=1)
nbr = shapes.synthetic_neighbours(200000, 200000, K, seed=1)
gen = DeviceDataGenerator(pool, pool, nbr, batch_size=B, shuffle=True, device="cuda:0")
it = iter(gen)
model.train()
for _ in range(steps):
    model.train_step(next(it))
torch.cuda.synchronize()
print("done")
