"""Time rat_gather_fwd alone (CUDA events, kkbox/tmall/ml shape): rotating output buffers (> L2) vs one buffer, dropout on/off."""
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
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
fm = shapes.make_feature_map(shape)
params = shapes.model_params(shape, K=K, gpu=0)
os.makedirs(os.path.join(params["model_root"], fm.dataset_id), exist_ok=True)
model = models.RAT_m2(fm, **params)
eng = model._engine
pool = shapes.synthetic_array(fm.feature_specs, 500000, seed=1)
nbr = shapes.synthetic_neighbours(500000, 500000, K, seed=1)
gen = DeviceDataGenerator(pool, pool, nbr, batch_size=B, shuffle=True, device="cuda:0")
it = iter(gen)
s = eng.spec
T, L, F, D = K + 1, s.L, s.F, s.embedding_dim
N = F + 1
nbuf = 6
blocks = [torch.empty(B, T, N, D, device="cuda") for _ in range(nbuf)]
batches = []
for _ in range(nbuf):
    b = next(it)
    ws = model._load_batch(b, training=False) if hasattr(model, "_load_batch") else None
    batches.append(b)
algo = B * (K * 8 + T * L * 4 + T + T * L * D * 4 + L * 4 + T * N * D * 4 + F * D * 4)

def run(drop, rotate, iters=60):
    ws = eng._workspace(B, T, False)
    st = rn.current_stream()
    def one(i):
        blk = blocks[i % nbuf] if rotate else blocks[0]
        rn.call("rat_gather_fwd", eng.store.emb_W, eng.store.lr_W, eng.p["label_embedding_layer.weight"], ws["ids"],
                ws["labels"], eng.col_off, eng.col_vocab, eng.field_col0, eng.field_width, blk, ws["x_emb"],
                ws["lr_out"], B, T, L, F, D, float(drop), 2021, 7, eng.err_flag, st)
    for i in range(10): one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): one(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{shape} B={B} K={K} drop={drop} rotate={rotate}: {us:.1f} us/launch  {algo/us/1e3:.0f} GB/s algorithmic ({algo/1e6:.1f} MB)")

# ids of one real batch in the workspace
model.eval()
model.forward(batches[0])
for drop in (0.0, 0.1):
    for rot in (False, True):
        run(drop, rot)

# ---- context: pure-write and copy streams of the same size as the gather's output (torch fill / copy kernels)
def ctx(name, fn, nbytes, iters=60):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"  [{name}] {us:.1f} us  {nbytes/us/1e3:.0f} GB/s")
nb = blocks[0].numel() * 4
ctx("fill %.0f MB (write only)" % (nb / 1e6), lambda i: blocks[i % nbuf].fill_(1.0), nb)
ctx("copy %.0f MB (read + write)" % (nb / 1e6), lambda i: blocks[i % nbuf].copy_(blocks[(i + 3) % nbuf]), 2 * nb)
big = [torch.empty(256 << 20, dtype=torch.float32, device="cuda") for _ in range(2)]
ctx("fill 1 GiB", lambda i: big[i % 2].fill_(1.0), big[0].numel() * 4, 10)
ctx("copy 1 GiB", lambda i: big[1].copy_(big[0]), big[0].numel() * 8, 10)
