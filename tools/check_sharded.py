"""2+ GPU check of the row-sharded embedding mode (torchrun): sharded == replicated data-parallel on the same weights.
   torchrun --nproc-per-node 2 tools/check_sharded.py [shape] [B_per_gpu] [vocab_scale]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
import numpy as np, torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from rat_native import shapes
from rat_native.engine import set_precision
from fuxictr.pytorch import models
from fuxictr.pytorch.data_generator import DeviceDataGenerator

shape = sys.argv[1] if len(sys.argv) > 1 else "tmall"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
vscale = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
set_precision("fp16")
K = 5
fm = shapes.make_feature_map(shape, vocab_scale=vscale)


def build(shard):
    torch.manual_seed(2021)
    params = shapes.model_params(shape, K=K, gpu=local, emb_dropout=0.0, net_dropout=0.0)
    params["shard_embeddings"] = shard
    os.makedirs(os.path.join(params["model_root"], fm.dataset_id), exist_ok=True)
    return models.RAT_m2(fm, **params)


A = build(False)
dist.broadcast(A._engine.store.W, 0)
Bm = build(True)
sdA = {k: v.clone() for k, v in A.state_dict().items()}
Bm.load_state_dict(sdA)
sdB = Bm.state_dict()
for k in sdA:
    assert torch.equal(sdA[k], sdB[k]), f"state_dict round trip through the shards differs: {k}"
pool = shapes.synthetic_array(fm.feature_specs, 50000, seed=7)
nbr = shapes.synthetic_neighbours(50000, 50000, K, seed=7)
gen = DeviceDataGenerator(pool, pool, nbr, batch_size=B * world, shuffle=True, device=f"cuda:{local}", seed=11, rank=rank,
                          world=world, drop_last=True)
it = iter(gen)
batches = [next(it) for _ in range(4)]
A.eval(); Bm.eval()
ya = A.forward(batches[0])["y_pred"].clone()
yb = Bm.forward(batches[0])["y_pred"].clone()
assert torch.equal(ya, yb), f"sharded forward differs from replicated: max {float((ya - yb).abs().max())}"
A.train(); Bm.train()
for i in range(3):
    la = float(A.train_step(batches[i + 1]).item())
    lb = float(Bm.train_step(batches[i + 1]).item())
    assert abs(la - lb) <= 1e-5 * max(1.0, abs(la)), (i, la, lb)
A._engine.check_errors(); Bm._engine.check_errors()
sdA, sdB = A.state_dict(), Bm.state_dict()
worst = 0.0
for k in sdA:
    if sdA[k].dtype.is_floating_point:
        worst = max(worst, float((sdA[k].float() - sdB[k].float()).abs().max()))
assert worst < 2e-5, f"weights after 3 steps differ by {worst}"
# timing (tables resident, device batches)
def timeit(m, n=10):
    for i in range(3): m.train_step(batches[1 + i % 3])
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for i in range(n): m.train_step(batches[1 + i % 3])
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / n
ta, tb = timeit(A), timeit(Bm)
if rank == 0:
    print(f"check_sharded OK: {shape} x{vscale} vocab, world={world}, B={B}/GPU: forward bit-identical, 3 train steps: "
          f"loss equal, max |dW| {worst:.2e}; step replicated {ta*1e3:.3f} ms, row-sharded {tb*1e3:.3f} ms "
          f"({B*world/tb:,.0f} samples/s)", flush=True)
dist.destroy_process_group()
