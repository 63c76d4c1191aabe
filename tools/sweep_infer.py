"""BASELINE configs[4]: K-sweep (retrieved neighbours) x batch-size sweep of RAT_m2 inference on one GPU (kkbox shape).
   python tools/sweep_infer.py [shape]   -> markdown table of samples/s (device-resident inputs, CUDA events)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
import numpy as np, torch
from rat_native import shapes
from fuxictr.pytorch import models
from fuxictr.pytorch.data_generator import DeviceDataGenerator

shape = sys.argv[1] if len(sys.argv) > 1 else "kkbox"
Ks = [1, 2, 4, 5, 8, 16, 32, 64]
Bs = [256, 1024, 4096, 16384, 65536]
fm = shapes.make_feature_map(shape)
rows_pool = 400_000
pool = shapes.synthetic_array(fm.feature_specs, rows_pool, seed=1)
print(f"# RAT_m2 {shape} shape, inference samples/s on 1 x B200 (fp16 tensor-core mode; K > 15 => cross-attention sequences "
      f"longer than 16 tokens use the long-sequence core of k_attn_fwd_tc)\n")
print("| K \\\\ B | " + " | ".join(str(b) for b in Bs) + " |")
print("|---|" + "---|" * len(Bs))
for K in Ks:
    params = shapes.model_params(shape, K=K, gpu=0)
    os.makedirs(os.path.join(params["model_root"], fm.dataset_id), exist_ok=True)
    model = models.RAT_m2(fm, **params)
    model.eval()
    nbr = shapes.synthetic_neighbours(rows_pool, rows_pool, K, seed=1)
    cells = []
    for B in Bs:
        if B * (K + 1) * (fm.num_fields + 1) * params["embedding_dim"] * 4 * 3 > 40e9:
            cells.append("-"); continue
        gen = DeviceDataGenerator(pool, pool, nbr, batch_size=B, shuffle=True, device="cuda:0", seed=3)
        it = iter(gen)
        batches = [next(it) for _ in range(min(6, len(gen)))]
        with torch.no_grad():
            for i in range(3): model.forward(batches[i % len(batches)])
            torch.cuda.synchronize()
            n = 10 if B >= 16384 else 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n): model.forward(batches[i % len(batches)])
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        cells.append(f"{B / ms * 1e3:,.0f}")
        model._engine._ws.clear()
        model._engine._graphs.clear()
        torch.cuda.empty_cache()
    print(f"| {K} | " + " | ".join(cells) + " |", flush=True)
    del model
    torch.cuda.empty_cache()
