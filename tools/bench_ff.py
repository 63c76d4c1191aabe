"""Time rat_ff_fwd / rat_ff_bwd alone (CUDA events, rotating buffers larger than L2) at a dataset shape.
   python tools/bench_ff.py [kkbox|tmall|ml] [B] [K]      (RAT_RR=0 / RAT_RR_FF_OFF=1 select the tile kernels)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
import torch
import rat_native as rn
from rat_native.engine import set_precision

shape = sys.argv[1] if len(sys.argv) > 1 else "kkbox"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
F, D = {"kkbox": (13, 40), "tmall": (8, 10), "ml": (3, 10)}[shape]
rows, M = B * (K + 1) * (F + 1), 2 * D
set_precision("fp16")
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
nbuf = 5
xs = [torch.randn(rows, D, device=dev, generator=g) for _ in range(nbuf)]
outs = [torch.empty(rows, D, device=dev) for _ in range(nbuf)]
douts = [torch.randn(rows, D, device=dev, generator=g) * 1e-3 for _ in range(nbuf)]
w1 = torch.randn(M, D, device=dev, generator=g) * 0.3; b1 = 0.1 * torch.randn(M, device=dev, generator=g)
w2 = torch.randn(D, M, device=dev, generator=g) * 0.3; b2 = 0.1 * torch.randn(D, device=dev, generator=g)
gw1, gb1, gw2, gb2 = torch.zeros_like(w1), torch.zeros_like(b1), torch.zeros_like(w2), torch.zeros_like(b2)
amax_in = torch.zeros(1, device=dev); amax_out = torch.zeros(1, device=dev)
st = rn.current_stream()
bw = torch.empty(int(rn.query("rat_ff_bwd_workspace_bytes", rows, D, M)) // 4 + 4, device=dev)
amax_in.fill_(float(douts[0].abs().max()))


def fwd(i):
    rn.call("rat_ff_fwd", xs[i % nbuf], xs[i % nbuf], outs[i % nbuf], None, None, w1, b1, w2, b2, rows, D, M, st)


def bwd(i):
    rn.call("rat_ff_bwd", xs[i % nbuf], douts[i % nbuf], douts[i % nbuf], outs[i % nbuf], None, None, w1, b1, w2, gw1, gb1, gw2, gb2,
            None, None, rows, D, M, amax_in, amax_out, bw, bw.numel() * 4, st)


for name, fn, passes in (("fwd", fwd, 2), ("bwd", bwd, 3)):
    for i in range(6): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 40
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    mb = rows * D * 4 * passes / 1e6
    print(f"{shape} B={B} K={K} ff_{name}: {us:.1f} us/call  ({mb:.0f} MB algorithmic -> {mb / us * 1e3:.0f} GB/s)")
