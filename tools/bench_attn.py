"""Time rat_attn_fwd / rat_attn_bwd alone (CUDA events, rotating buffers larger than L2) at a dataset shape.
   python tools/bench_attn.py [kkbox|tmall|ml] [B] [K]      (RAT_TC2=1 selects the second-generation forward kernel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
import torch
import rat_native as rn
from rat_native.engine import set_precision

shape = sys.argv[1] if len(sys.argv) > 1 else "kkbox"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
F, D, H = {"kkbox": (13, 40, 8), "tmall": (8, 10, 32), "ml": (3, 10, 2)}[shape]
T, N, dh = K + 1, F + 1, 10
I = H * dh
set_precision("fp16")
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
nbuf = 5
xs = [torch.randn(B, T, N, D, device=dev, generator=g) for _ in range(nbuf)]
outs = [torch.empty(B, T, N, D, device=dev) for _ in range(nbuf)]
douts = [torch.randn(B, T, N, D, device=dev, generator=g) * 1e-3 for _ in range(nbuf)]
lnw = 1 + 0.1 * torch.randn(D, device=dev, generator=g); lnb = 0.1 * torch.randn(D, device=dev, generator=g)
wqkv = torch.randn(3 * I, D, device=dev, generator=g) * 0.3
wo = torch.randn(D, I, device=dev, generator=g) * 0.2; bo = 0.1 * torch.randn(D, device=dev, generator=g)
gq = torch.zeros(I, D, device=dev); gk = torch.zeros(I, D, device=dev); gv = torch.zeros(I, D, device=dev)
gwo = torch.zeros(D, I, device=dev); gbo = torch.zeros(D, device=dev); glw = torch.zeros(D, device=dev); glb = torch.zeros(D, device=dev)
amax_in = torch.zeros(1, device=dev); amax_out = torch.zeros(1, device=dev)
st = rn.current_stream()
tokens = B * T * N
for mode in (0, 1):
    def fwd(i):
        rn.call("rat_attn_fwd", xs[i % nbuf], xs[i % nbuf], outs[i % nbuf], lnw, lnb, wqkv[:I], wqkv[I:2 * I], wqkv[2 * I:], wo, bo,
                B, T, N, D, H, dh, dh ** -0.5, 1.0, mode, st)
    nb = int(rn.query("rat_attn_bwd_workspace_bytes", B, T, N, D, H, dh, mode))
    bw = torch.empty(nb // 4 + 4, device=dev)
    amax_in.fill_(float(douts[0].abs().max()))
    def bwd(i):
        rn.call("rat_attn_bwd", xs[i % nbuf], douts[i % nbuf], douts[i % nbuf], outs[i % nbuf], lnw, lnb, wqkv[:I], wqkv[I:2 * I],
                wqkv[2 * I:], wo, gq, gk, gv, gwo, gbo, glw, glb, 0, B, T, N, D, H, dh, dh ** -0.5, 1.0, mode, amax_in, amax_out,
                bw, bw.numel() * 4, st)
    for name, fn, passes in ((("fwd", fwd, 2),) if os.environ.get("ONLY_FWD") else (("fwd", fwd, 2), ("bwd", bwd, 3))):
        for i in range(int(os.environ.get("WARM", "6"))): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = int(os.environ.get("ITERS", "40"))
        e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        mb = tokens * D * 4 * passes / 1e6
        print(f"{shape} B={B} K={K} mode={mode} attn_{name}: {us:.1f} us/call  ({mb:.0f} MB algorithmic -> {mb / us * 1e3:.0f} GB/s)"
              f"  gen={'2' if os.environ.get('RAT_TC2') == '1' else '1'}")
