import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
import torch, rat_native as rn
from rat_native.engine import set_precision
set_precision("fp16")
dev="cuda:0"
B=4096
for (M,N,K,ta,tb) in [(4096,400,520,0,0),(4096,520,400,0,1),(400,520,4096,1,1)]:
    A=torch.randn((K,M) if ta else (M,K),device=dev); Bm=torch.randn((K,N) if tb else (N,K),device=dev); C=torch.empty(M,N,device=dev)
    nb=int(rn.query("rat_sgemm_workspace_bytes",M,N,K)); ws=torch.empty(max(nb//4,4),device=dev)
    lda=A.shape[1]; ldb=Bm.shape[1]
    def f(): rn.call("rat_sgemm",A,Bm,C,None,M,N,K,lda,ldb,N,ta,tb,ws,ws.numel()*4,rn.current_stream())
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    print(f"M={M} N={N} K={K} ta={ta} tb={tb}: {e0.elapsed_time(e1)/50*1e3:.1f} us")
