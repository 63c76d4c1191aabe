"""Quick per-kernel timing of the forward path at a BASELINE shape (dev tool; not the bench)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
import numpy as np, torch
from oracle import rat_oracle as O
from tests.gpu_util import make_engine, rand_params_nontrivial
import rat_native as rn

shape = sys.argv[1] if len(sys.argv) > 1 else "kkbox"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
prec = sys.argv[4] if len(sys.argv) > 4 else "tf32"
from rat_native.engine import set_precision
set_precision(prec)
spec = O.shape_spec(shape)
params = O.init_params(spec, 0)
eng = make_engine(spec, params, O.init_buffers(spec))
pool = O.synthetic_pool(spec, 100000, seed=1)
nbr = O.synthetic_neighbours(B, 100000, K, seed=1)
X, y = O.assemble_batch(pool[:B], pool, nbr, np.arange(B))
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
ws = eng.load_wire(Xd, yd, False)
T = K + 1

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t_all = timeit(lambda: eng.forward_ids(ws, B, T, False))
print(f"{shape} B={B} K={K} precision={prec}: forward {t_all:.3f} ms  -> {B / t_all * 1e3:,.0f} samples/s")
s = spec; N = s.num_fields + 1; D = s.embedding_dim
a = ws["acts"]
pre = "encoder.encoder.0."
st = rn.current_stream()
t = timeit(lambda: rn.call("rat_gather_fwd", eng.store.emb_W, eng.store.lr_W, eng.p["label_embedding_layer.weight"], ws["ids"], ws["labels"], eng.col_off, eng.col_vocab, eng.field_col0, eng.field_width, a[0], ws["x_emb"], ws["lr_out"], B, T, s.input_length, s.num_fields, D, 0.0, 1, 0, eng.err_flag, st))
gb = B * (K * 8 + T * s.input_length * 4 + T + T * s.input_length * D * 4 + s.input_length * 4 + T * N * D * 4 + s.num_fields * D * 4) / 1e9
print(f"  gather      {t*1e3:8.1f} us   {gb / (t * 1e-3):8.0f} GB/s algorithmic")
t = timeit(lambda: eng._attn(a[0], a[0], a[1], pre + "intra_attention.", 0, B, T, N)); print(f"  attn intra  {t*1e3:8.1f} us")
t = timeit(lambda: eng._attn(a[1], a[1], a[2], pre + "cross_attention.", 1, B, T, N)); print(f"  attn cross  {t*1e3:8.1f} us")
t = timeit(lambda: eng._ff(a[2], a[2], a[0], pre + "mlp.", B * T * N)); print(f"  ff          {t*1e3:8.1f} us")
t = timeit(lambda: eng._dnn_forward(ws, B, False)); print(f"  dnn         {t*1e3:8.1f} us")
