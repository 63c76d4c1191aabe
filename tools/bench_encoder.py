"""csv -> id preparation: wall time of FeatureEncoder.fit + transform on a kkbox-like synthetic frame, this
package vs the reference (when /root/reference is importable: build container only).  CPU, no GPU involved.
    python tools/bench_encoder.py [rows]"""
import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_encoder as G

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = np.random.default_rng(0)
frame = G.kkbox_frame(rng, n)


def run(datasets_pkg, tag):
    with tempfile.TemporaryDirectory() as tmp:
        enc = datasets_pkg.kkbox.FeatureEncoder(feature_cols=[dict(c) for c in G.KKBOX_COLS], label_col=dict(G.LABEL),
                                                dataset_id="bench", data_root=tmp)
        df = enc.preprocess(frame.copy())
        t0 = time.perf_counter(); enc.fit(df, min_categr_count=2); t1 = time.perf_counter()
        arr = enc.transform(df); t2 = time.perf_counter()
    print(f"{tag}: rows {n}  fit {t1 - t0:.2f} s  transform {t2 - t1:.2f} s  -> {n / (t2 - t0):,.0f} rows/s  array {arr.shape}")
    return arr


if os.path.isdir("/root/reference") and "--ours-only" not in sys.argv:
    from make_golden import import_reference
    import_reference()
    from fuxictr import datasets as ref
    a_ref = run(ref, "reference")
    for k in [k for k in sys.modules if k == "fuxictr" or k.startswith("fuxictr.")]:
        del sys.modules[k]
    sys.path.remove("/root/reference")
else:
    a_ref = None
sys.path.insert(0, os.path.join(ROOT, "www24-rat_b200"))
from fuxictr import datasets as ours
a = run(ours, "this package")
if a_ref is not None:
    print("identical arrays:", bool(np.array_equal(a, a_ref)))
