#!/bin/bash
# quick GPU check: selected tests (-k expr) + a short bench; prints the roofline objects.  Usage: gpu_quick.sh "<pytest -k expr>" [bench args]
mkdir -p gpurun_out
K="$1"; shift
( timeout 900 python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -15 ) | tee gpurun_out/pytest_quick.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -3 ) > gpurun_out/bench_quick.log
python - <<'PY'
import json
for l in open("gpurun_out/bench_quick.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("train", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "infer", d["infer"]["value"])
        for k in d:
            if k.startswith("roofline"):
                r = d[k]; print(k, r["kernel"], r["achieved"], r["unit"], "frac", r["frac"], "ms", r.get("avg_launch_ms"))
        for k, v in list(d["kernels"].items())[:12]:
            print("  ", k, v)
    elif l.strip():
        print(l.rstrip()[-300:])
PY
