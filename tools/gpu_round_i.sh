#!/bin/bash
# final verification of the round: GPU suite, smoke, default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/i_pytest_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -3 gpurun_out/i_pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/i_smoke.log; cat gpurun_out/i_smoke.log
timeout 900 python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/i_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "infer", d["infer"]["value"], d["infer"]["e2e"]["value"], "launches", d["gpu_launches"])
for k, v in d["kernels"].items(): print(f"  {k:28s} {v['calls_per_step']:3d} {v['ms_per_step']:.4f}")
for k in ("roofline", "roofline_gather", "roofline_scatter"): print(k, d[k]["frac"], d[k]["avg_launch_ms"], d[k]["traffic"])
PY
