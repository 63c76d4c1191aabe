#!/bin/bash
# deferred record reductions together with programmatic dependent launch: backward / API suites with the option on, A/B bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
RAT_DEFER_REDUCE=1 timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/k_pytest_defer.log 2>&1
echo "suites with RAT_DEFER_REDUCE=1 rc=$?"; tail -2 gpurun_out/k_pytest_defer.log
run() { local n=$1 v=$2
  RAT_DEFER_REDUCE=$v timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/k_bench_$n.json 2> gpurun_out/k_bench_$n.err; }
run nodefer 0
run defer 1
run nodefer2 0
run defer2 1
python - <<'PY'
import json
for n in ("nodefer", "defer", "nodefer2", "defer2"):
    try:
        d = json.loads(open(f"gpurun_out/k_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["infer"]["value"])
    except Exception as e:
        print(n, "failed", e)
PY
