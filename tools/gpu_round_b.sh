#!/bin/bash
# one gpurun call: full GPU suite with the deferred reductions (default on), then A/B/C of the stream options
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/b_pytest_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -3 gpurun_out/b_pytest_gpu.log
RAT_DNN_SIDE=1 timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/b_pytest_side.log 2>&1
echo "side-stream suite rc=$?"; tail -3 gpurun_out/b_pytest_side.log
run() { # name, env...
  local n=$1; shift
  env "$@" timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/b_bench_$n.json 2> gpurun_out/b_bench_$n.err
}
run base RAT_DEFER_REDUCE=0 RAT_DNN_SIDE=0
run defer RAT_DEFER_REDUCE=1 RAT_DNN_SIDE=0
run side RAT_DEFER_REDUCE=1 RAT_DNN_SIDE=1
run sideonly RAT_DEFER_REDUCE=0 RAT_DNN_SIDE=1
python - <<'PY'
import json
for n in ("base", "defer", "side", "sideonly"):
    try:
        d = json.loads(open(f"gpurun_out/b_bench_{n}.json").read().strip().splitlines()[-1])
        k = d["kernels"]
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["infer"]["value"], d.get("gpu_launches"),
              "bn_fwd", k.get("rat_bn_act_fwd_train", {}).get("ms_per_step"), "bn_bwd", k.get("rat_bn_act_bwd_fused", {}).get("ms_per_step"))
    except Exception as e:
        print(n, "failed", e)
PY
