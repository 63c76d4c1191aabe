#!/bin/bash
# one gpurun call: new-kernel tests, full GPU suite, A/B of the fused DNN kernels, full default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_backward.py -x -q -m gpu -k "fused_cluster or head_bwd_cluster" > gpurun_out/a_pytest_new.log 2>&1
echo "new tests rc=$?"; tail -3 gpurun_out/a_pytest_new.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/a_pytest_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -3 gpurun_out/a_pytest_gpu.log
RAT_DNN_FUSED=0 timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/a_bench_split.json 2> gpurun_out/a_bench_split.err
RAT_DNN_FUSED=1 timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/a_bench_fused.json 2> gpurun_out/a_bench_fused.err
python - <<'PY'
import json
for n in ("split", "fused"):
    try:
        d = json.loads(open(f"gpurun_out/a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["infer"]["value"], d.get("gpu_launches"))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 900 python bench.py > gpurun_out/a_bench_full.json 2> gpurun_out/a_bench_full.err
echo "full bench rc=$?"; cut -c1-400 gpurun_out/a_bench_full.json
