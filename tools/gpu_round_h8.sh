#!/bin/bash
# 8 x B200: data-parallel bench line with the final code (in-kernel BatchNorm exchange at world 8), no secondary workloads
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --no-cpu-baseline --no-secondary > gpurun_out/h_bench_g8.json 2> gpurun_out/h_bench_g8.err
echo "bench g8 rc=$?"; grep -n "Error\|error" gpurun_out/h_bench_g8.err | head -5
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/h_bench_g8.json").read().strip().splitlines()[-1])
    print("g8", d["value"], d["ms_per_step"], d["e2e"]["value"], d["infer"]["value"], d.get("gpu_launches"))
except Exception as e:
    print("failed", e)
PY
