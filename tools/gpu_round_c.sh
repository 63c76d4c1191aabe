#!/bin/bash
# one gpurun --gpus 2 call: 2-rank NCCL parity tests (in-kernel BatchNorm exchange), then the default bench line at N=2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/c_pytest_multi.log 2>&1
echo "multi suite rc=$?"; tail -5 gpurun_out/c_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/c_bench_g2.json 2> gpurun_out/c_bench_g2.err
echo "bench g2 rc=$?"; tail -c 600 gpurun_out/c_bench_g2.err
RAT_DNN_FUSED=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --no-cpu-baseline --no-secondary > gpurun_out/c_bench_g2_split.json 2> gpurun_out/c_bench_g2_split.err
python - <<'PY'
import json
for n in ("g2", "g2_split"):
    try:
        d = json.loads(open(f"gpurun_out/c_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["infer"]["value"], d.get("gpu_launches"))
        s = d.get("secondary", {})
        for k in ("strong", "tmall_sharded", "tmall_sharded_x50"):
            if k in s: print(" ", k, json.dumps(s[k])[:300])
    except Exception as e:
        print(n, "failed", e)
PY
