#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/c_bench_g2.json 2> gpurun_out/c_bench_g2.err
echo "bench g2 rc=$?"; grep -n "File \"/\|Error" gpurun_out/c_bench_g2.err | head -30 | cut -c1-250
python - <<'PY'
import json
for n in ("g2",):
    try:
        d = json.loads(open(f"gpurun_out/c_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["infer"]["value"], d.get("gpu_launches"))
        s = d.get("secondary", {})
        for k in ("strong", "tmall_sharded", "tmall_sharded_x50", "variants"):
            if k in s: print(" ", k, json.dumps(s[k])[:400])
    except Exception as e:
        print(n, "failed", e)
PY
