#!/bin/bash
# programmatic dependent launch of the RAT-block kernels: full GPU suite + A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/e_pytest.log 2>&1
echo "gpu suite rc=$?"; tail -4 gpurun_out/e_pytest.log
run() { local n=$1; shift
  env "$@" timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/e_bench_$n.json 2> gpurun_out/e_bench_$n.err; }
run nopdl RAT_PDL=0
run pdl RAT_PDL=1
run nopdl2 RAT_PDL=0
run pdl2 RAT_PDL=1
python - <<'PY'
import json
for n in ("nopdl", "pdl", "nopdl2", "pdl2"):
    try:
        d = json.loads(open(f"gpurun_out/e_bench_{n}.json").read().strip().splitlines()[-1])
        k = d["kernels"]
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["infer"]["value"], "attn_bwd", k["rat_attn_bwd"]["ms_per_step"], "attn_fwd", k["rat_attn_fwd"]["ms_per_step"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/e_bench_{n}.err").read()[-800:])
PY
