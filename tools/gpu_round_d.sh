#!/bin/bash
# walk-form segment reduce: scatter / backward tests, dropout mask-replay test, A/B bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/d_pytest.log 2>&1
echo "backward+api suite rc=$?"; tail -5 gpurun_out/d_pytest.log
run() { local n=$1; shift
  env "$@" timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/d_bench_$n.json 2> gpurun_out/d_bench_$n.err; }
run scan RAT_SCAN_WALK=0
run walk RAT_SCAN_WALK=1
python - <<'PY'
import json
for n in ("scan", "walk"):
    try:
        d = json.loads(open(f"gpurun_out/d_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps(d["roofline_scatter"])[:400])
    except Exception as e:
        print(n, "failed", e)
PY
for sh in tmall ml; do
  for w in 0 1; do
    RAT_SCAN_WALK=$w timeout 300 python bench.py --shape $sh --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$sh walk=$w', d['value'], d['ms_per_step'], json.dumps(d['roofline_scatter'])[:300])"
  done
done
