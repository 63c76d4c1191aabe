#!/bin/bash
# round-2 final measurements on one B200: GPU suite, ncu launch list, ncu --set full of the hot kernels (+ DRAM-traffic
# database), smoke, the default bench line (after the traffic update) and the reference arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f_pytest_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -3 gpurun_out/f_pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/f_smoke.log; cat gpurun_out/f_smoke.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02b.csv python tools/prof_kernels.py kkbox 4096 3 2>&1 | tail -2 ) > gpurun_out/f_ncu_launches.log
python tools/summarize_launches.py gpurun_out/launches_r02b.csv > gpurun_out/r02b_launches_train_kkbox.md 2>&1; head -12 gpurun_out/r02b_launches_train_kkbox.md
timeout 900 ncu --set full --clock-control none -k regex:"k_gather_flat|k_segment_scan|k_fixup_items|k_attn_fwd_rr|k_attn_bwd_rr|k_ff_fwd_rr|k_ff_bwd_rr|k_gemm_tc|k_adam|k_bn_act_fwd_cl|k_bn_act_bwd_cl|k_head_bwd_cl" -c 48 -f -o /tmp/prof_r02b_step python tools/prof_kernels.py kkbox 4096 1 2>&1 | tail -2
python tools/ncu_summary.py /tmp/prof_r02b_step.ncu-rep > gpurun_out/r02b_step_kernels_ncu_full.txt 2>&1
SRC="www24-rat_b200/csrc"
python tools/ncu_traffic.py attn_bwd /tmp/prof_r02b_step.ncu-rep k_attn_bwd_rr kkbox 4096 5 $SRC/encoder_rr_bwd.cu $SRC/encoder_rr.cuh
python tools/ncu_traffic.py gather /tmp/prof_r02b_step.ncu-rep k_gather_flat kkbox 4096 5 $SRC/gather.cu
NCU_CALLS=1 python tools/ncu_traffic.py scatter /tmp/prof_r02b_step.ncu-rep "k_segment_scan|k_fixup_items" kkbox 4096 5 $SRC/scatter.cu
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/f_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
cut -c1-300 gpurun_out/f_bench_ref.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/f_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "serial", d["e2e"]["serial"]["value"], "infer", d["infer"]["value"], d["infer"]["e2e"]["value"], "launches", d["gpu_launches"])
print(json.dumps(d["roofline"])[:500])
PY
