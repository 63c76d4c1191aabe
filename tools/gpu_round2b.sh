#!/bin/bash
# ncu --set full captures of the hot kernels of one kkbox training step; only text summaries + the traffic database leave the box
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"k_gather_flat|k_segment_scan|k_fixup_items|k_attn_fwd_rr|k_attn_bwd_rr|k_ff_fwd_rr|k_ff_bwd_rr|k_gemm_tc|k_adam" -c 36 -f -o /tmp/prof_r02_step python tools/prof_kernels.py kkbox 4096 1 2>&1 | tail -3
python tools/ncu_summary.py /tmp/prof_r02_step.ncu-rep > gpurun_out/r02_step_kernels_ncu_full.txt 2>&1
SRC="www24-rat_b200/csrc"
python tools/ncu_traffic.py attn_bwd /tmp/prof_r02_step.ncu-rep k_attn_bwd_rr kkbox 4096 5 $SRC/encoder_rr_bwd.cu $SRC/encoder_rr.cuh
python tools/ncu_traffic.py gather /tmp/prof_r02_step.ncu-rep k_gather_flat kkbox 4096 5 $SRC/gather.cu
NCU_CALLS=1 python tools/ncu_traffic.py scatter /tmp/prof_r02_step.ncu-rep "k_segment_scan|k_fixup_items" kkbox 4096 5 $SRC/scatter.cu
timeout 600 ncu --set full --clock-control none -k regex:"k_gather_flat" -c 2 -f -o /tmp/prof_r02_gather_x20 python tools/prof_kernels.py kkbox 4096 1 fp16 20 2>&1 | tail -2
python tools/ncu_summary.py /tmp/prof_r02_gather_x20.ncu-rep > gpurun_out/r02_gather_x20_ncu_full.txt 2>&1
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
WARM=1 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_attn_fwd_rr|k_attn_bwd_rr" -c 4 -f -o gpurun_out/prof_rr3 python tools/bench_attn.py kkbox 4096 5 2>&1 | tail -2
ls -la gpurun_out; du -sh gpurun_out; head -40 gpurun_out/r02_step_kernels_ncu_full.txt
