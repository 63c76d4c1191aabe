#!/bin/bash
# Round-2 measurement pass in ONE gpurun call: GPU parity tests, smoke, both bench arms, ncu launch list, ncu --set full captures
# (no source import: the whole gpurun_out/ must stay under 64 MiB to be copied back).
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 900 python bench.py 2>&1 | tail -3 ) > gpurun_out/bench_r02.log
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -3 ) > gpurun_out/bench_r02_ref.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02.csv python tools/prof_kernels.py kkbox 4096 3 2>&1 | tail -3 ) > gpurun_out/ncu_launches.log
( timeout 900 ncu --set full --clock-control none -k regex:"k_gather_flat|k_segment_scan|k_fixup_items|k_attn_fwd_rr|k_attn_bwd_rr|k_ff_fwd_rr|k_ff_bwd_rr|k_gemm_tc|k_adam" -c 36 -f -o gpurun_out/prof_r02_step python tools/prof_kernels.py kkbox 4096 1 2>&1 | tail -3 ) > gpurun_out/ncu_full_step.log
( timeout 600 ncu --set full --clock-control none -k regex:"k_gather_flat" -c 2 -f -o gpurun_out/prof_r02_gather_x20 python tools/prof_kernels.py kkbox 4096 1 fp16 20 2>&1 | tail -3 ) > gpurun_out/ncu_full_gather_x20.log
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 600 gpurun_out/bench_r02.log; echo; ls -la gpurun_out | tail -12; du -sh gpurun_out
