#!/bin/bash
# bench (both arms) + ncu launch list of 3 training steps
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 900 python bench.py 2>&1 | tail -3 ) > gpurun_out/bench_r02.log
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -3 ) > gpurun_out/bench_r02_ref.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02.csv python tools/prof_kernels.py kkbox 4096 3 2>&1 | tail -3 ) > gpurun_out/ncu_launches.log
cat gpurun_out/smoke.log; head -c 400 gpurun_out/bench_r02.log; echo; du -sh gpurun_out
