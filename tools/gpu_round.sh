#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -20 ) > gpurun_out/smoke.log
( timeout 600 python bench.py 2>&1 | tail -30 ) > gpurun_out/bench.log
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -10 ) > gpurun_out/bench_ref.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python tools/prof_kernels.py kkbox 4096 3 2>&1 | tail -5 ) > gpurun_out/ncu_launches.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.log
