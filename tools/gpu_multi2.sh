#!/bin/bash
# round-2 multi-GPU validation (gpurun --gpus N): NCCL parity tests + the default bench line under torchrun (as the driver runs it)
N=${1:-2}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_multi_$N.log
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -2 ) > gpurun_out/bench_r02_final_g$N.log
cat gpurun_out/pytest_multi_$N.log; head -c 500 gpurun_out/bench_r02_final_g$N.log
