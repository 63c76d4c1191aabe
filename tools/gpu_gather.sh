#!/bin/bash
mkdir -p gpurun_out
{
python -m pytest tests -m gpu -x -q -k "gather or api" 2>&1 | tail -4
for cfg in "kkbox 4096 5" "tmall 4096 5" "ml 4096 5" "kkbox 16384 5"; do python tools/bench_gather.py $cfg 2>&1 | grep -E "drop=|Error"; done
bash tools/gpu_quick.sh "golden" | grep -E "roofline_gather|train"
} > gpurun_out/gather.log 2>&1
cat gpurun_out/gather.log
