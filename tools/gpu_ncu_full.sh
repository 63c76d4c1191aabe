#!/bin/bash
# ncu --set full capture of selected kernels of one training step (kkbox shape).  Usage: gpu_ncu_full.sh <regex> <count> <outname> [skip]
mkdir -p gpurun_out
REGEX=${1:-"k_attn|k_ff|k_gather"}; COUNT=${2:-8}; OUT=${3:-prof_full}; SKIP=${4:-0}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT -f -o gpurun_out/$OUT python tools/prof_kernels.py kkbox 4096 1 2>&1 | tail -5
ls -la gpurun_out/
