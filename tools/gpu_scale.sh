#!/bin/bash
# default bench at N GPUs exactly as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -E "^\{|Error|error" | tail -2 > gpurun_out/scale_$N.log
python - <<PY
import json
for l in open("gpurun_out/scale_$N.log"):
    if l.startswith("{"):
        d = json.loads(l); print("N=$N", d["config"]["parallelism"], "train", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "infer", d["infer"]["value"])
    else:
        print(l.rstrip()[-300:])
PY
