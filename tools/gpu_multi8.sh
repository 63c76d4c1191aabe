#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
{
echo "== bench kkbox N=$N"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -E "^\{|Error" | tail -2
echo "== bench tmall N=$N row-sharded"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --shape tmall --shard-tables 2>&1 | grep -E "^\{|Error" | tail -2
echo "== bench tmall N=$N row-sharded x50 vocab"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --shape tmall --shard-tables --vocab-scale 50 2>&1 | grep -E "^\{|Error" | tail -2
} > gpurun_out/multi_$N.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/multi_$N.log"):
    if l.startswith("{"):
        d = json.loads(l); print("  ", d["config"]["workload"][:40], d["config"]["parallelism"][:30], "params", d["config"]["params"], "train", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "infer", d["infer"]["value"], "gather", d["roofline_gather"]["avg_launch_ms"], "adam", d["roofline_adam"]["avg_launch_ms"])
    else:
        print(l.rstrip()[-400:])
PY
