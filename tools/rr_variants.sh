timeout 600 python -m pytest tests -m gpu -x -q -k "attn_bwd" 2>&1 | tail -8
timeout 100 python tools/bench_attn.py kkbox 4096 5 2>&1 | grep bwd
timeout 100 python tools/bench_attn.py tmall 4096 5 2>&1 | grep bwd
