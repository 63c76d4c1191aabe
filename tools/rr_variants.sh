timeout 600 python -m pytest tests -m gpu -x -q -k "attn_bwd" 2>&1 | tail -3
timeout 100 python tools/bench_attn.py kkbox 4096 5 2>&1 | grep bwd
T='tests/test_gpu_backward.py::test_auc_logloss_after_fixed_steps_match_oracle'
timeout 300 python -m pytest "$T" -m gpu -q -k "tmall" -s 2>&1 | grep -E "AUC oracle|passed|failed"
RAT_RR=0 timeout 300 python -m pytest "$T" -m gpu -q -k "tmall and fp16" -s 2>&1 | grep -E "AUC oracle|passed|failed"
