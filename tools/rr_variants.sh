timeout 600 python -m pytest tests -m gpu -x -q -k "attn_fwd or attn_bwd" 2>&1 | tail -15
T='tests/test_gpu_backward.py::test_auc_logloss_after_fixed_steps_match_oracle'
for i in 1 2 3; do timeout 300 python -m pytest "$T" -m gpu -q -k "fp16 and tmall" -s 2>&1 | grep -E "AUC oracle"; done
for i in 1 2; do RAT_RR=0 timeout 300 python -m pytest "$T" -m gpu -q -k "fp16 and tmall" -s 2>&1 | grep -E "AUC oracle"; done
