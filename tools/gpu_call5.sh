WCA_EA_SKIP=0 timeout 100 python tools/test_enc_attn.py 2>&1 | tail -12
timeout 100 python tools/trace_enc_attn.py 16 2>&1 | tail -16
