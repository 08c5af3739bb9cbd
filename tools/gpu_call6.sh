#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_native.log 2>&1; echo "tests native rc=$?"; tail -3 gpurun_out/t_native.log
python bench.py --cpu-sample 0 > gpurun_out/b_attn2.json 2> gpurun_out/b_attn2.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("b_attn2",):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["stages_ms_per_step"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
ncu --set full --clock-control none --import-source on -k regex:enc_attn -s 3 -c 1 -f -o gpurun_out/prof_enc_attn_r01 python tools/test_enc_attn.py time > gpurun_out/ncu_enc_attn.log 2>&1; echo "ncu rc=$?"
