#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_native.log 2>&1; echo "tests native rc=$?"; tail -3 gpurun_out/t_native.log
WCA_FP32_GEMM=bf16x9 python -m pytest tests -m gpu -x -q > gpurun_out/t_emu.log 2>&1; echo "tests emu rc=$?"; tail -3 gpurun_out/t_emu.log
python bench.py --cpu-sample 0 > gpurun_out/b_attn.json 2> gpurun_out/b_attn.err; echo "bench rc=$?"
WCA_ENCODER_ATTENTION=sdpa python bench.py --cpu-sample 0 > gpurun_out/b_sdpa.json 2> gpurun_out/b_sdpa.err; echo "bench sdpa rc=$?"
python - <<'PY'
import json
for f in ("b_attn","b_sdpa"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["stages_ms_per_step"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/b_attn.err
