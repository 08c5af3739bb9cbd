#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_native.log 2>&1; echo "tests native rc=$?"; tail -8 gpurun_out/t_native.log
WCA_FP32_GEMM=bf16x9 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_emu.log 2>&1; echo "tests emu rc=$?"; tail -3 gpurun_out/t_emu.log
python bench.py --cpu-sample 0 > gpurun_out/b_9.json 2> gpurun_out/b_9.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/b_9.json")); print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["stages_ms_per_step"], d["clocks"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/b_9.err
