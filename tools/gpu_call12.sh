#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_native.log 2>&1; echo "tests native rc=$?"; tail -5 gpurun_out/t_native.log
python bench.py --steps 4 --warmup 3 --cpu-sample 0 > gpurun_out/b_12a.json 2> gpurun_out/b_12a.err; echo "timit rc=$?"
python bench.py --workload librispeech --batch 8 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/b_12b.json 2> gpurun_out/b_12b.err; echo "libri rc=$?"
python - <<'PY'
import json
for f in ("b_12a","b_12b"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],2), round(d["ms_per_step"],1), {k:round(v,3) for k,v in d["stages_ms_per_step"].items()}, round(d["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
