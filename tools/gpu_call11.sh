#!/bin/bash
for b in 32 64; do
python bench.py --batch $b --steps 4 --warmup 3 --cpu-sample 0 > gpurun_out/b_b$b.json 2> gpurun_out/b_b$b.err; echo "batch $b rc=$?"
done
python - <<'PY'
import json
for f in ("b_b32","b_b64"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],2), round(d["e2e"]["value"],2), round(d["ms_per_step"],1), {k:round(v,3) for k,v in d["stages_ms_per_step"].items()}, round(d["roofline"]["frac"],3), d["clocks"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
