#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload librispeech --batch 8 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/b_libri.json 2> gpurun_out/b_libri.err; echo "libri rc=$?"
python bench.py --workload ami --model large-v3 --batch 16 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/b_ami.json 2> gpurun_out/b_ami.err; echo "ami rc=$?"
python bench.py --workload probe --batch 8 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/b_probe.json 2> gpurun_out/b_probe.err; echo "probe rc=$?"
python - <<'PY'
import json
for f in ("b_libri","b_ami","b_probe"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],2), round(d["ms_per_step"],1), {k:round(v,3) for k,v in d["stages_ms_per_step"].items()}, round(d["roofline"]["frac"],3), d["dtw_cells_per_s"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
