#!/bin/bash
# Round 2 development iteration on one B200: GPU suite, DTW micro-bench, headline bench step (tag = $1).
mkdir -p gpurun_out
tag=${1:-iter}
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/t_$tag.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/t_$tag.log | tail -15
python tools/bench_dtw.py 2>&1 | tail -9
for i in 1 2; do
python bench.py --configs '' --cpu-sample 0 --steps 10 > gpurun_out/b_${tag}_$i.json 2> gpurun_out/b_${tag}_$i.err || tail -5 gpurun_out/b_${tag}_$i.err
python - <<PY
import json
l=json.load(open("gpurun_out/b_${tag}_$i.json"))
print("run $i: %.2f ms/step, capture %.4f ms (frac %.3f), attention %.2f ms, %s, sm %.0f MHz" % (l["ms_per_step"], l["roofline"]["ms_per_step"], l["roofline"]["frac"], l["stages_ms_per_step"]["wca_full_attention"], {k: round(v, 4) for k, v in l["stages_ms_per_step"].items() if "score" in k or "dtw" in k}, l["clocks"]["sm_mhz"]))
PY
done
