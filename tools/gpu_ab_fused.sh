#!/bin/bash
# A/B on one box: fused projections on/off (headline step only), twice each, interleaved.
mkdir -p gpurun_out
for i in 1 2; do
  for f in 1 0; do
    WCA_FUSED_PROJECTIONS=$f python bench.py --configs '' --cpu-sample 0 --steps 10 > gpurun_out/ab_fused${f}_$i.json 2> gpurun_out/ab_fused${f}_$i.err
    python - <<PY
import json
l=json.load(open("gpurun_out/ab_fused${f}_$i.json"))
print("fused=$f run $i: %.2f ms/step, capture %.4f ms (frac %.3f), attention %.2f ms, sm %.0f MHz" % (l["ms_per_step"], l["roofline"]["ms_per_step"], l["roofline"]["frac"], l["stages_ms_per_step"]["wca_full_attention"], l["clocks"]["sm_mhz"]))
PY
  done
done
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -x -q -k "fp32_gemm_mode or bf16x9_mode or fused" 2>&1 | tail -15
