#!/bin/bash
# A/B on one box: head-score partials policy (auto / 0 / 1) on the LibriSpeech drain (800 utterances) and the headline step.
mkdir -p gpurun_out
for mode in auto 0 1; do
  WCA_SCORE_PARTIALS=$mode python bench.py --steps 4 --warmup 3 --cpu-sample 0 --configs librispeech --libri-utts 800 > gpurun_out/ab_part_$mode.json 2> gpurun_out/ab_part_$mode.err || tail -3 gpurun_out/ab_part_$mode.err
  python - <<PY
import json
l=json.load(open("gpurun_out/ab_part_$mode.json"))
c=l["configs"]["librispeech"]; st=c["stages_ms_rank0"]
print("partials=$mode: headline %.2f ms/step capture %.4f (frac %.3f) scores %s | libri %.1f utt/s capture %.1f ms (frac %.3f) scores %.1f ms dtw %.1f ms" % (
  l["ms_per_step"], l["roofline"]["ms_per_step"], l["roofline"]["frac"], {k: round(v,4) for k,v in l["stages_ms_per_step"].items() if "score" in k},
  c["value"], c["capture_rank0"]["ms"], c["capture_rank0"]["frac"], sum(v for k,v in st.items() if "score" in k), st["wca_dtw_align"]))
PY
done
