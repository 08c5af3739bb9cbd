#!/bin/bash
# Round-end check on one B200: GPU suite, ncu capture of the capture launch inside bench.py (-> tools/ncu_traffic.py), bench, reference arm.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_native.log 2>&1; echo "tests native rc=$?"; tail -3 gpurun_out/t_native.log
python bench.py --steps 1 --warmup 3 --cpu-sample 0 > /dev/null 2>&1; echo "plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:capture_tc -s 3 -c 1 -f -o gpurun_out/prof_capture_bench_r01c \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python bench.py > gpurun_out/b_final2.json 2> gpurun_out/b_final2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_ref2.json 2> gpurun_out/b_ref2.err; echo "ref rc=$?"
cat gpurun_out/b_final2.json; cat gpurun_out/b_ref2.json | cut -c1-300
