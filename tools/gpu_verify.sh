#!/bin/bash
# Round-end check on one B200: GPU suite, smoke, bench (with the configs block) + reference arm, ncu launch list of a bench step,
# ncu --set full of the capture launch inside bench.py (-> tools/ncu_traffic.py -> profiles/traffic.json).
# Every ncu pass runs after its command exited 0 plain.  usage: gpu_verify.sh <tag>
mkdir -p gpurun_out
tag=${1:-r02}
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/t_$tag.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/t_$tag.log | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_ref_$tag.json 2> gpurun_out/b_ref_$tag.err; echo "ref rc=$?"
python bench.py --steps 1 --warmup 3 --cpu-sample 0 --configs '' > /dev/null 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 --configs '' --profile-range > gpurun_out/ncu_list_$tag.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:capture_tc -s 3 -c 1 -f -o gpurun_out/prof_capture_bench_$tag \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 --configs '' > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/b_$tag.json; cat gpurun_out/b_ref_$tag.json | cut -c1-400
