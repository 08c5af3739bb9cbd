#!/bin/bash
# Round-end check on one B200: GPU suite, smoke, bench + reference arm, ncu launch list of a bench step, ncu --set full of the
# capture launch inside bench.py (-> tools/ncu_traffic.py -> profiles/traffic.json).  Every ncu pass runs after its command exited 0 plain.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_native.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_native.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/b_final5.json 2> gpurun_out/b_final5.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_ref5.json 2> gpurun_out/b_ref5.err; echo "ref rc=$?"
python bench.py --steps 1 --warmup 3 --cpu-sample 0 > /dev/null 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01_final3.csv \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 --profile-range > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:capture_tc -s 3 -c 1 -f -o gpurun_out/prof_capture_bench_r01e \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/b_final5.json; cat gpurun_out/b_ref5.json | cut -c1-300
