#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/b_final1.json 2> gpurun_out/b_final1.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 3 --cpu-sample 0 > /dev/null 2>&1; echo "plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:capture_tc -s 3 -c 1 -f -o gpurun_out/prof_capture_bench_r01b \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01b.csv \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 --profile-range > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
cat gpurun_out/b_final1.json
