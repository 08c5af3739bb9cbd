#!/bin/bash
mkdir -p gpurun_out
python tools/profile_forward.py 16 > gpurun_out/prof_fwd_emu_b16.txt 2>&1; echo "profile rc=$?"
python bench.py --steps 1 --warmup 3 --cpu-sample 0 > gpurun_out/b1.json 2>gpurun_out/b1.err; echo "bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:capture_tc -s 3 -c 1 -f -o gpurun_out/prof_capture_bench_r01 \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_emu_r01.csv \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
head -70 gpurun_out/prof_fwd_emu_b16.txt | cut -c1-100,200-330
