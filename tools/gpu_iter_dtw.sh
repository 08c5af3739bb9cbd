#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dtw or align or probe or boundar" > gpurun_out/t_dtw.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_dtw.log
python tools/bench_dtw.py 2>&1 | tail -12
python tools/ncu_dtw.py 16 41 150 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtw_align -c 1 -f -o gpurun_out/prof_dtw_small_r01b python tools/ncu_dtw.py 16 41 150 > gpurun_out/ncu_dtw_small.log 2>&1; echo "ncu rc=$?"
