#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dtw or align or probe or boundar" > gpurun_out/t_dtw.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_dtw.log
python tools/bench_dtw.py 2>&1 | tail -12
