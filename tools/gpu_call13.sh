#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_native.log 2>&1; echo "tests native rc=$?"; tail -5 gpurun_out/t_native.log
timeout 300 python tools/bench_dtw.py 2>&1 | tail -9
