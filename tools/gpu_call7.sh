#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_native.log 2>&1; echo "tests native rc=$?"; tail -5 gpurun_out/t_native.log
timeout 100 python tools/ncu_capture.py timit 16
timeout 100 python tools/ncu_capture.py libri 8
timeout 100 python tools/trace_capture.py timit 16 3 2>&1 | tail -16
