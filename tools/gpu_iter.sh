#!/bin/bash
# One development iteration on a B200: GPU suite, capture timing (against the previous build when tools/libwca_prev.so exists), role timeline.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_iter.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_iter.log
python tools/ncu_capture.py timit 16
[ -f tools/libwca_prev.so ] && WCA_LIB=tools/libwca_prev.so python tools/ncu_capture.py timit 16
python tools/ncu_capture.py timit 32
python tools/ncu_capture.py libri 8
[ -f tools/libwca_prev.so ] && WCA_LIB=tools/libwca_prev.so python tools/ncu_capture.py libri 8
python tools/trace_capture.py timit 16 3 2>&1 | tail -17
