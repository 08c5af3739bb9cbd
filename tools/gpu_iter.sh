#!/bin/bash
# One development iteration on a B200: GPU suite, capture timing (both splitter orders), role timeline.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_iter.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_iter.log
python tools/ncu_capture.py timit 16
WCA_DBG=0x100 python tools/ncu_capture.py timit 16
python tools/ncu_capture.py timit 32
python tools/ncu_capture.py libri 8
python tools/trace_capture.py timit 16 3 2>&1 | tail -18
