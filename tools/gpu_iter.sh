#!/bin/bash
# One development iteration on a B200: GPU suite, capture timing (against other builds of the library under tools/), role timeline.
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_iter.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_iter.log
fi
for lib in "" tools/libwca_head.so tools/libwca_prev.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  echo "== lib: ${lib:-current}"
  WCA_LIB=$lib python tools/ncu_capture.py timit 16
  WCA_LIB=$lib python tools/ncu_capture.py timit 32
  WCA_LIB=$lib python tools/ncu_capture.py libri 8
done
python tools/trace_capture.py ${2:-timit} ${3:-16} 3 2>&1 | tail -17
