#!/bin/bash
# capture-kernel A/B on one B200 (tools/ncu_capture.py: median of 20, L2 flushed); WCA_LIB=<other build> compares libraries
for lib in "" tools/libwca_prev.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  echo "== lib: ${lib:-current}"
  for shape in "timit 32" "libri 8"; do
    WCA_LIB=$lib WCA_PARTIALS=1 python tools/ncu_capture.py $shape
  done
done
