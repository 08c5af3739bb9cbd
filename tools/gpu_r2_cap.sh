#!/bin/bash
# capture-kernel A/B on one B200 (tools/ncu_capture.py: median of 20, L2 flushed): the current build against every other
# build of the library left as tools/libwca_*.so (git-ignored)
for lib in "" tools/libwca_*.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  echo "== lib: ${lib:-current}"
  for shape in "timit 32" "libri 8" "ami 32"; do
    WCA_LIB=$lib WCA_PARTIALS=1 python tools/ncu_capture.py $shape
  done
done
