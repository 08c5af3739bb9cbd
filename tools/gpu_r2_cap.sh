#!/bin/bash
# capture-kernel experiments on one B200 (tools/ncu_capture.py: median of 20, L2 flushed)
for shape in "timit 32" "libri 8"; do
  echo "== $shape"
  WCA_PARTIALS=0 python tools/ncu_capture.py $shape
  WCA_PARTIALS=1 python tools/ncu_capture.py $shape
  WCA_PARTIALS=1 WCA_DBG=0x600 python tools/ncu_capture.py $shape
done
