#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/t_native.log 2>&1; echo "tests native rc=$?"; tail -15 gpurun_out/t_native.log
