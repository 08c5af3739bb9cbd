#!/bin/bash
# Round 2, first look on one B200: GPU suite (native SGEMM, with the BF16x9 subprocess test), bench with the configs block.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2b.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/t_r2b.log
timeout 900 python bench.py > gpurun_out/b_r2b.json 2> gpurun_out/b_r2b.err; echo "bench rc=$?"; tail -5 gpurun_out/b_r2b.err
cat gpurun_out/b_r2b.json
