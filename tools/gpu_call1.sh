#!/bin/bash
# round-1 re-entry check: tests in both GEMM modes, bench in both, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_native.log 2>&1; echo "tests native rc=$?"
WCA_FP32_GEMM=bf16x9 python -m pytest tests -m gpu -x -q > gpurun_out/t_emu.log 2>&1; echo "tests emu rc=$?"
tail -3 gpurun_out/t_native.log gpurun_out/t_emu.log
python bench.py > gpurun_out/b_emu.json 2> gpurun_out/b_emu.err; echo "bench emu rc=$?"
python bench.py --fp32-gemm native --cpu-sample 0 > gpurun_out/b_native.json 2> gpurun_out/b_native.err; echo "bench native rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_ref.json 2> gpurun_out/b_ref.err; echo "bench ref rc=$?"
cat gpurun_out/b_emu.json gpurun_out/b_native.json gpurun_out/b_ref.json
tail -3 gpurun_out/b_emu.err
