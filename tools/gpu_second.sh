#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --verbose 1 > gpurun_out/bench_second.log 2>&1
grep -v "check N=" gpurun_out/bench_second.log | tail -5
