#!/bin/bash
# first GPU session: kernel parity, solver parity, micro-benchmarks
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
python - > gpurun_out/microbench.log 2>&1 <<'PY'
import sys; sys.path.insert(0,'.')
import rbl_b200
names={0:'copy GB/s',1:'read GB/s',2:'FFMA TF',3:'DFMA TF',4:'mma.sync tf32 TF',5:'mma.sync f64 TF'}
for w in range(6):
    print(names[w], rbl_b200.microbench(w, 1<<31, 10 if w<2 else 4000))
PY
cat gpurun_out/pytest_gpu.log gpurun_out/microbench.log
