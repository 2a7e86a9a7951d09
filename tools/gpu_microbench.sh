#!/bin/bash
# pipe rates and copy bandwidth of the box (profiles/r01_microbench.txt): gpurun -- bash tools/gpu_microbench.sh
mkdir -p gpurun_out
python - > gpurun_out/microbench.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import rbl_b200
names = {0: 'copy GB/s', 1: 'read GB/s', 2: 'FFMA TFLOP/s', 3: 'DFMA TFLOP/s', 4: 'mma.sync tf32 TFLOP/s',
         5: 'mma.sync f64 TFLOP/s', 6: 'mma.sync f16 k16 TFLOP/s', 7: 'mma.sync bf16 k16 TFLOP/s'}
for w, nm in names.items():
    print(nm, rbl_b200.microbench(w, 1 << 30, 4000 if w >= 2 else 20))
PY
cat gpurun_out/microbench.log
