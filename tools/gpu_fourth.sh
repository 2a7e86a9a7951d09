#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/microbench2.log 2>&1 <<'PY'
import sys; sys.path.insert(0,'.')
import rbl_b200
for w,nm in ((4,'mma.sync tf32'),(6,'mma.sync f16 k16'),(7,'mma.sync bf16 k16')):
    print(nm, rbl_b200.microbench(w, 1<<30, 4000))
PY
cat gpurun_out/microbench2.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --verbose 2 > gpurun_out/bench_v2.log 2>&1
grep "timeline\|Iterations" gpurun_out/bench_v2.log | head -4
tail -1 gpurun_out/bench_v2.log | cut -c1-300
