#!/bin/bash
set -o pipefail
mkdir -p gpurun_out
python tools/profile_solve.py --cap 4096 > gpurun_out/plain_cap.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:reorth_gram_h -s 120 -c 1 -o gpurun_out/prof_gram_h -f python tools/profile_solve.py --cap 4096 > gpurun_out/ncu_gram_h.log 2>&1
cat gpurun_out/plain_cap.log
python tools/profile_solve.py > gpurun_out/plain_full2.log 2>&1; cat gpurun_out/plain_full2.log
