#!/bin/bash
# plain runs first (must exit 0), then ncu: launch list of one whole solve, and full captures of the hot kernels
set -o pipefail
mkdir -p gpurun_out
python tools/profile_solve.py > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full.csv python tools/profile_solve.py > gpurun_out/ncu_full.log 2>&1
python tools/profile_solve.py --cap 4096 > gpurun_out/plain_cap.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:reorth_gram_h -s 120 -c 1 -o gpurun_out/prof_gram_h -f python tools/profile_solve.py --cap 4096 > gpurun_out/ncu_gram.log 2>&1
python tools/profile_solve.py --cap 4096 > gpurun_out/plain_cap2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:reorth_update_h -s 120 -c 1 -o gpurun_out/prof_update_h -f python tools/profile_solve.py --cap 4096 > gpurun_out/ncu_update.log 2>&1

python tools/profile_solve.py --cap 4096 > gpurun_out/plain_cap3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ritz_h_kernel -c 1 -o gpurun_out/prof_ritz_h -f python tools/profile_solve.py --cap 4096 > gpurun_out/ncu_ritz.log 2>&1
cat gpurun_out/plain_full.log; for f in ncu_full ncu_gram ncu_update ncu_ritz; do tail -n 3 gpurun_out/$f.log; done
