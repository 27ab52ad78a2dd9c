#!/bin/bash
# round 2 GPU job 27: ncu launch list of the bench command on the final code (CTA-cooperative PES fold as default)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 600 --csv --log-file gpurun_out/r02j_launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-synth --sustained-steps 0 > gpurun_out/j27_ncu_bench.log 2>&1
echo "rc $?" >> gpurun_out/j27_ncu_bench.log
