#!/bin/bash
# round 2 GPU job 6: where does the PES time go (finer kinds, chunk sweep, ncu hot spots of k_pes_defer)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=kinds timeout 600 python scripts/dev_perf.py > gpurun_out/j6_perf.log 2>&1
SSB_PES_CHUNKS=8 B=1024 STEPS=64 KERNELS=1 TAG=chunks8 timeout 600 python scripts/dev_perf.py > gpurun_out/j6_perf_c8.log 2>&1
SSB_PES_CHUNKS=16 B=1024 STEPS=64 KERNELS=1 TAG=chunks16 timeout 600 python scripts/dev_perf.py > gpurun_out/j6_perf_c16.log 2>&1
SSB_PES_CHUNKS=2 B=1024 STEPS=64 KERNELS=1 TAG=chunks2 timeout 600 python scripts/dev_perf.py > gpurun_out/j6_perf_c2.log 2>&1
B=1024 STEPS=216 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pes --launch-skip 440 --launch-count 6 \
   -o /tmp/r02c_pes -f python scripts/dev_prof.py > gpurun_out/j6_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/r02c_pes.ncu-rep gpurun_out/r02c_ncu_pes_summary.csv > gpurun_out/j6_ncu_summary.log 2>&1
python scripts/ncu_hot.py /tmp/r02c_pes.ncu-rep k_pes_defer 40 > gpurun_out/j6_hot_pes_defer.log 2>&1
ls -la gpurun_out | tail -8
