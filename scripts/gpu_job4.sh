#!/bin/bash
# round 2 GPU job 4: fused deferred-PES decode (parity + A/B timing)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "slam or deferred or inverse or alternate or batch" > gpurun_out/j4_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j4_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=fused timeout 600 python scripts/dev_perf.py > gpurun_out/j4_perf_fused.log 2>&1
SSB_PES_FUSE=0 B=1024 STEPS=64 TAG=unfused_predicated timeout 600 python scripts/dev_perf.py > gpurun_out/j4_perf_unfused.log 2>&1
B=512 STEPS=64 TAG=fused_b512 timeout 600 python scripts/dev_perf.py > gpurun_out/j4_perf_b512.log 2>&1
B=2048 STEPS=64 TAG=fused_b2048 timeout 600 python scripts/dev_perf.py > gpurun_out/j4_perf_b2048.log 2>&1
CONFIG=slamview97 B=1024 STEPS=64 KERNELS=1 TAG=fused timeout 600 python scripts/dev_perf.py > gpurun_out/j4_perf_view97.log 2>&1
ls -la gpurun_out | tail -8
