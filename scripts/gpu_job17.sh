#!/bin/bash
# round 2 GPU job 17: early end-of-step rows (VCO filters right after level 0's narrow ensembles, on their own stream)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slam or alternate or pathint or deferred or surface or probe" > gpurun_out/j17_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j17_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=lin_early timeout 600 python scripts/dev_perf.py > gpurun_out/j17_perf_early.log 2>&1
SSB_LIN_EARLY=0 B=1024 STEPS=64 KERNELS=1 TAG=lin_early_off timeout 600 python scripts/dev_perf.py > gpurun_out/j17_perf_early_off.log 2>&1
CONFIG=pathint97 B=1024 STEPS=64 KERNELS=1 TAG=lin_early timeout 600 python scripts/dev_perf.py > gpurun_out/j17_perf_pi97_early.log 2>&1
CONFIG=pathint97 SSB_LIN_EARLY=0 B=1024 STEPS=64 KERNELS=1 TAG=lin_early_off timeout 600 python scripts/dev_perf.py > gpurun_out/j17_perf_pi97_early_off.log 2>&1
B=1024 STEPS0=208 STEPS=24 timeout 600 python scripts/dev_timeline.py > gpurun_out/j17_timeline.log 2>&1
ls -la gpurun_out | tail -6
