#!/bin/bash
# round 2 GPU job 7: PES decode / fold with lanes = output columns; streaming Voja kernel for d = 649
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_recorded_path.py -q -m gpu > gpurun_out/j7_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j7_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=pes_cols timeout 600 python scripts/dev_perf.py > gpurun_out/j7_perf.log 2>&1
B=512 STEPS=64 TAG=b512 timeout 600 python scripts/dev_perf.py > gpurun_out/j7_perf_b512.log 2>&1
B=2048 STEPS=64 TAG=b2048 timeout 600 python scripts/dev_perf.py > gpurun_out/j7_perf_b2048.log 2>&1
CONFIG=slamview97 B=1024 STEPS=64 KERNELS=1 TAG=view97 timeout 600 python scripts/dev_perf.py > gpurun_out/j7_perf_view97.log 2>&1
unset DISTINCT
B=512 STEPS=16 KERNELS=1 timeout 1200 python scripts/dev_cfg5.py > gpurun_out/j7_cfg5.log 2>&1
ls -la gpurun_out | tail -8
