#!/bin/bash
# round 2 GPU job 8: PES spike-list compaction + fold ILP; voja stream update; K-blocked tcgen05 wide encode (cfg5)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_recorded_path.py -q -m gpu > gpurun_out/j8_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j8_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=pes_list timeout 600 python scripts/dev_perf.py > gpurun_out/j8_perf.log 2>&1
CONFIG=slamview97 B=1024 STEPS=64 KERNELS=1 TAG=view97 timeout 600 python scripts/dev_perf.py > gpurun_out/j8_perf_view97.log 2>&1
unset DISTINCT
B=512 STEPS=16 KERNELS=1 timeout 1200 python scripts/dev_cfg5.py > gpurun_out/j8_cfg5.log 2>&1
SSB_ENCODE=ffma B=64 STEPS=16 timeout 1200 python scripts/dev_cfg5.py > gpurun_out/j8_cfg5_ffma_b64.log 2>&1
ls -la gpurun_out | tail -6
