#!/bin/bash
# round 2 GPU job 12: per-trial network seeds for wide ensembles (SLAM), configs[4] full-size test, perf of per-trial SLAM,
# the reference's 100-per-axis 3-D clean-up grid (G = 10^6) on the K-blocked scan
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_per_trial_seeds.py -q -m gpu > gpurun_out/j12_pytest_pt.log 2>&1
echo "rc $?" >> gpurun_out/j12_pytest_pt.log
timeout 900 python -m pytest tests/test_gpu_config5_parity.py -q -m gpu > gpurun_out/j12_pytest_cfg5.log 2>&1
echo "rc $?" >> gpurun_out/j12_pytest_cfg5.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slam or inverse or loihi or deferred" > gpurun_out/j12_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j12_pytest.log
DISTINCT=256 PER_TRIAL_SEEDS=32 B=1024 STEPS=64 KERNELS=1 TAG=per_trial_seeds timeout 1200 python scripts/dev_perf.py > gpurun_out/j12_perf_slam55_pt.log 2>&1
GRID=100 B=128 STEPS=16 KERNELS=1 ORACLE=0 timeout 1500 python scripts/dev_cfg5.py > gpurun_out/j12_cfg5_grid100.log 2>&1
ls -la gpurun_out | tail -6
