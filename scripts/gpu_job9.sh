#!/bin/bash
# round 2 GPU job 9: per-trial network seeds (k_ens_small_pt), PES fold fill, cfg5 after the voja revert
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_per_trial_seeds.py -q -m gpu > gpurun_out/j9_pytest_pt.log 2>&1
echo "rc $?" >> gpurun_out/j9_pytest_pt.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slam or deferred or inverse or pathint" > gpurun_out/j9_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j9_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=fold_fill timeout 600 python scripts/dev_perf.py > gpurun_out/j9_perf.log 2>&1
CONFIG=pathint97 B=1024 STEPS=64 KERNELS=1 TAG=shared timeout 600 python scripts/dev_perf.py > gpurun_out/j9_perf_pi97_shared.log 2>&1
CONFIG=pathint97 PER_TRIAL_SEEDS=64 B=1024 STEPS=64 KERNELS=1 TAG=per_trial_seeds timeout 900 python scripts/dev_perf.py > gpurun_out/j9_perf_pi97_pt.log 2>&1
CONFIG=pathint97 SSP_DIM=55 PER_TRIAL_SEEDS=64 B=1024 STEPS=64 TAG=per_trial_seeds_d55 timeout 900 python scripts/dev_perf.py > gpurun_out/j9_perf_pi55_pt.log 2>&1
unset DISTINCT
B=512 STEPS=16 KERNELS=1 ORACLE=0 timeout 1200 python scripts/dev_cfg5.py > gpurun_out/j9_cfg5.log 2>&1
ls -la gpurun_out | tail -8
