#!/bin/bash
# round 2 GPU job 21: CTA-cooperative PES fold (k_pes_fold_cta)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_per_trial_seeds.py -q -m gpu -k "slam or alternate or deferred or inverse or loihi or weights or probe" > gpurun_out/j21_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j21_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=fold_cta timeout 600 python scripts/dev_perf.py > gpurun_out/j21_perf.log 2>&1
SSB_PES_FOLD=tasks B=1024 STEPS=64 KERNELS=1 TAG=fold_tasks timeout 600 python scripts/dev_perf.py > gpurun_out/j21_perf_tasks.log 2>&1
ls -la gpurun_out | tail -4
