#!/bin/bash
# round 2 GPU job 25: CTA-cooperative PES fold with the history factors staged by TMA next to the tiles
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
SSB_PES_FOLD=cta timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slam_rate or deferred or weights" > gpurun_out/j25_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j25_pytest.log
export DISTINCT=256
SSB_PES_FOLD=cta B=1024 STEPS=64 KERNELS=1 TAG=fold_cta_tma_factors timeout 600 python scripts/dev_perf.py > gpurun_out/j25_perf.log 2>&1
ls -la gpurun_out | tail -3
