#!/bin/bash
# round 2 GPU job 23: CTA-cooperative PES fold with the history factors prefetched two neurons ahead, 4-neuron tiles
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slam_rate or deferred or alternate or weights" > gpurun_out/j23_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j23_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=fold_cta_4neuron_tiles timeout 600 python scripts/dev_perf.py > gpurun_out/j23_perf.log 2>&1
ls -la gpurun_out | tail -3
