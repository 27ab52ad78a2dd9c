#!/bin/bash
# round 2 GPU job 10: k_lin_tck (large dense row-program blocks on tcgen05): d = 295 parity test, cfg5 at full size with the oracle, cfg2 perf
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "d295 or generic_width or alternate" > gpurun_out/j10_pytest_d259.log 2>&1
echo "rc $?" >> gpurun_out/j10_pytest_d259.log
B=512 STEPS=16 KERNELS=1 ORACLE=1 timeout 1500 python scripts/dev_cfg5.py > gpurun_out/j10_cfg5.log 2>&1
SSB_LIN=ffma B=512 STEPS=16 KERNELS=1 ORACLE=0 timeout 900 python scripts/dev_cfg5.py > gpurun_out/j10_cfg5_linffma.log 2>&1
DISTINCT=256 B=1024 STEPS=64 KERNELS=1 TAG=lin_lb6 timeout 600 python scripts/dev_perf.py > gpurun_out/j10_perf.log 2>&1
ls -la gpurun_out | tail -6
