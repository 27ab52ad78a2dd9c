#!/bin/bash
# round 2 GPU job 20: k_lin as one resident wave with a grid stride
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slam or alternate or pathint or deferred or surface or probe or d295" > gpurun_out/j20_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j20_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=lin_wave timeout 600 python scripts/dev_perf.py > gpurun_out/j20_perf.log 2>&1
SSB_LIN_WAVE=0 B=1024 STEPS=64 TAG=lin_wave_off timeout 600 python scripts/dev_perf.py > gpurun_out/j20_perf_wave_off.log 2>&1
SSB_LIN_MINB=6 B=1024 STEPS=64 KERNELS=1 TAG=lin_wave_minb6 timeout 600 python scripts/dev_perf.py > gpurun_out/j20_perf_minb6.log 2>&1
CONFIG=pathint97 B=1024 STEPS=64 TAG=lin_wave timeout 600 python scripts/dev_perf.py > gpurun_out/j20_perf_pi97.log 2>&1
CONFIG=slamview97 B=1024 STEPS=64 TAG=lin_wave timeout 600 python scripts/dev_perf.py > gpurun_out/j20_perf_view97.log 2>&1
ls -la gpurun_out | tail -6
