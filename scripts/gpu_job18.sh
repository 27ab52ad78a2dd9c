#!/bin/bash
# round 2 GPU job 18: the serial tail of the step (lin1 -> ens_small(1) -> final lin -> next lin0): k_lin variant by wave count,
# priority of the wide-encode -> decode chain
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slam_rate or slam_spiking or alternate or pathint" > gpurun_out/j18_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j18_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=lin_minb_adaptive timeout 600 python scripts/dev_perf.py > gpurun_out/j18_perf.log 2>&1
SSB_LIN_MINB=6 B=1024 STEPS=64 KERNELS=1 TAG=lin_minb6 timeout 600 python scripts/dev_perf.py > gpurun_out/j18_perf_minb6.log 2>&1
SSB_PRIO_D=1 B=1024 STEPS=64 TAG=prio_d timeout 600 python scripts/dev_perf.py > gpurun_out/j18_perf_prio_d.log 2>&1
SSB_PRIO_D=1 SSB_LIN_MINB=6 B=1024 STEPS=64 TAG=prio_d_minb6 timeout 600 python scripts/dev_perf.py > gpurun_out/j18_perf_prio_d_minb6.log 2>&1
SSB_PRIO_D=1 B=1024 STEPS0=208 STEPS=24 timeout 600 python scripts/dev_timeline.py > gpurun_out/j18_timeline_prio_d.log 2>&1
CONFIG=pathint97 B=1024 STEPS=64 KERNELS=1 TAG=minb_adaptive timeout 600 python scripts/dev_perf.py > gpurun_out/j18_perf_pi97.log 2>&1
ls -la gpurun_out | tail -6
