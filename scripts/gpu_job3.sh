#!/bin/bash
# round 2 GPU job 3: parity after the scheduling / aflag / K-blocked scan changes, perf variants, timeline, sub-batch experiment, cfg5
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "d649" > gpurun_out/j3_pytest_d649.log 2>&1
echo "rc $?" >> gpurun_out/j3_pytest_d649.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_recorded_path.py -q -m gpu --durations=8 > gpurun_out/j3_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j3_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=default timeout 600 python scripts/dev_perf.py > gpurun_out/j3_perf_default.log 2>&1
SSB_LEVEL_DEPS=0 B=1024 STEPS=64 TAG=level_barriers timeout 600 python scripts/dev_perf.py > gpurun_out/j3_perf_barriers.log 2>&1
SSB_SCAN_TR=64 B=1024 STEPS=64 TAG=scan_tr64 timeout 600 python scripts/dev_perf.py > gpurun_out/j3_perf_tr64.log 2>&1
B=512 STEPS=64 TAG=b512 timeout 600 python scripts/dev_perf.py > gpurun_out/j3_perf_b512.log 2>&1
B=2048 STEPS=64 TAG=b2048 timeout 600 python scripts/dev_perf.py > gpurun_out/j3_perf_b2048.log 2>&1
B=1024 STEPS0=208 STEPS=24 timeout 600 python scripts/dev_timeline.py > gpurun_out/j3_timeline.log 2>&1
B=1024 SPLITS=1,2,4 timeout 900 python scripts/dev_split.py > gpurun_out/j3_split.log 2>&1
unset DISTINCT
B=512 STEPS=16 KERNELS=1 timeout 1200 python scripts/dev_cfg5.py > gpurun_out/j3_cfg5.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -15
