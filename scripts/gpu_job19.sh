#!/bin/bash
# round 2 GPU job 19: step fusion (the next step's level-0 rows evaluated by the end-of-step launch), stream priorities
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_per_trial_seeds.py -q -m gpu > gpurun_out/j19_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j19_pytest.log
export DISTINCT=256
B=1024 STEPS=64 TAG=fused timeout 600 python scripts/dev_perf.py > gpurun_out/j19_perf_fused.log 2>&1
SSB_LIN_FUSE=0 B=1024 STEPS=64 TAG=unfused timeout 600 python scripts/dev_perf.py > gpurun_out/j19_perf_unfused.log 2>&1
SSB_PRIOS=0,0,1,0 B=1024 STEPS=64 TAG=fused_prio_b timeout 600 python scripts/dev_perf.py > gpurun_out/j19_perf_fused_prio_b.log 2>&1
SSB_PRIOS=0,0,0,0 B=1024 STEPS=64 TAG=fused_prio_all timeout 600 python scripts/dev_perf.py > gpurun_out/j19_perf_fused_prio_all.log 2>&1
SYNTH=1 B=1024 STEPS=64 TAG=fused_synth timeout 600 python scripts/dev_perf.py > gpurun_out/j19_perf_fused_synth.log 2>&1
CONFIG=pathint97 B=1024 STEPS=64 TAG=fused timeout 600 python scripts/dev_perf.py > gpurun_out/j19_perf_pi97_fused.log 2>&1
CONFIG=pathint97 SSB_LIN_FUSE=0 B=1024 STEPS=64 TAG=unfused timeout 600 python scripts/dev_perf.py > gpurun_out/j19_perf_pi97_unfused.log 2>&1
CONFIG=slamview97 B=1024 STEPS=64 TAG=fused timeout 600 python scripts/dev_perf.py > gpurun_out/j19_perf_view97_fused.log 2>&1
CONFIG=slamview97 SSB_LIN_FUSE=0 B=1024 STEPS=64 TAG=unfused timeout 600 python scripts/dev_perf.py > gpurun_out/j19_perf_view97_unfused.log 2>&1
B=1024 STEPS0=208 STEPS=24 timeout 600 python scripts/dev_timeline.py > gpurun_out/j19_timeline.log 2>&1
ls -la gpurun_out | tail -6
