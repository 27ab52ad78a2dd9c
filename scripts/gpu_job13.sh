#!/bin/bash
# round 2 GPU job 13: configs[4] full-size test, G = 10^6 clean-up grid at 512 trials, Voja ring depth experiment, bench.py --workload cfg5 on one GPU
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config5_parity.py -q -m gpu > gpurun_out/j13_pytest_cfg5.log 2>&1
echo "rc $?" >> gpurun_out/j13_pytest_cfg5.log
DISTINCT=256 B=1024 STEPS=64 KERNELS=1 TAG=voja_nb3 timeout 600 python scripts/dev_perf.py > gpurun_out/j13_perf_nb3.log 2>&1
SSB_VOJA_NB=2 DISTINCT=256 B=1024 STEPS=64 KERNELS=1 TAG=voja_nb2 timeout 600 python scripts/dev_perf.py > gpurun_out/j13_perf_nb2.log 2>&1
SSB_VOJA_NB=1 DISTINCT=256 B=1024 STEPS=64 KERNELS=1 TAG=voja_nb1 timeout 600 python scripts/dev_perf.py > gpurun_out/j13_perf_nb1.log 2>&1
( time timeout 1200 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-synth --sustained-steps 0 > gpurun_out/j13_bench_cfg5.json 2> gpurun_out/j13_bench_cfg5.err ) 2> gpurun_out/j13_bench_cfg5.time
GRID=100 B=512 STEPS=16 KERNELS=1 ORACLE=0 timeout 1500 python scripts/dev_cfg5.py > gpurun_out/j13_cfg5_grid100_b512.log 2>&1
ls -la gpurun_out | tail -8
