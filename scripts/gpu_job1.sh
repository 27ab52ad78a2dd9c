#!/bin/bash
# round 2, first GPU job: new parity tests, bench line, launch list, steady-state full capture
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j1_smi.log 2>&1
nproc > gpurun_out/j1_nproc.log; free -g >> gpurun_out/j1_nproc.log
timeout 1500 python -m pytest tests/test_gpu_config2_parity.py -x -q -s -m gpu > gpurun_out/j1_pytest_cfg2.log 2>&1
echo "pytest rc $?" >> gpurun_out/j1_pytest_cfg2.log
timeout 900 python bench.py > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err
echo "bench rc $?" >> gpurun_out/j1_bench.err
timeout 600 python bench.py --distinct 8 --sustained-steps 0 --no-cpu-baseline > gpurun_out/j1_bench_tiled8.json 2> gpurun_out/j1_bench_tiled8.err
B=1024 STEPS=216 timeout 600 python scripts/dev_prof.py > gpurun_out/j1_prof_plain.log 2>&1 && \
B=1024 STEPS=216 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 3000 --launch-count 64 \
   -o gpurun_out/r02a_steady -f python scripts/dev_prof.py > gpurun_out/j1_ncu_full.log 2>&1
ls -la gpurun_out
