#!/bin/bash
# round 2 GPU job: new parity tests, bench line, timeline, steady-state full capture (summaries only: .ncu-rep is too big to pull)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j2_smi.log 2>&1
nproc > gpurun_out/j2_nproc.log; free -g >> gpurun_out/j2_nproc.log
timeout 1500 python -m pytest tests/test_gpu_config2_parity.py -q -s -m gpu > gpurun_out/j2_pytest_cfg2.log 2>&1
echo "pytest rc $?" >> gpurun_out/j2_pytest_cfg2.log
timeout 900 python bench.py > gpurun_out/j2_bench.json 2> gpurun_out/j2_bench.err
echo "bench rc $?" >> gpurun_out/j2_bench.err
timeout 600 python bench.py --distinct 8 --sustained-steps 0 --no-cpu-baseline > gpurun_out/j2_bench_tiled8.json 2> gpurun_out/j2_bench_tiled8.err
B=1024 STEPS0=208 STEPS=24 timeout 600 python scripts/dev_timeline.py > gpurun_out/j2_timeline.log 2>&1
B=512 STEPS0=208 STEPS=24 timeout 600 python scripts/dev_timeline.py > gpurun_out/j2_timeline_b512.log 2>&1
B=1024 STEPS=64 KERNELS=1 timeout 600 python scripts/dev_perf.py > gpurun_out/j2_perf_kernels.log 2>&1
B=1024 STEPS=216 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 3000 --launch-count 36 \
   -o /tmp/r02a_steady -f python scripts/dev_prof.py > gpurun_out/j2_ncu_full.log 2>&1
python scripts/ncu_summary.py /tmp/r02a_steady.ncu-rep gpurun_out/r02a_ncu_full_steady_summary.csv > gpurun_out/j2_ncu_summary.log 2>&1
B=1024 python scripts/ncu_traffic.py /tmp/r02a_steady.ncu-rep gpurun_out/r02a_ncu_traffic.json > gpurun_out/j2_ncu_traffic.log 2>&1
for k in k_pes_defer k_ens_small k_lin; do python scripts/ncu_hot.py /tmp/r02a_steady.ncu-rep $k 30 > gpurun_out/j2_hot_$k.log 2>&1; done
du -sh gpurun_out; ls -la gpurun_out
