#!/bin/bash
# round 2 GPU job 11: measurement of record - whole GPU suite, bench line (+ reference arm), ncu launch list of the bench command,
# steady-state full capture (summaries only); compute-sanitizer is closed on this pool
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j11_smi.log 2>&1
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/j11_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/j11_pytest.log
timeout 900 python bench.py > gpurun_out/j11_bench.json 2> gpurun_out/j11_bench.err
echo "bench rc $?" >> gpurun_out/j11_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j11_bench_reference.json 2> gpurun_out/j11_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 600 --csv --log-file gpurun_out/r02f_launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-synth --sustained-steps 0 > gpurun_out/j11_ncu_bench.log 2>&1
B=1024 STEPS=216 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 3000 --launch-count 36 \
   -o /tmp/r02f_steady -f python scripts/dev_prof.py > gpurun_out/j11_ncu_full.log 2>&1
python scripts/ncu_summary.py /tmp/r02f_steady.ncu-rep gpurun_out/r02f_ncu_full_steady_summary.csv > gpurun_out/j11_ncu_summary.log 2>&1
B=1024 python scripts/ncu_traffic.py /tmp/r02f_steady.ncu-rep gpurun_out/r02f_ncu_traffic.json > gpurun_out/j11_ncu_traffic.log 2>&1
for k in k_ens_small k_wide_voja; do python scripts/ncu_hot.py /tmp/r02f_steady.ncu-rep $k 30 > gpurun_out/j11_hot_$k.log 2>&1; done
ls -la gpurun_out | tail -14
