#!/bin/bash
# round 2 GPU job 24: measurement of record on the final code - smoke, whole GPU suite, bench line (+ reference arm), ncu launch list of
# the bench command, steady-state full capture (summaries only)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j24_smi.log 2>&1
timeout 600 python __graft_entry__.py --smoke > gpurun_out/j24_smoke.log 2>&1
echo "smoke rc $?" >> gpurun_out/j24_smoke.log
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/j24_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/j24_pytest.log
( time timeout 900 python bench.py > gpurun_out/j24_bench.json 2> gpurun_out/j24_bench.err ) 2> gpurun_out/j24_bench.time
echo "bench rc $?" >> gpurun_out/j24_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j24_bench_reference.json 2> gpurun_out/j24_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 600 --csv --log-file gpurun_out/r02i_launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-synth --sustained-steps 0 > gpurun_out/j24_ncu_bench.log 2>&1
B=1024 STEPS=216 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 3000 --launch-count 40 \
   -o /tmp/r02i_steady -f python scripts/dev_prof.py > gpurun_out/j24_ncu_full.log 2>&1
python scripts/ncu_summary.py /tmp/r02i_steady.ncu-rep gpurun_out/r02i_ncu_full_steady_summary.csv > gpurun_out/j24_ncu_summary.log 2>&1
B=1024 FOLD_MB=445 python scripts/ncu_traffic.py /tmp/r02i_steady.ncu-rep gpurun_out/r02i_ncu_traffic.json > gpurun_out/j24_ncu_traffic.log 2>&1
for k in k_ens_small k_wide_voja; do python scripts/ncu_hot.py /tmp/r02i_steady.ncu-rep $k 30 > gpurun_out/j24_hot_$k.log 2>&1; done
ls -la gpurun_out | tail -14
