#!/bin/bash
# round 2 GPU job 26: CTA-cooperative PES fold as the default - parity subset and the bench line
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_per_trial_seeds.py -q -m gpu -x -k "slam or alternate or deferred or inverse or loihi or weights" > gpurun_out/j26_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j26_pytest.log
timeout 300 python bench.py > gpurun_out/j26_bench.json 2> gpurun_out/j26_bench.err
echo "bench rc $?" >> gpurun_out/j26_bench.err
