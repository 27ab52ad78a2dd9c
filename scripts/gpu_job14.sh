#!/bin/bash
# round 2 GPU job 14: input prefetch on its own stream (tables and on-device synthesis), sparse shared-weight decode as an alternative to
# the tensor-core decode, Voja ring depth 2 as default - parity subset, then perf
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slam or alternate or synth or input or deferred or surface" > gpurun_out/j14_pytest.log 2>&1
echo "rc $?" >> gpurun_out/j14_pytest.log
export DISTINCT=256
B=1024 STEPS=64 KERNELS=1 TAG=prefetch_stream timeout 600 python scripts/dev_perf.py > gpurun_out/j14_perf.log 2>&1
SYNTH=1 B=1024 STEPS=64 KERNELS=1 TAG=synth_prefetch_stream timeout 600 python scripts/dev_perf.py > gpurun_out/j14_perf_synth.log 2>&1
SSB_DECODE=sparse B=1024 STEPS=64 KERNELS=1 TAG=decode_sparse timeout 600 python scripts/dev_perf.py > gpurun_out/j14_perf_dec_sparse.log 2>&1
SSB_DECODE=sparse SYNTH=1 B=1024 STEPS=64 TAG=decode_sparse_synth timeout 600 python scripts/dev_perf.py > gpurun_out/j14_perf_dec_sparse_synth.log 2>&1
ls -la gpurun_out | tail -6
