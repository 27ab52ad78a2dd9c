#!/bin/bash
# round 2 GPU job 16 (8 GPUs): BASELINE configs[4] as stated - 3-D d = 649, 4 096 trials sharded over 8 GPUs (512 per GPU)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nproc > gpurun_out/j16_host.log; free -g >> gpurun_out/j16_host.log; nvidia-smi -L >> gpurun_out/j16_host.log
( time timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --workload cfg5 --distinct 128 --steps 4 --warmup 3 --no-cpu-baseline --no-synth --sustained-steps 0 \
    > gpurun_out/j16_bench_cfg5_8gpu.json 2> gpurun_out/j16_bench_cfg5_8gpu.err ) 2> gpurun_out/j16_time.log
echo "rc $?" >> gpurun_out/j16_time.log
tail -c 600 gpurun_out/j16_bench_cfg5_8gpu.json
