#!/bin/bash
# round 2 GPU job 15: CTA-cooperative Voja kernel for d = 649 (k_wide_voja_cta): parity at d = 295 (forced) and at full configs[4] size,
# then the configs[4] step with it and with the streaming kernel, rate mode and spiking
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "d295" > gpurun_out/j15_pytest_d295.log 2>&1
echo "rc $?" >> gpurun_out/j15_pytest_d295.log
timeout 900 python -m pytest tests/test_gpu_config5_parity.py -q -m gpu > gpurun_out/j15_pytest_cfg5.log 2>&1
echo "rc $?" >> gpurun_out/j15_pytest_cfg5.log
B=512 STEPS=16 KERNELS=1 ORACLE=0 timeout 900 python scripts/dev_cfg5.py > gpurun_out/j15_cfg5_cta.log 2>&1
NT=lif B=512 STEPS=16 KERNELS=1 ORACLE=0 timeout 900 python scripts/dev_cfg5.py > gpurun_out/j15_cfg5_cta_lif.log 2>&1
SSB_VOJA=stream NT=lif B=512 STEPS=16 KERNELS=1 ORACLE=0 timeout 900 python scripts/dev_cfg5.py > gpurun_out/j15_cfg5_stream_lif.log 2>&1
ls -la gpurun_out | tail -6
