#!/bin/bash
# Runs the per-kernel GPU checks in separate processes with hard timeouts so one hung kernel
# cannot hide the results of the others.  Output goes to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
(timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "not gemm and not conv3x3" 2>&1 | tail -25) > gpurun_out/ops_misc.log
(timeout 120 python tools/diag_gemm.py 2>&1 | tail -60) > gpurun_out/diag_gemm.log
(timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "gemm" 2>&1 | tail -30) > gpurun_out/ops_gemm.log
(timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "conv3x3" 2>&1 | tail -30) > gpurun_out/ops_conv.log
tail -n 12 gpurun_out/ops_misc.log gpurun_out/diag_gemm.log gpurun_out/ops_gemm.log gpurun_out/ops_conv.log
