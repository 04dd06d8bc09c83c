#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm or linear or conv or geglu or layernorm" 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_unet.py tests/test_golden.py tests/test_gpu_clip.py -q -m gpu -x 2>&1 | tail -3
{
for rep in 1 2 3; do
echo "== default (blocked LN tiles + vector prefetch)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GEMM_BLOCKED=0 DG_GEMM_VECPRE=0"; DG_GEMM_BLOCKED=0 DG_GEMM_VECPRE=0 timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
echo "== DG_GEMM_BLOCKED=0 only"; DG_GEMM_BLOCKED=0 timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GEMM_VECPRE=0 only"; DG_GEMM_VECPRE=0 timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== default"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
} > gpurun_out/r02_run23_blocked.log 2>&1
cat gpurun_out/r02_run23_blocked.log
