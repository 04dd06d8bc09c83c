#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm or linear or conv or geglu or layernorm" 2>&1 | tail -2
DG_GEMM_BLOCKED=2 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -q -x -k "gemm or linear or conv or geglu or layernorm or forward" 2>&1 | tail -2
{
for rep in 1 2 3 4; do
echo "== default (LayerNorm-folded GEMMs blocked)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GEMM_BLOCKED=2 (all multi-column-tile GEMMs)"; DG_GEMM_BLOCKED=2 timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
} > gpurun_out/r02_run25_blocked_all.log 2>&1
cat gpurun_out/r02_run25_blocked_all.log
DG_GEMM_BLOCKED=2 bash tools/profile_shapes.sh r02_p_blocked_all > /dev/null; head -40 gpurun_out/r02_p_blocked_all_shapes.txt
