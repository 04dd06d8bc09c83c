#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "gemm" 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_unet.py tests/test_golden.py tests/test_gpu_clip.py -q -m gpu 2>&1 | tail -2
{
for rep in 1 2 3; do
echo "== default (320-wide tiles from K = 320, row sums on 320-wide tiles)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GEMM_KB_THRESH=24 (before)"; DG_GEMM_KB_THRESH=24 timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
} > gpurun_out/r02_run13_ab.log 2>&1
cat gpurun_out/r02_run13_ab.log
bash tools/profile_shapes.sh r02_p_thresh4 | tail -1
