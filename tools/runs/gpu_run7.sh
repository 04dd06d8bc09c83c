#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
DG_GN_TWO_STEP=1 timeout 300 python -m pytest tests/test_gpu_unet.py -q -k "tiny_unet_forward or sd15_forward_full_width" 2>&1 | tail -2
{
for rep in 1 2 3; do
echo "== default (selective sets)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GEMM_SETS=1"; DG_GEMM_SETS=1 timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GN_TWO_STEP=1"; DG_GN_TWO_STEP=1 timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
} > gpurun_out/r02_run7_ab.log 2>&1
cat gpurun_out/r02_run7_ab.log
DG_GN_TWO_STEP=1 bash tools/profile_shapes.sh r02_p_gn2 | tail -8
