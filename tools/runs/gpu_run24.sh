#!/bin/bash
mkdir -p gpurun_out
{
for rep in 1 2 3 4; do
echo "== default (vector prefetch on)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GEMM_VECPRE=0"; DG_GEMM_VECPRE=0 timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
} > gpurun_out/r02_run24_vecpre.log 2>&1
cat gpurun_out/r02_run24_vecpre.log
bash tools/profile_shapes.sh r02_p_blocked > /dev/null; head -16 gpurun_out/r02_p_blocked_shapes.txt
DG_GEMM_VECPRE=0 bash tools/profile_shapes.sh r02_p_blocked_novecpre > /dev/null; head -16 gpurun_out/r02_p_blocked_novecpre_shapes.txt
