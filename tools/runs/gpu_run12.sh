#!/bin/bash
mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== default"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
for kv in "DG_GEMM_KB_THRESH=4" "DG_GEMM_KB_THRESH=9" "DG_GN_ITERS=4" "DG_GN_ITERS=16" "DG_GN_WAVES=4" "DG_GN_WAVES=16" "DG_GN_ITERS=4 DG_GN_WAVES=16"; do
echo "== $kv"; env $kv timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
done
} > gpurun_out/r02_run12_knobs.log 2>&1
cat gpurun_out/r02_run12_knobs.log
