#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_ops.py -q -k "attention" -x 2>&1 | tail -8
{
for rep in 1 2; do
echo "== tf32 PV (default)"; timeout 200 python tools/bench_ops.py attn 2>&1
echo "== DG_ATTN_TF32=0"; DG_ATTN_TF32=0 timeout 200 python tools/bench_ops.py attn 2>&1
done
} > gpurun_out/r02_run20_attn_tf32.log 2>&1
cat gpurun_out/r02_run20_attn_tf32.log
