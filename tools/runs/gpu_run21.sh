#!/bin/bash
mkdir -p gpurun_out
DG_ATTN_VAR=4 timeout 240 python -m pytest tests/test_gpu_ops.py -q -k "attention" -x 2>&1 | tail -4
{
for rep in 1 2; do
echo "== DG_ATTN_VAR=4 (4 query tiles x 32 keys, double-buffered scores)"; DG_ATTN_VAR=4 timeout 200 python tools/bench_ops.py attn 2>&1 | grep "d=40"
echo "== default (4 x 64 keys)"; timeout 200 python tools/bench_ops.py attn 2>&1 | grep "d=40"
done
echo "== forward DG_ATTN_VAR=4"; DG_ATTN_VAR=4 timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== forward default"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
} > gpurun_out/r02_run21_attn_var4.log 2>&1
cat gpurun_out/r02_run21_attn_var4.log
