#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "attention" -x 2>&1 | tail -3
{
echo "== wave-balanced grid (default)"; DG_TRACE=1 timeout 200 python tools/bench_ops.py attn 2>&1 | grep -v "^DG_TRACE" 
DG_TRACE=1 timeout 100 python tools/one_attn.py 2>&1 | grep "DG_TRACE attn" | sort | uniq -c
echo "== DG_ATTN_PART=0"; DG_ATTN_PART=0 timeout 200 python tools/bench_ops.py attn 2>&1
for rep in 1 2; do
echo "== forward, default"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== forward, DG_ATTN_PART=0"; DG_ATTN_PART=0 timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
} > gpurun_out/r02_run16_attn_part.log 2>&1
cat gpurun_out/r02_run16_attn_part.log
