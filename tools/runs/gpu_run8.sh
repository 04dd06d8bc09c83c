#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_edges.py -q -k "attention or unet" 2>&1 | tail -2
{
for rep in 1 2 3; do
echo "== default (x96 cross-attention, relaxed final store wait)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_ATTN_X96=0"; DG_ATTN_X96=0 timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
echo "== attn default"; timeout 120 python tools/bench_ops.py attn 2>&1 | tail -5
echo "== attn DG_ATTN_X96=0"; DG_ATTN_X96=0 timeout 120 python tools/bench_ops.py attn 2>&1 | tail -5
echo "== attn DG_ATTN_POLY=2"; DG_ATTN_POLY=2 timeout 120 python tools/bench_ops.py attn 2>&1 | tail -5
echo "== attn DG_ATTN_POLY=0"; DG_ATTN_POLY=0 timeout 120 python tools/bench_ops.py attn 2>&1 | tail -5
} > gpurun_out/r02_run8_ab.log 2>&1
cat gpurun_out/r02_run8_ab.log
