#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python -m pytest tests/test_gpu_unet.py tests/test_golden.py tests/test_gpu_edges.py -q -m gpu -x 2>&1 | tail -3
{
for rep in 1 2 3; do
echo "== default (GroupNorm sums accumulated by the producers)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GN_ATOMIC=0 (one entry per slab, folded by every apply CTA)"; DG_GN_ATOMIC=0 timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
echo "== loop default"; timeout 120 python tools/time_loop.py 2>&1 | tail -1
echo "== loop DG_GN_ATOMIC=0"; DG_GN_ATOMIC=0 timeout 120 python tools/time_loop.py 2>&1 | tail -1
} > gpurun_out/r02_run28_gn_atomic.log 2>&1
cat gpurun_out/r02_run28_gn_atomic.log
