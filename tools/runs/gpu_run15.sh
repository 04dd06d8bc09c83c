#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_golden.py tests/test_gpu_edges.py tests/test_gpu_driver.py -q -m gpu -x 2>&1 | tail -3
{
for rep in 1 2; do
echo "== default (CFG prefix once)"; timeout 120 python tools/time_loop.py 2>&1 | tail -1
echo "== DG_CFG_DEDUP=0"; DG_CFG_DEDUP=0 timeout 120 python tools/time_loop.py 2>&1 | tail -1
done
} > gpurun_out/r02_run15_ab.log 2>&1
cat gpurun_out/r02_run15_ab.log
