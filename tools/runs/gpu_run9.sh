#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "upsample" 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_unet.py tests/test_gpu_vae.py tests/test_golden.py -q -m gpu 2>&1 | tail -3
{
for rep in 1 2 3; do
echo "== default (phase upsample convs)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_UPCONV_PHASES=0"; DG_UPCONV_PHASES=0 timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
echo "== vae default"; timeout 120 python tools/time_vae.py 2>&1 | tail -2
echo "== vae DG_UPCONV_PHASES=0"; DG_UPCONV_PHASES=0 timeout 120 python tools/time_vae.py 2>&1 | tail -2
} > gpurun_out/r02_run9_ab.log 2>&1
cat gpurun_out/r02_run9_ab.log
