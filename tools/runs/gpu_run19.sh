#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "geglu" -x 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_unet.py tests/test_gpu_driver.py tests/test_gpu_from_pretrained.py -q -m gpu -x 2>&1 | tail -3
{
DG_TRACE=1 timeout 100 python tools/time_forward.py 2>&1 | grep "geglu=1" | sort | uniq -c | head
for rep in 1 2 3; do
echo "== forward"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
echo "== loop"; timeout 120 python tools/time_loop.py 2>&1 | tail -1
} > gpurun_out/r02_run19_gelu.log 2>&1
cat gpurun_out/r02_run19_gelu.log
for mbs in 8 4; do
timeout 400 python tools/c5_slice.py --cats 10 --images 32 --gpt_cats 2 --skip_unpacked --max_batch_size $mbs --out gpurun_out/r02_c5_slice_1gpu_mbs${mbs}_pinned.json 2>&1 | tail -1 | cut -c1-700
done
