#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/time_families.py > gpurun_out/r02_family_marginal.log 2>&1; cat gpurun_out/r02_family_marginal.log | tail -3
for mbs in 8 4; do
timeout 400 python tools/c5_slice.py --cats 10 --images 32 --gpt_cats 2 --skip_unpacked --max_batch_size $mbs --out gpurun_out/r02_c5_slice_1gpu_mbs$mbs.json 2>&1 | tail -1
done
