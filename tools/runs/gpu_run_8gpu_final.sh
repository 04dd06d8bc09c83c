#!/bin/bash
# 8 x B200, round-2 final state: configs[2] (sd15 weak scaling), configs[3] (SD-2.1 768x768, batch 16 over 8 GPUs), configs[4] slice (C5,
# --max_batch_size 8).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --no_extras > gpurun_out/r02_bench_8gpu_final.json 2> gpurun_out/r02_bench_8gpu_final.err; echo "bench8 rc=$?"
timeout 600 $TR --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --config sd21 --no_extras > gpurun_out/r02_bench_sd21_8gpu_final.json 2> gpurun_out/r02_bench_sd21_8gpu_final.err; echo "bench8 sd21 rc=$?"
export DG_C5_TMP=/tmp/dg_c5_8gpu
timeout 600 $TR --master-port 29513 tools/c5_slice.py --dist --only lvis --cats 12 --images 32 --max_batch_size 8 --out gpurun_out/r02_c5_slice_8gpu_lvis_final.json > gpurun_out/r02_c5_8gpu_lvis_final.log 2>&1; echo "c5 lvis rc=$?"
timeout 600 $TR --master-port 29514 tools/c5_slice.py --dist --only gpt --gpt_cats 2 --max_batch_size 8 --out gpurun_out/r02_c5_slice_8gpu_gpt_final.json > gpurun_out/r02_c5_8gpu_gpt_final.log 2>&1; echo "c5 gpt rc=$?"
find /tmp/dg_c5_8gpu/out_lvis -name "*.png" | wc -l; find /tmp/dg_c5_8gpu/out_gpt -name "*.png" | wc -l
python -c "
import json
for f in ('gpurun_out/r02_bench_8gpu_final.json','gpurun_out/r02_bench_sd21_8gpu_final.json'):
    d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['per_rank_ms_per_step'])
for f in ('gpurun_out/r02_c5_slice_8gpu_lvis_final.json','gpurun_out/r02_c5_slice_8gpu_gpt_final.json'):
    print(open(f).read()[:600])
"
tail -2 gpurun_out/r02_bench_8gpu_final.err
