#!/bin/bash
# GPU run #1 of round 2: tests, default bench (with c4 / e2e_png), attention variants, sd21 bench, C5 slice.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_run1_gpu.txt 2>&1
nproc >> gpurun_out/r02_run1_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r02_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
tail -5 gpurun_out/r02_run1_pytest.log
timeout 600 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
for v in 0 1; do echo "DG_ATTN64_VAR=$v"; DG_ATTN64_VAR=$v timeout 300 python tools/bench_ops.py attn; done > gpurun_out/r02_run1_attn.log 2>&1
DG_ATTN64_VAR=1 timeout 600 python bench.py --config sd21 --no_extras > gpurun_out/r02_bench_sd21_var1.json 2> gpurun_out/r02_bench_sd21_var1.err; echo "bench sd21 rc=$?"
timeout 900 python tools/c5_slice.py --out gpurun_out/r02_c5_slice.json > gpurun_out/r02_c5_slice.log 2>&1; echo "c5 rc=$?"
tail -3 gpurun_out/r02_c5_slice.log | cut -c1-600
cat gpurun_out/r02_run1_attn.log
