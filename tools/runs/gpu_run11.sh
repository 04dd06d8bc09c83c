#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_ops.py -q -k "groupnorm or attention" 2>&1 | tail -2
{
for rep in 1 2 3; do
echo "== default (GN fold unroll 16)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== GN fold unroll 4 (before)"; DG_LIB_PATH=$PWD/build_variants/lib_gnfold4.so timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
for v in 0 2; do echo "== attn DG_ATTN64_VAR=$v"; DG_ATTN64_VAR=$v timeout 120 python tools/bench_ops.py attn 2>&1 | tail -1; done
DG_ATTN64_VAR=2 timeout 200 python -m pytest tests/test_gpu_ops.py -q -k "attention" 2>&1 | tail -1
} > gpurun_out/r02_run11_ab.log 2>&1
cat gpurun_out/r02_run11_ab.log
