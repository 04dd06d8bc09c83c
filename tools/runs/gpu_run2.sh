#!/bin/bash
# GPU run #2: fused GroupNorm transform + two epilogue sets -- parity, then same-box A/B of the forward.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -s > gpurun_out/r02_run2_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/r02_run2_ops.log
timeout 900 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_ops.py > gpurun_out/r02_run2_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_run2_pytest.log
{
for rep in 1 2; do
echo "== default (XF on, two sets)"; timeout 200 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_FUSE_XF=0"; DG_FUSE_XF=0 timeout 200 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GEMM_SETS=1"; DG_GEMM_SETS=1 timeout 200 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_FUSE_XF=0 DG_GEMM_SETS=1 (round-1 structure)"; DG_FUSE_XF=0 DG_GEMM_SETS=1 timeout 200 python tools/time_forward.py 2>&1 | tail -1
done
echo "== unroll2 build"; DG_LIB_PATH=$PWD/build_variants/lib_unroll2.so timeout 200 python tools/time_forward.py 2>&1 | tail -1
echo "== gemm ops default"; timeout 200 python tools/bench_ops.py gemm 2>&1
echo "== gemm ops DG_GEMM_SETS=1"; DG_GEMM_SETS=1 timeout 200 python tools/bench_ops.py gemm 2>&1
echo "== attn bitpack build"; DG_LIB_PATH=$PWD/build_variants/lib_bitpack.so timeout 200 python tools/bench_ops.py attn 2>&1
echo "== attn default"; timeout 200 python tools/bench_ops.py attn 2>&1
} > gpurun_out/r02_run2_ab.log 2>&1
cat gpurun_out/r02_run2_ab.log
