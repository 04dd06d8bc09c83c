#!/bin/bash
# GPU run #5: is the wake-up of parked mbarrier waits what paces the short-K GEMMs (and the attention chain)?
mkdir -p gpurun_out
B=$PWD/build_variants
{
for lib in "" $B/lib_park0.so $B/lib_park200.so; do
  for rep in 1 2; do
    echo "== lib=${lib:-default(park 10ms)} two sets"; DG_LIB_PATH=$lib timeout 120 python tools/time_forward.py 2>&1 | tail -1
    echo "== lib=${lib:-default(park 10ms)} DG_GEMM_SETS=1"; DG_GEMM_SETS=1 DG_LIB_PATH=$lib timeout 120 python tools/time_forward.py 2>&1 | tail -1
  done
  echo "== lib=${lib:-default} attn"; DG_LIB_PATH=$lib timeout 120 python tools/bench_ops.py attn 2>&1 | tail -5
done
echo "##### stamps, spin waits, K=320 M=32768 two sets"; DG_LIB_PATH=$B/lib_park0_stamps.so DG_GEMM_DBG=1 timeout 120 python tools/dbg_epilogue.py 320 32768 2>&1 | grep -E "^==|DG_GEMM_DBG" | head -6
echo "##### stamps, spin waits, K=320 M=32768 one set"; DG_GEMM_SETS=1 DG_LIB_PATH=$B/lib_park0_stamps.so DG_GEMM_DBG=1 timeout 120 python tools/dbg_epilogue.py 320 32768 2>&1 | grep -E "^==|DG_GEMM_DBG" | head -6
} > gpurun_out/r02_run5_park.log 2>&1
cut -c1-1000 gpurun_out/r02_run5_park.log
