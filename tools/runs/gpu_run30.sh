#!/bin/bash
mkdir -p gpurun_out
{
run() { echo "== $*"; env "$@" timeout 120 python tools/time_forward.py 2>&1 | tail -1; }
run DG_NOP=1
run DG_ATTN_POLY=2
run DG_GEMM_KB_THRESH=2
run DG_GEMM_KB_THRESH=8
run DG_NOP=1
run DG_SK_COEF=3
run DG_SK_COEF=7
run DG_ATTN_PART_MIN_KEYS=64
run DG_SK_MINKB=12
run DG_NOP=1
} > gpurun_out/r02_run30_knobs.log 2>&1
cat gpurun_out/r02_run30_knobs.log
