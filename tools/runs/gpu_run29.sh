#!/bin/bash
# ncu --set full of the attention out-projection shape (32768 x 320 x 320 + residual) and the LayerNorm-folded qkv projection at the final state
mkdir -p gpurun_out
cap() { tag=$1; shift; timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -o gpurun_out/r02_final_full_$tag -f "$@" > gpurun_out/r02_final_full_$tag.log 2>&1; }
cap linear_res_320 -k regex:gemm2 --launch-skip 3 --launch-count 1 python tools/dbg_epilogue.py 320 32768
DG_TRACE=1 python tools_profile_forward.py 8 2>&1 | grep "DG_TRACE gemm" | head -40 | grep -n "N=960" | head -2
cap qkv --profile-from-start off -k "regex:gemm2_kernel<.int.2, .int.320" --launch-skip 4 --launch-count 1 python tools_profile_forward.py 8
ls -la gpurun_out/r02_final_full_linear_res_320.ncu-rep gpurun_out/r02_final_full_qkv.ncu-rep
