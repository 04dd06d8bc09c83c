#!/bin/bash
# Launch list of the bench command's timed region + full captures (.ncu-rep) of the kernels the round worked on.
mkdir -p gpurun_out
DG_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no_extras > gpurun_out/r02_bench_under_ncu.log 2>&1
cap() { tag=$1; shift; timeout 400 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -o gpurun_out/r02_full_$tag -f "$@" > gpurun_out/r02_full_$tag.log 2>&1; }
cap linear_shortk -k regex:gemm2 --launch-skip 10 --launch-count 1 python tools/dbg_epilogue.py 320 32768
cap attn_d40 -k regex:attn_tc --launch-skip 3 --launch-count 1 python tools/one_attn.py 5
cap attn_d64 -k regex:attn_tc --launch-skip 3 --launch-count 1 python tools/one_attn.py 5 4 5 9216 9216 64
cap geglu --profile-from-start off -k "regex:gemm2_kernel<2, 256" --launch-count 1 python tools_profile_forward.py 8
cap conv320 --profile-from-start off -k "regex:gemm2_kernel<2, 320" --launch-skip 2 --launch-count 1 python tools_profile_forward.py 8
cap gn_apply --profile-from-start off -k regex:gn_apply_blk --launch-skip 1 --launch-count 1 python tools_profile_forward.py 8
ls -la gpurun_out/*.ncu-rep
python tools/summarize_launches.py gpurun_out/r02_launches_bench.csv | head -12
