#!/bin/bash
# Per-shape GEMM times of one eager SD-1.5 batch-8 forward under ncu (warm caches: --cache-control none), joined with DG_TRACE.
#   tools/profile_shapes.sh <tag>        (environment switches / DG_LIB_PATH are inherited)
tag=$1; mkdir -p gpurun_out
DG_TRACE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv \
  --log-file gpurun_out/${tag}_launches.csv python tools_profile_forward.py 8 2> gpurun_out/${tag}_trace.log > /dev/null
python tools/join_trace.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_trace.log > gpurun_out/${tag}_shapes.txt
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_summary.txt
tail -1 gpurun_out/${tag}_shapes.txt; head -12 gpurun_out/${tag}_summary.txt
