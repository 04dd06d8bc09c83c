#!/bin/bash
# GPU run #4: where does a short-K GEMM launch spend its time?  Clock stamps + per-launch ncu durations, one vs two epilogue sets.
mkdir -p gpurun_out
L=$PWD/build_variants/lib_stamps.so
{
for K in 320 1280; do
echo "##### K=$K M=32768 two sets"; DG_LIB_PATH=$L DG_GEMM_DBG=1 timeout 120 python tools/dbg_epilogue.py $K 32768 2>&1 | grep -E "^==|DG_GEMM_DBG"
echo "##### K=$K M=32768 one set"; DG_GEMM_SETS=1 DG_LIB_PATH=$L DG_GEMM_DBG=1 timeout 120 python tools/dbg_epilogue.py $K 32768 2>&1 | grep -E "^==|DG_GEMM_DBG"
done
} > gpurun_out/r02_run4_stamps.log 2>&1
for sets in 2 1; do
  if [ $sets = 1 ]; then export DG_GEMM_SETS=1; else unset DG_GEMM_SETS; fi
  timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg --cache-control none --clock-control none --csv --log-file gpurun_out/r02_run4_ncu_sets$sets.csv python tools/dbg_epilogue.py 320 32768 > /dev/null 2>&1
done
unset DG_GEMM_SETS
cat gpurun_out/r02_run4_stamps.log | cut -c1-900
python - <<'PY'
import csv
for sets in (2, 1):
    lines = open(f"gpurun_out/r02_run4_ncu_sets{sets}.csv").read().splitlines()
    st = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[st:]))
    by = {}
    for r in rows:
        if "gemm2" in r["Kernel Name"]:
            by.setdefault(r["ID"], {})[r["Metric Name"]] = r["Metric Value"]
    print("sets", sets, [(v.get("gpu__time_duration.sum"), v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")) for v in by.values()])
PY
