#!/bin/bash
# Round-end measurement pass on one B200: tests, bench lines, launch lists, DRAM traffic of the GEMM family, full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_final_pytest.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --config sd21 --no_extras > gpurun_out/r02_bench_sd21_1gpu.json 2> gpurun_out/r02_bench_sd21_1gpu.err; echo "bench sd21 rc=$?"
# launch list of the bench command itself (graph-replayed forwards of the timed loop)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no_extras > gpurun_out/r02_bench_under_ncu.log 2>&1
# one eager forward, warm caches, joined with the shape trace
bash tools/profile_shapes.sh r02_fwd_b8_warmcache > /dev/null
# DRAM bytes of every GEMM launch of that forward
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm2 --clock-control none --csv \
  --log-file gpurun_out/r02_gemm2_dram.csv python tools_profile_forward.py 8 > /dev/null 2>&1
# full captures: short-K linear + residual + row sums, GEGLU, conv3x3 320-wide, upsample phase, attention d=40
cap() { tag=$1; shift; timeout 300 ncu --set full --import-source on --clock-control none "$@" -o gpurun_out/r02_full_$tag -f > gpurun_out/r02_full_$tag.log 2>&1; }
cap linear_shortk -k regex:gemm2 --launch-skip 6 --launch-count 1 python tools/dbg_epilogue.py 320 32768
cap attn_d40 -k regex:attn_tc --launch-skip 3 --launch-count 1 python tools/one_attn.py 8 8 4096 4096 40
cap geglu --profile-from-start off -k regex:gemm2_kernel.*256 --launch-skip 0 --launch-count 1 python tools_profile_forward.py 8
cap conv320 --profile-from-start off -k regex:gemm2_kernel.*320 --launch-skip 2 --launch-count 1 python tools_profile_forward.py 8
ls -la gpurun_out/*.ncu-rep
python -c "
import json
for f in ('gpurun_out/r02_bench_1gpu.json','gpurun_out/r02_bench_sd21_1gpu.json'):
    d=json.load(open(f)); r=d['roofline']
    print(f, d['value'], d['e2e']['value'], r['frac'], r['graph_family_ms'], d.get('c4',{}).get('value'), d.get('e2e_png',{}).get('value'))
"
