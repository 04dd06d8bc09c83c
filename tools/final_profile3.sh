#!/bin/bash
# Round-2 final state on one B200: GPU tests, bench lines (sd15 / sd21), launch list of the bench command, per-shape GEMM times,
# DRAM traffic of the GEMM family, full captures of the top kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_final_pytest.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err; echo "bench rc=$?"
timeout 600 python bench.py --config sd21 --no_extras > gpurun_out/r02_bench_sd21_1gpu_final.json 2> gpurun_out/r02_bench_sd21_1gpu_final.err; echo "bench sd21 rc=$?"
DG_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/r02_final_launches_bench.csv python bench.py --steps 1 --warmup 1 --no_extras > gpurun_out/r02_final_bench_under_ncu.log 2>&1
bash tools/profile_shapes.sh r02_final_fwd_b8_warmcache > /dev/null
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm2 --clock-control none --csv \
  --log-file gpurun_out/r02_gemm2_dram.csv python tools_profile_forward.py 8 > /dev/null 2>&1
cap() { tag=$1; shift; timeout 400 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -o gpurun_out/r02_final_full_$tag -f "$@" > gpurun_out/r02_final_full_$tag.log 2>&1; }
cap attn_d40 -k regex:attn_tc --launch-skip 3 --launch-count 1 python tools/one_attn.py 5
cap geglu --profile-from-start off -k "regex:gemm2_kernel<.int.2, .int.256" --launch-count 1 python tools_profile_forward.py 8
cap conv320 --profile-from-start off -k "regex:gemm2_kernel<.int.2, .int.320" --launch-skip 2 --launch-count 1 python tools_profile_forward.py 8
cap linear160 --profile-from-start off -k "regex:gemm2_kernel<.int.2, .int.160" --launch-skip 4 --launch-count 1 python tools_profile_forward.py 8
ls -la gpurun_out/r02_final_full_*.ncu-rep
python tools/summarize_launches.py gpurun_out/r02_final_launches_bench.csv | head -14
python -c "
import json
for f in ('gpurun_out/r02_bench_1gpu_final.json','gpurun_out/r02_bench_sd21_1gpu_final.json'):
    d=json.load(open(f)); r=d['roofline']
    print(f, d['value'], d['e2e']['value'], r['frac'], r.get('graph_family_ms'), d.get('c4',{}).get('value'), d.get('e2e_png',{}).get('value'))
"
