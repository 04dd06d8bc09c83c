#!/bin/bash
# compute-sanitizer racecheck over the GEMM epilogue changes (vector generations, blocked tiles) and the wave-balanced attention grid
mkdir -p gpurun_out
cd /root/repo
timeout 700 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_ops.py -q -x -k "geglu or layernorm_fold" > gpurun_out/r02_racecheck_gemm.log 2>&1; echo "gemm rc=$?"
grep -c "Race reported\|hazard" gpurun_out/r02_racecheck_gemm.log; grep "Race reported\|hazard" gpurun_out/r02_racecheck_gemm.log | sed 's/0x[0-9a-f]*/ADDR/g' | sort | uniq -c | sort -rn | head -12; tail -3 gpurun_out/r02_racecheck_gemm.log
timeout 700 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_ops.py -q -x -k "attention and (5-8-2048 or 10-10-384)" > gpurun_out/r02_racecheck_attn.log 2>&1; echo "attn rc=$?"
grep -c "Race reported\|hazard" gpurun_out/r02_racecheck_attn.log; grep "Race reported\|hazard" gpurun_out/r02_racecheck_attn.log | sed 's/0x[0-9a-f]*/ADDR/g' | sort | uniq -c | sort -rn | head -12; tail -3 gpurun_out/r02_racecheck_attn.log
