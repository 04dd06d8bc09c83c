#!/bin/bash
# compute-sanitizer memcheck over the paths this round added: wave-balanced attention grids, blocked tile assignment + vector
# prefetch in the LayerNorm-folded / GEGLU GEMMs, shortcut inside conv2, and the smoke() loop (CFG prefix once, table reuse).
mkdir -p gpurun_out
cd /root/repo
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -x -k "attention and (5-8-2048 or 5-8-1024 or 10-10-384 or 25-2-896)" > gpurun_out/r02_sanitizer_attn.log 2>&1; echo "attn rc=$?"; tail -4 gpurun_out/r02_sanitizer_attn.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -x -k "geglu or layernorm_fold or shortcut" > gpurun_out/r02_sanitizer_gemm.log 2>&1; echo "gemm rc=$?"; tail -4 gpurun_out/r02_sanitizer_gemm.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_sanitizer_smoke.log
