#!/bin/bash
# GPU run #3: two epilogue sets + 256-wide GEGLU tiles (default), fused GroupNorm transform (opt-in): parity, then same-box A/B.
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_run3_smoke.log 2>&1; rc=$?
tail -2 gpurun_out/r02_run3_smoke.log
if [ $rc -ne 0 ]; then echo "canary failed (rc=$rc): stopping"; DG_GEMM_SETS=1 timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_ops.py -q -s > gpurun_out/r02_run3_ops.log 2>&1; echo "ops rc=$?"; tail -4 gpurun_out/r02_run3_ops.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_ops.py > gpurun_out/r02_run3_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_run3_pytest.log | cut -c1-300
for xf in 1 2; do DG_FUSE_XF=$xf timeout 300 python -m pytest tests/test_gpu_unet.py -q -s -k "tiny_unet_forward or sd15_forward_full_width or sd21_forward_full_width_batch4 or tiny_denoise" > gpurun_out/r02_run3_xf$xf.log 2>&1; echo "xf=$xf rc=$?"; grep -E "passed|failed" gpurun_out/r02_run3_xf$xf.log | tail -1; done
{
for rep in 1 2; do
echo "== default (two sets, GEGLU 256)"; timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_GEMM_SETS=1"; DG_GEMM_SETS=1 timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== GEGLU wide build (320/160 single stage), two sets elsewhere"; DG_LIB_PATH=$PWD/build_variants/lib_gegluwide.so timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== GEGLU wide build, DG_GEMM_SETS=1 (round-1 structure)"; DG_GEMM_SETS=1 DG_LIB_PATH=$PWD/build_variants/lib_gegluwide.so timeout 120 python tools/time_forward.py 2>&1 | tail -1
done
echo "== DG_FUSE_XF=1"; DG_FUSE_XF=1 timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== DG_FUSE_XF=2"; DG_FUSE_XF=2 timeout 120 python tools/time_forward.py 2>&1 | tail -1
echo "== gemm ops default"; timeout 120 python tools/bench_ops.py gemm 2>&1 | tail -9
echo "== gemm ops DG_GEMM_SETS=1"; DG_GEMM_SETS=1 timeout 120 python tools/bench_ops.py gemm 2>&1 | tail -9
} > gpurun_out/r02_run3_ab.log 2>&1
cat gpurun_out/r02_run3_ab.log
