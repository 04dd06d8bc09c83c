#!/bin/bash
B=$PWD/build_variants
echo "### one set (GEGLU 256, double-buffered)"; DG_GEMM_SETS=1 bash tools/profile_shapes.sh r02_p_sets1
echo "### two sets"; bash tools/profile_shapes.sh r02_p_sets2
echo "### round-1 structure (GEGLU 320 single stage, one set)"; DG_GEMM_SETS=1 DG_LIB_PATH=$B/lib_gegluwide.so bash tools/profile_shapes.sh r02_p_r1
