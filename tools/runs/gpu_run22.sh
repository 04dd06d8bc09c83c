#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_final_pytest.log
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
for rep in 1 2; do
echo "== loop, default"; timeout 120 python tools/time_loop.py 2>&1 | tail -1
echo "== loop, DG_TEMB_CACHE=0"; DG_TEMB_CACHE=0 timeout 120 python tools/time_loop.py 2>&1 | tail -1
done
