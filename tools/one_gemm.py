"""One GEMM shape, a few launches (for ncu): python tools/one_gemm.py M K N [reps]"""
import math
import sys
import torch
sys.path.insert(0, ".")
from divergen_b200 import ops
M, K, N = (int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
x = torch.randn(M, K, device="cuda").half()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
for _ in range(reps):
    ops.linear(x, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.linear(x, w)
e1.record()
torch.cuda.synchronize()
print(f"M={M} K={K} N={N}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
