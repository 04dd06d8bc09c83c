"""Clock-stamp the GEMM epilogue variants (run with DG_GEMM_DBG=1): plain / +residual / +GroupNorm sums / +row sums."""
import ctypes as C
import math
import sys
import torch
sys.path.insert(0, ".")
from divergen_b200 import _lib, ops

M, K, N = int(sys.argv[2]) if len(sys.argv) > 2 else 18944, int(sys.argv[1]) if len(sys.argv) > 1 else 2880, 320
hw = M // 8
x = torch.randn(M, K, device="cuda").half()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
b = torch.randn(N, device="cuda").half()
r = torch.randn(M, N, device="cuda").half()
lib, ctx = _lib.load(), _lib.context(0)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = torch.empty(M, N, device="cuda", dtype=torch.float16)
rs = torch.empty(M, ops.row_parts(N), 2, device="cuda")
gs = torch.empty(8, hw // 32, N // 10, 2, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
for name, res, rso, gso in (("plain", None, None, None), ("+res", r, None, None), ("+gn", None, None, gs), ("+rs", None, rs, None),
                            ("+res+gn", r, None, gs), ("+res+rs", r, rs, None)):
    print("==", name, file=sys.stderr, flush=True)
    for _ in range(2):
        _lib.check(lib.dg_op_gemm_fused(ctx, P(x), P(w), P(b), P(None), P(None), P(None), 0, 0.0, P(res), P(out), M, K, N, N, 0,
                                        P(rso), P(gso), 10 if gso is not None else 0, hw, s))
    torch.cuda.synchronize()
