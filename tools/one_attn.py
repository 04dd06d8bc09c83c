"""One attention shape, a few launches (for ncu): python tools/one_attn.py [reps] [B heads S Sk d]"""
import sys
import torch
sys.path.insert(0, ".")
from divergen_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, heads, S, Sk, d = (int(v) for v in sys.argv[2:7]) if len(sys.argv) > 6 else (8, 8, 4096, 4096, 40)
g = torch.Generator().manual_seed(0)
q = torch.randn(B, S, heads * d, generator=g).half().cuda()
k, v = (torch.randn(B, Sk, heads * d, generator=g).half().cuda() for _ in range(2))
for _ in range(reps):
    o = ops.attention(q, k, v, heads)
torch.cuda.synchronize()
print("ok", o.float().abs().max().item())
