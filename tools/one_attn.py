"""One self-attention call at the SD-1.5 64x64 level (for ncu captures): python tools/one_attn.py [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from divergen_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, heads, S, d = 8, 8, 4096, 40
g = torch.Generator().manual_seed(0)
q, k, v = (torch.randn(B, S, heads * d, generator=g).half().cuda() for _ in range(3))
for _ in range(reps):
    o = ops.attention(q, k, v, heads)
torch.cuda.synchronize()
print("ok", o.float().abs().max().item())
