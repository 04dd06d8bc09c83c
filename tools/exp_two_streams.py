"""Experiment: forward graphs of independent batches replayed concurrently on two streams vs one after the other, and one
larger batch (timing only: the two model instances share the context's split-K workspace, so concurrent results are not
checked).  Usage: exp_two_streams.py [samples per instance, default 8]"""
import sys
import torch
sys.path.insert(0, ".")
from bench import fast_state_dict
from divergen_b200 import UNet2DConditionModel

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
u0 = UNet2DConditionModel(device="cuda:0")
sd = fast_state_dict(u0.expected_state_dict_shapes())
u0.load_state_dict(sd)
u1 = UNet2DConditionModel(device="cuda:0")
u1.load_state_dict(sd)
g = torch.Generator().manual_seed(0)
x = torch.randn(2 * nb, 4, 64, 64, generator=g).half().cuda()
ehs = torch.randn(2 * nb, 77, 768, generator=g).half().cuda()
xa, xb, ea, eb = x[:nb].contiguous(), x[nb:].contiguous(), ehs[:nb].contiguous(), ehs[nb:].contiguous()
oa, ob = torch.empty_like(xa), torch.empty_like(xb)
o2 = torch.empty_like(x)
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def big():
    u0(x, 981, ehs, out=o2)

def two():
    cur = torch.cuda.current_stream()
    s0.wait_stream(cur); s1.wait_stream(cur)
    with torch.cuda.stream(s0):
        u0(xa, 981, ea, out=oa)
    with torch.cuda.stream(s1):
        u1(xb, 981, eb, out=ob)
    cur.wait_stream(s0); cur.wait_stream(s1)

def seq():
    u0(xa, 981, ea, out=oa)
    u1(xb, 981, eb, out=ob)

for rep in range(2):
    print(f"2 x batch {nb}: one stream {timed(seq):.3f} ms | two streams {timed(two):.3f} ms | batch {2 * nb}, one stream {timed(big):.3f} ms", flush=True)
