"""Marginal cost of each kernel family inside the batch-8 SD-1.5 forward graph: the forward replayed with ALL families, and with
one family left out (dg_unet_set_family_mask; the outputs of the reduced graphs are meaningless, only their time is used).  The
difference to the full graph is what the family costs IN the launch chain (programmatic dependent launch overlaps its head and
tail with its neighbours), as opposed to the family captured alone (bench.py's `family_graph_ms`)."""
import sys
import torch
sys.path.insert(0, ".")
from bench import fast_state_dict
from divergen_b200 import UNet2DConditionModel, _lib

lib = _lib.load()
unet = UNet2DConditionModel(device="cuda:0")
unet.load_state_dict(fast_state_dict(unet.expected_state_dict_shapes()))
g = torch.Generator().manual_seed(0)
x = torch.randn(8, 4, 64, 64, generator=g).half().cuda()
ehs = torch.randn(8, 77, 768, generator=g).half().cuda()
out = torch.empty_like(x)

def timed(mask, reps=20):
    _lib.check(lib.dg_unet_set_family_mask(unet._h, mask))
    for _ in range(3):
        unet(x, 981, ehs, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        unet(x, 981, ehs, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

try:
    for rep in range(2):
        full = timed(15)
        parts = {name: timed(15 & ~bit) for name, bit in (("gemm", 1), ("attn", 2), ("norm", 4), ("other", 8))}
        alone = {name: timed(bit) for name, bit in (("gemm", 1), ("attn", 2), ("norm", 4), ("other", 8))}
        print(f"full {full:.3f} ms | marginal: " + "  ".join(f"{k} {full - v:.3f}" for k, v in parts.items())
              + " | alone: " + "  ".join(f"{k} {v:.3f}" for k, v in alone.items()), flush=True)
finally:
    _lib.check(lib.dg_unet_set_family_mask(unet._h, 15))
