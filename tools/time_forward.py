"""Event-timed SD-1.5 batch-8 UNet forward, eager launches vs replayed CUDA graph (A/B env switches: DG_PDL, DG_FUSE_*)."""
import sys
import torch
sys.path.insert(0, ".")
from bench import fast_state_dict
from divergen_b200 import UNet2DConditionModel

unet = UNet2DConditionModel(device="cuda:0")
unet.load_state_dict(fast_state_dict(unet.expected_state_dict_shapes()))
g = torch.Generator().manual_seed(0)
x = torch.randn(8, 4, 64, 64, generator=g).half().cuda()
ehs = torch.randn(8, 77, 768, generator=g).half().cuda()
for graphs in (False, True):
    unet.set_graphs(graphs)
    for _ in range(3):
        unet(x, 981, ehs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        unet(x, 981, ehs)
    e1.record()
    torch.cuda.synchronize()
    print(f"graphs={graphs}: {e0.elapsed_time(e1) / 20:.3f} ms per forward", flush=True)
