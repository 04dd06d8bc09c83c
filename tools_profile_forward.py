"""One eager SD-1.5 UNet forward (batch 8, 64x64) bracketed by cudaProfilerStart/Stop for ncu
(`ncu --profile-from-start off ...`).  DG_TRACE=1 prints the shape of every tcgen05 launch in order."""
import sys
import torch
sys.path.insert(0, ".")
from bench import fast_state_dict
from divergen_b200 import UNet2DConditionModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
unet = UNet2DConditionModel(device="cuda:0")
unet.load_state_dict(fast_state_dict(unet.expected_state_dict_shapes()))
unet.set_graphs(False)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 4, 64, 64, generator=g).half().cuda()
ehs = torch.randn(B, 77, 768, generator=g).half().cuda()
for _ in range(2):
    unet(x, 981, ehs)
torch.cuda.synchronize()
torch.cuda.profiler.start()
unet(x, 981, ehs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
