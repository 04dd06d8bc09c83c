"""Event-timed 50-step CFG loop (SD-1.5, 4 images = UNet batch 8) through StableDiffusionPipeline.__call__ with latents out:
ms per loop and per step (A/B env switches: DG_CFG_DEDUP, DG_FUSE_SC, ...)."""
import sys
import torch
sys.path.insert(0, ".")
from bench import fast_state_dict
from divergen_b200 import DDIMScheduler, StableDiffusionPipeline, UNet2DConditionModel

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 4
unet = UNet2DConditionModel(device="cuda:0")
unet.load_state_dict(fast_state_dict(unet.expected_state_dict_shapes()))
pipe = StableDiffusionPipeline(unet, DDIMScheduler())
g = torch.Generator().manual_seed(0)
pos = torch.randn(n_img, 77, 768, generator=g).half().cuda()
neg = torch.randn(n_img, 77, 768, generator=g).half().cuda()
lat = torch.randn(n_img, 4, 64, 64, generator=g).half().cuda()
call = lambda: pipe(prompt_embeds=pos, negative_prompt_embeds=neg, latents=lat.clone(), num_inference_steps=50, output_type="latent").images
out = call(); call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4):
    out = call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 4
print(f"loop {ms:.1f} ms  {ms / 50:.3f} ms per step  {n_img / ms * 1e3:.3f} images/s  checksum {out.float().abs().mean().item():.6f}", flush=True)
