"""Event-timed throughput of the rows either side of the loop: VAE decode, CLIP text encoder, CLIP scorer,
uint8 conversion.  Random-init weights of the real sizes.  python tools/time_rows_f.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from divergen_b200 import AutoencoderKL, CLIPScorer, CLIPTextModel, ops
from divergen_b200.generate import random_state_dict

dev = torch.device("cuda:0")


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


vae = AutoencoderKL(device=dev)
vae.load_state_dict(random_state_dict(vae, dev, 1))
z = torch.randn(4, 4, 64, 64, device=dev).half()
ms = timeit(lambda: vae.decode(z))
# decoder: 1.24 TMAC per 512x512 image
print(f"f1 VAE decode, batch 4 @512x512: {ms:.2f} ms ({ms / 4:.2f} ms/image, {4 * 2.48e12 / ms / 1e9:.0f} TFLOP/s)")
img = vae.decode(z).sample
ms = timeit(lambda: ops.image_to_uint8(img), reps=20)
print(f"f4 image_to_uint8, batch 4 @512x512: {ms * 1e3:.1f} us ({img.numel() * 2 / ms / 1e6:.0f} GB/s read)")

clip = CLIPTextModel(device=dev)
clip.load_state_dict(random_state_dict(clip, dev, 2))
ids = torch.randint(1, 49000, (64, 77))
ids[:, 0] = 49406
ids[:, 20:] = 49407
ms = timeit(lambda: clip(ids))
# 12 layers x (4 C^2 + 2 C I) MACs per token + attention
flops = 64 * 77 * 12 * 2 * (4 * 768 * 768 + 2 * 768 * 3072)
print(f"f2 CLIP text encoder, 64 prompts x 77 tokens: {ms:.2f} ms ({ms / 64 * 1e3:.0f} us/prompt, {flops / ms / 1e9:.0f} TFLOP/s, host sync included)")

sc = CLIPScorer(device=dev)
sd = random_state_dict(sc, dev, 3)
sc.load_state_dict(sd)
px = torch.randn(32, 3, 224, 224, device=dev).half()
tid = ids[:1]
ms = timeit(lambda: sc(px, tid), reps=3)
vflops = 32 * 257 * 24 * 2 * (4 * 1024 * 1024 + 2 * 1024 * 4096)
print(f"f3 CLIP scorer (ViT-L/14), 32 images + 1 prompt: {ms:.2f} ms ({ms / 32:.2f} ms/image, {vflops / ms / 1e9:.0f} TFLOP/s image tower)")
