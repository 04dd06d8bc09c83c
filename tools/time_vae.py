"""Time AutoencoderKL.decode (SD VAE, random-init) at the reference's image size: python tools/time_vae.py [batch]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from divergen_b200 import AutoencoderKL
from divergen_b200.generate import random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
vae = AutoencoderKL(device=dev)
vae.load_state_dict(random_state_dict(vae, dev, 1))
z = torch.randn(B, 4, 64, 64, device=dev).half()
for _ in range(3):
    y = vae.decode(z).sample
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = vae.decode(z).sample
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
# decoder MACs for 64x64 latents: ~1.24 TFLOP per image
print(f"vae decode batch {B} 512x512: {ms:.2f} ms ({ms / B:.2f} ms/image), finite={torch.isfinite(y.float()).all().item()} absmax={y.float().abs().max().item():.3f}")
