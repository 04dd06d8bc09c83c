"""Generate the golden input/output vectors under tests/golden/ from the oracle (run in the build container, CPU):

    python tests/golden/make_golden.py

PARITY UNPINNED: upstream diffusers is neither vendored nor pinned by the reference and cannot be imported here, so these
vectors pin the ORACLE (and through it the CUDA path) against regressions; they are not outputs of the reference itself.
The scalar known answers in ddim_known.json were derived independently of the oracle code.
Weights are not stored: `seeded_state_dict(cfg, seed)` regenerates them, and a checksum of the state dict is stored so a
change of torch's default initialisers is detected instead of silently changing the fixture.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ddim_oracle import DDIMOracle, denoise_loop  # noqa: E402
from oracle.unet_oracle import UNet2DConditionOracle, UNetConfig, seeded_state_dict, timestep_embedding  # noqa: E402


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().half().numpy().tobytes())
    return h.hexdigest()


def main():
    torch.set_num_threads(1)
    out = {}
    # ---- scheduler
    s = DDIMOracle()
    s.set_timesteps(50)
    out["alphas_cumprod"] = s.alphas_cumprod.numpy()
    out["timesteps50"] = s.timesteps.numpy()
    g = torch.Generator().manual_seed(7)
    x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    out["step_x"], out["step_model_out"] = x.numpy(), e.numpy()
    for pred in ("epsilon", "v_prediction"):
        sp = DDIMOracle(prediction_type=pred)
        sp.set_timesteps(50)
        for t in (981, 501, 1):
            out[f"step_{pred}_{t}"] = sp.step(e, t, x).prev_sample.numpy()
    out["temb_981_320"] = timestep_embedding(torch.tensor([981.0]), 320, True, 0)[0].numpy()
    # ---- tiny UNet forward + 2-step CFG loop (fp16-representable weights and inputs, fp32 arithmetic)
    meta = {}
    for name, linear in (("conv_proj", False), ("linear_proj", True)):
        cfg = UNetConfig.tiny(linear=linear)
        sd = {k: v.half().float() for k, v in seeded_state_dict(cfg, 0).items()}
        meta[f"state_dict_sha256_{name}"] = sd_checksum(sd)
        m = UNet2DConditionOracle(cfg).eval()
        m.load_state_dict(sd)
        g = torch.Generator().manual_seed(42)
        lat = torch.randn(1, 4, 16, 16, generator=g).half().float()
        pos = torch.randn(1, 77, 64, generator=g).half().float()
        neg = torch.randn(1, 77, 64, generator=g).half().float()
        with torch.no_grad():
            fwd = m(torch.cat([lat, lat]), 981, torch.cat([neg, pos])).sample
        loop = denoise_loop(m, DDIMOracle(), lat, pos, neg, num_inference_steps=2)
        out[f"{name}_lat"], out[f"{name}_pos"], out[f"{name}_neg"] = lat.numpy(), pos.numpy(), neg.numpy()
        out[f"{name}_forward_t981"] = fwd.numpy()
        out[f"{name}_loop2"] = loop.numpy()
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    known = {
        "source": "SURVEY.md 3.4 (derived numerically from the public SD scheduler config, independent of oracle/)",
        "alphas_cumprod": {"0": 0.999149978, "1": 0.998296022, "21": 0.980380654, "961": 0.007281722, "981": 0.005775496, "999": 0.004660095},
        "scalar_steps_x1_out0.5": {"epsilon_t981": {"x0": 6.598261, "prev": 1.061226}, "epsilon_t1": {"x0": 0.980196, "prev": 0.994357},
                                  "v_prediction_t981": {"x0": -0.422557, "eps": 1.035106, "prev": 0.995273}},
        "timestep_embedding_t981_dim320": {"0": 0.67995721, "1": -0.79842919, "2": 0.57806414, "160": 0.73325181, "161": 0.60208869,
                                           "162": 0.81599128, "319": 0.10372588},
        "param_counts": {"sd15": 859520964, "sd21": 865910724, "tensors": 686},
    }
    known.update(meta)
    json.dump(known, open(os.path.join(HERE, "ddim_known.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"), os.path.getsize(os.path.join(HERE, "oracle_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
