"""Known-answer tests for the DDIM oracle (constants re-derived in SURVEY.md 3.4)."""
import numpy as np
import torch

from oracle.ddim_oracle import DDIMOracle


def test_alphas_cumprod_known_answers():
    s = DDIMOracle()
    want = {0: 0.999149978, 1: 0.998296022, 21: 0.980380654, 961: 0.007281722, 981: 0.005775496, 999: 0.004660095}
    for i, v in want.items():
        assert abs(s.alphas_cumprod[i].item() - v) < 2e-7, (i, s.alphas_cumprod[i].item())


def test_timesteps_leading_offset1():
    s = DDIMOracle()
    s.set_timesteps(50)
    ts = s.timesteps.tolist()
    assert ts[:3] == [981, 961, 941] and ts[-3:] == [41, 21, 1] and len(ts) == 50


def _f64(s):
    s.alphas_cumprod = s.alphas_cumprod.double()
    s.final_alpha_cumprod = s.alphas_cumprod[0]
    return s


def test_step_scalar_known_answers():
    x = torch.ones(1, dtype=torch.float64)
    e = torch.full((1,), 0.5, dtype=torch.float64)
    s = _f64(DDIMOracle()); s.set_timesteps(50)
    o = s.step(e, 981, x)
    assert abs(o.pred_original_sample.item() - 6.598261) < 1e-5 and abs(o.prev_sample.item() - 1.061226) < 1e-5
    o = s.step(e, 1, x)  # prev_t < 0 -> final_alpha_cumprod
    assert abs(o.pred_original_sample.item() - 0.980196) < 1e-5 and abs(o.prev_sample.item() - 0.994357) < 1e-5
    s = _f64(DDIMOracle(prediction_type="v_prediction")); s.set_timesteps(50)
    o = s.step(e, 981, x)
    assert abs(o.pred_original_sample.item() + 0.422557) < 1e-5 and abs(o.prev_sample.item() - 0.995273) < 1e-5


def test_host_scheduler_matches_oracle_schedule():
    """The product's host-side schedule arithmetic (numpy) against the oracle's (torch)."""
    from divergen_b200.scheduler import DDIMScheduler
    p, o = DDIMScheduler(), DDIMOracle()
    np.testing.assert_allclose(p.alphas_cumprod, o.alphas_cumprod.numpy(), rtol=2e-6)
    for n in (50, 20, 1000):
        p.set_timesteps(n); o.set_timesteps(n)
        assert p.timesteps.tolist() == o.timesteps.tolist()
    p.set_timesteps(50)
    a_t, a_prev = p.alphas_for(1)
    assert abs(a_prev - 0.999149978) < 2e-7 and abs(a_t - 0.998296022) < 2e-7
