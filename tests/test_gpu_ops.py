"""GPU parity of every kernel on the hot path, called through the C ABI (divergen_b200.ops -> dg_op_*), against a
plain PyTorch fp32 reference of the same op on the same seeded inputs (floating-point kernels: SURVEY.md 8c).

Tolerances (fp16 storage, fp32 accumulate):  |err| <= atol + rtol*|ref|  with atol scaled to the output magnitude; DDIM
step <= 2 fp16 ulp.  The stats of the mismatch are printed on failure so a descriptor / layout bug shows its pattern.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from divergen_b200 import ops as _ops
    return _ops


def _report(name, got, ref, atol, rtol):
    got32, ref32 = got.float(), ref.float()
    err = (got32 - ref32).abs()
    tol = atol + rtol * ref32.abs()
    bad = err > tol
    nbad = int(bad.sum())
    if nbad or not torch.isfinite(got32).all():
        idx = bad.nonzero()[:8].tolist()
        flat = err.reshape(-1, err.shape[-1])
        col_err = flat.max(0).values
        row_err = flat.max(1).values
        msg = (f"{name}: {nbad}/{err.numel()} outside tol (atol {atol}, rtol {rtol}); max err {err.max().item():.4g}, "
               f"ref absmax {ref32.abs().max().item():.4g}, got absmax {got32.abs().max().item():.4g}, "
               f"finite={bool(torch.isfinite(got32).all())}; first bad idx {idx}; "
               f"bad cols(first 16 of {int((col_err > atol).sum())}): {(col_err > atol).nonzero().flatten()[:16].tolist()}; "
               f"bad rows(first 16 of {int((row_err > atol).sum())}): {(row_err > atol).nonzero().flatten()[:16].tolist()}")
        raise AssertionError(msg)


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV, torch.float16)


# ------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("B,H,W,C0,C1,groups,silu", [(2, 16, 16, 320, 0, 32, True), (2, 8, 8, 1280, 640, 32, True),
                                                     (1, 32, 32, 64, 0, 32, False), (3, 4, 4, 2560, 0, 32, True),
                                                     (2, 64, 64, 640, 320, 32, True)])
def test_groupnorm(ops, B, H, W, C0, C1, groups, silu):
    x0 = _rand(B, H, W, C0, seed=1) * 2 + 0.5
    x1 = _rand(B, H, W, C1, seed=2) if C1 else None
    C = C0 + C1
    gamma, beta = _rand(C, seed=3) * 0.2 + 1, _rand(C, seed=4) * 0.1
    got = ops.groupnorm_nhwc(x0, gamma, beta, groups, 1e-5, silu, x1)
    x = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), groups, gamma.float(), beta.float(), 1e-5)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 3, 1)
    _report("groupnorm", got, ref, 4e-3, 4e-3)


@pytest.mark.parametrize("rows,C", [(1000, 320), (257, 640), (64, 1280), (33, 64)])
def test_layernorm(ops, rows, C):
    x = _rand(rows, C, seed=5) * 3 + 1
    gamma, beta = _rand(C, seed=6) * 0.2 + 1, _rand(C, seed=7) * 0.1
    got = ops.layernorm(x, gamma, beta, 1e-5)
    ref = F.layer_norm(x.float(), (C,), gamma.float(), beta.float(), 1e-5)
    _report("layernorm", got, ref, 4e-3, 4e-3)


def test_time_embedding(ops):
    from oracle.unet_oracle import timestep_embedding
    w1, b1 = _rand(1280, 320, scale=0.05, seed=8), _rand(1280, scale=0.05, seed=9)
    w2, b2 = _rand(1280, 1280, scale=0.03, seed=10), _rand(1280, scale=0.05, seed=11)
    ts = [981.0, 1.0, 500.0]
    got = ops.time_embedding(ts, 320, w1, b1, w2, b2)
    t_emb = timestep_embedding(torch.tensor(ts), 320, True, 0).to(DEV).half().float()
    h = F.silu(t_emb @ w1.float().T + b1.float())
    ref = h.half().float() @ w2.float().T + b2.float()
    _report("time_embedding", got, ref, 3e-3, 5e-3)


@pytest.mark.parametrize("pred,guidance", [("epsilon", 7.5), ("v_prediction", 7.5), ("epsilon", 1.0)])
def test_cfg_ddim_step(ops, pred, guidance):
    """Against the oracle's DDIMOracle.step on the same inputs (fp32), tolerance 2 fp16 ulp of the result."""
    from oracle.ddim_oracle import DDIMOracle
    n = 4
    lat = _rand(n, 4, 64, 64, seed=12)
    noise = _rand(2 * n if guidance > 1 else n, 4, 64, 64, seed=13)
    s = DDIMOracle(prediction_type=pred)
    s.set_timesteps(50)
    for t in (981, 501, 1):
        prev_t = t - 20
        a_t = s.alphas_cumprod[t].item()
        a_prev = s.alphas_cumprod[prev_t].item() if prev_t >= 0 else s.final_alpha_cumprod.item()
        got = ops.cfg_ddim_step(noise, lat.clone(), a_t, a_prev, guidance, pred)
        nf = noise.float().cpu()
        if guidance > 1:
            u, c = nf.chunk(2)
            nf = u + guidance * (c - u)
        ref = s.step(nf, t, lat.float().cpu()).prev_sample.to(DEV)
        ulp = torch.clamp(ref.abs(), min=2.0 ** -14) * 2.0 ** -10
        err = (got.float() - ref).abs()
        assert (err <= 2 * ulp + 1e-6).all(), f"t={t}: max err {err.max().item()} ({(err / ulp).max().item():.2f} ulp)"


# ------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("M,K,N", [(128, 64, 160), (256, 320, 320), (1000, 640, 640), (4096, 320, 1280),
                                   (616, 768, 640), (77, 1024, 2560), (512, 2560, 1280), (130, 64, 64), (64, 128, 8)])
def test_gemm_plain(ops, M, K, N):
    x = _rand(M, K, seed=20)
    w = _rand(N, K, scale=1 / math.sqrt(K), seed=21)
    got = ops.linear(x, w)
    ref = x.float() @ w.float().T
    _report(f"gemm {M}x{K}x{N}", got, ref, 4e-3, 4e-3)


def test_gemm_bias_residual(ops):
    M, K, N = 2048, 1280, 640
    x, w = _rand(M, K, seed=22), _rand(N, K, scale=1 / math.sqrt(K), seed=23)
    b, r = _rand(N, seed=24), _rand(M, N, seed=25)
    got = ops.linear(x, w, b, r)
    ref = x.float() @ w.float().T + b.float() + r.float()
    _report("gemm+bias+res", got, ref, 5e-3, 4e-3)


@pytest.mark.parametrize("M,K,inner", [(256, 320, 1280), (1000, 640, 2560), (64, 128, 512)])
def test_gemm_geglu(ops, M, K, inner):
    x = _rand(M, K, seed=26)
    w, b = _rand(2 * inner, K, scale=1 / math.sqrt(K), seed=27), _rand(2 * inner, scale=0.1, seed=28)
    got = ops.geglu_linear(x, w, b)
    y = x.float() @ w.float().T + b.float()
    h, g = y.chunk(2, -1)
    ref = h * F.gelu(g)
    _report("geglu", got, ref, 5e-3, 5e-3)


@pytest.mark.parametrize("M,K,N", [(256, 320, 320), (1000, 640, 960), (4096, 320, 320), (77, 1280, 1280), (512, 1280, 3840)])
def test_gemm_layernorm_fold(ops, M, K, N):
    """LayerNorm folded into the GEMM epilogue (rstd*(x Wf^T - mean*colsum) + b32) vs F.layer_norm + matmul in fp32."""
    x = _rand(M, K, seed=50) * 2 + 0.7
    gamma, beta = _rand(K, seed=51) * 0.2 + 1, _rand(K, seed=52) * 0.1
    w, b = _rand(N, K, scale=1 / math.sqrt(K), seed=53), _rand(N, scale=0.1, seed=54)
    got = ops.layernorm_linear(x, gamma, beta, w, b)
    ref = F.layer_norm(x.float(), (K,), gamma.float(), beta.float(), 1e-5) @ w.float().T + b.float()
    _report(f"ln-fold gemm {M}x{K}x{N}", got, ref, 6e-3, 5e-3)


def test_gemm_layernorm_fold_geglu(ops):
    M, K, inner = 1000, 320, 1280
    x = _rand(M, K, seed=55) * 1.5 - 0.3
    gamma, beta = _rand(K, seed=56) * 0.2 + 1, _rand(K, seed=57) * 0.1
    w, b = _rand(2 * inner, K, scale=1 / math.sqrt(K), seed=58), _rand(2 * inner, scale=0.1, seed=59)
    got = ops.layernorm_linear(x, gamma, beta, w, b, geglu=True)
    y = F.layer_norm(x.float(), (K,), gamma.float(), beta.float(), 1e-5) @ w.float().T + b.float()
    h, g = y.chunk(2, -1)
    _report("ln-fold geglu", got, h * F.gelu(g), 6e-3, 6e-3)


@pytest.mark.parametrize("M,K,N,hw,blk", [(2048, 320, 320, 1024, 10), (512, 640, 1280, 64, 40), (8192, 64, 320, 4096, 10),
                                          (96, 128, 128, 32, 4), (192, 2560, 640, 64, 20)])
def test_gemm_fused_statistics(ops, M, K, N, hw, blk):
    """Row (LayerNorm) and slab/block (GroupNorm) sums produced by the GEMM epilogue match the sums of its fp16 output
    (they are taken on the fp32 values just before rounding: tolerance = the fp16 rounding noise of the summed elements)."""
    x, w = _rand(M, K, seed=60), _rand(N, K, scale=1 / math.sqrt(K), seed=61)
    b, r = _rand(N, seed=62), _rand(M, N, seed=63)
    out, rs, gs = ops.linear_stats(x, w, b, r, gn_blk=blk, hw=hw)
    ref = x.float() @ w.float().T + b.float() + r.float()
    _report("gemm(stats) value", out, ref, 5e-3, 4e-3)
    o = out.float()
    want_rs = torch.stack([o.sum(1), (o * o).sum(1)], 1)
    assert torch.allclose(rs, want_rs, rtol=2e-3, atol=5e-2), (rs - want_rs).abs().max().item()
    ob = o.view(M // hw, hw // 32, 32, N // blk, blk)
    want_gs = torch.stack([ob.sum((2, 4)), (ob * ob).sum((2, 4))], -1)
    assert torch.allclose(gs, want_gs, rtol=2e-3, atol=5e-2), (gs - want_gs).abs().max().item()
    # the producer's row partials feed the LayerNorm-folded consumer directly
    if K == N:
        pass


@pytest.mark.parametrize("B,H,W,C,N,blk", [(2, 32, 32, 320, 320, 10), (3, 8, 8, 128, 640, 20), (2, 16, 8, 64, 64, 2)])
def test_conv3x3_groupnorm_fused(ops, B, H, W, C, N, blk):
    """conv3x3 epilogue block sums -> GroupNorm apply from those sums == GroupNorm(SiLU) of the conv output."""
    x = _rand(B, H, W, C, seed=64)
    w, bias = _rand(N, C, 3, 3, scale=1 / math.sqrt(9 * C), seed=65), _rand(N, scale=0.1, seed=66)
    out, gs = ops.conv3x3_stats(x, w, bias, blk)
    ob = out.float().view(B, H * W, N // blk, blk)
    want = torch.stack([ob.sum((1, 3)), (ob * ob).sum((1, 3))], -1)     # slab order inside a sample is the kernel's business
    assert torch.allclose(gs.sum(1), want, rtol=2e-3, atol=1e-1), (gs.sum(1) - want).abs().max().item()
    gamma, beta = _rand(N, seed=67) * 0.2 + 1, _rand(N, seed=68) * 0.1
    got = ops.groupnorm_fused_nhwc(out, gs, gamma, beta, 32, 1e-5, True, blk)
    ref = F.silu(F.group_norm(out.float().permute(0, 3, 1, 2), 32, gamma.float(), beta.float(), 1e-5)).permute(0, 2, 3, 1)
    _report("groupnorm from fused sums", got, ref, 4e-3, 4e-3)


def _block_sums(x, blk):
    """[B, H*W // 32, C // blk, 2] (sum, sum of squares) of an NHWC tensor: what the GEMM epilogues leave for the next GroupNorm."""
    B, H, W, C = x.shape
    xb = x.float().view(B, H * W // 32, 32, C // blk, blk)
    return torch.stack([xb.sum((2, 4)), (xb * xb).sum((2, 4))], -1).contiguous()


@pytest.mark.parametrize("B,H,W,C0,C1,N,blk,silu,taps", [
    (2, 64, 64, 320, 0, 320, 10, True, 9),        # level-0 resnet conv (320-wide tiles, one sample per tile)
    (2, 32, 32, 640, 320, 640, 10, True, 9),      # up-block conv over cat([x, skip]): two sources, groups span both
    (2, 16, 16, 1280, 0, 1280, 40, True, 9),      # few-tile level: 160-wide tiles
    (3, 8, 8, 1280, 1280, 1280, 40, True, 9),     # 8x8 level: a tile holds two samples, split-K
    (8, 8, 8, 1280, 0, 1280, 40, True, 9),        # batch 8 at the bottom of the U (the benchmarked shape)
    (2, 64, 64, 320, 0, 4, 10, True, 9),          # conv_norm_out -> conv_out (4 output channels)
    (5, 4, 8, 64, 0, 64, 2, True, 9),             # tiny-UNet level: 32 pixels per sample, four samples per tile, ragged batch
    (2, 24, 24, 128, 0, 128, 4, True, 9),         # 96x96 latents' 24x24 level: boxes that do not divide the image evenly
    (2, 32, 32, 320, 0, 320, 10, False, 1),       # transformer GroupNorm (no SiLU) -> proj_in as a 1x1 conv
    (3, 8, 8, 128, 0, 256, 4, False, 1),          # ... with several samples per 128-row tile
])
def test_conv_groupnorm_silu_fused_into_operand_path(ops, B, H, W, C0, C1, N, blk, silu, taps):
    """conv(act(norm(x))) with GroupNorm + SiLU applied inside the GEMM's A-operand path (gemm2_kernel<..., kXf>) against
    the fp32 torch reference, and against the two-kernel path (GroupNorm apply, then conv) it replaces."""
    x0 = _rand(B, H, W, C0, seed=70) * 1.5 + 0.7            # a mean well away from zero: exercises the mean_h subtraction
    x1 = (_rand(B, H, W, C1, seed=71) * 0.5 - 1.0) if C1 else None
    C = C0 + C1
    gamma, beta = _rand(C, seed=72) * 0.2 + 1, _rand(C, seed=73) * 0.1
    if taps == 9:
        w = _rand(N, C, 3, 3, scale=1 / math.sqrt(9 * C), seed=74)
    else:
        w = _rand(N, C, scale=1 / math.sqrt(C), seed=74)
    bias = _rand(N, scale=0.1, seed=75)
    res = _rand(B, H, W, N, seed=76) if N % 8 == 0 else None
    got = ops.conv_gn(x0, _block_sums(x0, blk), gamma, beta, 32, 1e-5, silu, blk, w, bias, x1, _block_sums(x1, blk) if C1 else None, res)
    x = torch.cat([x0, x1], -1) if C1 else x0
    hn = F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma.float(), beta.float(), 1e-5)
    hn = F.silu(hn) if silu else hn
    wf = w.float() if taps == 9 else w.float()[:, :, None, None]
    ref = F.conv2d(hn, wf, bias.float(), padding=1 if taps == 9 else 0).permute(0, 2, 3, 1)
    if res is not None:
        ref = ref + res.float()
    _report(f"conv_gn {B}x{H}x{W}x{C}->{N} taps {taps}", got, ref, 8e-3, 6e-3)
    # the path it replaces, for scale: GroupNorm apply kernel -> conv kernel
    hn16 = ops.groupnorm_fused_nhwc(x0, _block_sums(x0, blk), gamma, beta, 32, 1e-5, silu, blk, x1, _block_sums(x1, blk) if C1 else None)
    if taps == 9:
        old = ops.conv3x3_nhwc(hn16, w, bias, None, None, res)
    else:
        old = ops.linear(hn16.view(-1, C), w, bias, res.view(-1, N) if res is not None else None).view(B, H, W, N)
    e_new = (got.float() - ref).pow(2).mean().sqrt().item()
    e_old = (old.float() - ref).pow(2).mean().sqrt().item()
    print(f"conv_gn {B}x{H}x{W}x{C}->{N}: rms error fused {e_new:.3g} vs two-kernel {e_old:.3g} (ref rms {ref.pow(2).mean().sqrt().item():.3g})")
    assert e_new <= 3 * e_old + 1e-4


@pytest.mark.parametrize("B,H,W,C0,C1,N", [(1, 64, 64, 64, 0, 160), (2, 64, 64, 320, 0, 320), (2, 32, 32, 640, 320, 640),
                                           (2, 16, 16, 1280, 0, 1280), (2, 8, 8, 1280, 1280, 1280), (3, 8, 8, 64, 0, 64),
                                           (2, 64, 64, 320, 0, 4), (1, 24, 24, 128, 0, 128), (2, 4, 4, 128, 0, 128)])
def test_conv3x3(ops, B, H, W, C0, C1, N):
    x0 = _rand(B, H, W, C0, seed=30)
    x1 = _rand(B, H, W, C1, seed=31) if C1 else None
    C = C0 + C1
    w = _rand(N, C, 3, 3, scale=1 / math.sqrt(9 * C), seed=32)
    bias = _rand(N, scale=0.1, seed=33)
    got = ops.conv3x3_nhwc(x0, w, bias, x1)
    x = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), padding=1).permute(0, 2, 3, 1)
    _report(f"conv3x3 {B}x{H}x{W}x{C}->{N}", got, ref, 5e-3, 4e-3)


@pytest.mark.parametrize("B,H,W,C,N,blk", [(2, 32, 32, 640, 640, 20), (8, 8, 8, 1280, 1280, 40), (3, 16, 16, 128, 64, 2),
                                           (1, 8, 16, 64, 192, 0), (2, 2, 2, 64, 64, 0), (2, 64, 64, 128, 128, 4)])
def test_upsample_conv3x3_phases(ops, B, H, W, C, N, blk):
    """Upsample2D (nearest x2 + conv3x3) as four 2x2 phase convolutions on the low-resolution input vs the torch fp32 reference;
    the fused GroupNorm block sums cover the whole upsampled output."""
    x = _rand(B, H, W, C, seed=80)
    w, bias = _rand(N, C, 3, 3, scale=1 / math.sqrt(9 * C), seed=81), _rand(N, scale=0.1, seed=82)
    res = ops.upsample_conv3x3(x, w, bias, blk)
    out, gs = res if blk else (res, None)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.float(), bias.float(), padding=1).permute(0, 2, 3, 1)
    _report(f"upsample conv {B}x{H}x{W}x{C}->{N}", out, ref, 6e-3, 5e-3)     # (phase weights are sums of up to 4 fp16 weights, rounded once)
    if blk:
        ob = out.float().reshape(B, 4 * H * W, N // blk, blk)
        want = torch.stack([ob.sum((1, 3)), (ob * ob).sum((1, 3))], -1)
        assert torch.isfinite(gs).all()
        assert torch.allclose(gs.sum(1), want, rtol=2e-3, atol=1e-1), (gs.sum(1) - want).abs().max().item()


@pytest.mark.parametrize("B,H,W,C,N", [(2, 64, 64, 320, 320), (8, 16, 16, 1280, 1280), (3, 8, 8, 64, 128), (2, 4, 4, 64, 64),
                                       (1, 96, 96, 64, 64)])
def test_conv3x3_stride2(ops, B, H, W, C, N):
    """Downsample2D (conv3x3, stride 2, pad 1) through the strided operand map vs the torch fp32 reference."""
    x = _rand(B, H, W, C, seed=85)
    w, bias = _rand(N, C, 3, 3, scale=1 / math.sqrt(9 * C), seed=86), _rand(N, scale=0.1, seed=87)
    got = ops.conv3x3_stride2(x, w, bias)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    _report(f"conv3x3 stride 2 {B}x{H}x{W}x{C}->{N}", got, ref, 5e-3, 4e-3)


@pytest.mark.parametrize("B,H,W,C,N,Cs0,Cs1", [(2, 32, 32, 640, 640, 640, 320), (2, 64, 64, 320, 320, 320, 320), (8, 8, 8, 1280, 1280, 1280, 1280),
                                               (3, 16, 16, 128, 128, 64, 0), (2, 16, 16, 1280, 1280, 640, 0)])
def test_conv3x3_with_shortcut_in_the_k_loop(ops, B, H, W, C, N, Cs0, Cs1):
    """ResnetBlock2D tail: conv2(h) + conv_shortcut(cat[x, skip]) as one GEMM whose K loop runs over the nine taps of h and then
    over the channels of x / skip (no shortcut tensor, no residual read) vs the torch fp32 reference."""
    h = _rand(B, H, W, C, seed=90)
    x0 = _rand(B, H, W, Cs0, seed=91)
    x1 = _rand(B, H, W, Cs1, seed=92) if Cs1 else None
    w, bias = _rand(N, C, 3, 3, scale=1 / math.sqrt(9 * C), seed=93), _rand(N, scale=0.1, seed=94)
    ws, bs = _rand(N, Cs0 + Cs1, scale=1 / math.sqrt(Cs0 + Cs1), seed=95), _rand(N, scale=0.1, seed=96)
    got = ops.conv3x3_shortcut(h, w, bias, x0, ws, bs, x1)
    xs = torch.cat([x0, x1], -1) if Cs1 else x0
    ref = F.conv2d(h.float().permute(0, 3, 1, 2), w.float(), bias.float(), padding=1).permute(0, 2, 3, 1)
    ref = ref + xs.float() @ ws.float().T + bs.float()
    _report(f"conv3x3 + shortcut {B}x{H}x{W} {C}+{Cs0}+{Cs1}->{N}", got, ref, 6e-3, 4e-3)


def test_conv3x3_temb_residual(ops):
    B, H, W, C, N = 2, 32, 32, 320, 320
    x = _rand(B, H, W, C, seed=34)
    w, bias = _rand(N, C, 3, 3, scale=1 / math.sqrt(9 * C), seed=35), _rand(N, scale=0.1, seed=36)
    table = _rand(B, 1000, seed=37)          # a wider per-sample table; the layer reads a column window of it
    res = _rand(B, H, W, N, seed=38)
    rv = table[:, 160:160 + N]
    from divergen_b200 import _lib
    import ctypes as C_
    lib, ctx = _lib.load(), _lib.context(0)
    wp = torch.empty((N, 9 * C), dtype=torch.float16, device=DEV)
    s = C_.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dg_op_pack_conv3x3(ctx, w.data_ptr(), wp.data_ptr(), N, C, s))
    out = torch.empty((B, H, W, N), dtype=torch.float16, device=DEV)
    _lib.check(lib.dg_op_conv3x3(ctx, x.data_ptr(), C, None, 0, wp.data_ptr(), bias.data_ptr(), rv.data_ptr(), 1000,
                                 res.data_ptr(), out.data_ptr(), B, H, W, N, N, s))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), padding=1).permute(0, 2, 3, 1)
    ref = ref + rv.float()[:, None, None, :] + res.float()
    _report("conv3x3+temb+res", out, ref, 6e-3, 4e-3)


# ------------------------------------------------------------------ tcgen05 attention
def _attn_ref(q, k, v, heads):
    B, Sq, C = q.shape
    d = C // heads
    qh = q.float().view(B, Sq, heads, d).transpose(1, 2)
    kh = k.float().view(B, -1, heads, d).transpose(1, 2)
    vh = v.float().view(B, -1, heads, d).transpose(1, 2)
    w = torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, -1)
    return (w @ vh).transpose(1, 2).reshape(B, Sq, C)


@pytest.mark.parametrize("B,heads,Sq,Sk,d", [(1, 1, 256, 128, 64), (2, 5, 576, 576, 64), (2, 8, 1024, 1024, 40),
                                             (2, 8, 256, 256, 80), (2, 8, 256, 256, 160), (2, 8, 64, 64, 160),
                                             (2, 8, 1024, 77, 40), (2, 8, 256, 77, 80), (2, 8, 64, 77, 160),
                                             (1, 2, 256, 256, 32), (1, 4, 16, 77, 32), (1, 8, 4096, 4096, 40),
                                             # wave-balanced grids (plan_attn_grid: CTAs of kQTiles and kQTiles - 1 query tiles)
                                             (5, 8, 2048, 512, 40), (4, 8, 4096, 640, 40), (5, 8, 1024, 512, 80),
                                             (10, 10, 384, 512, 64), (25, 2, 896, 577, 160), (8, 8, 4096, 77, 40)])
def test_attention(ops, B, heads, Sq, Sk, d):
    C = heads * d
    qkv_self = Sq == Sk
    if qkv_self:
        qkv = _rand(B, Sq, 3 * C, seed=40)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = _rand(B, Sq, C, seed=41)
        kv = _rand(B, Sk, 2 * C, seed=42)
        k, v = kv[..., :C], kv[..., C:]
    got = ops.attention(q, k, v, heads)
    ref = _attn_ref(q, k, v, heads)
    _report(f"attention B{B} h{heads} {Sq}x{Sk} d{d}", got, ref, 3e-3, 1e-2)


def test_attention_large_logits(ops):
    """Exercises the lazy-rescale path: logits with a wide dynamic range and a late-arriving maximum."""
    B, heads, S, d = 1, 2, 512, 64
    C = heads * d
    q, k, v = _rand(B, S, C, scale=3.0, seed=43), _rand(B, S, C, scale=3.0, seed=44), _rand(B, S, C, seed=45)
    k[:, -7:, :] *= 3.0
    got = ops.attention(q, k, v, heads)
    ref = _attn_ref(q, k, v, heads)
    _report("attention large logits", got, ref, 4e-3, 1e-2)
