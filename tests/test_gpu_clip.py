"""GPU parity of the CLIP text encoder through the transformers-shaped surface / `dg_clip_encode`
against `transformers.CLIPTextModel` itself (the library the reference's pipeline calls; transformers is part of the image,
so this row's parity IS pinned to upstream): random-init weights rounded to fp16, same token ids.
Same statement as tests/test_gpu_unet.py: max|err| <= 2e-2 * max|ref| + 1e-3, cosine >= 0.999,
rms(E_new) <= 2 * rms(E_torch) + 1e-4 with E_torch = upstream run in fp16 on the GPU against upstream in fp32 on the CPU.
"""
import pytest
import torch

from tests.test_gpu_unet import DEV, _check, _need_gpu

pytestmark = pytest.mark.gpu


def _models(seed=0, **cfg_kw):
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HFCLIPTextModel
    from divergen_b200 import CLIPTextModel
    cfg_kw = dict(cfg_kw)
    cfg = CLIPTextConfig(hidden_act=cfg_kw.pop("hidden_act", "quick_gelu"), layer_norm_eps=1e-5, **cfg_kw)
    cfg_kw["hidden_act"] = cfg.hidden_act
    torch.manual_seed(seed)
    ref = HFCLIPTextModel(cfg).eval()
    sd = {k: v.half().float() for k, v in ref.state_dict().items() if not k.endswith("position_ids")}
    ref.load_state_dict(sd, strict=False)
    ours = CLIPTextModel(device=DEV, **cfg_kw)
    res = ours.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    return ref, ours


def _ids(b, s, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, vocab - 2, (b, s), generator=g)
    ids[:, 0] = vocab - 2                    # BOS
    for i in range(b):                       # EOS somewhere, padding (= EOS id, as CLIP's tokenizer pads) after it
        e = int(torch.randint(2, s, (1,), generator=g))
        ids[i, e:] = vocab - 1
    return ids


TINY = dict(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
            max_position_embeddings=77)


@pytest.mark.parametrize("b,s", [(1, 77), (3, 77), (2, 16)])
def test_tiny_clip_text_encoder(b, s):
    _need_gpu()
    ref, ours = _models(0, **TINY)
    ids = _ids(b, s, TINY["vocab_size"], 1)
    with torch.no_grad():
        ref32 = ref(input_ids=ids).last_hidden_state
        ref16 = ref.half().to(DEV)(input_ids=ids.to(DEV)).last_hidden_state
    out = ours(ids)
    got = out.last_hidden_state
    assert got.shape == (b, s, TINY["hidden_size"]) and out[0] is got
    _check(got, ref32, ref16, name=f"clip tiny b={b} s={s}")


def test_clip_is_causal():
    """Changing a later token must not change the embeddings of earlier positions (CLIP's causal mask)."""
    _need_gpu()
    _, ours = _models(0, **TINY)
    ids = _ids(1, 77, TINY["vocab_size"], 2)
    a = ours(ids).last_hidden_state.clone()
    ids2 = ids.clone()
    ids2[0, 40] = (ids2[0, 40] + 17) % 900 + 1
    b = ours(ids2).last_hidden_state
    assert torch.equal(a[0, :40], b[0, :40])
    assert not torch.equal(a[0, 40:], b[0, 40:])


def test_sd_clip_text_encoder_full_width():
    """The SD-1.x text encoder (ViT-L/14 text tower: 12 x 768, 123 M parameters), prompt + empty-prompt pair."""
    _need_gpu()
    from divergen_b200 import SD_CLIP_CONFIG
    ref, ours = _models(3, **SD_CLIP_CONFIG)
    assert len(ours.expected_state_dict_shapes()) == 196
    ids = _ids(2, 77, SD_CLIP_CONFIG["vocab_size"], 4)
    with torch.no_grad():
        ref32 = ref(input_ids=ids).last_hidden_state
        ref16 = ref.half().to(DEV)(input_ids=ids.to(DEV)).last_hidden_state
    got = ours(ids).last_hidden_state
    _check(got, ref32, ref16, name="clip sd 2x77")


def test_sd21_text_encoder_activation():
    """SD-2.x's OpenCLIP text tower uses erf GELU (`hidden_act: "gelu"`): same kernels, other activation (tiny width here;
    the full 23 x 1024 configuration is SD21_CLIP_CONFIG)."""
    _need_gpu()
    ref, ours = _models(5, hidden_act="gelu", **TINY)
    ids = _ids(2, 77, TINY["vocab_size"], 9)
    with torch.no_grad():
        ref32 = ref(input_ids=ids).last_hidden_state
        ref16 = ref.half().to(DEV)(input_ids=ids.to(DEV)).last_hidden_state
    _check(ours(ids).last_hidden_state, ref32, ref16, name="clip tiny gelu")
    ref_q, _ = _models(5, **TINY)                 # the activation really is different from quick_gelu
    with torch.no_grad():
        assert (ref_q(input_ids=ids).last_hidden_state - ref32).abs().max().item() > 1e-2


def test_clip_errors():
    _need_gpu()
    from divergen_b200 import CLIPTextModel
    with pytest.raises((RuntimeError, ValueError)):
        CLIPTextModel(device=DEV, hidden_size=96, num_attention_heads=2)          # head dim 48
    m = CLIPTextModel(device=DEV, **TINY)
    with pytest.raises((RuntimeError, ValueError)):
        m(torch.zeros(1, 77, dtype=torch.long))                                  # weights not set
    with pytest.raises(ValueError):
        m(torch.zeros(1, 78, dtype=torch.long))


def test_driver_embedding_table_with_device_encoder():
    """generate.encode_prompt_table: one row per distinct prompt plus the unconditional row 0, through the device encoder."""
    _need_gpu()
    from types import SimpleNamespace
    from divergen_b200.generate import encode_prompt_table
    ref, ours = _models(0, **TINY)

    class StubTokenizer:                       # CLIPTokenizer's calling convention; its vocabulary files are not shipped
        model_max_length = 77

        def __call__(self, texts, padding=None, max_length=77, truncation=True, return_tensors="pt"):
            rows = []
            for t in texts:
                toks = [998] + [1 + (ord(c) % 900) for c in t][:max_length - 2] + [999]
                rows.append(toks + [999] * (max_length - len(toks)))
            return SimpleNamespace(input_ids=torch.tensor(rows))

    prompts = ["a photo of a cat", "a photo of a single aerosol can", "x"]
    table = encode_prompt_table(prompts, StubTokenizer(), ours, batch=2)
    assert table.shape == (4, 77, TINY["hidden_size"]) and table.dtype == torch.float16
    with torch.no_grad():
        want = ref(input_ids=StubTokenizer()([""] + prompts).input_ids).last_hidden_state
    _check(table, want, name="driver embedding table")


# ---------------------------------------------------------------- row f3: CLIP similarity scores (both towers)
def _clip_models(seed, text_kw, vision_kw, proj):
    from transformers import CLIPConfig, CLIPModel
    from divergen_b200 import CLIPScorer
    v = text_kw["vocab_size"]             # the end-of-text token is the largest id, as in CLIP's vocabulary (49407)
    cfg = CLIPConfig(text_config=dict(hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=proj, eos_token_id=v - 1,
                                      bos_token_id=v - 2, pad_token_id=v - 1, **text_kw),
                     vision_config=dict(hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=proj, **vision_kw),
                     projection_dim=proj)
    torch.manual_seed(seed)
    ref = CLIPModel(cfg).eval()
    sd = {k: (v if k == "logit_scale" else v.half().float()) for k, v in ref.state_dict().items() if not k.endswith("position_ids")}
    ref.load_state_dict(sd, strict=False)
    ours = CLIPScorer(device=DEV, text_config=text_kw, vision_config=vision_kw, projection_dim=proj)
    res = ours.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    return ref, ours


TINY_VISION = dict(image_size=56, patch_size=14, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2)


@pytest.mark.parametrize("b,t", [(3, 1), (2, 2)])
def test_tiny_clip_scores(b, t):
    """logits_per_text against transformers.CLIPModel (fp32): |err| <= 1e-2 * max|ref| + 2e-2 (fp16 features, logits O(10))."""
    _need_gpu()
    ref, ours = _clip_models(0, TINY, TINY_VISION, 64)
    g = torch.Generator().manual_seed(5)
    px = torch.randn(b, 3, 56, 56, generator=g).half()
    ids = _ids(t, 77, TINY["vocab_size"], 6)
    with torch.no_grad():
        want = ref(input_ids=ids, pixel_values=px.float()).logits_per_text
    got = ours(px.to(DEV), ids).cpu()
    assert got.shape == (t, b)
    err = (got - want).abs().max().item()
    print(f"clip scores tiny b={b} t={t}: max err {err:.4g}, ref {want.flatten().tolist()}")
    assert err <= 1e-2 * want.abs().max().item() + 2e-2


def test_vit_l14_clip_scores_full_width():
    """ViT-L/14 (24 x 1024 image tower at 257 tokens + 12 x 768 text tower, 427.6 M parameters): 2 images, 1 prompt."""
    _need_gpu()
    from divergen_b200.clip import VIT_L14_TEXT, VIT_L14_VISION
    ref, ours = _clip_models(1, VIT_L14_TEXT, VIT_L14_VISION, 768)
    g = torch.Generator().manual_seed(7)
    px = torch.randn(2, 3, 224, 224, generator=g).half()
    ids = _ids(1, 77, VIT_L14_TEXT["vocab_size"], 8)
    with torch.no_grad():
        want = ref(input_ids=ids, pixel_values=px.float()).logits_per_text
    got = ours(px.to(DEV), ids).cpu()
    err = (got - want).abs().max().item()
    print(f"clip scores ViT-L/14: max err {err:.4g}, ref {want.flatten().tolist()}, got {got.flatten().tolist()}")
    assert err <= 1e-2 * want.abs().max().item() + 5e-3


def test_scoring_before_writing_equals_scoring_the_png(tmp_path):
    """Row f3 fused into generation: the score taken from the device uint8 image equals the score the filteration script
    would get from the written PNG (PIL open -> clip preprocess -> CLIPModel), here with a tiny random-init CLIP whose image
    size is 56 (so `preprocess` resizes 160x128 -> 70x56 and crops)."""
    _need_gpu()
    import numpy as np
    from types import SimpleNamespace
    from PIL import Image
    from divergen_b200.preprocess import CLIP_MEAN, CLIP_STD, clip_preprocess, resize_size
    ref, ours = _clip_models(0, TINY, TINY_VISION, 64)
    imgs = np.random.default_rng(3).integers(0, 256, (2, 160, 128, 3), dtype=np.uint8)
    ids = _ids(1, 77, TINY["vocab_size"], 11)
    # what the driver does (clip_scores_for, with the tokenizer output given directly)
    got = ours(clip_preprocess(torch.from_numpy(imgs).cuda(), n_px=56), ids).view(-1).cpu()
    # what filteration does after the fact
    mean, std = torch.tensor(CLIP_MEAN).view(3, 1, 1), torch.tensor(CLIP_STD).view(3, 1, 1)
    px = []
    for b in range(2):
        path = tmp_path / f"{b}.png"
        Image.fromarray(imgs[b]).save(path)
        im = Image.open(path).convert("RGB")
        oh, ow = resize_size(im.height, im.width, 56)
        im = im.resize((ow, oh), Image.BICUBIC)
        top, left = int(round((oh - 56) / 2.0)), int(round((ow - 56) / 2.0))
        t = torch.from_numpy(np.asarray(im)[top:top + 56, left:left + 56].copy()).permute(2, 0, 1).float() / 255.0
        px.append((t - mean) / std)
    with torch.no_grad():
        want = ref(input_ids=ids, pixel_values=torch.stack(px)).logits_per_text.view(-1)
    err = (got - want).abs().max().item()
    print(f"fused scoring: max err {err:.4g}, ref {want.tolist()}")
    assert err <= 1e-2 * want.abs().max().item() + 2e-2
