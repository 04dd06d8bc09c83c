"""`CLIPTextModel` -- the transformers class surface over `dg_clip_*` (include/divergen_b200.h), SURVEY.md 8f row f2.

The reference obtains `prompt_embeds, negative_embeds = stage_1.encode_prompt(prompt)`
(DiverGen/generation/txt2img_diffusers_stages_from_txt.py:242); inside diffusers that is
`text_encoder(tokenizer(prompt, padding="max_length", max_length=77, truncation=True).input_ids)[0]`.
Keeps: `.config.{vocab_size, hidden_size, intermediate_size, num_hidden_layers, num_attention_heads,
max_position_embeddings}`, `.dtype`, `.device`, `.to()`, `.eval()`, `load_state_dict` with transformers key names
(196 tensors, 123 060 480 parameters for SD-1.x; `position_ids` buffers are accepted and ignored) and
`forward(input_ids) -> .last_hidden_state` (also `[0]`).  Attention masks other than CLIP's built-in causal mask, pooled
output and hidden-state lists are not provided: the Stable-Diffusion pipeline uses none of them.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, Tuple

import torch

from . import _lib

SD_CLIP_CONFIG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                      num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu")
# SD-2.x: OpenCLIP ViT-H/14 text tower as shipped in `stabilityai/stable-diffusion-2-1/text_encoder/config.json` (the
# penultimate-layer trick is baked in: the checkpoint has 23 layers)
SD21_CLIP_CONFIG = dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                        num_attention_heads=16, max_position_embeddings=77, hidden_act="gelu")


class CLIPTextOutput(tuple):
    """`BaseModelOutputWithPooling`-like: `out.last_hidden_state` and `out[0]`."""
    def __new__(cls, last_hidden_state):
        self = super().__new__(cls, (last_hidden_state,))
        self.last_hidden_state = last_hidden_state
        return self


class CLIPTextModel:
    def __init__(self, device="cuda:0", **config):
        cfg = dict(SD_CLIP_CONFIG)
        cfg.update(config)
        act = cfg.setdefault("hidden_act", "quick_gelu")
        if act not in ("quick_gelu", "gelu"):
            raise ValueError("hidden_act must be 'quick_gelu' (SD-1.x text encoder) or 'gelu' (SD-2.x OpenCLIP text encoder)")
        self.config = SimpleNamespace(**cfg)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("divergen_b200 has no CPU path; device must be a CUDA (sm_100) device")
        self.dtype = torch.float16
        self._lib = _lib.load()
        self._ctx = _lib.context(self.device.index or 0)
        h = C.c_void_p()
        _lib.check(self._lib.dg_clip_create(self._ctx, cfg["vocab_size"], cfg["hidden_size"], cfg["intermediate_size"],
                                            cfg["num_hidden_layers"], cfg["num_attention_heads"],
                                            cfg["max_position_embeddings"], C.byref(h)), "dg_clip_create")
        self._h = h
        _lib.check(self._lib.dg_clip_set_activation(h, {"quick_gelu": 0, "gelu": 1}[act]), "dg_clip_set_activation")
        self._prepared = 0

    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.dg_clip_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def expected_state_dict_shapes(self) -> Dict[str, Tuple[int, ...]]:
        out = {}
        shp, nd = (C.c_int64 * 4)(), C.c_int32()
        for i in range(self._lib.dg_clip_num_weights(self._h)):
            name = self._lib.dg_clip_weight_name(self._h, i).decode()
            _lib.check(self._lib.dg_clip_weight_shape(self._h, i, shp, C.byref(nd)))
            out[name] = tuple(shp[k] for k in range(nd.value))
        return out

    def load_state_dict(self, state_dict, strict: bool = True):
        expected = self.expected_state_dict_shapes()
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected and not k.endswith("position_ids")]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} (+{max(0, len(missing) - 5)}), "
                               f"unexpected {unexpected[:5]} (+{max(0, len(unexpected) - 5)})")
        for k, v in state_dict.items():
            if k not in expected:
                continue
            t = v.detach().to(device=self.device, dtype=torch.float16).contiguous()
            shp = (C.c_int64 * max(1, t.dim()))(*t.shape)
            _lib.check(self._lib.dg_clip_set_weight(self._h, k.encode(), C.c_void_p(t.data_ptr()), t.dim(), shp),
                       f"dg_clip_set_weight({k})")
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def forward(self, input_ids, attention_mask=None, position_ids=None, output_attentions=None,
                output_hidden_states=None, return_dict=None):
        for name, val in (("attention_mask", attention_mask), ("position_ids", position_ids),
                          ("output_attentions", output_attentions), ("output_hidden_states", output_hidden_states)):
            if val is not None and val is not False:
                raise ValueError(f"CLIPTextModel.forward: `{name}` is not supported by divergen_b200")
        ids = torch.as_tensor(input_ids)
        if ids.dim() == 1:
            ids = ids[None]
        if ids.dim() != 2 or ids.shape[1] > self.config.max_position_embeddings:
            raise ValueError(f"input_ids must be [batch, <= {self.config.max_position_embeddings}], got {tuple(ids.shape)}")
        ids = ids.to("cpu", torch.int32).contiguous()
        b, s = ids.shape
        if b > self._prepared:
            _lib.check(self._lib.dg_clip_prepare(self._h, b), "dg_clip_prepare")
            self._prepared = b
        out = torch.empty((b, s, self.config.hidden_size), dtype=torch.float16, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.dg_clip_encode(self._h, C.cast(ids.data_ptr(), C.POINTER(C.c_int32)), b, s,
                                            C.c_void_p(out.data_ptr()), C.c_void_p(stream)), "dg_clip_encode")
        torch.cuda.current_stream(self.device).synchronize()      # `ids` is pageable host memory: keep it alive until copied
        return CLIPTextOutput(out)

    __call__ = forward


VIT_L14_TEXT = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                    max_position_embeddings=77)
VIT_L14_VISION = dict(image_size=224, patch_size=14, hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                      num_attention_heads=16)


class CLIPScorer:
    """`_, logits_per_text = clip_model(images, text)` (OpenAI CLIP ViT-L/14 in DiverGen/filteration/get_clip_score.py:176-180;
    SURVEY.md 8f row f3) over `dg_clipscore_*`.  Weights: a transformers `CLIPModel` state dict (text_model.*, vision_model.*,
    visual_projection.weight, text_projection.weight, logit_scale).  `pixel_values` are preprocessed images
    [B, 3, image, image] (clip's `preprocess`: bicubic resize, centre crop, mean / std normalisation -- host side), `input_ids`
    the tokenizer output [T, <= 77]; the end-of-text position is `input_ids.argmax(-1)` as in OpenAI CLIP."""

    def __init__(self, device="cuda:0", text_config=None, vision_config=None, projection_dim: int = 768):
        t = dict(VIT_L14_TEXT); t.update(text_config or {})
        v = dict(VIT_L14_VISION); v.update(vision_config or {})
        self.text_config, self.vision_config = SimpleNamespace(**t), SimpleNamespace(**v)
        self.projection_dim = projection_dim
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("divergen_b200 has no CPU path; device must be a CUDA (sm_100) device")
        self._lib = _lib.load()
        self._ctx = _lib.context(self.device.index or 0)
        tc = (C.c_int32 * 6)(t["vocab_size"], t["hidden_size"], t["intermediate_size"], t["num_hidden_layers"],
                             t["num_attention_heads"], t["max_position_embeddings"])
        vc = (C.c_int32 * 6)(v["image_size"], v["patch_size"], v["hidden_size"], v["intermediate_size"], v["num_hidden_layers"],
                             v["num_attention_heads"])
        h = C.c_void_p()
        _lib.check(self._lib.dg_clipscore_create(self._ctx, tc, vc, projection_dim, C.byref(h)), "dg_clipscore_create")
        self._h = h
        self._prepared = (0, 0)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.dg_clipscore_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def expected_state_dict_shapes(self) -> Dict[str, Tuple[int, ...]]:
        out = {}
        shp, nd = (C.c_int64 * 4)(), C.c_int32()
        for i in range(self._lib.dg_clipscore_num_weights(self._h)):
            name = self._lib.dg_clipscore_weight_name(self._h, i).decode()
            _lib.check(self._lib.dg_clipscore_weight_shape(self._h, i, shp, C.byref(nd)))
            out[name] = tuple(shp[k] for k in range(nd.value))
        return out

    def load_state_dict(self, state_dict, strict: bool = True):
        expected = self.expected_state_dict_shapes()
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected and not k.endswith("position_ids") and k != "logit_scale"]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} (+{max(0, len(missing) - 5)}), "
                               f"unexpected {unexpected[:5]} (+{max(0, len(unexpected) - 5)})")
        for k, v in state_dict.items():
            if k == "logit_scale":
                _lib.check(self._lib.dg_clipscore_set_logit_scale(self._h, float(v)), "dg_clipscore_set_logit_scale")
                continue
            if k not in expected:
                continue
            t = v.detach().to(device=self.device, dtype=torch.float16).contiguous()
            shp = (C.c_int64 * max(1, t.dim()))(*t.shape)
            _lib.check(self._lib.dg_clipscore_set_weight(self._h, k.encode(), C.c_void_p(t.data_ptr()), t.dim(), shp),
                       f"dg_clipscore_set_weight({k})")
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def __call__(self, pixel_values: torch.Tensor, input_ids) -> torch.Tensor:
        """-> logits_per_text [T, B] fp32 on the device (`logits_per_image` is its transpose)."""
        v = self.vision_config
        if pixel_values.dim() != 4 or tuple(pixel_values.shape[1:]) != (3, v.image_size, v.image_size):
            raise ValueError(f"pixel_values must be [B, 3, {v.image_size}, {v.image_size}], got {tuple(pixel_values.shape)}")
        px = pixel_values.to(self.device, torch.float16).contiguous()
        ids = torch.as_tensor(input_ids)
        if ids.dim() == 1:
            ids = ids[None]
        if ids.dim() != 2 or ids.shape[1] > self.text_config.max_position_embeddings:
            raise ValueError(f"input_ids must be [T, <= {self.text_config.max_position_embeddings}]")
        ids = ids.to("cpu", torch.int32).contiguous()
        eos = ids.argmax(dim=-1).to(torch.int32).contiguous()
        b, t = px.shape[0], ids.shape[0]
        if b > self._prepared[0] or t > self._prepared[1]:
            self._prepared = (max(b, self._prepared[0]), max(t, self._prepared[1]))
            _lib.check(self._lib.dg_clipscore_prepare(self._h, *self._prepared), "dg_clipscore_prepare")
        out = torch.empty((t, b), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device)
        _lib.check(self._lib.dg_clipscore_score(self._h, C.c_void_p(px.data_ptr()), b, C.cast(ids.data_ptr(), C.POINTER(C.c_int32)),
                                                C.cast(eos.data_ptr(), C.POINTER(C.c_int32)), t, ids.shape[1],
                                                C.cast(out.data_ptr(), C.POINTER(C.c_float)), C.c_void_p(stream.cuda_stream)),
                   "dg_clipscore_score")
        stream.synchronize()                   # ids / eos are pageable host memory
        return out
