"""Row f3 remainder: the filteration CLIP-score job (`divergen_b200.clip_score`, the reference's
filteration/get_clip_score.py) end to end on PNG folders, with and without mask compositing, against the reference's own
arithmetic restated with PIL / numpy / transformers.CLIPModel; the device mask compositing alone, bit for bit; and the
generation driver's packed calls against unpacked ones."""
import json
import os

import numpy as np
import pytest
import torch

from tests.test_gpu_clip import TINY, TINY_VISION, _clip_models
from tests.test_gpu_unet import DEV, _need_gpu

pytestmark = pytest.mark.gpu


def test_mask_composite_matches_the_reference_formula():
    """get_clip_score.py:139-146: mask_im = mask > 128; image*mask_im + ones_like(image)*(1-mask_im); area = sum/H/W."""
    _need_gpu()
    from divergen_b200.preprocess import mask_composite
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (3, 37, 53, 3), dtype=np.uint8)
    mask = rng.integers(0, 256, (3, 37, 53), dtype=np.uint8)
    mask[1] = 0
    mask[2, :5] = 129
    mask[2, 5:7] = 128
    got, areas = mask_composite(torch.from_numpy(img).to(DEV), torch.from_numpy(mask).to(DEV))
    for b in range(3):
        m = np.expand_dims(mask[b], axis=2) > 128
        want = (img[b] * m + np.ones_like(img[b]) * (1 - m)).astype(np.uint8)
        assert np.array_equal(got[b].cpu().numpy(), want)
        assert areas[b].item() == np.sum(m) / m.shape[0] / m.shape[1]


def _write_clip_dir(d, ref):
    """A transformers-format CLIP folder: config.json + model.safetensors + tokenizer files."""
    from safetensors.torch import save_file
    from tests.hf_layout import write_tokenizer
    os.makedirs(d, exist_ok=True)
    sd = {k: v.contiguous() for k, v in ref.state_dict().items() if not k.endswith("position_ids")}
    save_file(sd, os.path.join(d, "model.safetensors"))
    json.dump({"text_config": dict(TINY), "vision_config": dict(TINY_VISION), "projection_dim": 64}, open(os.path.join(d, "config.json"), "w"))
    write_tokenizer(d)


@pytest.mark.parametrize("use_mask", [False, True])
def test_clip_score_job_against_the_reference_flow(tmp_path, use_mask):
    _need_gpu()
    from PIL import Image
    from transformers import CLIPTokenizer
    from divergen_b200 import clip_score
    from divergen_b200.generate import clip_prompt_text
    from divergen_b200.preprocess import CLIP_MEAN, CLIP_STD, resize_size
    tiny_text = dict(TINY, vocab_size=514)                 # the byte-level test vocabulary
    ref, _ = _clip_models(0, tiny_text, TINY_VISION, 64)
    clip_dir = tmp_path / "clip"
    saved_cfg = {"text_config": tiny_text, "vision_config": dict(TINY_VISION), "projection_dim": 64}
    _write_clip_dir(str(clip_dir), ref)
    json.dump(saved_cfg, open(clip_dir / "config.json", "w"))
    cats = [{"id": 1, "name": "aerosol_can"}, {"id": 7, "name": "alligator"}, {"id": 9, "name": "antenna"}]
    (tmp_path / "cats.json").write_text(json.dumps(cats))
    rng = np.random.default_rng(1)
    n_img = {"aerosol_can": 3, "alligator": 3, "antenna": 2}      # antenna has the wrong count -> skipped with empty lists
    for name, n in n_img.items():
        os.makedirs(tmp_path / "in" / name)
        os.makedirs(tmp_path / "masks" / "sam" / name)
        for k in range(n):
            Image.fromarray(rng.integers(0, 256, (96, 80, 3), dtype=np.uint8)).save(tmp_path / "in" / name / f"{k}_000000{k}.png")
            Image.fromarray((rng.integers(0, 2, (96, 80), dtype=np.uint8) * 255)).save(tmp_path / "masks" / "sam" / name / f"{k}_000000{k}.png")
    argv = ["--indir", str(tmp_path / "in"), "--outdir", str(tmp_path / "out"), "--n_samples", "3", "--max_batch_size", "2",
            "--in_lvis_json_path", str(tmp_path / "cats.json"), "--clip_ckpt_dir", str(clip_dir), "--stages", "sd"]
    if use_mask:
        argv += ["--use_mask", "--in_mask_dir", str(tmp_path / "masks"), "--seg_name", "sam"]
    assert clip_score.main(argv) == 0
    out_dir = tmp_path / "out" / "sam" if use_mask else tmp_path / "out"
    data = json.load(open(out_dir / "results.json"))
    assert [c["name"] for c in data] == [c["name"] for c in cats]
    assert data[2]["clip_scores"] == [] and (not use_mask or data[2]["areas"] == [])

    # the reference flow, restated: PIL open -> (mask) -> Resize/CenterCrop/ToTensor/Normalize -> CLIPModel fp32
    tok = CLIPTokenizer.from_pretrained(str(clip_dir))
    mean, std = torch.tensor(CLIP_MEAN).view(3, 1, 1), torch.tensor(CLIP_STD).view(3, 1, 1)
    n_px = TINY_VISION["image_size"]
    for cat in data[:2]:
        name = cat["name"]
        paths = sorted((tmp_path / "in" / name).glob("*.png"))
        px, areas = [], []
        for p in paths:
            image = Image.open(p).convert("RGB")
            if use_mask:
                mask = np.expand_dims(np.array(Image.open(tmp_path / "masks" / "sam" / name / p.name).convert("L")), axis=2)
                mask_im = mask > 128
                image = Image.fromarray((image * mask_im + np.ones_like(image) * (1 - mask_im)).astype(np.uint8))
                areas.append(np.sum(mask_im) / mask_im.shape[0] / mask_im.shape[1])
            oh, ow = resize_size(image.height, image.width, n_px)
            image = image.resize((ow, oh), Image.BICUBIC)
            top, left = int(round((oh - n_px) / 2.0)), int(round((ow - n_px) / 2.0))
            t = torch.from_numpy(np.asarray(image)[top:top + n_px, left:left + n_px].copy()).permute(2, 0, 1).float() / 255.0
            px.append((t - mean) / std)
        ids = tok([clip_prompt_text(name)], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
        with torch.no_grad():
            want = ref(input_ids=ids, pixel_values=torch.stack(px)).logits_per_text.view(-1)
        got = torch.tensor(cat["clip_scores"])
        assert got.shape == want.shape
        err = (got - want).abs().max().item()
        print(f"clip score job ({'masked' if use_mask else 'plain'}) {name}: max err {err:.4g}")
        assert err <= 1e-2 * want.abs().max().item() + 2e-2
        if use_mask:
            assert cat["areas"] == areas


def test_packed_driver_writes_the_same_files_as_unpacked(tmp_path):
    """Cross-prompt packing (gpt-prompt recipe: 1 image per prompt per rank) leaves the same files with the same content, up
    to the batch-size dependence of the kernels' tile / split-K choices."""
    _need_gpu()
    from divergen_b200.generate import main
    prompts = tmp_path / "prompts"
    prompts.mkdir()
    (prompts / "3.txt").write_text("".join("a photo of a single thing number {}\n".format(k) for k in range(5)))
    common = ["--from_file", str(prompts), "--n_samples", "1", "--random_init", "--num_inference_steps", "2", "--offset", "0"]
    assert main(common + ["--outdir", str(tmp_path / "packed"), "--max_batch_size", "4", "--stats_json", str(tmp_path / "s.json")]) == 0
    assert main(common + ["--outdir", str(tmp_path / "plain"), "--max_batch_size", "4", "--no_pack"]) == 0
    a = sorted(os.listdir(tmp_path / "packed" / "samples" / "sd"))
    b = sorted(os.listdir(tmp_path / "plain" / "samples" / "sd"))
    assert a == b == ["3_{:07d}.latent.pt".format(k) for k in range(5)]
    for f in a:
        x = torch.load(tmp_path / "packed" / "samples" / "sd" / f).float()
        y = torch.load(tmp_path / "plain" / "samples" / "sd" / f).float()
        assert (x - y).abs().max().item() <= 2e-2 * y.abs().max().item() + 1e-3
    st = json.load(open(tmp_path / "s.json"))
    assert st["images"] == 5 and st["packed"] and st["images_per_s"] > 0
