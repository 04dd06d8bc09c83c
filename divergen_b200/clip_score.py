"""CLIP scoring job: the reference's `DiverGen/filteration/get_clip_score.py`, re-hosted on the B200 path.

Keeps the reference's CLI (`--indir --outdir --use_mask --in_mask_dir --seg_name --dist --n_samples --max_batch_size
--in_lvis_json_path --clip_ckpt_dir --stages`, :42-54), its work split (image i of a category goes to rank i % world_size,
:113-115), its prompt (`'a photo of a single <name>'`, :175), its mask compositing (`mask > 128`, background value 1, area =
mask fraction, :133-146), its gather (all_gather of indices / scores / areas, sorted by image index, :183-203) and its output
(`<outdir>[/<stage>][/<seg_name>]/results.json`: the LVIS category list with `clip_scores` (and `areas`) per category,
:205-212).  Categories whose image count is not one of `--n_samples` are skipped with empty lists (:104-111).

What runs where: PNG decoding stays on the host (PIL); resize / crop / normalise (`clip_preprocess`: Pillow's bicubic, bit for
bit), the mask compositing (`dg_op_mask_composite_u8`) and both CLIP towers (`dg_clipscore_score`) run on the GPU.
`--clip_ckpt_dir` is a transformers-format `openai/clip-vit-large-patch14` folder (model.safetensors + tokenizer files); the
reference's `clip.load("ViT-L/14")` .pt archive holds the same weights under OpenAI's names.
"""
from __future__ import annotations

import argparse
import json
import os
from glob import glob
from typing import List, Optional, Sequence, Tuple


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="CLIP scores of generated instance images (filteration stage 1)")
    p.add_argument("--indir", type=str)
    p.add_argument("--outdir", type=str, nargs="?")
    p.add_argument("--use_mask", action="store_true", default=False)
    p.add_argument("--in_mask_dir", type=str)
    p.add_argument("--seg_name", type=str)
    p.add_argument("--dist", action="store_true", default=False)
    p.add_argument("--n_samples", nargs="+", type=int)
    p.add_argument("--max_batch_size", type=int, default=1)
    p.add_argument("--in_lvis_json_path", type=str,
                   default="output/220714_lvis_v1_t5_with_original_prompts_id/prompts/lvis_v1_id_to_prompt.json")
    p.add_argument("--clip_ckpt_dir", type=str, default=None)
    p.add_argument("--stages", type=str, nargs="+")
    return p


def stage_dirs(args, stage: str) -> Tuple[str, str]:
    """(input dir, output dir) of a stage -- reference :87-101."""
    if stage == "sd":
        cur_in = args.indir
        cur_out = os.path.join(args.outdir, args.seg_name) if args.use_mask else args.outdir
    else:
        cur_in = os.path.join(args.indir, stage)
        cur_out = os.path.join(args.outdir, stage, args.seg_name) if args.use_mask else os.path.join(args.outdir, stage)
    return cur_in, cur_out


def mask_path_for(args, stage: str, category_name: str, sample_path: str) -> str:
    """Reference :134-137."""
    if stage == "sd":
        return os.path.join(args.in_mask_dir, args.seg_name, category_name, os.path.basename(sample_path))
    return os.path.join(args.in_mask_dir, stage, args.seg_name, category_name, os.path.basename(sample_path))


def picked_for_rank(sample_paths: Sequence[str], rank: int, world: int) -> List[Tuple[int, str]]:
    return [(i, p) for i, p in enumerate(sample_paths) if i % world == rank]      # reference :113-115


def merge_gathered(indices: Sequence[Sequence[int]], values: Sequence[Sequence[float]]) -> List[float]:
    """Reference :192-200: concatenate every rank's (indices, values) and order the values by image index."""
    flat = [(i, v) for idx, val in zip(indices, values) for i, v in zip(idx, val)]
    return [v for _, v in sorted(flat, key=lambda t: t[0])]


def score_category(scorer, tokenizer, device, category_name: str, picked, max_batch_size: int, mask_paths=None):
    """Scores (and mask areas) of this rank's images of one category, in `picked` order."""
    import numpy as np
    import torch
    from PIL import Image
    from .generate import clip_prompt_text
    from .preprocess import clip_preprocess, mask_composite
    ids = tokenizer([clip_prompt_text(category_name)], padding="max_length", max_length=getattr(tokenizer, "model_max_length", 77),
                    truncation=True, return_tensors="pt").input_ids
    clips: List[float] = []
    areas: List[float] = []
    for b0 in range(0, len(picked), max(1, max_batch_size)):
        chunk = picked[b0:b0 + max(1, max_batch_size)]
        imgs = [np.asarray(Image.open(p).convert("RGB")) for _, p in chunk]
        # images of one category share a size (they come from one generation run); mixed sizes fall back to singles
        groups = [list(range(len(imgs)))] if len({im.shape for im in imgs}) == 1 else [[k] for k in range(len(imgs))]
        for g in groups:
            x = torch.from_numpy(np.stack([imgs[k] for k in g])).to(device)
            if mask_paths is not None:
                m = np.stack([np.asarray(Image.open(mask_paths[chunk[k][0]]).convert("L")) for k in g])
                x, a = mask_composite(x, torch.from_numpy(m).to(device))
                areas.extend(a.tolist())
            clips.extend(scorer(clip_preprocess(x, n_px=scorer.vision_config.image_size), ids).view(-1).cpu().tolist())
    return clips, areas


def main(argv: Optional[Sequence[str]] = None) -> int:
    import torch
    from .generate import init_distributed, load_clip_scorer
    args = build_parser().parse_args(argv)
    if args.dist:
        rank, local_rank, world, device = init_distributed()
    else:
        rank, local_rank, world, device = 0, 0, 1, torch.device("cuda:0")
    print("local rank: {}, global rank: {}, world size: {}, device: {}".format(local_rank, rank, world, device))
    torch.cuda.set_device(device)
    if not args.clip_ckpt_dir:
        raise FileNotFoundError("--clip_ckpt_dir is required: there is no network to download ViT-L/14 from")
    print(">>> Loading CLIP...")
    loaded = load_clip_scorer(args.clip_ckpt_dir, device)
    if loaded is None:
        raise FileNotFoundError("{}: no model(.fp16).safetensors".format(args.clip_ckpt_dir))
    tokenizer, scorer = loaded
    with open(args.in_lvis_json_path, "r") as f:
        data = json.load(f)
    for stage in args.stages:
        cur_in, cur_out = stage_dirs(args, stage)
        for category in data:
            name = category["name"]
            sample_paths = sorted(glob(os.path.join(cur_in, name, "*.png")))
            if len(sample_paths) not in args.n_samples:
                print(">>>Skip {}, it has {} images, but expected to have {}".format(name, len(sample_paths), args.n_samples))
                category["clip_scores"] = []
                if args.use_mask:
                    category["areas"] = []
                continue
            picked = picked_for_rank(sample_paths, rank, world)
            masks = {i: mask_path_for(args, stage, name, p) for i, p in picked} if args.use_mask else None
            print(">>> Processing {}...".format(name))
            clips, areas = score_category(scorer, tokenizer, device, name, picked, args.max_batch_size, masks)
            indices = [i for i, _ in picked]
            if world > 1:
                import torch.distributed as dist
                parts = [None] * world
                dist.all_gather_object(parts, (indices, clips, areas))
                clips = merge_gathered([p[0] for p in parts], [p[1] for p in parts])
                if args.use_mask:
                    areas = merge_gathered([p[0] for p in parts], [p[2] for p in parts])
            category["clip_scores"] = clips
            if args.use_mask:
                category["areas"] = areas
        if rank == 0:
            os.makedirs(cur_out, exist_ok=True)
            with open(os.path.join(cur_out, "results.json"), "w") as f:
                json.dump(data, f)
    if args.dist:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
