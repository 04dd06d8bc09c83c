"""Generation driver: the reference's `DiverGen/generation/txt2img_diffusers_stages_from_txt.py`, re-hosted on the B200 path.

Keeps what `convert_dir_structure.py`, `segmentation/` and `filteration/` depend on:
  * the CLI flags (`--from_file --outdir --n_samples --max_batch_size --seed --dist --ckpt_dir --stages --offset
    --disable_overwrite`, reference :27-110) plus model / sampler flags the reference hard-codes inside diffusers;
  * rank sharding: every rank makes `n_samples // world_size` images of every prompt (:124-125), file indices offset by
    `total_batch_size * rank` (:262), seed `seed + rank` (:200), skip-if-exists resume (:245-252);
  * file naming `<outdir>/samples/<stage>/<category_id>_<count:07d>.png` (:262-266).
Differences, all deliberate: the stage is Stable Diffusion (`sd`, the pool the downstream stages call 'sd') instead of the
DeepFloyd-IF cascade; text embeddings are encoded ONCE on rank 0 for the whole prompt set and broadcast (the path's only
data-carrying collective) instead of per micro-batch on every rank; micro-batches default to 4 images per pipeline call.

Also deliberate: pipeline calls are PACKED across prompts up to `--max_batch_size` images (`pack_calls`): the gpt-prompt
recipe (`--n_samples 8` on 8 ranks, DiverGen/DATA.md:24-39) is one image per prompt per rank, i.e. UNet batch 2 when every
prompt is its own call; packed it runs at UNet batch 2 * max_batch_size.  Latents are still drawn per reference call from the
rank's generator, so which seed makes which file does not change.  The resume test (`--disable_overwrite`) looks at the
file the call's LAST image is written to (the save index, :262).  The reference's own skip test (:246) computes
`i * current_num_images_per_prompt` where the save index uses the running `tmp`; the two agree only for a file's first
prompt without a remainder batch, so the save index -- the on-disk contract -- is what resume checks here.

Host-side only: every tensor operation is a C-ABI call (StableDiffusionPipeline -> dg_denoise_loop, AutoencoderKL ->
dg_vae_decode, CLIPTextModel -> dg_clip_encode).  The checkpoint is a Hugging Face pipeline folder read by
`DiffusionPipeline.from_pretrained(ckpt_dir, variant='fp16', torch_dtype=torch.float16)` as in the reference (:139); with a
VAE (or `--random_init --decode`) PNGs are written exactly where the reference writes them; without one, latents go to
`<cid>_<count:07d>.latent.pt` under the same index contract.  Synthetic text embeddings exist for `--random_init` runs only:
a real checkpoint without its tokenizer / text encoder is an error.
"""
from __future__ import annotations

import argparse
import os
from dataclasses import dataclass
from glob import glob
from typing import Iterator, List, Optional, Sequence, Tuple


# ------------------------------------------------------------------ pure index arithmetic (no torch, unit-tested on CPU)
@dataclass(frozen=True)
class BatchPlan:
    """Reference :124-131.  `batch_size` is the number of pipeline calls per prompt, the first of which carries the
    remainder (`remainder_batch_size` images) when `total_batch_size` is not a multiple of `max_batch_size`."""
    total_batch_size: int
    batch_size: int
    remainder_batch_size: int
    max_batch_size: int


def plan_batches(n_samples: int, world_size: int, max_batch_size: int) -> BatchPlan:
    total = n_samples // world_size
    if total * world_size != n_samples:
        raise ValueError("n_samples must be divisible by world_size")          # reference :125
    if max_batch_size <= 0:
        raise ValueError("max_batch_size must be positive")
    calls, rem = divmod(total, max_batch_size)
    if rem > 0:
        calls += 1
    return BatchPlan(total, calls, rem, max_batch_size)


@dataclass(frozen=True)
class Call:
    """One pipeline call: `prompt` -> `num_images` images written at file indices `counts`."""
    index: int            # position in the sorted, repeated prompt list (reference `i`)
    prompt: str
    num_images: int
    counts: Tuple[int, ...]


def iter_calls(prompt_lines: Sequence[str], plan: BatchPlan, rank: int, n_samples: int, offset: int) -> Iterator[Call]:
    """The reference's per-category loop (:221-327) as a generator of (prompt, images, file indices).

    data = sorted(batch_size * lines); call i belongs to prompt number i // batch_size; the first call of a prompt takes the
    remainder; count = j + tmp + total_batch_size*rank + offset + (i // batch_size) * n_samples (:262)."""
    data = sorted(plan.batch_size * [l for l in prompt_lines])
    tmp = 0
    for i, prompt in enumerate(data):
        if i % plan.batch_size == 0:
            tmp = 0
            cur = plan.remainder_batch_size if plan.remainder_batch_size != 0 else plan.max_batch_size
        else:
            cur = plan.max_batch_size
        base = tmp + plan.total_batch_size * rank + offset + (i // plan.batch_size) * n_samples
        yield Call(i, prompt.strip(), cur, tuple(base + j for j in range(cur)))
        tmp += cur


def pack_calls(calls: Sequence[Call], max_images: int) -> Iterator[List[Call]]:
    """Group consecutive pipeline calls (of one category file) so that one UNet pass serves up to `max_images` images, whatever
    prompt each belongs to.  A call is never split; order is kept, so file indices and the order in which latents are drawn
    from the rank's generator are those of the unpacked loop."""
    group: List[Call] = []
    n = 0
    for c in calls:
        if group and n + c.num_images > max_images:
            yield group
            group, n = [], 0
        group.append(c)
        n += c.num_images
    if group:
        yield group


def output_name(category_id: str, count: int, ext: str = "png") -> str:
    return "{}_{:07d}.{}".format(category_id, count, ext)                     # reference :263


def load_category_names(lvis_json_path: str) -> dict:
    """`id_to_name = {d['id']: d['name'] for d in data}` of generation/convert_dir_structure.py:88-91 (a JSON list of
    LVIS category records)."""
    import json
    with open(lvis_json_path, "r") as f:
        data = json.load(f)
    return {int(d["id"]): d["name"] for d in data}


def output_dir_for(outdir: str, stage: str, category_id: str, id_to_name: Optional[dict]) -> str:
    """Where a category's files go.  Without category names: `<outdir>/samples/<stage>/` (reference generation script
    :144-146,262-266).  With them: `<outdir>/<stage>/<category_name>/` -- the layout generation/convert_dir_structure.py
    :113-121 produces by COPYING every PNG (file names unchanged), which segmentation/ and filteration/ read; writing it
    directly removes that pass over 1.2 M files."""
    if id_to_name is None:
        return os.path.join(outdir, "samples", stage)
    try:
        name = id_to_name[int(category_id)]
    except (KeyError, ValueError):
        raise ValueError("category id '{}' (from the prompt file name) is not in the LVIS category JSON".format(category_id))
    return os.path.join(outdir, stage, name)


def list_prompt_files(from_file: Sequence[str]) -> List[str]:
    """Reference :207-209: a directory expands to its *.txt files.  (The reference discards the result of `sorted`, so its
    category order is glob order; sorting here only fixes the order, not the set of files or any file name.)"""
    if len(from_file) == 1 and os.path.isdir(from_file[0]):
        return sorted(glob(os.path.join(from_file[0], "*.txt")))
    return list(from_file)


def category_id_of(path: str) -> str:
    return os.path.basename(path).split(".")[0]                               # reference :219


# ------------------------------------------------------------------ distributed plumbing
def init_distributed(backend: str = "nccl"):
    """Reference :13-24 (env-var rendezvous, one process per GPU)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world_size = int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if not dist.is_initialized():
        from datetime import timedelta
        # reference :20 -- ranks only meet at the first barrier / broadcast and at the last barrier; with resume or uneven
        # work one of them may wait there for hours
        dist.init_process_group(backend=backend, timeout=timedelta(seconds=720000))
    device = torch.device("cuda:{}".format(local_rank)) if backend == "nccl" else torch.device("cpu")
    return rank, local_rank, world_size, device


def broadcast_embedding_table(table, src: int = 0):
    """The path's single data-carrying collective: rank `src` holds the [P, 77, D] text-embedding table
    (+ the unconditional row), every other rank passes a same-shaped buffer or None and receives it."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return table
    rank = dist.get_rank()
    dev = table.device if table is not None else torch.device("cpu")
    meta = torch.zeros(4, dtype=torch.int64, device=dev)
    if rank == src:
        meta[:3] = torch.tensor(table.shape, dtype=torch.int64)
        meta[3] = {torch.float16: 0, torch.float32: 1, torch.bfloat16: 2}[table.dtype]
    dist.broadcast(meta, src=src)
    if rank != src:
        dtype = {0: torch.float16, 1: torch.float32, 2: torch.bfloat16}[int(meta[3])]
        table = torch.empty(tuple(int(v) for v in meta[:3]), dtype=dtype, device=dev)
    dist.broadcast(table, src=src)
    return table


def synthetic_text_embeddings(prompts: Sequence[str], dim: int, tokens: int = 77, seed: int = 0):
    """Stand-in for CLIPTextModel when no text encoder weights are present (there are none offline): a seeded N(0,1)
    embedding per distinct prompt string.  Row 0 is the unconditional ("") embedding."""
    import hashlib
    import torch
    rows = []
    for p in [""] + list(prompts):
        h = int.from_bytes(hashlib.sha256(p.encode()).digest()[:8], "little") ^ seed
        g = torch.Generator().manual_seed(h % (2 ** 63))
        rows.append(torch.randn(tokens, dim, generator=g))
    return torch.stack(rows).half()


class AsyncImageWriter:
    """Output path: uint8 images leave the device through pinned memory on a side copy, and PNG
    encoding (zlib releases the GIL) runs on worker threads while the GPU denoises the next micro-batch.  File names and
    PNG content are what `pt_to_pil(image)[j].save(out_path)` (reference :267) produces."""

    def __init__(self, workers: int = 4, max_in_flight: int = 0):
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self._pending = []
        self._written = 0
        # back-pressure: at most this many batches (pinned buffers) queued or being encoded; the GPU loop blocks beyond it
        self._max_in_flight = max_in_flight if max_in_flight > 0 else 4 * max(1, workers)

    def _reap(self, block_until: int):
        """Collect finished batches -- a failed save (disk full, bad path) raises HERE, at the next submit, not after the
        whole run -- and wait until at most `block_until` batches are outstanding."""
        keep = []
        for f in self._pending:
            if f.done():
                self._written += f.result()
            else:
                keep.append(f)
        self._pending = keep
        while len(self._pending) > block_until:
            self._written += self._pending.pop(0).result()

    def submit(self, images_u8, paths: Sequence[str]):
        import torch
        self._reap(self._max_in_flight - 1)
        host = torch.empty(images_u8.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(images_u8, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        self._pending.append(self._pool.submit(self._save, host, done, list(paths)))

    @staticmethod
    def _save(host, done, paths):
        from PIL import Image
        done.synchronize()
        arr = host.numpy()
        for j, path in enumerate(paths):
            Image.fromarray(arr[j]).save(path)
        return len(paths)

    def close(self) -> int:
        try:
            self._reap(0)
        finally:
            self._pool.shutdown(wait=True)
        return self._written


def clip_prompt_text(category_name: str) -> str:
    """filteration/get_clip_score.py:175."""
    return "a photo of a single {}".format(" ".join(category_name.split("_")))


def clip_scores_for(scorer, tokenizer, images_u8, category_name: str):
    """Row f3 fused into generation: the CLIP score filteration/get_clip_score.py:174-180 would compute after re-reading the
    PNG -- `text = 'a photo of a single {}'.format(' '.join(category_name.split('_')))`, `logits_per_text` of ViT-L/14 --
    taken from the uint8 image while it is still on the device (PNG is lossless and the resize is Pillow's, bit for bit, so
    the pixels are the same).  Returns a list of floats, one per image."""
    from .preprocess import clip_preprocess
    text = clip_prompt_text(category_name)
    ids = tokenizer([text], padding="max_length", max_length=getattr(tokenizer, "model_max_length", 77), truncation=True,
                    return_tensors="pt").input_ids
    return scorer(clip_preprocess(images_u8, n_px=scorer.vision_config.image_size), ids).view(-1).cpu().tolist()


def load_clip_scorer(clip_dir: str, device):
    """(tokenizer, scorer) from a transformers-format CLIP directory (`openai/clip-vit-large-patch14`: model.safetensors +
    tokenizer files; `config.json`, when present, overrides the ViT-L/14 defaults), or None when no weights are there."""
    path = _first_existing(clip_dir, ("model.fp16.safetensors", "model.safetensors")) if clip_dir else None
    if not path:
        return None
    import json
    from safetensors.torch import load_file
    from transformers import CLIPTokenizer
    from . import CLIPScorer
    kw = {}
    cfg_path = os.path.join(clip_dir, "config.json")
    if os.path.exists(cfg_path):
        cfg = json.load(open(cfg_path))
        tk = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "max_position_embeddings")
        vk = ("image_size", "patch_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads")
        t, v = cfg.get("text_config") or {}, cfg.get("vision_config") or {}
        kw = dict(text_config={k: t[k] for k in tk if k in t}, vision_config={k: v[k] for k in vk if k in v})
        if "projection_dim" in cfg:
            kw["projection_dim"] = cfg["projection_dim"]
    sc = CLIPScorer(device=device, **kw)
    sc.load_state_dict(load_file(path))
    return CLIPTokenizer.from_pretrained(clip_dir), sc


def encode_prompt_table(prompts: Sequence[str], tokenizer, text_encoder, batch: int = 64):
    """Text embeddings of every distinct prompt, once, on rank 0 (row 0 = the unconditional "" prompt): what the reference
    recomputes per micro-batch on every rank through `stage_1.encode_prompt(prompt)` (:242).  `tokenizer` is a transformers
    `CLIPTokenizer`-like callable, `text_encoder` a `CLIPTextModel` (divergen_b200.clip or any `.forward(ids)[0]`)."""
    import torch
    texts = [""] + list(prompts)
    rows = []
    for i in range(0, len(texts), batch):
        ids = tokenizer(texts[i:i + batch], padding="max_length", max_length=getattr(tokenizer, "model_max_length", 77),
                        truncation=True, return_tensors="pt").input_ids
        rows.append(text_encoder(ids)[0].to(torch.float16))
    return torch.cat(rows, dim=0)


def load_text_encoder(ckpt_dir: str, device):
    """(tokenizer, text_encoder) from `<ckpt_dir>/tokenizer` + `<ckpt_dir>/text_encoder/{config.json, model*.safetensors}`
    (the configuration decides between the SD-1.x and SD-2.x towers), or None when either folder is missing -- the caller
    decides whether that is an error (it is, unless --random_init)."""
    tok_dir = os.path.join(ckpt_dir, "tokenizer")
    te_dir = os.path.join(ckpt_dir, "text_encoder")
    te_path = _first_existing(te_dir, ("model.fp16.safetensors", "model.safetensors"))
    if not (os.path.isdir(tok_dir) and te_path):
        return None
    import json
    from safetensors.torch import load_file
    from transformers import CLIPTokenizer
    from . import CLIPTextModel
    from .loading import text_encoder_kwargs_from_config
    cfg_path = os.path.join(te_dir, "config.json")
    kw = text_encoder_kwargs_from_config(json.load(open(cfg_path))) if os.path.exists(cfg_path) else {}
    enc = CLIPTextModel(device=device, **kw)
    enc.load_state_dict(load_file(te_path))
    return CLIPTokenizer.from_pretrained(tok_dir), enc


# ------------------------------------------------------------------ CLI
def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="DiverGen generation driver on the B200-native Stable-Diffusion path")
    p.add_argument("--prompt", type=str, nargs="?", default="a painting of a virus monster playing guitar")
    p.add_argument("--from_file", type=str, action="append", help="prompt file(s) or a directory of <category_id>.txt")
    p.add_argument("--outdir", type=str, nargs="?", default="output/txt2img-samples")
    p.add_argument("--n_samples", type=int, default=1, help="images per prompt over ALL ranks")
    p.add_argument("--max_batch_size", type=int, default=4, help="images per pipeline call")
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--dist", action="store_true", default=False, help="one process per GPU (torchrun / torch.distributed.launch)")
    p.add_argument("--ckpt_dir", type=str, default="models/ldm/stable-diffusion-v1/")
    p.add_argument("--dataset_json_path", type=str, default="/data/datasets/lvis/lvis_v1_val.json")   # parsed, unused (as upstream)
    p.add_argument("--stages", type=str, nargs="+", default=["sd"])
    p.add_argument("--offset", type=int, default=1024)
    p.add_argument("--disable_overwrite", action="store_true", default=False)
    # additions (fixed inside diffusers / the checkpoint in the reference)
    p.add_argument("--model", choices=["sd15", "sd21"], default="sd15")
    p.add_argument("--num_inference_steps", type=int, default=50)
    p.add_argument("--guidance_scale", type=float, default=7.5)
    p.add_argument("--random_init", action="store_true", help="random-init UNet weights (benchmarks; no checkpoint offline)")
    p.add_argument("--png_workers", type=int, default=4, help="PNG encoder threads per rank")
    p.add_argument("--clip_dir", type=str, default=None,
                   help="transformers-format CLIP ViT-L/14 directory: score every image on the device (needs --in_lvis_json_path "
                        "for the category names) and write <outdir>/clip_scores_<stage>_rank<r>.json")
    p.add_argument("--in_lvis_json_path", type=str, default=None,
                   help="LVIS category JSON (as for convert_dir_structure.py): write <outdir>/<stage>/<category_name>/ directly")
    p.add_argument("--decode", action="store_true", help="with --random_init: also build a random-init VAE and write PNGs")
    p.add_argument("--max_prompt_files", type=int, default=0, help="process only the first N category files (0 = all)")
    p.add_argument("--no_pack", action="store_true", help="one pipeline call per reference call (no packing across prompts)")
    p.add_argument("--stats_json", type=str, default=None, help="write loop timing (images/s, embedding broadcast) here")
    return p


def random_state_dict(model, device, seed: int = 0):
    """Random-init weights with the checkpoint's shapes and sane scales (no checkpoint is reachable offline)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for k, shp in model.expected_state_dict_shapes().items():
        fan_in = 1
        for d in shp[1:]:
            fan_in *= d
        if "norm" in k and k.endswith("weight"):
            sd[k] = torch.ones(shp, device=device).half()
        elif k.endswith("bias"):
            sd[k] = torch.zeros(shp, device=device).half()
        else:
            sd[k] = (torch.randn(shp, generator=g, device=device) / max(1.0, fan_in) ** 0.5).half()
    return sd


def _first_existing(directory: str, names: Sequence[str]) -> Optional[str]:
    for n in names:
        if os.path.exists(os.path.join(directory, n)):
            return os.path.join(directory, n)
    return None


def _build_pipeline(args, device, local_rank):
    """The reference's `DiffusionPipeline.from_pretrained(...)`, `.to(device)`, `.enable_model_cpu_offload(local_rank)`
    (:139-143) -- or, for `--random_init` benchmark runs (no checkpoint is reachable offline), seeded random weights."""
    import torch
    from . import (AutoencoderKL, DDIMScheduler, DiffusionPipeline, SD15_CONFIG, SD21_CONFIG, StableDiffusionPipeline,
                   UNet2DConditionModel)
    if args.random_init:
        cfg = SD15_CONFIG if args.model == "sd15" else SD21_CONFIG
        unet = UNet2DConditionModel(device=device, **{k: v for k, v in cfg.items() if k != "time_cond_proj_dim"})
        unet.load_state_dict(random_state_dict(unet, device))
        vae = None
        if args.decode:
            vae = AutoencoderKL(device=device)
            vae.load_state_dict(random_state_dict(vae, device, seed=1))
        sched = DDIMScheduler(prediction_type="v_prediction" if args.model == "sd21" else "epsilon")
        pipe = StableDiffusionPipeline(unet, sched, vae=vae)
    else:
        if not os.path.exists(os.path.join(args.ckpt_dir, "model_index.json")):
            raise FileNotFoundError("{} is not a Hugging Face pipeline folder (no model_index.json); use --random_init for "
                                    "benchmark runs without a checkpoint".format(args.ckpt_dir))
        print("==> Loading stage {} from {}...".format(args.stages[0], args.ckpt_dir))
        pipe = DiffusionPipeline.from_pretrained(args.ckpt_dir, variant="fp16", torch_dtype=torch.float16)
        pipe.to(device)
    pipe.enable_model_cpu_offload(local_rank)                                    # no-ops kept for call parity (:143,186)
    pipe.enable_xformers_memory_efficient_attention()
    return pipe


def gather_clip_results(scores: dict, id_to_name: dict, lvis_json_path: str, out_path: str, rank: int, world: int):
    """What filteration/get_clip_score.py:183-212 leaves behind: the LVIS category list with `clip_scores` per category, in
    the order of the category directory's sorted PNG names (its `sorted(glob('*.png'))` index), gathered from every rank and
    written once by rank 0 as `results.json`.  `scores` = {category_id: {file name: score}} of this rank."""
    import json
    merged = [scores]
    if world > 1:
        import torch.distributed as dist
        merged = [None] * world
        dist.all_gather_object(merged, scores)
    if rank != 0:
        return None
    per_cat = {}
    for part in merged:
        for cid, d in part.items():
            per_cat.setdefault(int(cid), {}).update(d)
    with open(lvis_json_path, "r") as f:
        data = json.load(f)
    for category in data:
        d = per_cat.get(int(category["id"]), {})
        category["clip_scores"] = [d[name] for name in sorted(d)]
    with open(out_path, "w") as f:
        json.dump(data, f)
    return data


def main(argv: Optional[Sequence[str]] = None) -> int:
    import json
    import time
    import torch

    args = build_parser().parse_args(argv)
    if args.dist:
        rank, local_rank, world, device = init_distributed()
    else:
        rank, local_rank, world, device = 0, 0, 1, torch.device("cuda:0")
    print("local rank: {}, global rank: {}, world size: {}, device: {}".format(local_rank, rank, world, device))
    torch.cuda.set_device(device)
    plan = plan_batches(args.n_samples, world, args.max_batch_size)
    stage = args.stages[0]
    id_to_name = load_category_names(args.in_lvis_json_path) if args.in_lvis_json_path else None
    sample_dir = os.path.join(args.outdir, "samples", stage) if id_to_name is None else os.path.join(args.outdir, stage)
    if rank == 0:
        os.makedirs(sample_dir, exist_ok=True)
    if args.dist:
        torch.distributed.barrier()                                               # reference :152-153

    pipe = _build_pipeline(args, device, local_rank)
    vae = pipe.vae
    if args.clip_dir and (vae is None or id_to_name is None):
        raise ValueError("--clip_dir scores decoded images by category name: it needs a VAE (a checkpoint with vae/, or "
                         "--random_init --decode) and --in_lvis_json_path")
    ext = "png" if vae is not None else "latent.pt"
    dim = pipe.unet.config.cross_attention_dim
    lat_hw = pipe.unet.config.sample_size

    generator = torch.manual_seed(args.seed + rank)                             # reference :200 (global CPU generator)
    files = list_prompt_files(args.from_file) if args.from_file else []
    if args.max_prompt_files > 0:
        files = files[:args.max_prompt_files]
    per_file = [(f, open(f).read().splitlines()) for f in files] if files else [("prompt.txt", [args.prompt])]
    # ---- text embeddings: once, on rank 0, for every distinct prompt; one broadcast
    prompts = sorted({l.strip() for _, lines in per_file for l in lines})
    t_emb0 = time.perf_counter()
    table = None
    if rank == 0:
        if pipe.text_encoder is not None and pipe.tokenizer is not None:
            table = encode_prompt_table(prompts, pipe.tokenizer, pipe.text_encoder).to(device)
        elif args.random_init:
            table = synthetic_text_embeddings(prompts, dim).to(device)
        else:
            raise FileNotFoundError("{} has no tokenizer/ + text_encoder/: refusing to generate from synthetic text embeddings "
                                    "with real UNet weights (that is what --random_init is for)".format(args.ckpt_dir))
    table = broadcast_embedding_table(table if table is not None else torch.empty(0, device=device))
    torch.cuda.synchronize()
    t_emb = time.perf_counter() - t_emb0
    row = {p: i + 1 for i, p in enumerate(prompts)}

    n_done = 0
    writer = AsyncImageWriter(args.png_workers) if vae is not None else None
    clip = load_clip_scorer(args.clip_dir, device) if args.clip_dir else None
    if args.clip_dir and clip is None:
        raise FileNotFoundError("--clip_dir {}: no model(.fp16).safetensors found".format(args.clip_dir))
    scores_path = os.path.join(args.outdir, "clip_scores_{}_rank{}.json".format(stage, rank))
    scores = {}
    if clip is not None and args.disable_overwrite and os.path.exists(scores_path):
        with open(scores_path, "r") as f:                                         # resume: keep the scores of the calls we skip
            scores = json.load(f)
    max_pack = args.max_batch_size if not args.no_pack else 0
    t_loop0 = time.perf_counter()
    for fi, (path, lines) in enumerate(per_file):
        cid = category_id_of(path)
        cat_dir = output_dir_for(args.outdir, stage, cid, id_to_name)
        if id_to_name is not None:
            os.makedirs(cat_dir, exist_ok=True)                                   # convert_dir_structure.py:94-99 (every rank: idempotent)
        print("==> Reading prompts from {}, {}/{}".format(path, fi + 1, len(per_file)))
        todo = []
        for call in iter_calls(lines, plan, rank, args.n_samples, args.offset):
            last = os.path.join(cat_dir, output_name(cid, call.counts[-1], ext))
            if args.disable_overwrite and os.path.exists(last):
                print("==> Skipping stage {} for {}...".format(stage, os.path.basename(last)))
                continue
            todo.append(call)
        groups = pack_calls(todo, max_pack) if max_pack > 0 else ([c] for c in todo)
        for group in groups:
            # one UNet pass for the whole group: one embedding row and one latent per image; the latents are drawn call by
            # call from the rank's generator exactly as the reference's per-call `pipe(..., generator=generator)` would
            pos = torch.cat([table[row[c.prompt]][None].expand(c.num_images, -1, -1) for c in group])
            neg = table[0][None].expand(pos.shape[0], -1, -1)
            lat = torch.cat([torch.randn((c.num_images, 4, lat_hw, lat_hw), generator=generator, dtype=torch.float16) for c in group])
            if lat.device.type == "cpu":
                lat = lat.pin_memory()          # the pipeline copies pinned inputs without blocking: the host stays one call ahead of the GPU
            out = pipe(prompt_embeds=pos, negative_prompt_embeds=neg, latents=lat,
                       output_type="uint8" if vae is not None else "latent",
                       num_images_per_prompt=1, num_inference_steps=args.num_inference_steps,
                       guidance_scale=args.guidance_scale).images
            dsts = [os.path.join(cat_dir, output_name(cid, count, ext)) for c in group for count in c.counts]
            if vae is not None:
                writer.submit(out, dsts)                                          # reference :267 (pt -> PIL -> .save), asynchronously
                if clip is not None:
                    for dst, sc in zip(dsts, clip_scores_for(clip[1], clip[0], out, id_to_name[int(cid)])):
                        scores.setdefault(cid, {})[os.path.basename(dst)] = sc
            else:
                for j, dst in enumerate(dsts):
                    torch.save(out[j].cpu(), dst)
            n_done += len(dsts)
    if writer is not None:
        writer.close()
    torch.cuda.synchronize()
    t_loop = time.perf_counter() - t_loop0
    if clip is not None:
        with open(scores_path, "w") as f:
            json.dump(scores, f)
        gather_clip_results(scores, id_to_name, args.in_lvis_json_path, os.path.join(sample_dir, "results.json"), rank, world)
    stats = {"rank": rank, "world": world, "images": n_done, "loop_s": round(t_loop, 3), "embed_broadcast_s": round(t_emb, 4),
             "distinct_prompts": len(prompts), "images_per_s": round(n_done / t_loop, 4) if t_loop > 0 else None,
             "max_batch_size": args.max_batch_size, "packed": not args.no_pack, "png": vae is not None, "scored": clip is not None}
    if args.stats_json:
        with open("{}.rank{}".format(args.stats_json, rank) if world > 1 else args.stats_json, "w") as f:
            json.dump(stats, f)
    if args.dist:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    print("rank {} wrote {} outputs under {} in {:.1f} s ({:.2f} images/s)".format(rank, n_done, sample_dir, t_loop,
                                                                                  n_done / t_loop if t_loop > 0 else 0.0))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
