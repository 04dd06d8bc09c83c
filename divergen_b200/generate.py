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

Host-side only: every tensor operation is a C-ABI call (StableDiffusionPipeline -> dg_denoise_loop, AutoencoderKL ->
dg_vae_decode).  With a VAE (`<ckpt_dir>/vae/*.safetensors`, or `--random_init --decode`) PNGs are written exactly where
the reference writes them; without one, latents go to `<cid>_<count:07d>.latent.pt` under the same index contract.  The
CLIP text encoder is the remaining "next" row: embeddings are synthetic unless supplied.
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
        dist.init_process_group(backend=backend)
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

    def __init__(self, workers: int = 4):
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self._pending = []

    def submit(self, images_u8, paths: Sequence[str]):
        import torch
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
        n = sum(f.result() for f in self._pending)
        self._pending = []
        self._pool.shutdown(wait=True)
        return n


def clip_scores_for(scorer, tokenizer, images_u8, category_name: str):
    """Row f3 fused into generation: the CLIP score filteration/get_clip_score.py:174-180 would compute after re-reading the
    PNG -- `text = 'a photo of a single {}'.format(' '.join(category_name.split('_')))`, `logits_per_text` of ViT-L/14 --
    taken from the uint8 image while it is still on the device (PNG is lossless and the resize is Pillow's, bit for bit, so
    the pixels are the same).  Returns a list of floats, one per image."""
    from .preprocess import clip_preprocess
    text = "a photo of a single {}".format(" ".join(category_name.split("_")))
    ids = tokenizer([text], padding="max_length", max_length=getattr(tokenizer, "model_max_length", 77), truncation=True,
                    return_tensors="pt").input_ids
    return scorer(clip_preprocess(images_u8), ids).view(-1).cpu().tolist()


def load_clip_scorer(clip_dir: str, device):
    """(tokenizer, scorer) from a transformers-format CLIP directory (`openai/clip-vit-large-patch14`: model.safetensors +
    tokenizer files), or None."""
    path = _first_existing(clip_dir, ("model.fp16.safetensors", "model.safetensors")) if clip_dir else None
    if not path:
        return None
    from safetensors.torch import load_file
    from transformers import CLIPTokenizer
    from . import CLIPScorer
    sc = CLIPScorer(device=device)
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
    """(tokenizer, text_encoder) from `<ckpt_dir>/tokenizer` + `<ckpt_dir>/text_encoder/model*.safetensors`, or None when
    either is missing (no checkpoint is reachable offline: the driver then falls back to synthetic embeddings)."""
    tok_dir = os.path.join(ckpt_dir, "tokenizer")
    te_path = _first_existing(os.path.join(ckpt_dir, "text_encoder"), ("model.fp16.safetensors", "model.safetensors"))
    if not (os.path.isdir(tok_dir) and te_path):
        return None
    from safetensors.torch import load_file
    from transformers import CLIPTokenizer
    from . import CLIPTextModel
    enc = CLIPTextModel(device=device)
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


def main(argv: Optional[Sequence[str]] = None) -> int:
    import torch
    from . import AutoencoderKL, DDIMScheduler, SD15_CONFIG, SD21_CONFIG, StableDiffusionPipeline, UNet2DConditionModel

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

    cfg = SD15_CONFIG if args.model == "sd15" else SD21_CONFIG
    unet = UNet2DConditionModel(device=device, **{k: v for k, v in cfg.items() if k != "time_cond_proj_dim"})
    unet_dir = os.path.join(args.ckpt_dir, "unet")
    st_path = os.path.join(unet_dir, "diffusion_pytorch_model.fp16.safetensors")
    if os.path.exists(st_path) and not args.random_init:
        from safetensors.torch import load_file
        unet.load_state_dict(load_file(st_path))
    elif args.random_init:
        unet.load_state_dict(random_state_dict(unet, device))
    else:
        raise FileNotFoundError("{} not found (use --random_init for benchmarks)".format(st_path))
    sched = DDIMScheduler(prediction_type="v_prediction" if args.model == "sd21" else "epsilon")
    vae = None
    vae_path = _first_existing(os.path.join(args.ckpt_dir, "vae"), ("diffusion_pytorch_model.fp16.safetensors",
                                                                     "diffusion_pytorch_model.safetensors"))
    if vae_path and not args.random_init:
        from safetensors.torch import load_file
        vae = AutoencoderKL(device=device)
        vae.load_state_dict(load_file(vae_path))
    elif args.random_init and args.decode:
        vae = AutoencoderKL(device=device)
        vae.load_state_dict(random_state_dict(vae, device, seed=1))
    ext = "png" if vae is not None else "latent.pt"
    pipe = StableDiffusionPipeline(unet, sched, vae=vae)
    pipe.enable_model_cpu_offload(local_rank)                                    # no-ops kept for call parity (:143,186)

    generator = torch.manual_seed(args.seed + rank)                             # reference :200 (global CPU generator)
    files = list_prompt_files(args.from_file) if args.from_file else []
    if args.max_prompt_files > 0:
        files = files[:args.max_prompt_files]
    per_file = [(f, open(f).read().splitlines()) for f in files] if files else [("prompt.txt", [args.prompt])]
    # ---- text embeddings: once, on rank 0, for every distinct prompt; one broadcast
    prompts = sorted({l.strip() for _, lines in per_file for l in lines})
    table = None
    if rank == 0:
        te = None if args.random_init else load_text_encoder(args.ckpt_dir, device)
        table = (encode_prompt_table(prompts, te[0], te[1]) if te is not None
                 else synthetic_text_embeddings(prompts, cfg["cross_attention_dim"])).to(device)
    table = broadcast_embedding_table(table if table is not None else torch.empty(0, device=device))
    row = {p: i + 1 for i, p in enumerate(prompts)}

    n_done = 0
    writer = AsyncImageWriter(args.png_workers) if vae is not None else None
    clip = load_clip_scorer(args.clip_dir, device) if (vae is not None and id_to_name is not None) else None
    scores = {}
    for fi, (path, lines) in enumerate(per_file):
        cid = category_id_of(path)
        cat_dir = output_dir_for(args.outdir, stage, cid, id_to_name)
        if id_to_name is not None:
            os.makedirs(cat_dir, exist_ok=True)                                   # convert_dir_structure.py:94-99 (every rank: idempotent)
        print("==> Reading prompts from {}, {}/{}".format(path, fi + 1, len(per_file)))
        for call in iter_calls(lines, plan, rank, args.n_samples, args.offset):
            last = os.path.join(cat_dir, output_name(cid, call.counts[-1], ext))
            if args.disable_overwrite and os.path.exists(last):
                print("==> Skipping stage {} for {}...".format(stage, os.path.basename(last)))
                continue
            pos = table[row[call.prompt]][None]
            neg = table[0][None]
            out = pipe(prompt_embeds=pos, negative_prompt_embeds=neg, generator=generator,
                       output_type="uint8" if vae is not None else "latent",
                       num_images_per_prompt=call.num_images, num_inference_steps=args.num_inference_steps,
                       guidance_scale=args.guidance_scale).images
            dsts = [os.path.join(cat_dir, output_name(cid, count, ext)) for count in call.counts]
            if vae is not None:
                writer.submit(out, dsts)                                          # reference :267 (pt -> PIL -> .save), asynchronously
                if clip is not None:
                    for dst, sc in zip(dsts, clip_scores_for(clip[1], clip[0], out, id_to_name[int(cid)])):
                        scores.setdefault(cid, {})[os.path.basename(dst)] = sc
            else:
                for j, dst in enumerate(dsts):
                    torch.save(out[j].cpu(), dst)
            n_done += call.num_images
    if writer is not None:
        writer.close()
    if clip is not None:
        import json
        with open(os.path.join(args.outdir, "clip_scores_{}_rank{}.json".format(stage, rank)), "w") as f:
            json.dump(scores, f)
    if args.dist:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    print("rank {} wrote {} outputs under {}".format(rank, n_done, sample_dir))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
