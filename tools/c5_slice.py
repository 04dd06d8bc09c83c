"""BASELINE config 5 (LVIS 1203-category prompt sweep, 1.23 M images on 8 x B200), measured on a fixed slice through the
real driver (`python -m divergen_b200.generate`): random-init SD-1.5 UNet + VAE, PNGs written, text-embedding table built and
broadcast, both recipes of DiverGen/DATA.md:9-39 --

  lvis : one prompt per category, `--n_samples 1024` over 8 ranks = 128 images per rank per category;
         slice = C categories x I images per rank (defaults 64 x 32), micro-batches of 4;
  gpt  : 128 prompts per category, `--n_samples 8` over 8 ranks = ONE image per prompt per rank;
         slice = G categories x 128 prompts, packed across prompts (UNet batch 8) and, for comparison, unpacked (UNet batch 2).

Usage (single GPU, or under torchrun for N ranks):
  python tools/c5_slice.py --out profiles/r02_c5_slice.json [--cats 64 --images 32 --gpt_cats 2] [--dist]
Prompt text is synthetic (`a photo of a single category_<i>, ...`): throughput does not depend on it; the reference's prompt
files live in /root/reference, which does not exist on the GPU box.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOTAL_LVIS = 1059 * 1024      # images of recipe 1.1 (DiverGen/input/lvis_prompt: 1059 files, n_samples 1024)
TOTAL_GPT = 144 * 128 * 8     # images of recipe 1.2 (DiverGen/input/gpt_prompt: 144 files x 128 prompts, n_samples 8)


def write_prompts(d, cats, prompts_per_cat, first_id=1):
    os.makedirs(d, exist_ok=True)
    names = []
    for c in range(cats):
        cid = first_id + c
        with open(os.path.join(d, "{}.txt".format(cid)), "w") as f:
            for k in range(prompts_per_cat):
                f.write("a photo of a single category_{} number {}, in a white background\n".format(cid, k))
        names.append({"id": cid, "name": "category_{}".format(cid)})
    return names


def run(argv, stats_path, world, rank):
    from divergen_b200.generate import main
    t0 = time.perf_counter()
    assert main(argv + ["--stats_json", stats_path]) == 0
    wall = time.perf_counter() - t0
    p = "{}.rank{}".format(stats_path, rank) if world > 1 else stats_path
    st = json.load(open(p))
    st["wall_s_incl_model_build"] = round(wall, 2)
    return st


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_c5_slice.json"))
    ap.add_argument("--cats", type=int, default=64)
    ap.add_argument("--images", type=int, default=32, help="images per rank per category (lvis recipe)")
    ap.add_argument("--gpt_cats", type=int, default=2)
    ap.add_argument("--dist", action="store_true")
    ap.add_argument("--skip_unpacked", action="store_true")
    ap.add_argument("--max_batch_size", type=int, default=4, help="images per pipeline call (UNet batch = 2x with CFG)")
    ap.add_argument("--only", choices=["lvis", "gpt"], default=None, help="run one recipe (one process group per torchrun launch)")
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    # one scratch directory shared by the ranks of a node (rank 0's name is passed through the environment by the launcher)
    tmp = os.environ.get("DG_C5_TMP") or tempfile.mkdtemp(prefix="dg_c5_")
    os.makedirs(tmp, exist_ok=True)
    res = {"world": world, "max_batch_size": a.max_batch_size, "model": "SD-1.5 512x512, 50 DDIM steps, CFG 7.5, random-init UNet + VAE, PNGs written"}
    dist_flag = ["--dist"] if a.dist else []
    common = ["--random_init", "--decode", "--seed", "42", "--offset", "0", "--max_batch_size", str(a.max_batch_size)] + dist_flag
    # ---- recipe 1.1 (lvis prompts)
    cats = write_prompts(os.path.join(tmp, "lvis_prompt"), a.cats, 1)
    json.dump(cats, open(os.path.join(tmp, "cats.json"), "w"))
    if a.only in (None, "lvis"):
      st = run(common + ["--from_file", os.path.join(tmp, "lvis_prompt"), "--outdir", os.path.join(tmp, "out_lvis"),
                       "--n_samples", str(a.images * world), "--in_lvis_json_path", os.path.join(tmp, "cats.json")],
             os.path.join(tmp, "lvis_stats.json"), world, rank)
      res["lvis"] = dict(st, categories=a.cats, images_per_rank_per_category=a.images)
    # ---- recipe 1.2 (gpt prompts): packed, then unpacked
    write_prompts(os.path.join(tmp, "gpt_prompt"), a.gpt_cats, 128, first_id=2001)
    if a.only in (None, "gpt"):
      st = run(common + ["--from_file", os.path.join(tmp, "gpt_prompt"), "--outdir", os.path.join(tmp, "out_gpt"), "--n_samples", str(world)],
             os.path.join(tmp, "gpt_stats.json"), world, rank)
      res["gpt_packed"] = dict(st, categories=a.gpt_cats, prompts_per_category=128)
    if not a.skip_unpacked and a.only is None:
        st = run(common + ["--no_pack", "--from_file", os.path.join(tmp, "gpt_prompt"), "--outdir", os.path.join(tmp, "out_gpt_np"),
                           "--n_samples", str(world)], os.path.join(tmp, "gpt_np_stats.json"), world, rank)
        res["gpt_unpacked"] = dict(st, categories=a.gpt_cats, prompts_per_category=128)
    if rank == 0 and a.only is not None:
        json.dump(res, open(a.out, "w"), indent=1)
        print(json.dumps(res))
    elif rank == 0:
        n_png = sum(len(fs) for _, _, fs in os.walk(os.path.join(tmp, "out_lvis")))
        res["lvis"]["png_files_on_disk_all_ranks"] = n_png
        # projection: every rank runs at this rank's measured rate (no data-path collective: SCALE shows 0.98-1.0 weak scaling)
        r_l, r_g = res["lvis"]["images_per_s"], res["gpt_packed"]["images_per_s"]
        gpus = 8
        hours = (TOTAL_LVIS / (r_l * gpus) + TOTAL_GPT / (r_g * gpus)) / 3600.0
        res["projection"] = {"images": TOTAL_LVIS + TOTAL_GPT, "gpus": gpus, "hours": round(hours, 2),
                             "assumes": "8 ranks at this run's per-rank rates ({:.2f} / {:.2f} images/s lvis / gpt recipe)".format(r_l, r_g),
                             "roofline_hours": round((TOTAL_LVIS + TOTAL_GPT) / (17.2 * gpus) / 3600.0, 2)}
        json.dump(res, open(a.out, "w"), indent=1)
        print(json.dumps(res))
        if not os.environ.get("DG_C5_TMP"):
            shutil.rmtree(tmp, ignore_errors=True)
