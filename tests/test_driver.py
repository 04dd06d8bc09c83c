"""Host logic of the generation driver (divergen_b200/generate.py) against the reference's index arithmetic
(DiverGen/generation/txt2img_diffusers_stages_from_txt.py:124-131, :221-238, :262-263) -- pure Python, no GPU."""
import itertools
import os

import pytest

from divergen_b200.generate import (build_parser, category_id_of, iter_calls, list_prompt_files, output_name, plan_batches)


def _reference_counts(lines, n_samples, world_size, max_batch_size, rank, offset):
    """Line-by-line restatement of the reference loop (returns [(prompt, [counts...]), ...])."""
    total_batch_size = n_samples // world_size
    assert total_batch_size * world_size == n_samples
    batch_size = total_batch_size // max_batch_size
    remainder_batch_size = total_batch_size % max_batch_size
    if remainder_batch_size > 0:
        batch_size += 1
    data = sorted(batch_size * list(lines))
    out, tmp = [], 0
    for i, prompt in enumerate(data):
        prompt = prompt.strip()
        if i % batch_size == 0:
            tmp = 0
            cur = remainder_batch_size if remainder_batch_size != 0 else max_batch_size
        else:
            cur = max_batch_size
        out.append((prompt, [j + tmp + total_batch_size * rank + offset + (i // batch_size) * n_samples for j in range(cur)]))
        tmp += cur
    return out


@pytest.mark.parametrize("n_samples,world,mbs", [(1024, 8, 1), (1024, 8, 4), (8, 8, 1), (24, 2, 5), (7, 1, 3), (32, 4, 8)])
def test_calls_match_reference_loop(n_samples, world, mbs):
    lines = ["a photo of a single cat, in a white background", "a photo of a single aerosol can "]
    plan = plan_batches(n_samples, world, mbs)
    for rank in range(world):
        got = [(c.prompt, list(c.counts)) for c in iter_calls(lines, plan, rank, n_samples, 1024)]
        assert got == _reference_counts(lines, n_samples, world, mbs, rank, 1024)


@pytest.mark.parametrize("n_samples,world,mbs", [(1024, 8, 4), (24, 2, 5), (16, 4, 3)])
def test_ranks_tile_the_index_range_exactly_once(n_samples, world, mbs):
    """convert_dir_structure.py:133-138 checks count == n_samples per category: all ranks together must write every index of
    [offset + k*n_samples, offset + (k+1)*n_samples) for prompt k exactly once."""
    lines = ["p%d" % i for i in range(3)]
    plan = plan_batches(n_samples, world, mbs)
    per_prompt = {}
    for rank in range(world):
        for c in iter_calls(lines, plan, rank, n_samples, 1024):
            per_prompt.setdefault(c.prompt, []).extend(c.counts)
            assert 1 <= c.num_images <= mbs
    for k, p in enumerate(sorted(lines)):
        assert sorted(per_prompt[p]) == list(range(1024 + k * n_samples, 1024 + (k + 1) * n_samples))


def test_plan_rejects_indivisible():
    with pytest.raises(ValueError):
        plan_batches(10, 4, 1)


def test_names_and_category_ids(tmp_path):
    assert output_name("17", 1029) == "17_0001029.png"            # parsed back by convert_dir_structure.py:116-121
    for cid in ("3", "12", "1203"):
        (tmp_path / (cid + ".txt")).write_text("a photo of a single thing\n")
    files = list_prompt_files([str(tmp_path)])
    assert [category_id_of(f) for f in files] == sorted(["3", "12", "1203"])
    assert list_prompt_files(["a.txt", "b.txt"]) == ["a.txt", "b.txt"]


def test_cli_keeps_reference_flags():
    a = build_parser().parse_args(["--from_file", "input/lvis_prompt/", "--outdir", "o", "--n_samples", "1024", "--max_batch_size", "1",
                                   "--seed", "7", "--dist", "--ckpt_dir", "c", "--stages", "sd", "--offset", "0", "--disable_overwrite"])
    assert (a.from_file, a.n_samples, a.max_batch_size, a.seed, a.dist, a.offset, a.disable_overwrite) == (
        ["input/lvis_prompt/"], 1024, 1, 7, True, 0, True)


def test_category_layout_matches_convert_dir_structure(tmp_path):
    """Row f4: with the LVIS category JSON the driver's destination is exactly what generation/convert_dir_structure.py
    :113-121 produces from `samples/<stage>/<id>_<index>.png`: `<outdir>/<stage>/<category_name>/<same file name>`."""
    import json
    from divergen_b200.generate import load_category_names, output_dir_for, output_name
    cats = [{"id": 1, "name": "aerosol_can"}, {"id": 7, "name": "alligator"}]
    jp = tmp_path / "cats.json"
    jp.write_text(json.dumps(cats))
    id_to_name = load_category_names(str(jp))
    assert id_to_name == {1: "aerosol_can", 7: "alligator"}
    # the reference's conversion of one generated file
    filename = output_name("7", 1034, "png")
    category_id = int(filename.split("_")[0])                                   # convert_dir_structure.py:116
    want = os.path.join("out", "sd", id_to_name[category_id], filename)         # :118-120
    assert os.path.join(output_dir_for("out", "sd", "7", id_to_name), filename) == want
    assert output_dir_for("out", "sd", "7", None) == os.path.join("out", "samples", "sd")


def test_category_layout_rejects_unknown_ids():
    from divergen_b200.generate import output_dir_for
    with pytest.raises(ValueError):
        output_dir_for("out", "sd", "prompt", {1: "aerosol_can"})        # --prompt runs have no category id
    with pytest.raises(ValueError):
        output_dir_for("out", "sd", "9", {1: "aerosol_can"})


def test_clip_score_text_is_the_filteration_prompt():
    """`text='a photo of a single {}'.format(' '.join(category_name.split('_')))` (get_clip_score.py:175)."""
    from divergen_b200.generate import clip_prompt_text
    assert clip_prompt_text("aerosol_can") == "a photo of a single aerosol can"
    assert clip_prompt_text("alligator") == "a photo of a single alligator"
    assert clip_prompt_text("bow_(decorative_ribbons)") == "a photo of a single bow (decorative ribbons)"


def test_pack_calls_keeps_order_indices_and_limits():
    """Packing across prompts (the gpt-prompt recipe: `--n_samples 8` on 8 ranks = 1 image per prompt per rank): groups are
    consecutive, never split a call, never exceed max_batch_size, and together are exactly the unpacked call sequence."""
    from divergen_b200.generate import iter_calls, pack_calls, plan_batches
    lines = ["prompt {}".format(k) for k in range(11)]
    plan = plan_batches(8, 8, 4)                                   # 1 image per prompt per rank
    calls = list(iter_calls(lines, plan, rank=3, n_samples=8, offset=1024))
    groups = list(pack_calls(calls, 4))
    assert [len(g) for g in groups] == [4, 4, 3]
    assert [c for g in groups for c in g] == calls
    assert all(sum(c.num_images for c in g) <= 4 for g in groups)
    # remainder batches: 6 images per prompt per rank at max_batch_size 4 -> calls of 2, 4, 2, 4, ...; a group holds one
    # 4-call or a 2-call alone (2 + 4 > 4), i.e. packing never reorders or splits
    plan = plan_batches(6, 1, 4)
    calls = list(iter_calls(["a", "b"], plan, rank=0, n_samples=6, offset=0))
    assert [c.num_images for c in calls] == [2, 4, 2, 4]
    groups = list(pack_calls(calls, 4))
    assert [[c.num_images for c in g] for g in groups] == [[2], [4], [2], [4]]
    groups = list(pack_calls(calls, 8))
    assert [[c.num_images for c in g] for g in groups] == [[2, 4, 2], [4]]
    assert [n for g in groups for c in g for n in c.counts] == [n for c in calls for n in c.counts]
    assert list(pack_calls([], 4)) == []


def test_writer_surfaces_save_errors_early(tmp_path):
    """A failed save raises at a later submit / close instead of only after the whole run; back-pressure bounds the queue."""
    import numpy as np
    from divergen_b200.generate import AsyncImageWriter

    class FakeImages:                       # stands in for a device tensor: AsyncImageWriter only copies and records an event
        pass

    w = AsyncImageWriter(workers=1, max_in_flight=2)
    import concurrent.futures as cf
    # drive the reaping logic directly with futures (no GPU here)
    ok = cf.Future(); ok.set_result(2)
    bad = cf.Future(); bad.set_exception(OSError("disk full"))
    w._pending = [ok, bad]
    with pytest.raises(OSError):
        w._reap(0)
    assert w._written == 2
    w._pending = []
    assert w.close() == 2


def test_clip_score_job_helpers(tmp_path):
    """filteration/get_clip_score.py bookkeeping: stage directories (:87-101), mask paths (:134-137), rank striding (:113-115),
    index-sorted merge of the gathered scores (:192-200)."""
    from types import SimpleNamespace
    from divergen_b200.clip_score import build_parser, mask_path_for, merge_gathered, picked_for_rank, stage_dirs
    a = SimpleNamespace(indir="in", outdir="out", use_mask=False, in_mask_dir="m", seg_name="sam")
    assert stage_dirs(a, "sd") == ("in", "out") and stage_dirs(a, "III") == (os.path.join("in", "III"), os.path.join("out", "III"))
    a.use_mask = True
    assert stage_dirs(a, "sd") == ("in", os.path.join("out", "sam"))
    assert stage_dirs(a, "III") == (os.path.join("in", "III"), os.path.join("out", "III", "sam"))
    assert mask_path_for(a, "sd", "cat", "/x/y/1_0000001.png") == os.path.join("m", "sam", "cat", "1_0000001.png")
    assert mask_path_for(a, "III", "cat", "/x/y/1_0000001.png") == os.path.join("m", "III", "sam", "cat", "1_0000001.png")
    paths = ["p{}".format(i) for i in range(7)]
    got = [picked_for_rank(paths, r, 3) for r in range(3)]
    assert [[i for i, _ in g] for g in got] == [[0, 3, 6], [1, 4], [2, 5]]
    merged = merge_gathered([[0, 3, 6], [1, 4], [2, 5]], [[0.0, 3.0, 6.0], [1.0, 4.0], [2.0, 5.0]])
    assert merged == [float(i) for i in range(7)]
    p = build_parser().parse_args(["--indir", "i", "--outdir", "o", "--use_mask", "--in_mask_dir", "m", "--seg_name", "s", "--dist",
                                   "--n_samples", "1024", "1280", "--max_batch_size", "8", "--stages", "sd"])
    assert p.n_samples == [1024, 1280] and p.use_mask and p.max_batch_size == 8


def test_gather_clip_results_single_rank(tmp_path):
    """The fused-scoring job leaves the reference's results.json: the category list with `clip_scores` in sorted-file order."""
    import json
    from divergen_b200.generate import gather_clip_results
    cats = tmp_path / "cats.json"
    cats.write_text(json.dumps([{"id": 1, "name": "aerosol_can"}, {"id": 7, "name": "alligator"}]))
    scores = {"7": {"7_0000002.png": 2.0, "7_0000000.png": 0.5, "7_0000001.png": 1.0}}
    data = gather_clip_results(scores, {1: "aerosol_can", 7: "alligator"}, str(cats), str(tmp_path / "results.json"), 0, 1)
    assert json.load(open(tmp_path / "results.json")) == data
    assert data[0]["clip_scores"] == [] and data[1]["clip_scores"] == [0.5, 1.0, 2.0]
