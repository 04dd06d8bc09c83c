"""World-size-2 `gloo` test of the multi-GPU plumbing (no GPU): the single data-carrying collective of the path (broadcast of
the text-embedding table from rank 0) and the rank sharding of file indices."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from divergen_b200.generate import (broadcast_embedding_table, gather_clip_results, iter_calls, plan_batches,
                                    synthetic_text_embeddings)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prompts = ["a photo of a single cat", "a photo of a single dog"]
    table = synthetic_text_embeddings(prompts, 64) if rank == 0 else torch.empty(0)
    table = broadcast_embedding_table(table)
    plan = plan_batches(8, world, 3)
    counts = [c for call in iter_calls(prompts, plan, rank, 8, 100) for c in call.counts]
    torch.save({"table": table, "counts": counts}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_broadcast_and_sharding_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(world)]
    want = synthetic_text_embeddings(["a photo of a single cat", "a photo of a single dog"], 64)
    for x in r:
        assert x["table"].dtype == torch.float16 and torch.equal(x["table"], want)
    assert sorted(r[0]["counts"] + r[1]["counts"]) == list(range(100, 116))
    assert not set(r[0]["counts"]) & set(r[1]["counts"])


def _gather_worker(rank, world, port, out_dir):
    import json
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cats = os.path.join(out_dir, "cats.json")
    if rank == 0:
        json.dump([{"id": 1, "name": "aerosol_can"}, {"id": 7, "name": "alligator"}], open(cats, "w"))
    dist.barrier()
    # every rank scored its own images (file indices offset by total_batch_size * rank): uneven on purpose
    plan = plan_batches(6, world, 2)
    mine = {"7": {"7_{:07d}.png".format(c): float(c) for call in iter_calls(["x"], plan, rank, 6, 0) for c in call.counts}}
    if rank == 1:
        mine["1"] = {"1_0000009.png": 9.0}
    gather_clip_results(mine, {1: "aerosol_can", 7: "alligator"}, cats, os.path.join(out_dir, "results.json"), rank, world)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_clip_scores_gathered_and_index_sorted_world2(tmp_path):
    """filteration/get_clip_score.py:183-212 on two ranks: rank 0 writes one results.json with the scores of BOTH ranks in
    file-index order."""
    import json
    world, port = 2, _free_port()
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    data = json.load(open(tmp_path / "results.json"))
    assert [c["name"] for c in data] == ["aerosol_can", "alligator"]
    assert data[0]["clip_scores"] == [9.0]
    assert data[1]["clip_scores"] == [float(i) for i in range(6)]
