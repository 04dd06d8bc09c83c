"""Host logic of the wave-balanced attention grid (divergen_b200/csrc/host_common.cuh: plan_attn_grid; attn_tc.cuh: the
CTA -> (batch, head, first query tile, tiles) decomposition).  No GPU: the planner is a host function behind the C ABI
(dg_plan_attention_grid); the kernel's index arithmetic is restated here and checked to cover every query tile once."""
import ctypes as C
import heapq

import pytest


def _plan(lib, n_bh, q_tiles, big, sms=148):
    plan = (C.c_int32 * 7)()
    mk = (C.c_double * 2)()
    assert lib.dg_plan_attention_grid(n_bh, q_tiles, big, sms, plan, mk) == 0
    return list(plan), list(mk)


def _decompose(idx, n_bh, big, n_g1, a1, b1, a2, b2):
    """attn_tc_kernel's `if (p.part_on)` block."""
    f1, f2, t1 = n_g1 * a1, (n_bh - n_g1) * a2, n_g1 * b1
    if idx < f1:
        return idx // a1, (idx % a1) * big, big
    if idx < f1 + f2:
        idx -= f1
        return n_g1 + idx // a2, (idx % a2) * big, big
    if idx < f1 + f2 + t1:
        idx -= f1 + f2
        return idx // b1, a1 * big + (idx % b1) * (big - 1), big - 1
    idx -= f1 + f2 + t1
    return n_g1 + idx // b2, a2 * big + (idx % b2) * (big - 1), big - 1


def _simulate(costs, sms):
    """Block scheduler: the next CTA in grid order goes to the first SM that becomes free (one CTA per SM)."""
    free = [0.0] * sms
    heapq.heapify(free)
    end = 0.0
    for c in costs:
        t = heapq.heappop(free) + c
        end = max(end, t)
        heapq.heappush(free, t)
    return end


CASES = [(64, 32, 4), (32, 32, 4), (40, 16, 4), (40, 8, 2), (100, 3, 2), (50, 7, 2), (56, 16, 4), (24, 32, 4),
         (64, 8, 2), (20, 72, 2), (8, 32, 4), (128, 32, 4), (1, 1, 2), (3, 5, 4)]


@pytest.mark.parametrize("n_bh,q_tiles,big", CASES)
def test_plan_covers_every_tile_once_and_matches_the_simulation(built_lib, n_bh, q_tiles, big):
    (on, n_g1, a1, b1, a2, b2, ctas), (makespan, uniform) = _plan(built_lib, n_bh, q_tiles, big)
    if not on:
        assert (n_g1, a1, b1, a2, b2, ctas) == (0, 0, 0, 0, 0, 0)
        return
    assert 0 < n_g1 <= n_bh
    assert a1 * big + b1 * (big - 1) == q_tiles
    if n_g1 < n_bh:
        assert a2 * big + b2 * (big - 1) == q_tiles
    assert ctas == n_g1 * (a1 + b1) + (n_bh - n_g1) * (a2 + b2)
    seen = set()
    sizes = []
    for idx in range(ctas):
        bh, tile0, nq = _decompose(idx, n_bh, big, n_g1, a1, b1, a2, b2)
        assert 0 <= bh < n_bh and nq in (big, big - 1)
        sizes.append(nq)
        for t in range(tile0, tile0 + nq):
            assert 0 <= t < q_tiles and (bh, t) not in seen
            seen.add((bh, t))
    assert len(seen) == n_bh * q_tiles
    assert sizes == sorted(sizes, reverse=True), "full-size CTAs come first in grid order (longest first)"
    sim = _simulate([float(big) if s == big else (big - 1) * 1.04 for s in sizes], 148)
    assert abs(sim - makespan) < 1e-6
    assert makespan <= 0.97 * uniform + 1e-9


def test_sd15_level0_shapes(built_lib):
    """UNet batch 8 x 8 heads x 32 query tiles, 4 tiles per CTA: 512 uniform CTAs = 4 waves = 16 tile times on 148 SMs; the
    balanced grid needs 14.24 (2048 tiles / 148 SMs = 13.84 is the floor).  The CFG-prefix half batch: 8 -> 7.12."""
    plan, (makespan, uniform) = _plan(built_lib, 64, 32, 4)
    assert plan[0] == 1 and uniform == 16.0 and makespan == pytest.approx(14.24)
    plan, (makespan, uniform) = _plan(built_lib, 32, 32, 4)
    assert plan[0] == 1 and uniform == 8.0 and makespan == pytest.approx(7.12)
    # SD-2.1 level 0 (UNet batch 4 x 5 heads x 72 tiles, 2 per CTA): 720 CTAs = 4.86 waves, nothing to gain
    assert _plan(built_lib, 20, 72, 2)[0][0] == 0
