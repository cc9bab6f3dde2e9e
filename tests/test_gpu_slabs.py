"""Row-slab partition (multi-GPU path) exercised on ONE GPU: K slab contexts in one process.
Default: the library's own multi-GPU entry points (weed_group_create / weed_group_step: pack kernels
write into the neighbours' receive buffers, the waits are on the device — the same kernels the
multi-process path runs over CUDA IPC mappings).  The partitioned world must reproduce the
unpartitioned engine BIT FOR BIT for every entity, every frame."""
import numpy as np
import pytest

from helpers import bits
from multithreadedgameengine_b200 import binding as B, scenes
from multithreadedgameengine_b200.slabs import SlabGroup, plan_slabs

pytestmark = pytest.mark.gpu

KEYS = ("T.x", "T.y", "RB.px", "RB.py", "RB.vx", "RB.vy", "RB.speed", "RB.collisionCount")


def run_case(cfg, cols, world, frames, plan=None, **kw):
    from multithreadedgameengine_b200.engine import GameEngine
    N = cfg["entityCount"]
    ref = GameEngine(cfg, host_neighbor_rows=False)
    ref.load_columns(cols)
    grp = SlabGroup(cfg, cols, world, plan=plan, **kw)
    for f in range(frames):
        ref.step(1.0, 0, B.COLS_INPUT_ALL | B.COL_COLLISIONS)
        grp.step(1.0)
        got, seen = grp.gather(N, KEYS)
        act = ref.col["T.active"] != 0
        assert (seen[act] == 1).all(), f"frame {f}: {int((seen[act] != 1).sum())} entities not owned exactly once"
        for k in KEYS:
            a, b = bits(ref.col[k][act]), bits(got[k][act])
            assert np.array_equal(a, b), f"frame {f} {k}: {int((a != b).sum())} mismatches"
    ref.close()
    return grp


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_reproduce_single_context_balls(world):
    cfg, cols = scenes.scaled("config4", 60000)
    run_case(cfg, cols, world, frames=6).close()


def test_slabs_with_fast_movers_migrate():
    """Large start velocities push entities across the cuts every frame."""
    cfg, cols = scenes.scaled("config3", 40000)
    rng = np.random.default_rng(3)
    v = ((rng.random((2, cfg["entityCount"])) - 0.5) * 60).astype(np.float32)
    cols["RB.px"] = (cols["T.x"].astype(np.float64) - v[0]).astype(np.float32)
    cols["RB.py"] = (cols["T.y"].astype(np.float64) - v[1]).astype(np.float32)
    cfg["physics"]["gravity"] = dict(x=0.0, y=0.0)
    cfg["physics"]["verletDamping"] = 1.0
    run_case(cfg, cols, 4, frames=8).close()


def test_observer_on_a_cut_and_capped_rows():
    """The Mouse (trigger, visualRange 150, capped row) parked exactly on a cut, in a dense
    scene whose rows hit the cap: exercises the 2*h_obs part of the halo depth and the
    explicit-pair paths across slabs."""
    cfg, cols = scenes.balls_synthetic(30000, (1600.0, 1600.0), 16.0, 6, 2, (2.0, 6.0), 16.0, seed=9)
    blocks, H = plan_slabs(cfg, cols, 2)
    cols["T.x"][0] = 800.0
    cols["T.y"][0] = blocks[0][1] * 16.0 + 1.0
    run_case(cfg, cols, 2, frames=5).close()


def test_plan_balances_entities():
    from multithreadedgameengine_b200.slabs import halo_rows
    cfg, cols = scenes.scaled("config4", 50000)
    blocks, H = plan_slabs(cfg, cols, 4)
    assert blocks[0][0] == 0 and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    assert H == 20                         # clusters piled on the y = 0 wall pull the first cut next to the Mouse: 2 * ceil(150/16)
    cfg, cols = scenes.scaled("config3", 50000)
    blocks, H = plan_slabs(cfg, cols, 2)
    assert H == 3                          # (S+1) * ceil(16/16): the Mouse sits in a corner, far from the cut
    assert halo_rows(cfg, cols) == 20      # without the cuts its range counts


@pytest.mark.parametrize("mode", ["c", "py", "nccl-buffers"])
def test_every_transport_reproduces_the_single_context(mode):
    """weed_group_* (mode c), SlabEngines connected through weed_slab_exchange_connect (py) and the
    staging path weed_slab_pack / weed_slab_apply (what the NCCL fallback uses)."""
    cfg, cols = scenes.scaled("config4", 50000)
    run_case(cfg, cols, 3, frames=5, mode=mode).close()


def test_reach_guard_reports_an_observer_that_comes_near_a_cut():
    """The halo is planned for the entities near the cuts at the start (3 rows here: the Mouse is
    far away).  Move the Mouse next to a cut: its capped row would need 20 rows of halo, the
    partition could silently diverge — the slab must say so instead."""
    cfg, cols = scenes.scaled("config3", 50000)
    plan = plan_slabs(cfg, cols, 2)
    assert plan[1] == 3
    cols["T.x"][0] = 300.0
    cols["T.y"][0] = plan[0][0][1] * 16.0 - 40.0          # two rows below the cut
    grp = SlabGroup(cfg, cols, 2, plan=plan)
    with pytest.raises(B.WeedError) as e:
        grp.step(1.0)
    assert e.value.code == B.WEED_E_OVERFLOW and "reach" in str(e.value)
    grp.close()


def test_dynamic_cuts_follow_the_load_and_stay_bit_exact():
    """weed_slab_balance: start from deliberately bad cuts (equal row counts over a scene whose
    entities sit in the lower half), let the cuts move by up to 2 rows per frame.  Every frame
    must still equal the unpartitioned world bit for bit; the cuts must stay contiguous, must
    have moved, and must have moved toward the crowded half."""
    cfg, cols = scenes.balls_synthetic(60000, (2048.0, 2048.0), 16.0, 24, 2, (2.0, 5.0), 16.0, seed=31)
    cols["T.y"][1:] = (cols["T.y"][1:] * np.float32(0.5) + np.float32(1000.0)).astype(np.float32)   # rows 62..127 of 128
    cols["RB.py"][1:] = cols["T.y"][1:]
    cols["C.visualRange"][0] = 16.0         # a short-sighted Mouse: halo 3 rows, so 43-row slabs may shrink
    rows = 128
    _, H = plan_slabs(cfg, cols, 3)
    plan = ([(0, 43), (43, 86), (86, rows)], H)
    grp = run_case(cfg, cols, 3, frames=25, plan=plan, balance_rows=2, balance_hysteresis=3)
    st = [s.status() for s in grp.slabs]
    cuts = [(x["rowBegin"], x["rowEnd"]) for x in st]
    assert cuts[0][0] == 0 and cuts[-1][1] == rows and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:])), cuts
    assert sum(x["cutMoves"] for x in st) > 0, st
    assert cuts[0][1] > 43, cuts            # slab 0 was almost empty: it has to grow
    grp.close()
