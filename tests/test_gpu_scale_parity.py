"""Bit-exact parity at benchmark scale (the GPU-vs-oracle physics cases stopped at a few
thousand entities).  Every frame of these runs compares the FULL state (x, y, px, py, vx, vy, ax, ay,
speed, collisionCount bit for bit, velocityAngle within 1 float32 ulp), collisionData (count, pairs,
order) and, where they fit in host memory, all neighbor rows (ids, order, cap, float32 d² bits)
against the C oracle in its J-order mode — the code paths only large worlds reach (multi-tile
decoupled look-back in k_cell_scan, many-tile pair scan, weed_step's pipelined copies above 2^18
entities) are therefore compared with the oracle, not only with themselves.

  config 4 at 400 k   mixed radii, clusters piled on the walls (cap path, explicit lists), 5 frames
  config 3 at 1 M     its full size, 5 frames
  config 2            10 000 prey + 500 predators, maxNeighbors 1500 (its real size), 4 frames
  piles               200 / 1500 / 5000 entities in single cells over a dense scene: sorted cell lists, the reverse-edge
                      cap path of the dense regime, the warp-per-entity sweep of pool / resumed-scan entities, 3 frames
  config 4 at 16 M    its full size, 2 frames: `pytest -m gpu --runslow` (about two minutes of CPU oracle)

Tolerance: none for integer/position work (bit-exact); velocityAngle 1 float32 ulp (device atan2 vs
libm atan2, the only transcendental on the path).
"""
import numpy as np
import pytest

from helpers import active_rows_equal, make_oracle
from multithreadedgameengine_b200 import binding as B, scenes
from oracle.oracle_c import OracleC
from test_gpu_parity import ALL_DL, compare_rows, compare_state, make_engine

pytestmark = pytest.mark.gpu


def run_and_compare(cfg, cols, frames, rows=True, flags=0, via_step=False):
    eng = make_engine(cfg, cols, flags=flags, host_neighbor_rows=rows)
    ora = make_oracle(OracleC, cfg, cols)
    dl = ALL_DL if rows else (B.COLS_INPUT_ALL | B.COL_COLLISIONS)
    for f in range(frames):
        if via_step:       # the host-facing call with copies (weed_step: pipelined above 2^18 entities)
            eng.step(1.0, eng.mask("RB.ax", "RB.ay"), dl)
        else:
            eng.run(1)
            eng.download(dl)
        ora.step(1.0, 1)
        compare_state(eng, ora, f"frame {f}")
        if rows:
            compare_rows(eng, ora, cfg)
    st = eng.stats()
    eng.close()
    return st


def test_config4_scaled_400k_5_frames_bit_exact():
    cfg, cols = scenes.scaled("config4", 400_000)
    cfg["physics"]["maxCollisionPairs"] = 2_000_000          # log every pair, not the first 10 000
    st = run_and_compare(cfg, cols, 5)
    assert st["cappedRows"] > 0 and st["explicitPairs"] > 0   # the cap path was exercised


def test_config4_scaled_400k_through_weed_step_with_copies():
    cfg, cols = scenes.scaled("config4", 400_000)
    run_and_compare(cfg, cols, 3, via_step=True)


def test_config3_1M_5_frames_bit_exact():
    cfg, cols = scenes.config3()
    cfg["physics"]["maxCollisionPairs"] = 4_000_000
    run_and_compare(cfg, cols, 5)


def test_config2_full_size_bit_exact():
    cfg, cols = scenes.boids()                                # 10 000 prey + 500 predators, M 1500
    st = run_and_compare(cfg, cols, 4)
    assert st["neighborsTotal"] > 0


def test_tiled_sweep_is_bit_identical_at_scale():
    """The TMA-tiled sweep stays in the library for A/B measurements: it must produce the oracle's bits
    too (300 k entities, clusters: capped rows, explicit lists, multi-group rows, tiles that straddle
    grid rows)."""
    cfg, cols = scenes.scaled("config4", 300_000)
    cfg["physics"]["maxCollisionPairs"] = 1_500_000
    run_and_compare(cfg, cols, 3, flags=B.FLAG_K6_TILE)


def test_collapsed_bed_every_row_capped():
    """The state the balls scenes settle into: far more than maxNeighbors candidates in range of everybody.
    Every row is capped, most entities have lower-id partners past their cap (the extended internal rows)
    and the densest ones more than those hold (F_XOVER: the sweep resumes the scan).  Small cap, dense
    start, a few frames: state, rows and pairs against the oracle."""
    cfg, cols = scenes.balls_synthetic(60_000, (640.0, 640.0), 16.0, 12, 2, (2.0, 5.0), 16.0, seed=5)
    cfg["physics"]["maxCollisionPairs"] = 3_000_000
    st = run_and_compare(cfg, cols, 4)
    assert st["cappedRows"] > 50_000


def test_piles_of_hundreds_to_thousands_in_one_cell():
    """What a settled bed does to the grid (config 4 after 300 frames: cells of 500-1900 entities).  Cell lists
    above 128 entities are sorted by a block and ranked by binary search, above 4096 they keep the linear
    count; with most rows capped the lower-id partners past a cap are searched cell by cell, skipping the
    cells whose rows all closed before the searcher's slot.  Piles of 200, 1500 and 5000 entities in single
    cells on top of a dense uniform scene, small cap: state, rows and pairs against the oracle."""
    cfg, cols = scenes.balls_synthetic(26_700, (320.0, 320.0), 16.0, 12, 2, (2.0, 5.0), 16.0, seed=11)
    rng = np.random.Generator(np.random.PCG64(3))
    at = 20_001
    for n, (cx, cy) in ((200, (3, 5)), (1500, (10, 10)), (5000, (17, 19))):
        sl = slice(at, at + n)
        cols["T.x"][sl] = (cx * 16.0 + 0.25 + rng.random(n) * 15.5).astype(np.float32)
        cols["T.y"][sl] = (cy * 16.0 + 0.25 + rng.random(n) * 15.5).astype(np.float32)
        cols["RB.px"][sl] = cols["T.x"][sl]
        cols["RB.py"][sl] = cols["T.y"][sl]
        at += n
    cfg["physics"]["maxCollisionPairs"] = 3_000_000
    st = run_and_compare(cfg, cols, 3)
    assert st["cappedRows"] > 20_000 and st["maxCellOccupancy"] > 4096


@pytest.mark.parametrize("flags,M,vr", [(B.FLAG_K4_WIDE, 24, 16.0), (B.FLAG_K4_THREAD, 300, 40.0)], ids=["wide-short-rows", "thread-long-rows"])
def test_both_scan_forms_on_the_other_side_of_their_threshold(flags, M, vr):
    """The library picks the warp-per-entity scan for maxNeighbors >= 256 and the thread-per-entity one
    below; forcing each onto the other's ground (capped rows, rows of several hundred entries walked 64 at
    a time by the sweep, windows of 7 x 7 cells) must not change a bit."""
    cfg, cols = scenes.balls_synthetic(30_000, (800.0, 800.0), 16.0, M, 2, (2.0, 5.0), vr, seed=17)
    cfg["physics"]["maxCollisionPairs"] = 2_000_000
    st = run_and_compare(cfg, cols, 3, flags=flags)
    assert st["neighborsTotal"] > 0


def test_tiled_sweep_with_four_substeps():
    cfg, cols = scenes.scaled("config5", 300_000)            # S = 4: FIRST / middle / LAST instantiations
    run_and_compare(cfg, cols, 2, flags=B.FLAG_K6_TILE)
    run_and_compare(cfg, cols, 2)


@pytest.mark.slow
def test_config4_full_16M_2_frames_bit_exact():
    """The benchmark scene itself.  Rows are compared on three blocks of 200 000 entities (the full
    rows would need 2 x 8.3 GB of host memory twice); state and collisionData in full."""
    cfg, cols = scenes.config4()
    cfg["physics"]["maxCollisionPairs"] = 40_000_000
    eng = make_engine(cfg, cols, host_neighbor_rows=False)
    ora = make_oracle(OracleC, cfg, cols)
    N, M = cfg["entityCount"], cfg["spatial"]["maxNeighbors"]
    stride = 1 + M
    for f in range(2):
        eng.run(1)
        eng.download(B.COLS_INPUT_ALL | B.COL_COLLISIONS)
        ora.step(1.0, 1)
        compare_state(eng, ora, f"frame {f}")
        cellOf, _, _ = ora.grid_csr()
        for first in (0, N // 2, N - 200_000):
            nd = np.zeros(200_000 * stride, np.int32)
            dd = np.zeros(200_000 * stride, np.float32)
            B.check(eng.ctx, B.lib().weed_fetch_neighbors_to(eng.ctx, first, 200_000, nd.ctypes.data, dd.ctypes.data))
            rows = np.nonzero(cellOf[first:first + 200_000] >= 0)[0]
            o = first * stride
            active_rows_equal(nd, dd, ora.neighborData[o:o + 200_000 * stride], ora.distanceData[o:o + 200_000 * stride],
                              200_000, M, rows)
    eng.close()
