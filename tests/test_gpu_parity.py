"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star / SURVEY Appendix A):
  * grid assignment and neighbor rows: BIT-EXACT (ids, order, cap, float32 d² bits);
  * physics state vs the oracle's J-order mode (the documented deterministic resolution
    order of the GPU): bit-exact for x, y, px, py, vx, vy, ax, ay, speed, collisionCount and
    collisionData; velocityAngle within 1 float32 ulp (device atan2 vs libm atan2);
  * vs the reference's sequential Gauss-Seidel order: statistical (tests/test_gpu_longrun.py).
"""
import numpy as np
import pytest

from helpers import active_rows_equal, assert_cols_equal, bits, make_oracle, random_scene
from multithreadedgameengine_b200 import binding as B, scenes
from oracle.oracle_c import OracleC

pytestmark = pytest.mark.gpu

EXACT = ["T.x", "T.y", "RB.px", "RB.py", "RB.vx", "RB.vy", "RB.ax", "RB.ay", "RB.speed", "RB.collisionCount"]
ALL_DL = B.COLS_INPUT_ALL | B.COL_NEIGHBORS | B.COL_COLLISIONS


def make_engine(cfg, cols, **kw):
    from multithreadedgameengine_b200.engine import GameEngine
    eng = GameEngine(cfg, **kw)
    eng.load_columns(cols)
    return eng


def ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def compare_state(eng, ora, what=""):
    assert_cols_equal(eng.col, ora.col, EXACT, what)
    for k in ("T.active", "RB.active", "RB.static", "C.active", "C.isTrigger", "RB.maxVel", "C.radius", "C.visualRange"):
        assert np.array_equal(bits(eng.col[k]), bits(ora.col[k])), k
    a, b = eng.col["RB.velocityAngle"], ora.col["RB.velocityAngle"]
    ok = (ulp_diff(a, b) <= 1) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), f"{what} velocityAngle beyond 1 ulp at {np.nonzero(~ok)[0][:5]}"
    n = int(ora.collisionData[0])
    assert int(eng.collisionData[0]) == n, f"{what} pairCount {int(eng.collisionData[0])} vs {n}"
    assert np.array_equal(eng.collisionData[1:1 + 2 * n], ora.collisionData[1:1 + 2 * n]), f"{what} collision pairs"


def compare_rows(eng, ora, cfg):
    cellOf, start, idx = ora.grid_csr()
    rows = np.nonzero(cellOf >= 0)[0]
    N, M = cfg["entityCount"], cfg["spatial"]["maxNeighbors"]
    active_rows_equal(eng.neighborData, eng.distanceData, ora.neighborData, ora.distanceData, N, M, rows)
    return cellOf


@pytest.mark.parametrize("seed", range(8))
def test_neighbor_rows_bit_exact_adversarial(seed):
    rng = np.random.default_rng(100 + seed)
    cs = [50.0, 33.3, 30.0, 66.5, 128.0, 17.0, 80.0, 100.0][seed]
    M = [16, 8, 400, 5, 32, 12, 1, 64][seed]
    cfg, cols = random_scene(rng, N=700, cellSize=cs, M=M)
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    eng.spatial.update()
    ora.spatial()
    eng.download(B.COL_NEIGHBORS)
    cellOf = compare_rows(eng, ora, cfg)
    # rows of entities that are not in the grid are never written
    stride = 1 + M
    nd = eng.neighborData.reshape(-1, stride)
    assert not nd[cellOf < 0].any()
    # words past 1+count keep their old contents (zero here)
    for i in np.nonzero(cellOf >= 0)[0]:
        assert not nd[i, 1 + nd[i, 0]:].any()
    s = eng.stats()
    assert s["activeInGrid"] == int((cellOf >= 0).sum())
    assert (s["gridCols"], s["gridRows"]) == (ora.gridCols, ora.gridRows)
    cnt = np.bincount(cellOf[cellOf >= 0])
    assert s["maxCellOccupancy"] == cnt.max()
    assert s["neighborsTotal"] == int(nd[cellOf >= 0, 0].sum())
    eng.close()


@pytest.mark.parametrize("seed", range(6))
def test_frames_bit_exact_vs_jorder_oracle(seed):
    rng = np.random.default_rng(200 + seed)
    cs = [50.0, 33.3, 30.0, 66.5, 128.0, 17.0][seed]
    cfg, cols = random_scene(rng, N=900, cellSize=cs, M=[16, 8, 400, 5, 32, 12][seed], S=[2, 1, 3, 2, 4, 2][seed])
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    for frame in range(5):
        dt = [1.0, 0.73, 1.0, 1.31, 1.0][frame]
        eng.step(dt, 0, ALL_DL)
        ora.step(dt, 1)
        compare_rows(eng, ora, cfg)
        compare_state(eng, ora, f"seed {seed} frame {frame}")
    eng.close()


def test_split_workers_equal_fused_step():
    rng = np.random.default_rng(7)
    cfg, cols = random_scene(rng, N=800, M=24)
    a = make_engine(cfg, cols)
    b = make_engine(cfg, cols)
    for _ in range(3):
        a.step(1.0, 0, ALL_DL)
        b.spatial.update()
        b.physics_worker.update(16.67, 1.0)
        b.download(ALL_DL)
        assert_cols_equal(a.col, b.col)
        assert np.array_equal(a.neighborData, b.neighborData)
        assert np.array_equal(bits(a.distanceData), bits(b.distanceData))
        n = int(a.collisionData[0])
        assert np.array_equal(a.collisionData[:1 + 2 * n], b.collisionData[:1 + 2 * n])
    with pytest.raises(B.WeedError):
        b.physics_worker.update(16.67, 1.0)   # needs the rows of a preceding spatial update
    a.close(); b.close()


def test_split_workers_with_a_host_tick_in_between():
    """The flow that swaps ONE worker: weed_spatial, a host tick() that reads the fresh rows and writes
    ax / ay, weed_upload(ax | ay), weed_physics.  Uploading accelerations must not invalidate the rows
    (ADVICE r1); uploading a position must."""
    rng = np.random.default_rng(11)
    cfg, cols = random_scene(rng, N=900, M=24)
    a = make_engine(cfg, cols)
    b = make_engine(cfg, cols)
    acc = a.mask("RB.ax", "RB.ay")
    for f in range(3):
        tick = np.random.default_rng(50 + f)
        ax = ((tick.random(cfg["entityCount"]) - 0.5) * 2).astype(np.float32)
        ay = ((tick.random(cfg["entityCount"]) - 0.5) * 2).astype(np.float32)
        a.col["RB.ax"][:] = ax; a.col["RB.ay"][:] = ay
        a.step(1.0, acc, ALL_DL)                         # fused: upload, spatial, physics
        b.spatial.update()
        b.fetch_neighbors()                              # what tick() would read
        b.col["RB.ax"][:] = ax; b.col["RB.ay"][:] = ay
        b.upload(acc)
        b.physics_worker.update(16.67, 1.0)
        b.download(ALL_DL)
        assert_cols_equal(a.col, b.col)
        assert np.array_equal(a.neighborData, b.neighborData)
        n = int(a.collisionData[0])
        assert np.array_equal(a.collisionData[:1 + 2 * n], b.collisionData[:1 + 2 * n])
    b.spatial.update()
    b.upload(b.mask("T.x"))                              # a teleport: the grid is stale now
    with pytest.raises(B.WeedError):
        b.physics_worker.update(16.67, 1.0)
    a.close(); b.close()


def test_config1_readme_scene_30_frames():
    cfg, cols = scenes.balls_readme()
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    for frame in range(30):
        eng.step(1.0, 0, ALL_DL)
        ora.step(1.0, 1)
        compare_state(eng, ora, f"frame {frame}")
    compare_rows(eng, ora, cfg)
    assert int(ora.collisionData[0]) > 0
    eng.close()


def test_config2_boids_heterogeneous_ranges():
    """Mixed visual ranges make rows asymmetric: pairs the higher-id side cannot infer go
    through the explicit list."""
    cfg, cols = scenes.boids(n_prey=3000, n_pred=150, seed=11)
    cfg["worldWidth"], cfg["worldHeight"] = 1800.0, 900.0
    for k in ("T.x", "RB.px"):
        cols[k] = (cols[k] * np.float32(1800.0 / 5000.0)).astype(np.float32)
    for k in ("T.y", "RB.py"):
        cols[k] = (cols[k] * np.float32(900.0 / 2000.0)).astype(np.float32)
    cfg["spatial"]["maxNeighbors"] = 300
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    for frame in range(6):
        eng.step(1.0, 0, ALL_DL)
        ora.step(1.0, 1)
        compare_state(eng, ora, f"frame {frame}")
    compare_rows(eng, ora, cfg)
    assert eng.stats()["explicitPairs"] > 0
    eng.close()


def test_capped_rows_take_the_explicit_path():
    cfg, cols = scenes.balls_synthetic(3000, (600.0, 300.0), 16.0, 4, 2, (2.0, 6.0), 16.0, seed=5)
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    for frame in range(6):
        eng.step(1.0, 0, ALL_DL)
        ora.step(1.0, 1)
        compare_state(eng, ora, f"frame {frame}")
        compare_rows(eng, ora, cfg)
    s = eng.stats()
    assert s["cappedRows"] > 0 and s["explicitPairs"] > 0
    eng.close()


def test_coincident_after_bounds_uses_hash_nudge():
    """Two out-of-world balls of equal radius are clamped into the same corner by the boundary
    pass: distance exactly 0, the documented hash nudge (include/weed_nudge.h) applies."""
    cfg, cols = scenes.balls_readme(n_balls=6, seed=2, world=(400.0, 300.0))
    for i, (x, y) in enumerate([(-30.0, -20.0), (-12.0, -25.0), (500.0, 400.0), (450.0, 380.0), (200.0, 100.0), (203.0, 100.0)], start=1):
        cols["T.x"][i] = cols["RB.px"][i] = x
        cols["T.y"][i] = cols["RB.py"][i] = y
        cols["C.radius"][i] = 10.0
    cols["RB.static"][4] = 1
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    for frame in range(4):
        eng.step(1.0, 0, ALL_DL)
        ora.step(1.0, 1)
        compare_state(eng, ora, f"frame {frame}")
    assert eng.col["T.x"][1] != eng.col["T.x"][2] or eng.col["T.y"][1] != eng.col["T.y"][2]
    eng.close()


def test_update_physics_config_and_validation():
    cfg, cols = scenes.balls_readme(n_balls=500, seed=9, world=(900.0, 500.0))
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    eng.step(1.0, 0, ALL_DL); ora.step(1.0, 1)
    new = dict(subStepCount=0, boundaryElasticity=1.7, collisionResponseStrength=-0.2, verletDamping=0.9,
               minSpeedForRotation=0.5, gravity=dict(x=0.3, y=-0.1))
    eng.updatePhysicsConfig(new)
    st = eng.physics_settings()      # validatePhysicsConfig clamps (utils.js:269-301)
    assert (st["subStepCount"], st["boundaryElasticity"], st["collisionResponseStrength"]) == (1, 1.0, 0.0)
    ora.set_physics(0, 1.7, -0.2, 0.9, 0.5, 0.3, -0.1)
    for frame in range(3):
        eng.step(1.0, 0, ALL_DL); ora.step(1.0, 1)
        compare_state(eng, ora, f"after reconfig frame {frame}")
    eng.close()


def test_host_written_columns_and_partial_row_fetch():
    """tick()-style use: the host writes ax/ay (and teleports an entity), uploads just those
    columns, and reads single neighbor rows back (gameObject.js:700-729)."""
    cfg, cols = scenes.balls_readme(n_balls=600, seed=4, world=(1000.0, 500.0))
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    rng = np.random.default_rng(0)
    up = eng.mask("RB.ax", "RB.ay", "T.x", "T.y", "RB.px", "RB.py")
    for frame in range(4):
        ax = ((rng.random(601) - 0.5) * 2).astype(np.float32)
        ay = ((rng.random(601) - 0.5) * 2).astype(np.float32)
        for tgt in (eng.col, ora.col):
            tgt["RB.ax"][:] = ax
            tgt["RB.ay"][:] = ay
            tgt["T.x"][0] = 100.0 + 50 * frame       # the Mouse entity moves every frame
            tgt["T.y"][0] = 120.0
            tgt["T.x"][17] = tgt["RB.px"][17] = 333.0  # teleport: x setter also sets px (gameObject.js:230-237)
        eng.step(1.0, up, B.COLS_OUTPUT_ALL | B.COL_COLLISIONS)
        ora.step(1.0, 1)
        compare_state(eng, ora, f"frame {frame}")
        stride = 1 + eng.maxNeighbors
        for i in (0, 17, 300, 600):
            ids, d2 = eng.neighbors_of(i)
            n = int(ora.neighborData[i * stride])
            assert np.array_equal(ids, ora.neighborData[i * stride + 1:i * stride + 1 + n])
            assert np.array_equal(bits(d2), bits(ora.distanceData[i * stride + 1:i * stride + 1 + n]))
    eng.close()


def test_graph_and_direct_launch_agree():
    cfg, cols = scenes.balls_readme(n_balls=700, seed=6, world=(1000.0, 500.0))
    a = make_engine(cfg, cols)
    b = make_engine(cfg, cols, flags=B.FLAG_NO_GRAPH)
    c = make_engine(cfg, cols, flags=B.FLAG_KERNEL_TIMING)
    a.run(5); b.run(5); c.run(5)
    for e in (a, b, c):
        e.download(ALL_DL)
    assert_cols_equal(a.col, b.col); assert_cols_equal(a.col, c.col)
    assert np.array_equal(a.neighborData, b.neighborData) and np.array_equal(a.neighborData, c.neighborData)
    assert sum(c.stats()["ms"][:8]) > 0
    for e in (a, b, c):       # the device-clock frame time needs no flag
        assert 0 < e.stats()["ms"][8] < 50
    for e in (a, b, c):
        e.close()


def test_pipelined_step_equals_plain_step():
    """Above 2^18 entities weed_step overlaps the ax/ay upload with K1-K3 and the early columns'
    download with K4-K7 on a second stream; the host must see exactly what the plain path gives."""
    cfg, cols = scenes.balls_synthetic(300_000, (8192.0, 4096.0), 16.0, 32, 2, (2.0, 5.0), 16.0, seed=21)
    a = make_engine(cfg, cols, host_neighbor_rows=False)                       # pipelined
    b = make_engine(cfg, cols, host_neighbor_rows=False, flags=B.FLAG_NO_GRAPH)  # plain, direct launches
    rng = np.random.default_rng(1)
    up = a.mask("RB.ax", "RB.ay")
    dl = B.COLS_OUTPUT_ALL | B.COL_COLLISIONS
    for frame in range(4):
        ax = rng.normal(0, 0.3, cfg["entityCount"]).astype(np.float32)
        for e in (a, b):
            e.col["RB.ax"][:] = ax
            e.col["RB.ay"][:] = -ax
            e.step(1.0, up, dl)
        assert_cols_equal(a.col, b.col, what=f"frame {frame}")
        n = int(a.collisionData[0])
        assert n == int(b.collisionData[0]) and np.array_equal(a.collisionData[:1 + 2 * n], b.collisionData[:1 + 2 * n])
        assert not a.col["RB.ax"][1:].any()      # balls: zeroed by the integration, downloaded early
    # a full upload (spawn / teleport) takes the ordered path of the same function
    for e in (a, b):
        e.col["T.x"][5] = 100.0; e.col["RB.px"][5] = 100.0
        e.step(1.0, B.COLS_INPUT_ALL, dl)
    assert_cols_equal(a.col, b.col, what="after full upload")
    a.close(); b.close()


def test_errors_are_codes_not_crashes():
    from multithreadedgameengine_b200.engine import GameEngine
    cfg, cols = scenes.balls_readme(n_balls=50, seed=1, world=(400.0, 300.0))
    eng = GameEngine(cfg)
    L = B.lib()
    small = np.zeros(16, dtype=np.uint8)
    assert L.weed_bind(eng.ctx, B.BUF_TRANSFORM, small.ctypes.data, small.nbytes) == B.WEED_E_SIZE
    assert L.weed_bind(eng.ctx, 99, small.ctypes.data, small.nbytes) == B.WEED_E_INVALID
    assert L.weed_fetch_neighbors(eng.ctx, 40, 100) == B.WEED_E_INVALID
    assert L.weed_physics(eng.ctx, 1.0) == B.WEED_E_STATE
    assert L.weed_bind(eng.ctx, B.BUF_COLLISION, None, 0) == B.WEED_OK
    assert L.weed_download(eng.ctx, B.COL_COLLISIONS) == B.WEED_E_NOT_BOUND
    eng.close()


# ---- hand-derived micro-cases (tests/golden/micro_cases.json) on the GPU ------------------------
from test_golden_cases import CASES, build_case, check_case  # noqa: E402


class _EngineAsSim:
    """Adapter: drive a GameEngine like an oracle object for check_case()."""

    def __init__(self, eng):
        self.eng = eng

    def spatial(self):
        self.eng.spatial.update()
        self.eng.download(B.COL_NEIGHBORS)

    def step(self, dt, order):
        self.eng.step(dt, 0, ALL_DL)

    col = property(lambda self: self.eng.col)
    neighborData = property(lambda self: self.eng.neighborData)
    distanceData = property(lambda self: self.eng.distanceData)
    collisionData = property(lambda self: self.eng.collisionData)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_micro_case_on_gpu(case):
    cfg, cols, _ = build_case(case)
    cfg["physics"]["maxCollisionPairs"] = 100
    eng = make_engine(cfg, cols)
    check_case(case, _EngineAsSim(eng), order=1)
    eng.close()


# ---- full-size properties (BASELINE config 3: 1M entities) -------------------------------------
@pytest.fixture(scope="module")
def big():
    cfg, cols = scenes.config3()
    eng = make_engine(cfg, cols)
    yield cfg, cols, eng
    eng.close()


def test_full_size_rows_match_bruteforce_sample(big):
    """At 1M entities the oracle is too slow to run in a test; instead 300 sampled rows are
    checked against a brute-force float64 evaluation of spatial_worker.js:211-270 over ALL
    entities (independent of any grid)."""
    cfg, cols, eng = big
    eng.spatial.update()
    N, M = cfg["entityCount"], cfg["spatial"]["maxNeighbors"]
    cs = cfg["spatial"]["cellSize"]
    inv = 1.0 / cs
    cols_n, rows_n = int(np.ceil(cfg["worldWidth"] / cs)), int(np.ceil(cfg["worldHeight"] / cs))
    x = cols["T.x"].astype(np.float64)
    y = cols["T.y"].astype(np.float64)
    col = np.clip(np.trunc(x * inv).astype(np.int64), 0, cols_n - 1)
    row = np.clip(np.trunc(y * inv).astype(np.int64), 0, rows_n - 1)
    order_key = (row * cols_n + col) * (N + 1) + np.arange(N)     # (cell, id) lexicographic
    rng = np.random.default_rng(0)
    sample = np.concatenate([[0, 1, N - 1], rng.integers(0, N, 300)])
    stride = 1 + M
    for i in sample:
        ids, d2 = eng.neighbors_of(int(i))
        vr = float(cols["C.visualRange"][i])
        cr = int(np.ceil(vr * inv))
        c0, r0 = int(np.trunc(x[i] * inv)), int(np.trunc(y[i] * inv))
        dx, dy = x - x[i], y - y[i]
        dist2 = dx * dx + dy * dy
        ok = (dist2 < vr * vr) & (dist2 > 0) & (np.abs(row - r0) <= cr) & (np.abs(col - c0) <= cr)
        cand = np.nonzero(ok)[0]
        cand = cand[np.argsort(order_key[cand], kind="stable")][:M]
        assert np.array_equal(ids, cand.astype(np.int32)), f"row {i}"
        assert np.array_equal(bits(d2), bits(dist2[cand].astype(np.float32))), f"row {i} d2"
    s = eng.stats()
    assert s["activeInGrid"] == N and s["neighborsTotal"] > 0


def test_full_size_spatial_is_idempotent_and_step_keeps_invariants(big):
    cfg, cols, eng = big
    N, M = cfg["entityCount"], cfg["spatial"]["maxNeighbors"]
    eng.load_columns(cols)
    eng.spatial.update()
    eng.download(B.COL_NEIGHBORS)
    first = eng.neighborData.copy(), eng.distanceData.copy()
    eng.spatial.update()
    eng.download(B.COL_NEIGHBORS)
    assert np.array_equal(first[0], eng.neighborData) and np.array_equal(bits(first[1]), bits(eng.distanceData))
    nd = eng.neighborData.reshape(N, 1 + M)
    assert nd[:, 0].min() >= 0 and nd[:, 0].max() <= M
    # listed ids are valid, never the entity itself
    k = np.arange(M)[None, :] < nd[:, :1]
    assert (nd[:, 1:][k] >= 0).all() and (nd[:, 1:][k] < N).all()
    assert not (nd[:, 1:] == np.arange(N)[:, None])[k].any()
    before = {key: eng.col[key].copy() for key in ("T.x", "T.y")}
    for _ in range(3):
        eng.step(1.0, 0, B.COLS_OUTPUT_ALL | B.COL_COLLISIONS)
    c = eng.col
    assert np.isfinite(c["T.x"]).all() and np.isfinite(c["T.y"]).all()
    # vx == fround((x_new - x_old)/dtRatio) before constraints; entities that never collided
    # and stayed off the walls moved exactly by the integration
    n = int(eng.collisionData[0])
    assert 0 < n <= eng.maxCollisionPairs
    pairs = eng.collisionData[1:1 + 2 * n].reshape(-1, 2)
    assert (pairs[:, 0] < pairs[:, 1]).all()                       # i < j (physics_worker.js:444)
    assert (np.diff(pairs[:, 0]) >= 0).all()                       # sweep order: i ascending
    assert (c["RB.ax"] == 0).all() and (c["RB.ay"] == 0).all()     # :313-314
    assert (c["RB.speed"] >= 0).all()
    assert not np.array_equal(before["T.y"], c["T.y"])


def test_config2_boids_with_host_tick_consumer():
    """BASELINE config 2 as the reference runs it: neighbor rows are fetched to the host,
    tick() (examples/boids_tick.py, a restatement of demos/predators/boid.js) writes ax/ay,
    the next frame uploads exactly those two columns."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    from boids_tick import tick_all

    cfg, cols = scenes.boids(n_prey=1200, n_pred=60, seed=21)
    cfg["worldWidth"], cfg["worldHeight"] = 1400.0, 700.0
    for k, f in (("T.x", 1400 / 5000), ("RB.px", 1400 / 5000), ("T.y", 700 / 2000), ("RB.py", 700 / 2000)):
        cols[k] = (cols[k] * np.float32(f)).astype(np.float32)
    cfg["spatial"]["maxNeighbors"] = 200
    N, M = cfg["entityCount"], 200
    etype = np.zeros(N, dtype=np.uint8)
    etype[1:1201] = 1
    etype[1201:] = 2
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    eng.Transform.entityType[:] = etype
    up = eng.mask("RB.ax", "RB.ay")
    for frame in range(5):
        eng.step(1.0, up if frame else 0, B.COLS_OUTPUT_ALL | B.COL_NEIGHBORS | B.COL_COLLISIONS)
        ora.step(1.0, 1)
        compare_state(eng, ora, f"frame {frame}")
        compare_rows(eng, ora, cfg)
        tick_all(eng.col, etype, eng.neighborData, eng.distanceData, M, cfg["worldWidth"], cfg["worldHeight"])
        tick_all(ora.col, etype, ora.neighborData, ora.distanceData, M, cfg["worldWidth"], cfg["worldHeight"])
        assert np.array_equal(bits(eng.col["RB.ax"]), bits(ora.col["RB.ax"]))
        assert np.abs(eng.col["RB.ax"]).max() > 0
    eng.close()


def test_no_neighbor_rows_flag_and_device_pointers():
    """Physics-only consumers can skip the API rows (WEED_FLAG_NO_NEIGHBOR_ROWS): the state must
    not change, row fetches must fail with a code; device pointers are exposed for in-process
    consumers."""
    import ctypes as C
    cfg, cols = scenes.scaled("config4", 20000)
    a = make_engine(cfg, cols)
    b = make_engine(cfg, cols, flags=B.FLAG_NO_NEIGHBOR_ROWS, host_neighbor_rows=False)
    a.run(4); b.run(4)
    a.download(B.COLS_INPUT_ALL | B.COL_COLLISIONS); b.download(B.COLS_INPUT_ALL | B.COL_COLLISIONS)
    assert_cols_equal(a.col, b.col)
    n = int(a.collisionData[0])
    assert n == int(b.collisionData[0]) and np.array_equal(a.collisionData[:1 + 2 * n], b.collisionData[:1 + 2 * n])
    assert B.lib().weed_fetch_neighbors(b.ctx, 0, 1) == B.WEED_E_STATE
    ptr, nbytes = a.device_ptr(B.DEV_NEIGHBOR)
    pitch = B.lib().weed_row_pitch(a.ctx)                  # device rows are slot-major planes of `pitch` words
    assert pitch == (cfg["entityCount"] + 127) // 128 * 128 and ptr and nbytes == (a.maxNeighbors + 7) // 8 * 8 * pitch * 4
    for which in (B.DEV_NEIGHBOR_COUNT, B.DEV_SLOT_OF):
        ptr, nbytes = a.device_ptr(which)
        assert ptr and nbytes == cfg["entityCount"] * 4
    ptr, nbytes = a.device_ptr(B.DEV_STATE)
    assert ptr and nbytes == cfg["entityCount"] * 16
    assert b.device_ptr(B.DEV_NEIGHBOR)[0] is None
    a.close(); b.close()


def test_device_side_boids_system_matches_host_tick():
    """the boids tick() as a device-side system (weed_system_boids) must write the
    same RigidBody.ax/ay, bit for bit, as the host-side restatement of demos/predators/boid.js
    fed with the fetched neighbor rows — and the simulation driven by either must stay identical."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    from boids_tick import tick_all

    cfg, cols = scenes.boids(n_prey=1500, n_pred=80, seed=5)
    cfg["worldWidth"], cfg["worldHeight"] = 1500.0, 750.0
    for k, f in (("T.x", 1500 / 5000), ("RB.px", 1500 / 5000), ("T.y", 750 / 2000), ("RB.py", 750 / 2000)):
        cols[k] = (cols[k] * np.float32(f)).astype(np.float32)
    cfg["spatial"]["maxNeighbors"] = 160
    N, M = cfg["entityCount"], 160
    etype = np.zeros(N, dtype=np.uint8)
    etype[1:1501] = 1
    etype[1501:] = 2
    dev = make_engine(cfg, cols)       # tick() on the device
    host = make_engine(cfg, cols)      # tick() on the host, rows fetched
    for e in (dev, host):
        e.Transform.entityType[:] = etype
        e.upload(e.mask("T.entityType"))
    up = host.mask("RB.ax", "RB.ay")
    for frame in range(5):
        dev.step(1.0, 0, 0)                                      # nothing crosses PCIe
        host.step(1.0, up if frame else 0, B.COLS_OUTPUT_ALL | B.COL_NEIGHBORS)
        dev.system_boids(1.0)
        tick_all(host.col, etype, host.neighborData, host.distanceData, M, cfg["worldWidth"], cfg["worldHeight"])
        dev.download(B.COLS_OUTPUT_ALL)
        for k in ("RB.ax", "RB.ay", "T.x", "T.y", "RB.vx", "RB.vy"):
            assert np.array_equal(bits(dev.col[k]), bits(host.col[k])), f"frame {frame} {k}"
    assert np.abs(dev.col["RB.ax"]).max() > 0
    dev.close(); host.close()


def test_device_side_predators_demo_tick_matches_host_tick():
    """Config 2 complete (SURVEY §8 d): Prey flee, Predators hunt, the Mouse repels while its
    button is down — weed_system_flock against the host-side restatement of prey.js /
    predator.js / boid.js fed with fetched rows, bit for bit over several frames."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    from boids_tick import tick_classes

    cfg, cols = scenes.boids(n_prey=1500, n_pred=80, seed=9)
    cfg["worldWidth"], cfg["worldHeight"] = 1500.0, 750.0
    for k, f in (("T.x", 1500 / 5000), ("RB.px", 1500 / 5000), ("T.y", 750 / 2000), ("RB.py", 750 / 2000)):
        cols[k] = (cols[k] * np.float32(f)).astype(np.float32)
    for k, v in (("T.x", 700.0), ("RB.px", 700.0), ("T.y", 380.0), ("RB.py", 380.0)):
        cols[k][0] = v                       # the Mouse in the middle of the herd
    cfg["spatial"]["maxNeighbors"] = 160
    N, M = cfg["entityCount"], 160
    etype = np.zeros(N, dtype=np.uint8)
    etype[1:1501] = 1
    etype[1501:] = 2
    dev = make_engine(cfg, cols)
    host = make_engine(cfg, cols)
    for e in (dev, host):
        e.Transform.entityType[:] = etype
        e.upload(e.mask("T.entityType"))
    up = host.mask("RB.ax", "RB.ay")
    fled = hunted = 0
    for frame in range(6):
        down = frame >= 2                    # button pressed from the third frame on
        dev.step(1.0, 0, 0)
        host.step(1.0, up if frame else 0, B.COLS_OUTPUT_ALL | B.COL_NEIGHBORS)
        before = host.col["RB.ax"].copy()
        dev.system_flock(scenes.PREDATORS_DEMO_CLASSES, 1.0, mouseDown=down)
        tick_classes(host.col, etype, host.neighborData, host.distanceData, M, cfg["worldWidth"], cfg["worldHeight"],
                     scenes.PREDATORS_DEMO_CLASSES, 1.0, mouseDown=down)
        dev.download(B.COLS_OUTPUT_ALL)
        for k in ("RB.ax", "RB.ay", "T.x", "T.y", "RB.vx", "RB.vy"):
            assert np.array_equal(bits(dev.col[k]), bits(host.col[k])), f"frame {frame} {k}"
        fled += int((host.col["RB.ax"][1:1501] != before[1:1501]).sum())
        hunted += int((host.col["RB.ax"][1501:] != before[1501:]).sum())
    assert fled > 0 and hunted > 0
    dev.close(); host.close()


def test_full_size_config4_16M_properties():
    """BASELINE config 4 at its full size (16M entities, clustered, capped rows): sampled rows
    against a brute-force float64 evaluation over ALL entities, plus frame invariants.  Rows are
    fetched one at a time into scratch buffers (the full neighborData would be 8.3 GB)."""
    cfg, cols = scenes.config4()
    N, M = cfg["entityCount"], cfg["spatial"]["maxNeighbors"]
    eng = make_engine(cfg, cols, host_neighbor_rows=False)
    eng.spatial.update()
    cs = cfg["spatial"]["cellSize"]
    inv = 1.0 / cs
    cols_n, rows_n = int(np.ceil(cfg["worldWidth"] / cs)), int(np.ceil(cfg["worldHeight"] / cs))
    x = cols["T.x"].astype(np.float64)
    y = cols["T.y"].astype(np.float64)
    col = np.clip(np.trunc(x * inv).astype(np.int64), 0, cols_n - 1)
    row = np.clip(np.trunc(y * inv).astype(np.int64), 0, rows_n - 1)
    rng = np.random.default_rng(1)
    dense = np.argsort(np.bincount(row * cols_n + col, minlength=rows_n * cols_n)[row * cols_n + col])[-8:]   # entities of the fullest cells
    sample = np.concatenate([[0, N - 1], rng.integers(0, N, 30), dense])
    capped = 0
    for i in sample:
        ids, d2 = eng.neighbors_of(int(i))
        vr = float(cols["C.visualRange"][i])
        cr = int(np.ceil(vr * inv))
        c0, r0 = int(np.trunc(x[i] * inv)), int(np.trunc(y[i] * inv))
        near = np.nonzero((np.abs(row - r0) <= cr) & (np.abs(col - c0) <= cr))[0]
        dx, dy = x[near] - x[i], y[near] - y[i]
        dist2 = dx * dx + dy * dy
        ok = (dist2 < vr * vr) & (dist2 > 0)
        cand, cd2 = near[ok], dist2[ok]
        order = np.lexsort((cand, (row[cand] * cols_n + col[cand])))
        cand, cd2 = cand[order][:M], cd2[order][:M]
        assert np.array_equal(ids, cand.astype(np.int32)), f"row {i}"
        assert np.array_equal(bits(d2), bits(cd2.astype(np.float32))), f"row {i} d2"
        capped += len(ids) == M
    assert capped > 0                                  # the fullest cells hit the cap
    s = eng.stats()
    assert s["activeInGrid"] == N and s["cappedRows"] > 0
    eng.run(2)
    eng.download(B.COLS_OUTPUT_ALL | B.COL_COLLISIONS)
    c = eng.col
    assert np.isfinite(c["T.x"]).all() and np.isfinite(c["T.y"]).all()
    n = int(eng.collisionData[0])
    assert n == eng.maxCollisionPairs                  # 2e7 pairs exist, the log is capped at 10000
    pairs = eng.collisionData[1:1 + 2 * n].reshape(-1, 2)
    assert (pairs[:, 0] < pairs[:, 1]).all() and (np.diff(pairs[:, 0]) >= 0).all()
    assert (c["RB.ax"] == 0).all() and (c["RB.speed"] >= 0).all()
    eng.close()
