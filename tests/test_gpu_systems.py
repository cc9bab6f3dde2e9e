"""Device-side consumers against the CPU restatements, through the C ABI.

Bit-exact for the Enter/Stay/Exit stream, isItOnScreen, screenX/Y and every shadow-sprite field
except rotation (device atan2 vs libm atan2: within 1 float32 ulp)."""
import numpy as np
import pytest

from multithreadedgameengine_b200 import binding as B, scenes
from oracle.oracle_c import CollisionEventsC, screen_visibility_c, shadow_sprites_c
from test_gpu_parity import make_engine, ulp_diff

pytestmark = pytest.mark.gpu


def dense_scene(seed=3, n=4000):
    cfg, cols = scenes.balls_synthetic(n, (1600.0, 800.0), 24.0, 48, 2, (4.0, 9.0), 24.0, seed=seed)
    cfg["physics"]["maxCollisionPairs"] = 200000
    return cfg, cols


def test_collision_events_stream_matches_reference_sequence():
    cfg, cols = dense_scene()
    eng = make_engine(cfg, cols)
    ora = CollisionEventsC()
    seen = set()
    for frame in range(12):
        eng.step(1.0, 0, B.COL_COLLISIONS)
        ev = eng.collision_events()
        want = ora.process(eng.collisionData)
        got = eng.collision_callbacks(ev)
        assert got == want, f"frame {frame}: {len(got)} vs {len(want)} calls"
        assert ev["entered"] + ev["stayed"] == len(ev["pairs"])
        seen |= {t for t, _, _ in got}
    assert seen == {1, 2, 3}
    # forgetting the previous frame turns every pair into Enter and produces no Exit
    eng.step(1.0, 0, B.COL_COLLISIONS)
    ev = eng.collision_events(forget_previous=True)
    assert ev["stayed"] == 0 and ev["exited"] == 0 and (ev["state"] == 1).all()
    eng.close()


def test_collision_events_respect_the_pair_cap():
    """A truncated collisionData is what the reference's logic worker sees too: pairs beyond
    maxCollisionPairs do not exist for Enter/Stay/Exit."""
    cfg, cols = dense_scene(seed=5)
    cfg["physics"]["maxCollisionPairs"] = 300
    eng = make_engine(cfg, cols)
    ora = CollisionEventsC()
    for frame in range(5):
        eng.step(1.0, 0, B.COL_COLLISIONS)
        assert int(eng.collisionData[0]) == 300
        assert eng.collision_callbacks(eng.collision_events()) == ora.process(eng.collisionData), frame
    eng.close()


def test_screen_visibility_and_shadows_match_oracle():
    cfg, cols = dense_scene(seed=9, n=6000)
    N, M = cfg["entityCount"], cfg["spatial"]["maxNeighbors"]
    rng = np.random.default_rng(2)
    cols["T.active"][rng.random(N) < 0.05] = 0
    eng = make_engine(cfg, cols)
    sx = np.zeros(N, np.float32); sy = np.zeros(N, np.float32); on = np.zeros(N, np.uint8)
    light = (rng.random(N) < 0.02).astype(np.uint8)
    inten = rng.choice([0.0, 120.0, 800.0, np.nan], N).astype(np.float32)     # NaN is not <= 0: such a light is processed (:918)
    caster = (rng.random(N) < 0.7).astype(np.uint8)
    rad = rng.choice([0.0, 6.0, 11.5], N).astype(np.float32)
    hgt = rng.choice([0.0, 25.0, 40.0], N).astype(np.float32)
    eng.shadows_upload(light, inten, caster, rad, hgt)
    for frame in range(3):
        eng.step(1.0, 0, B.COLS_INPUT_ALL | B.COL_NEIGHBORS)
        cam = (1.25, 100.0 + 40 * frame, 60.0, 1280.0, 720.0)
        gx, gy, gon = eng.screen_visibility(*cam)
        screen_visibility_c(eng.col["T.active"], eng.col["T.x"], eng.col["T.y"], *cam, sx, sy, on)
        assert np.array_equal(gon, on) and 0 < on.sum() < N
        assert np.array_equal(gx.view(np.uint32), sx.view(np.uint32)) and np.array_equal(gy.view(np.uint32), sy.view(np.uint32))
        for caps in ((20, 15, None), (4, 15, None), (20, 3, None), (20, 15, 11)):
            got = eng.shadows(*caps)
            want = shadow_sprites_c(M, eng.neighborData, eng.distanceData, eng.col["T.active"], eng.col["T.x"], eng.col["T.y"],
                                    light, inten, caster, rad, hgt, on, *caps)
            n = want["count"]
            assert got["count"] == n and n > 0, (frame, caps)
            assert np.array_equal(got["active"], want["active"])
            for k in ("radius", "x", "y", "scaleX", "scaleY", "alpha"):
                a, b = got[k][:n], want[k][:n]
                same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))   # NaN payloads differ CPU/GPU
                assert same.all(), (k, frame, caps)
            assert (ulp_diff(got["rotation"][:n], want["rotation"][:n]) <= 1).all()
    eng.close()


def test_systems_report_state_errors():
    cfg, cols = dense_scene(n=500)
    eng = make_engine(cfg, cols)
    with pytest.raises(B.WeedError):
        eng.shadows()            # no columns uploaded, no visibility pass yet
    eng.close()


def test_pools_spawn_and_despawn_like_the_reference():
    """batches of GameObject.spawn / despawn / despawnAll on the device against the
    sequential CPU restatement — same indices, same columns, same free-list order — and the
    simulation keeps matching the oracle afterwards."""
    from helpers import make_oracle
    from oracle.oracle_c import OracleC, PoolC
    from test_gpu_parity import ALL_DL, compare_state

    cfg, cols = scenes.balls_synthetic(3000, (1200.0, 600.0), 24.0, 48, 2, (4.0, 9.0), 24.0, seed=17)
    N = cfg["entityCount"]
    # entities 1001..3000 form one class that starts inactive (the engine spawns them later)
    for k in ("T.active", "RB.active", "C.active"):
        cols[k][1001:] = 0
    eng = make_engine(cfg, cols)
    ora = make_oracle(OracleC, cfg, cols)
    pool = eng.create_pool(1001, 2000)
    opool = PoolC(ora.col, 1001, 2000)
    rng = np.random.default_rng(23)
    live = []
    for step in range(10):
        rec = np.column_stack([rng.uniform(20, 1180, 150), rng.uniform(20, 580, 150), rng.normal(0, 2, 150),
                               rng.normal(0, 2, 150)]).astype(np.float32)
        got, want = eng.spawn(pool, rec), opool.spawn(rec)
        assert np.array_equal(got, want), step
        live += got.tolist()
        if step % 3 == 2:
            pick = rng.choice(live, 120).tolist()          # repeats and already-despawned entities
            assert eng.despawn(pool, pick) == opool.despawn(pick)
            live = [i for i in live if i not in set(pick)]
        assert eng.pool_stats(pool)["available"] == opool.available()
        eng.step(1.0, 0, ALL_DL)
        ora.step(1.0, 1)
        compare_state(eng, ora, f"step {step}")
    assert eng.despawn_all(pool) == opool.despawn_all() > 0
    assert eng.pool_stats(pool) == {"total": 2000, "available": 2000, "active": 0}
    # drain: the whole free-list order must agree, then the pool is exhausted
    z = np.zeros((2005, 4), np.float32)
    got, want = eng.spawn(pool, z), opool.spawn(z)
    assert np.array_equal(got, want) and (got[-5:] == -1).all() and (got[:2000] >= 1001).all()
    with pytest.raises(B.WeedError):
        eng.despawn(pool, [5])                               # not an index of this pool
    eng.close()
