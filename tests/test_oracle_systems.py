"""The two CPU restatements of the output consumers agree with each other
and with hand-derived cases written from the reference source.

  f2  LogicWorker.processCollisionCallbacks       logic_worker.js:429-526
  f3  ParticleWorker.updateEntityScreenVisibility  particle_worker.js:1012-1062
      ParticleWorker.updateShadowSprites           particle_worker.js:861-1003
"""
import numpy as np

from oracle.oracle_c import CollisionEventsC, screen_visibility_c, shadow_sprites_c
from oracle.oracle_np import CollisionEventsNP, screen_visibility_np, shadow_sprites_np


def cd(pairs, maxPairs=64):
    a = np.zeros(1 + 2 * maxPairs, np.int32)
    a[0] = len(pairs)
    a[1:1 + 2 * len(pairs)] = np.asarray(pairs, np.int32).reshape(-1)
    return a


def test_enter_stay_exit_hand_derived():
    """Frame 1: (1,2) new -> Enter on both objects.  Frame 2: still there -> Stay.  Frame 3: gone
    -> the previous Set holds keyAB then keyBA, both absent now, so Exit fires for [1,2] and
    again for [2,1] (logic_worker.js:493-516)."""
    for cls in (CollisionEventsC, CollisionEventsNP):
        ev = cls()
        assert ev.process(cd([(1, 2)])) == [(1, 1, 2), (1, 2, 1)]
        assert ev.process(cd([(1, 2)])) == [(2, 1, 2), (2, 2, 1)]
        assert ev.process(cd([])) == [(3, 1, 2), (3, 2, 1), (3, 2, 1), (3, 1, 2)]
        assert ev.process(cd([])) == []


def test_exit_order_is_previous_frame_list_order():
    for cls in (CollisionEventsC, CollisionEventsNP):
        ev = cls()
        ev.process(cd([(5, 9), (0, 3), (2, 7)]))
        calls = ev.process(cd([(0, 3), (4, 6)]))
        assert calls[:4] == [(2, 0, 3), (2, 3, 0), (1, 4, 6), (1, 6, 4)]
        assert calls[4:] == [(3, 5, 9), (3, 9, 5), (3, 9, 5), (3, 5, 9), (3, 2, 7), (3, 7, 2), (3, 7, 2), (3, 2, 7)]


def test_collision_events_c_equals_np_random_frames():
    rng = np.random.default_rng(5)
    a, b = CollisionEventsC(), CollisionEventsNP()
    live = set()
    for frame in range(40):
        # pairs persist with probability 0.7, new ones arrive; list order shuffles every frame
        live = {p for p in live if rng.random() < 0.7}
        while len(live) < 30 and rng.random() < 0.9:
            i, j = sorted(rng.integers(0, 200, 2).tolist())
            if i != j:
                live.add((i, j))
        pairs = sorted(live)
        rng.shuffle(pairs)
        data = cd(pairs)
        assert a.process(data) == b.process(data), frame


def test_screen_visibility_c_equals_np_and_keeps_inactive_entries():
    rng = np.random.default_rng(7)
    N = 5000
    active = (rng.random(N) < 0.8).astype(np.uint8)
    x = rng.uniform(-500, 4000, N).astype(np.float32)
    y = rng.uniform(-500, 2500, N).astype(np.float32)
    outs = []
    for fn in (screen_visibility_c, screen_visibility_np):
        sx = np.full(N, 123.0, np.float32); sy = np.full(N, -7.0, np.float32); on = np.full(N, 9, np.uint8)
        fn(active, x, y, 1.3, 400.0, 250.0, 1920.0, 1080.0, sx, sy, on)
        outs.append((sx, sy, on))
    for u, v in zip(*outs):
        assert np.array_equal(u.view(np.uint8), v.view(np.uint8))
    sx, sy, on = outs[0]
    assert (sx[active == 0] == 123.0).all() and (on[active == 0] == 9).all()      # :1046 `continue`
    assert 0 < on[active == 1].sum() < active.sum()
    # an entity exactly on the margin is outside (strict comparisons, :1055-1056)
    sx1 = np.zeros(1, np.float32); sy1 = np.zeros(1, np.float32); on1 = np.zeros(1, np.uint8)
    screen_visibility_c(np.ones(1, np.uint8), np.array([-150.0], np.float32), np.array([10.0], np.float32), 1.0, 0.0, 0.0,
                        1000.0, 1000.0, sx1, sy1, on1)
    assert on1[0] == 0 and sx1[0] == -150.0


def random_shadow_scene(rng, N=400, M=12):
    stride = 1 + M
    nd = np.zeros(N * stride, np.int32); dd = np.zeros(N * stride, np.float32)
    for i in range(N):
        c = int(rng.integers(0, M + 1))
        nd[i * stride] = c
        nd[i * stride + 1:i * stride + 1 + c] = rng.integers(0, N, c)
        dd[i * stride + 1:i * stride + 1 + c] = rng.choice([0.25, 4.0, 90.0, 2500.0, 70000.0, 1e5], c).astype(np.float32)
    tact = (rng.random(N) < 0.9).astype(np.uint8)
    x = rng.uniform(0, 2000, N).astype(np.float32); y = rng.uniform(0, 1000, N).astype(np.float32)
    light = (rng.random(N) < 0.15).astype(np.uint8)
    inten = rng.choice([0.0, -1.0, 50.0, 900.0], N).astype(np.float32)
    caster = (rng.random(N) < 0.6).astype(np.uint8)
    rad = rng.choice([0.0, 8.0, 14.5], N).astype(np.float32)
    hgt = rng.choice([0.0, 20.0, 40.0], N).astype(np.float32)
    on = (rng.random(N) < 0.85).astype(np.uint8)
    return M, nd, dd, tact, x, y, light, inten, caster, rad, hgt, on


def test_shadow_sprites_c_equals_np_with_all_three_caps():
    rng = np.random.default_rng(11)
    hit = set()
    for trial in range(6):
        args = random_shadow_scene(rng)
        for caps in ((20, 15, None), (3, 15, None), (20, 2, None), (20, 15, 7)):
            a = shadow_sprites_c(*args, *caps)
            b = shadow_sprites_np(*args, *caps)
            assert a["count"] == b["count"]
            n = a["count"]
            assert np.array_equal(a["active"], b["active"]) and a["active"][:n].all() and not a["active"][n:].any()
            for k in ("radius", "x", "y", "rotation", "scaleX", "scaleY", "alpha"):
                assert np.array_equal(a[k][:n].view(np.uint32), b[k][:n].view(np.uint32)), (k, caps)
            if caps[2] is not None and n == caps[2]:
                hit.add("sprites")
            if caps[0] == 3:
                hit.add("lights")
            if caps[1] == 2:
                hit.add("perlight")
    assert hit == {"sprites", "lights", "perlight"}


def test_shadow_hand_derived():
    """One light at (0,0), intensity 200; caster 1 at (30,40): distSq 2500, dist 50; radius 0 ->
    10 (`|| 10`), height 0 -> radius.  pos = caster - dir*10 = (24, 32); scaleX = 10*0.0714;
    scaleY = (0.3 + (50/256)*0.9) * (10*0.025); alpha = 200/5000; rotation = atan2(40,30) - pi/2.
    Neighbor 2 is closer than 1 unit (dist < 1): skipped (:951)."""
    M = 4
    nd = np.zeros(3 * 5, np.int32); dd = np.zeros(3 * 5, np.float32)
    nd[0] = 2; nd[1] = 2; nd[2] = 1; dd[1] = 0.25; dd[2] = 2500.0
    ones = np.ones(3, np.uint8)
    x = np.array([0, 30, 0.5], np.float32); y = np.array([0, 40, 0], np.float32)
    light = np.array([1, 0, 0], np.uint8); inten = np.array([200, 0, 0], np.float32)
    z = np.zeros(3, np.float32)
    for fn in (shadow_sprites_c, shadow_sprites_np):
        o = fn(M, nd, dd, ones, x, y, light, inten, ones, z, z, ones)
        assert o["count"] == 1 and o["active"][0] == 1 and o["active"][1] == 0
        assert o["radius"][0] == 10.0 and o["x"][0] == np.float32(24.0) and o["y"][0] == np.float32(32.0)
        assert o["scaleX"][0] == np.float32(10 * 0.0714)
        assert o["scaleY"][0] == np.float32((0.3 + (50 * 0.00390625) * 0.9) * (10 * 0.025))
        assert o["alpha"][0] == np.float32(200 / 5000)
        assert o["rotation"][0] == np.float32(np.arctan2(40.0, 30.0) - 1.5707963267948966)


# ---- f4: spawn / despawn pools (gameObject.js:794-951, 668-690, 1001-1034) ----------------------
def blank_cols(N):
    c = {k: np.zeros(N, np.uint8) for k in ("T.active", "RB.active", "C.active")}
    for k in ("T.x", "T.y", "RB.vx", "RB.vy", "RB.ax", "RB.ay", "RB.px", "RB.py", "RB.speed", "RB.velocityAngle"):
        c[k] = np.full(N, 7.0, np.float32)      # stale garbage spawn() has to reset
    return c


def test_pool_hand_derived_interleaved_order_and_spawn_state():
    """count 20 at startIndex 100: the free list is written offset by offset (0,8,16, 1,9,17, ...
    7,15) and popped from the END: 115, 107, 114, 106, 113 (gameObject.js:818-832, :876)."""
    from oracle.oracle_c import PoolC
    from oracle.oracle_np import PoolNP
    for cls in (PoolC, PoolNP):
        col = blank_cols(200)
        p = cls(col, 100, 20)
        got = p.spawn([[10.5, 20.25, 1.5, -2.0]] * 5)
        assert got.tolist() == [115, 107, 114, 106, 113]
        i = 115
        assert col["T.active"][i] == 1 and col["RB.active"][i] == 1 and col["C.active"][i] == 1
        assert col["T.x"][i] == 10.5 and col["T.y"][i] == 20.25 and col["RB.vx"][i] == 1.5 and col["RB.vy"][i] == -2.0
        assert col["RB.px"][i] == 9.0 and col["RB.py"][i] == 22.25          # px = x - vx (:936-939)
        assert col["RB.ax"][i] == 0 and col["RB.speed"][i] == 0 and col["RB.velocityAngle"][i] == 0
        assert col["T.x"][116] == 7.0 and col["T.active"][116] == 0           # untouched
        assert p.available() == 15
        # despawn pushes in call order; a second despawn of the same entity is ignored (:669-670)
        assert p.despawn([107, 115, 107]) == 2
        assert p.spawn([[0, 0, 0, 0]] * 2).tolist() == [115, 107]
        assert p.despawn_all() == 5 and p.available() == 20
        # despawnAll walks the index range upward (:1013), so the stack now ends ..., 113, 114, 115
        assert p.spawn([[0, 0, 0, 0]] * 3).tolist() == [115, 114, 113]


def test_pool_c_equals_np_random_sequences_and_exhaustion():
    from oracle.oracle_c import PoolC
    from oracle.oracle_np import PoolNP
    rng = np.random.default_rng(13)
    ca, cb = blank_cols(600), blank_cols(600)
    a, b = PoolC(ca, 50, 333), PoolNP(cb, 50, 333)
    live = []
    for step in range(60):
        if rng.random() < 0.55:
            rec = rng.normal(0, 50, (int(rng.integers(1, 80)), 4)).astype(np.float32)
            ia, ib = a.spawn(rec), b.spawn(rec)
            assert np.array_equal(ia, ib)
            live += [int(i) for i in ia if i >= 0]
        elif live:
            k = int(rng.integers(1, len(live) + 1))
            pick = rng.choice(live, k).tolist()            # with repeats
            assert a.despawn(pick) == b.despawn(pick)
            live = [i for i in live if i not in set(pick)]
        assert a.available() == b.available() and np.array_equal(a.free_list(), b.free_list())
        for key in ca:
            assert np.array_equal(ca[key].view(np.uint8), cb[key].view(np.uint8)), (step, key)
    ia = a.spawn(np.zeros((400, 4), np.float32))
    assert np.array_equal(ia, b.spawn(np.zeros((400, 4), np.float32))) and (ia == -1).any() and a.available() == 0
