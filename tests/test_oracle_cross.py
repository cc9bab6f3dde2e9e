"""The two independent CPU restatements (C and numpy/Python) must agree bit-for-bit.
This is the only pin available: the reference has no tests/golden vectors for this path and
no JS engine exists in the image ("parity unpinned", SURVEY §8 c)."""
import numpy as np
import pytest

from oracle.oracle_c import OracleC, lib
from oracle.oracle_np import OracleNP, SeededRandom, layout, nudge_dir, nudge_hash, to_int32
from multithreadedgameengine_b200 import scenes
from helpers import active_rows_equal, assert_cols_equal, bits, make_oracle, random_scene


def test_toint32_matches():
    L = lib()
    vals = [0.0, -0.0, 0.9, -0.9, 1.5, -1.5, 2147483647.0, 2147483648.0, -2147483648.0, -2147483649.0,
            4294967296.0, 4294967295.0, 1e10, -1e10, 3e9 / 50, 1e300, -1e300, float("nan"), float("inf"),
            float("-inf"), 2.0 ** 53, -(2.0 ** 53) - 2, 6e37, 1.5e38]
    rng = np.random.default_rng(0)
    vals += list((rng.random(2000) - 0.5) * 10 ** rng.integers(0, 25, 2000))
    for v in vals:
        assert L.wo_js_toint32(float(v)) == to_int32(float(v)), v
    # known answers (ECMA-262 ToInt32)
    assert to_int32(4294967296.0 + 5) == 5
    assert to_int32(2147483648.0) == -2147483648
    assert to_int32(-1.9) == -1
    assert to_int32(float("nan")) == 0


def test_seeded_random_matches():
    L = lib()
    for seed in (1.0, 1234.0, 0.5, 42.0, 2 ** 31 + 7.0):
        r = SeededRandom(seed)
        seq = [r() for _ in range(50)]
        assert all(0.0 <= v < 1.0 for v in seq)
        for n in (1, 2, 17, 50):
            assert L.wo_seeded_random(seed, n) == seq[n - 1]


def test_nudge_matches_and_is_unit():
    import ctypes as C
    L = lib()
    rng = np.random.default_rng(1)
    for h in list(rng.integers(0, 2 ** 32, 500)) + [0, 2 ** 32 - 1, 0x20000000, 0xDFFFFFFF, 0xE0000000]:
        c, s = C.c_double(), C.c_double()
        L.wo_nudge_dir(int(h), C.byref(c), C.byref(s))
        pc, ps = nudge_dir(int(h))
        assert (c.value, s.value) == (pc, ps)
        ang = 2 * np.pi * int(h) / 2 ** 32
        assert abs(pc - np.cos(ang)) < 1e-9 and abs(ps - np.sin(ang)) < 1e-9
    assert L.wo_nudge_hash(3, 9, 4, 1, 77) == nudge_hash(3, 9, 4, 1, 77)


def test_layout_matches_survey_a1():
    # SURVEY §8 a1 worked offsets for N = 1001
    t, ts = layout("Transform", 1001)
    assert (t["active"], t["entityType"], t["x"], t["y"], t["rotation"], ts) == (0, 1001, 2004, 6008, 10012, 14016)
    r, rs = layout("RigidBody", 1001)
    assert r["vx"] == 2004 and r["collisionCount"] == 82084 and rs == 83085
    c, cs = layout("Collider", 1001)
    assert (c["radius"], c["isTrigger"], c["restitution"], c["collisionLayer"], c["collisionMask"],
            c["aabbMinX"], c["visualRange"], cs) == (10012, 22024, 23028, 27032, 29034, 31036, 47052, 51056)
    L = lib()
    for comp, name in enumerate(("Transform", "RigidBody", "Collider")):
        for N in (1, 2, 3, 7, 1001, 4096, 10001):
            lay, size = layout(name, N)
            assert L.wo_buffer_size(comp, N) == size
            for col, off in lay.items():
                assert L.wo_column_offset(comp, col.encode(), N) == off


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("order", [0, 1])
def test_random_scenes_bit_identical(seed, order):
    rng = np.random.default_rng(seed)
    cs = [50.0, 33.3, 30.0, 66.5, 128.0, 17.0][seed]
    cfg, cols = random_scene(rng, N=220, cellSize=cs, M=[16, 8, 400, 5, 32, 12][seed], S=[2, 1, 3, 2, 4, 2][seed])
    a = make_oracle(OracleC, cfg, cols)
    b = make_oracle(OracleNP, cfg, cols)
    for frame in range(3):
        a.spatial()
        b.spatial()
        cellOf, start, idx = a.grid_csr()
        bstart, bidx = b.grid_csr()
        assert np.array_equal(cellOf, b.cellOf)
        assert np.array_equal(start, bstart) and np.array_equal(idx, bidx)
        rows = np.nonzero(cellOf >= 0)[0]
        active_rows_equal(a.neighborData, a.distanceData, b.neighborData, b.distanceData,
                          cfg["entityCount"], cfg["spatial"]["maxNeighbors"], rows)
        a.physics(1.0 if frame != 1 else 0.73, order)
        b.physics(1.0 if frame != 1 else 0.73, order)
        assert_cols_equal(a.col, b.col, what=f"frame {frame}")
        n = int(a.collisionData[0])
        assert n == int(b.collisionData[0])
        assert np.array_equal(a.collisionData[:1 + 2 * n], b.collisionData[:1 + 2 * n])


def test_config1_frames_bit_identical():
    cfg, cols = scenes.balls_readme(n_balls=400, seed=5, world=(1200.0, 600.0))
    for order in (0, 1):
        a = make_oracle(OracleC, cfg, cols)
        b = make_oracle(OracleNP, cfg, cols)
        for _ in range(4):
            a.step(1.0, order)
            b.step(1.0, order)
        assert_cols_equal(a.col, b.col)
        assert int(a.collisionData[0]) > 0
        assert np.array_equal(a.collisionData[:1 + 2 * int(a.collisionData[0])],
                              b.collisionData[:1 + 2 * int(b.collisionData[0])])
