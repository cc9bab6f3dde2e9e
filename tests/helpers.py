"""Shared test helpers: build oracles / engines from a (config, columns) scene."""
from __future__ import annotations

import numpy as np

from oracle.oracle_c import OracleC
from oracle.oracle_np import OracleNP

PHYS_DEFAULTS = dict(subStepCount=4, boundaryElasticity=0.8, collisionResponseStrength=0.5,
                     verletDamping=0.995, minSpeedForRotation=0.1)


def phys_kwargs(cfg):
    """Flatten config.physics the way gameEngine.js:34-49 + validatePhysicsConfig would."""
    p = dict(PHYS_DEFAULTS)
    src = dict(cfg.get("physics") or {})
    g = src.pop("gravity", None) or cfg.get("gravity") or {"x": 0.0, "y": 0.0}
    src.pop("maxCollisionPairs", None)
    src.pop("noLimitFPS", None)
    p.update(src)
    p["gravityX"] = float(g.get("x", 0.0))
    p["gravityY"] = float(g.get("y", 0.0))
    return p


def max_pairs(cfg):
    return int((cfg.get("physics") or {}).get("maxCollisionPairs") or cfg.get("maxCollisionPairs") or 10000)


def make_oracle(cls, cfg, cols):
    o = cls(cfg["entityCount"], cfg["worldWidth"], cfg["worldHeight"], cfg["spatial"]["cellSize"],
            cfg["spatial"]["maxNeighbors"], max_pairs(cfg), cfg.get("seed", 1.0), phys_kwargs(cfg))
    if isinstance(o, OracleC):
        o.load(cols)
    else:
        for k, v in cols.items():
            o.col[k][:] = v
    return o


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    return a


def assert_cols_equal(a, b, keys=None, what=""):
    keys = keys or a.keys()
    for k in keys:
        x, y = bits(a[k]), bits(b[k])
        if not np.array_equal(x, y):
            bad = np.nonzero(x != y)[0]
            raise AssertionError(f"{what} column {k}: {bad.size} mismatches, first at {bad[0]}: "
                                 f"{a[k][bad[0]]!r} vs {b[k][bad[0]]!r}")


def active_rows_equal(nd_a, dd_a, nd_b, dd_b, N, M, rows):
    """Compare count + the first count entries of each listed row (ids and float32 d² bits).
    Vectorised (benchmark-scale scenes have millions of rows); the first mismatch is reported by
    row."""
    stride = 1 + M
    rows = np.asarray(rows, dtype=np.int64)
    if rows.size == 0:
        return
    A = np.asarray(nd_a).reshape(-1, stride)
    Bn = np.asarray(nd_b).reshape(-1, stride)
    DA = bits(np.asarray(dd_a)).reshape(-1, stride)
    DB = bits(np.asarray(dd_b)).reshape(-1, stride)
    for lo in range(0, rows.size, 1 << 18):                      # bounded temporaries
        r = rows[lo:lo + (1 << 18)]
        a, b, da, db = A[r], Bn[r], DA[r], DB[r]
        ca, cb = a[:, 0], b[:, 0]
        if not np.array_equal(ca, cb):
            k = int(np.nonzero(ca != cb)[0][0])
            raise AssertionError(f"row {int(r[k])}: count {int(ca[k])} vs {int(cb[k])}")
        valid = np.arange(stride)[None, :] <= ca[:, None]        # header + count entries
        bad = ((a != b) | (da != db)) & valid
        if bad.any():
            k, w = (int(v[0]) for v in np.nonzero(bad))
            what = "ids" if a[k, w] != b[k, w] else "d2"
            raise AssertionError(f"row {int(r[k])}: {what} differ at word {w}: {a[k, w]} / {da[k, w]:#x} vs {b[k, w]} / {db[k, w]:#x}")


def random_scene(rng, N=300, W=1000.0, H=600.0, cellSize=50.0, M=16, S=2, weird=True):
    """A small adversarial scene: mixed visual ranges, triggers, statics, inactive slots,
    out-of-world, NaN/Inf, coincident points, entities on cell edges."""
    from multithreadedgameengine_b200 import scenes
    c = scenes._blank(N)
    F32 = np.float32
    c["T.active"][:] = rng.random(N) < 0.9
    c["T.x"][:] = (rng.random(N) * W).astype(F32)
    c["T.y"][:] = (rng.random(N) * H).astype(F32)
    c["RB.active"][:] = rng.random(N) < 0.9
    c["RB.static"][:] = rng.random(N) < 0.1
    c["C.active"][:] = rng.random(N) < 0.9
    c["C.isTrigger"][:] = rng.random(N) < 0.1
    c["C.radius"][:] = (rng.random(N) * 25 + 2).astype(F32)
    c["C.visualRange"][:] = (rng.random(N) * 2.2 * cellSize).astype(F32)
    c["RB.maxVel"][:] = np.where(rng.random(N) < 0.2, 0, rng.random(N) * 30).astype(F32)
    v = ((rng.random((2, N)) - 0.5) * 8).astype(F32)
    c["RB.px"][:] = (c["T.x"].astype(np.float64) - v[0]).astype(F32)
    c["RB.py"][:] = (c["T.y"].astype(np.float64) - v[1]).astype(F32)
    c["RB.ax"][:] = ((rng.random(N) - 0.5)).astype(F32)
    c["RB.ay"][:] = ((rng.random(N) - 0.5)).astype(F32)
    c["RB.velocityAngle"][:] = rng.random(N).astype(F32)
    c["RB.collisionCount"][:] = rng.integers(0, 255, N)
    if weird:
        k = max(1, N // 30)
        idx = rng.permutation(N)
        c["T.x"][idx[:k]] = (rng.integers(0, int(W / cellSize), k) * cellSize).astype(F32)  # on cell edges
        c["T.y"][idx[k:2 * k]] = np.nextafter((rng.integers(1, int(H / cellSize), k) * cellSize).astype(F32), F32(0))
        c["T.x"][idx[2 * k:3 * k]] = (-rng.random(k) * 300).astype(F32)                      # out of world
        c["T.y"][idx[3 * k:4 * k]] = (H + rng.random(k) * 300).astype(F32)
        c["T.x"][idx[4 * k]] = np.nan
        c["T.y"][idx[4 * k + 1]] = np.nan
        c["T.x"][idx[4 * k + 2]] = np.inf
        c["T.y"][idx[4 * k + 3]] = -np.inf
        c["T.x"][idx[4 * k + 4]] = 3e9        # ToInt32 wrap territory
        c["T.x"][idx[4 * k + 5]] = -1e12
        for a, b in zip(idx[5 * k:6 * k], idx[6 * k:7 * k]):  # coincident pairs
            c["T.x"][a] = c["T.x"][b]
            c["T.y"][a] = c["T.y"][b]
        c["C.visualRange"][idx[7 * k]] = np.nan
        c["C.visualRange"][idx[7 * k + 1]] = 0
        c["C.visualRange"][idx[7 * k + 2]] = np.inf
        c["C.visualRange"][idx[7 * k + 3]] = -5
    cfg = dict(entityCount=N, worldWidth=W, worldHeight=H, seed=7,
               spatial=dict(cellSize=cellSize, maxNeighbors=M),
               physics=dict(subStepCount=S, gravity=dict(x=0.1, y=0.5), verletDamping=0.99,
                            maxCollisionPairs=200))
    return cfg, c
