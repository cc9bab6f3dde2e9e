"""The documented J-order (what the GPU computes, bit for bit) against the reference's own
sequential Gauss-Seidel sweep (physics_worker.js:428-562), statistically — BASELINE.json:
"validated by per-step max-abs error and long-run overlap and energy statistics".

J-order is a Jacobi iteration of the same pair formulas: corrections of one sweep do not see
each other, so it relaxes overlaps more slowly than the in-place sweep (measured over 2000 frames:
mean overlap depth 1.19x the reference's in the README scene, 1.75x in a compressed bed) and the two
trajectories of a chaotic falling bed diverge; the bands below — the stated tolerance of DESIGN.md
section 4 — are what BOTH orders must satisfy to count as the same physical regime.  Both sides here are the CPU oracle (order 0 vs order 1); the GPU equals
order 1 exactly (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from helpers import make_oracle
from multithreadedgameengine_b200 import scenes
from oracle.oracle_c import OracleC


def overlaps(o, cfg):
    x = o.col["T.x"].astype(np.float64)
    y = o.col["T.y"].astype(np.float64)
    r = o.col["C.radius"].astype(np.float64)
    N, M = cfg["entityCount"], cfg["spatial"]["maxNeighbors"]
    nd = o.neighborData.reshape(N, 1 + M)
    deps = []
    for i in range(1, N):
        js = nd[i, 1:1 + nd[i, 0]]
        js = js[js > i]
        if len(js):
            d = r[i] + r[js] - np.hypot(x[i] - x[js], y[i] - y[js])
            deps.extend(d[d > 0])
    return np.array(deps) if deps else np.zeros(1)


SCENES = {
    "config1_readme": lambda: scenes.balls_readme(),
    "config3_density": lambda: scenes.scaled("config3", 3000),
}

FRAMES = 2000      # long-run statistics over 2000 frames

# Stated tolerance of the J-order against the reference's sequential sweep.  Measured
# values in brackets; the bands leave the margin a chaotic bed needs, no more.
#   config 1 (the README scene: 1000 balls, r 10-30): a loose bed, the regime the engine is used in
#   config 3 density (3000 balls, r 4): settles into a bed compressed far below contact distance, in BOTH
#   orders (position-based correction with 2 substeps cannot carry the column); the worst case for Jacobi
BANDS = {
    #                     mean overlap   p99.9 overlap  pair count      kinetic energy  centre of mass
    "config1_readme":  dict(mean=(0.8, 1.5), p999=(0.7, 1.4), n=(0.8, 1.25), ke=(0.5, 2.0), cm=0.02),   # [1.19, 0.98, 1.06, 0.96, 0.009]
    "config3_density": dict(mean=(0.8, 2.0), p999=(0.7, 1.7), n=(0.8, 1.25), ke=(0.4, 2.5), cm=0.06),   # [1.75, 1.38, 1.07, 1.60, 0.050]
}
# per-step (teacher-forced, both orders from the same state, first 40 frames), in contact diameters
STEP = {"config1_readme": dict(max=0.5, p999=0.45),        # [0.43, 0.40]
        "config3_density": dict(max=2.5, p999=1.6)}        # [1.85, 1.32]


def outside_world(o, cfg):
    x, y = o.col["T.x"][1:], o.col["T.y"][1:]
    W, H = np.float32(cfg["worldWidth"]), np.float32(cfg["worldHeight"])
    return int(((x < 0) | (x > W) | (y < 0) | (y > H) | ~np.isfinite(x) | ~np.isfinite(y)).sum())


@pytest.mark.parametrize("name", list(SCENES))
def test_long_run_statistics_same_regime(name):
    cfg, cols = SCENES[name]()
    runs = {}
    for order in (0, 1):
        o = make_oracle(OracleC, cfg, cols)
        out = 0
        for _ in range(FRAMES):
            o.step(1.0, order)
            out += outside_world(o, cfg)
        vx, vy = o.col["RB.vx"].astype(np.float64), o.col["RB.vy"].astype(np.float64)
        dep = overlaps(o, cfg)
        runs[order] = dict(ke=0.5 * (vx * vx + vy * vy).sum(), cmx=o.col["T.x"][1:].mean(), cmy=o.col["T.y"][1:].mean(),
                           mean_overlap=dep.mean(), p999=np.percentile(dep, 99.9), n=len(dep), outside=out)
    a, b = runs[0], runs[1]
    band = BANDS[name]
    # entity centres outside the world rectangle, summed over all frames.  (The sweep runs AFTER the boundary
    # pass of a substep, physics_worker.js:344-376 then :405-568, so a correction can push an entity past
    # [r, W - r] in either order; leaving the WORLD takes a compressed bed.)
    if name == "config1_readme":
        assert a["outside"] == 0 and b["outside"] == 0
    else:
        assert b["outside"] <= 1.5 * a["outside"] + 10, (a["outside"], b["outside"])     # [13 442 vs 16 601 entity-frames of 6 M]
    # the bed settles at the same place
    assert abs(a["cmx"] - b["cmx"]) < band["cm"] * cfg["worldWidth"]
    assert abs(a["cmy"] - b["cmy"]) < band["cm"] * cfg["worldHeight"]
    for key, ra in (("mean", b["mean_overlap"] / a["mean_overlap"]), ("p999", b["p999"] / a["p999"]),
                    ("n", b["n"] / a["n"]), ("ke", (b["ke"] + 1) / (a["ke"] + 1))):
        lo, hi = band[key]
        assert lo < ra < hi, (key, ra, a, b)


@pytest.mark.parametrize("name", list(SCENES))
def test_per_step_difference_is_local(name):
    """Teacher-forced single frames: both orders start from the same state.  Entities without
    a collision partner must agree exactly; the rest differ by a fraction of a contact diameter
    in a loose bed, by at most a couple of diameters in the compressed one."""
    cfg, cols = SCENES[name]()
    a = make_oracle(OracleC, cfg, cols)
    diam = 2.0 * float(cols["C.radius"].max())
    worst, worst999 = 0.0, 0.0
    for _ in range(40):
        b = make_oracle(OracleC, cfg, {k: a.col[k].copy() for k in a.col})
        a.step(1.0, 0)
        b.step(1.0, 1)
        d = np.maximum(np.abs(a.col["T.x"] - b.col["T.x"]), np.abs(a.col["T.y"] - b.col["T.y"]))
        free = a.col["RB.collisionCount"] == 0
        free &= b.col["RB.collisionCount"] == 0
        assert d[free].max() == 0.0          # no partner -> identical arithmetic
        assert np.median(d) < 0.1 * float(cols["C.radius"].max())   # typical entity: a small fraction of a radius
        worst = max(worst, float(d.max()) / diam)
        worst999 = max(worst999, float(np.percentile(d, 99.9)) / diam)
        # pair sets of the first substep are identical, so the velocity fields are too
        assert np.array_equal(a.col["RB.vx"], b.col["RB.vx"]) and np.array_equal(a.col["RB.vy"], b.col["RB.vy"])
    assert worst < STEP[name]["max"] and worst999 < STEP[name]["p999"], (worst, worst999)
