"""Both oracles against the hand-derived micro-cases of tests/golden/micro_cases.json
(worked out from the reference sources; see tests/golden/make_micro_cases.py)."""
import json
import os

import numpy as np
import pytest

from multithreadedgameengine_b200 import scenes
from oracle.oracle_c import OracleC
from oracle.oracle_np import OracleNP

CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "micro_cases.json")))


def build_case(case):
    ents = case["entities"]
    N = len(ents)
    c = scenes._blank(N)
    for i, (x, y, r, vr, fl) in enumerate(ents):
        c["T.x"][i] = float(x)
        c["T.y"][i] = float(y)
        c["RB.px"][i], c["RB.py"][i] = c["T.x"][i], c["T.y"][i]
        c["C.radius"][i] = r
        c["C.visualRange"][i] = vr
        c["T.active"][i] = "T" in fl
        c["RB.active"][i] = "R" in fl
        c["C.active"][i] = "C" in fl
        c["RB.static"][i] = "S" in fl
        c["C.isTrigger"][i] = "G" in fl
        c["RB.maxVel"][i] = 50
    for i, (px, py) in enumerate(case.get("prev", [])):
        c["RB.px"][i], c["RB.py"][i] = px, py
    for k, v in case.get("set", {}).items():
        c[k][:] = v
    p = dict(subStepCount=1, boundaryElasticity=0.8, collisionResponseStrength=0.5, verletDamping=0.995,
             minSpeedForRotation=0.1, gravityX=0.0, gravityY=0.0)
    p.update(case.get("physics", {}))
    cfg = dict(entityCount=N, worldWidth=float(case["world"][0]), worldHeight=float(case["world"][1]), seed=1,
               spatial=dict(cellSize=float(case["cellSize"]), maxNeighbors=case["maxNeighbors"]),
               physics=dict(subStepCount=p["subStepCount"], boundaryElasticity=p["boundaryElasticity"],
                            collisionResponseStrength=p["collisionResponseStrength"], verletDamping=p["verletDamping"],
                            minSpeedForRotation=p["minSpeedForRotation"], gravity=dict(x=p["gravityX"], y=p["gravityY"])))
    return cfg, c, p


def check_case(case, sim, order, spatial_only_ok=True):
    """sim must offer .spatial() .physics(dt, order)/.step, .col, .neighborData, .distanceData, .collisionData"""
    M = case["maxNeighbors"]
    stride = 1 + M
    frames = case.get("frames", 0)
    if frames == 0:
        sim.spatial()
    for _ in range(frames):
        sim.step(1.0, order)
    for i, (ids, d2) in case.get("rows", {}).items():
        o = int(i) * stride
        assert int(sim.neighborData[o]) == len(ids), f"{case['name']}: row {i} count"
        assert list(sim.neighborData[o + 1:o + 1 + len(ids)]) == ids, f"{case['name']}: row {i} ids"
        assert list(sim.distanceData[o + 1:o + 1 + len(ids)]) == [np.float32(v) for v in d2], f"{case['name']}: row {i} d2"
        assert float(sim.distanceData[o]) == len(ids)
    for i in case.get("untouched_rows", []):
        assert not sim.neighborData[i * stride:(i + 1) * stride].any()
    exp = dict(case.get("expect", {}))
    exp.update(case.get("expect_reference" if order == 0 else "expect_jorder", {}))
    for k, vals in exp.items():
        got = sim.col[k]
        want = np.array(vals, dtype=np.float64).astype(got.dtype)   # literal -> fround / uint8
        assert np.array_equal(got[:len(vals)], want), f"{case['name']}: {k} {got[:len(vals)]} != {want}"
    if "pairs" in case and frames:
        n = int(sim.collisionData[0])
        assert n == len(case["pairs"])
        assert sim.collisionData[1:1 + 2 * n].reshape(-1, 2).tolist() == case["pairs"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("impl", [OracleC, OracleNP], ids=["c", "np"])
@pytest.mark.parametrize("order", [0, 1], ids=["reference_order", "j_order"])
def test_micro_case(case, impl, order):
    cfg, cols, p = build_case(case)
    sim = impl(cfg["entityCount"], cfg["worldWidth"], cfg["worldHeight"], cfg["spatial"]["cellSize"],
               cfg["spatial"]["maxNeighbors"], 100, 1.0, p)
    if impl is OracleC:
        sim.load(cols)
    else:
        for k, v in cols.items():
            sim.col[k][:] = v
    check_case(case, sim, order)
    if "cells" in case:
        sim.spatial()
        cell_of = sim.grid_csr()[0] if impl is OracleC else sim.cellOf
        for i, cell in case["cells"].items():
            assert int(cell_of[int(i)]) == cell, f"{case['name']}: cell of {i}"
