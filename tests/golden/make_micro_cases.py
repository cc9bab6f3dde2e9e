#!/usr/bin/env python
"""Hand-derived micro-cases for the spatial+physics path.

The reference ships no tests or golden vectors for this path, and no JS engine
exists in this image, so these cases were worked out BY HAND from the reference sources
(src/workers/spatial_worker.js:122-278, src/workers/physics_worker.js:145-604); each case says
which lines it exercises.  Expected values are written as literals below — they are NOT
produced by running any oracle.  Running this script just (re)writes micro_cases.json.

Entity tuples: (x, y, radius, visualRange, flags) with flags a string of
  T = Transform.active, R = RigidBody.active, C = Collider.active, S = static, G = trigger.
px/py default to x/y (zero start velocity) unless given in "prev".
"""
import json
import os

CASES = [
    dict(
        name="row_strict_less_than",
        why="spatial_worker.js:257 d2 < vr2 is strict; :234-246 scan order; :249 self skipped",
        world=[200, 100], cellSize=50, maxNeighbors=4,
        entities=[(10, 10, 1, 30, "TRC"), (30, 10, 1, 30, "TRC"), (60, 10, 1, 30, "TRC")],
        # d(0,1)=20 -> 400 < 900 ; d(1,2)=30 -> 900 !< 900 ; d(0,2)=50
        rows={"0": [[1], [400.0]], "1": [[0], [400.0]], "2": [[], []]},
    ),
    dict(
        name="coincident_excluded_and_cell_edge",
        why=":257 d2 > 0 excludes coincident points; :157 x=50 with cellSize 50 lands in column 1",
        world=[200, 100], cellSize=50, maxNeighbors=4,
        entities=[(50, 10, 1, 20, "TRC"), (50, 10, 1, 20, "TRC"), (49, 10, 1, 20, "TRC")],
        cells={"0": 1, "1": 1, "2": 0},
        rows={"0": [[2], [1.0]], "1": [[2], [1.0]], "2": [[0, 1], [1.0, 1.0]]},
    ),
    dict(
        name="cap_keeps_scan_order",
        why=":264 stop at maxNeighbors; order = rows, then columns, then ascending id within a cell",
        world=[150, 150], cellSize=50, maxNeighbors=2,
        # entity 4 sits in the centre cell; candidates in cells (row0,col1)=id3, (row1,col0)=id2,
        # (row1,col1)=ids 0 and 4, (row1,col2)=id1 ... scan: row0 first -> 3, then row1 col0 -> 2
        entities=[(80, 80, 1, 60, "TRC"), (110, 75, 1, 60, "TRC"), (40, 75, 1, 60, "TRC"),
                  (75, 40, 1, 60, "TRC"), (75, 75, 1, 60, "TRC")],
        rows={"4": [[3, 2], [1225.0, 1225.0]]},
    ),
    dict(
        name="out_of_world_insert_clamped_query_unclamped",
        why=":157-160 insert clamps the cell, :214-215 the query centre does not; (-0.6)|0 = 0, (-1.2)|0 = -1",
        world=[200, 100], cellSize=50, maxNeighbors=4,
        entities=[(-30, 10, 1, 50, "TRC"), (-60, 10, 1, 50, "TRC"), (10, 10, 1, 50, "TRC"), (60, 10, 1, 100, "TRC")],
        cells={"0": 0, "1": 0, "2": 0, "3": 1},
        # e0: centre col 0, r=1 -> cols 0..1: sees 1 (d=30), 2 (d=40); 3 at d=90 no
        # e1: centre col -1, r=1 -> cols 0..0 only: 0 (d=30); 2 at d=70 > 50 no
        # e3: vr 100, r=2, centre col 1 -> cols 0..3: 0 (d=90 -> 8100), 1 (d=120 no), 2 (d=50 -> 2500)
        rows={"0": [[1, 2], [900.0, 1600.0]], "1": [[0], [900.0]], "2": [[0], [1600.0]],
              "3": [[0, 2], [8100.0, 2500.0]]},
    ),
    dict(
        name="inactive_and_nan_skipped",
        why=":148 inactive skipped, :153 NaN skipped; their rows are never written",
        world=[200, 100], cellSize=50, maxNeighbors=4,
        entities=[(10, 10, 1, 40, "TRC"), (20, 10, 1, 40, "RC"), ("nan", 10, 1, 40, "TRC"), (30, 10, 1, 40, "TRC")],
        cells={"0": 0, "1": -1, "2": -1, "3": 0},
        rows={"0": [[3], [400.0]], "3": [[0], [400.0]]},
        untouched_rows=[1, 2],
    ),
    dict(
        name="free_fall_two_frames",
        why="physics_worker.js:275-310: d = (x-px)*damping + dt^2*g + a*dt; p <- old x; v = d/dt; a <- 0",
        world=[1000, 1000], cellSize=50, maxNeighbors=4,
        physics=dict(subStepCount=1, gravityY=0.5, verletDamping=0.99),
        entities=[(100, 50, 5, 10, "TRC")],
        frames=2,
        # frame 1: dy = 0*0.99 + 0.5 = 0.5 ; frame 2: dy = 0.5*0.99 + 0.5 = 0.995
        expect={"T.y": [51.495], "RB.py": [50.5], "RB.vy": [0.995], "T.x": [100.0], "RB.vx": [0.0],
                "RB.speed": [0.995]},
        float_from_double=True,
    ),
    dict(
        name="acceleration_and_axis_clamp",
        why=":279 a*dtRatio added; :284 maxVel<=0 -> 100; :297-298 per-axis clamp; :313-314 a <- 0",
        world=[1000, 1000], cellSize=50, maxNeighbors=4,
        physics=dict(subStepCount=1, verletDamping=1.0),
        entities=[(500, 500, 5, 10, "TRC"), (300, 300, 5, 10, "TRC")],
        set={"RB.ax": [3.0, 500.0], "RB.ay": [-4.0, -500.0], "RB.maxVel": [2.0, 0.0]},
        frames=1,
        expect={"T.x": [502.0, 400.0], "T.y": [498.0, 200.0], "RB.vx": [2.0, 100.0], "RB.vy": [-2.0, -100.0],
                "RB.ax": [0.0, 0.0], "RB.ay": [0.0, 0.0], "RB.px": [500.0, 300.0], "RB.py": [500.0, 300.0]},
    ),
    dict(
        name="two_ball_overlap_split_evenly",
        why=":447-455 overlap test, :510-546 depth*strength split in halves along the normal; "
            ":551-559 counters and pair log",
        world=[1000, 1000], cellSize=50, maxNeighbors=4,
        physics=dict(subStepCount=1, collisionResponseStrength=0.5, verletDamping=1.0),
        entities=[(100, 100, 10, 40, "TRC"), (112, 100, 10, 40, "TRC")],
        frames=1,
        # dist 12, minDist 20, depth 8, corr 4, half 2, n = (-1, 0)
        expect={"T.x": [98.0, 114.0], "T.y": [100.0, 100.0], "RB.collisionCount": [1, 1]},
        pairs=[[0, 1]],
    ),
    dict(
        name="static_takes_nothing_trigger_moves_nothing",
        why=":532-539 the dynamic body takes the full correction against a static one; "
            ":514-517 trigger pairs are counted and logged but not moved",
        world=[1000, 1000], cellSize=50, maxNeighbors=4,
        physics=dict(subStepCount=1, collisionResponseStrength=0.5, verletDamping=1.0),
        entities=[(100, 100, 10, 40, "TRCS"), (112, 100, 10, 40, "TRC"),
                  (500, 500, 10, 40, "TRCG"), (506, 500, 10, 40, "TRC")],
        frames=1,
        # pair (0,1): i static -> j moves by -n*corr = +4 ; pair (2,3): trigger
        expect={"T.x": [100.0, 116.0, 500.0, 506.0], "RB.collisionCount": [1, 1, 1, 1]},
        pairs=[[0, 1], [2, 3]],
    ),
    dict(
        name="boundary_bounce",
        why=":353-375 clamp to [r, W-r] x [r, H-r]; p <- x + (x - p)*elasticity with the clamped x",
        world=[200, 100], cellSize=50, maxNeighbors=4,
        physics=dict(subStepCount=1, boundaryElasticity=0.8, verletDamping=1.0),
        entities=[(5, 50, 10, 10, "TRC"), (195, 97, 10, 10, "TRC")],
        prev=[(5, 50), (195, 97)],
        frames=1,
        # e0: x 5 < 10 -> x = 10, px = 10 + (10 - 5)*0.8 = 14 ; e1: x -> 190, px = 190 + (190-195)*.8 = 186;
        #     y 97 > 90 -> y = 90, py = 90 + (90-97)*.8 = 84.4
        expect={"T.x": [10.0, 190.0], "RB.px": [14.0, 186.0], "T.y": [50.0, 90.0], "RB.py": [50.0, 84.4]},
        float_from_double=True,
    ),
    dict(
        name="chain_of_three_reference_vs_jorder",
        why=":428-562 the reference sweep is in-place (pair (1,2) sees the correction of pair (0,1)); "
            "the documented J-order evaluates both pairs on the start positions",
        world=[1000, 1000], cellSize=50, maxNeighbors=4,
        physics=dict(subStepCount=1, collisionResponseStrength=0.5, verletDamping=1.0),
        entities=[(100, 100, 10, 40, "TRC"), (112, 100, 10, 40, "TRC"), (124, 100, 10, 40, "TRC")],
        frames=1,
        # reference: (0,1): 98 | 114 ; then (1,2): dist 10, depth 10, half 2.5 -> 111.5 | 126.5
        expect_reference={"T.x": [98.0, 111.5, 126.5]},
        # J-order: (0,1) and (1,2) both from the start positions: 98 | 112+2-2 | 126
        expect_jorder={"T.x": [98.0, 112.0, 126.0]},
        pairs=[[0, 1], [1, 2]],
    ),
]

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro_cases.json")
    json.dump(CASES, open(out, "w"), indent=1)
    print("wrote", out, len(CASES), "cases")
