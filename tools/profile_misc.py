#!/usr/bin/env python
"""Runs, once each with direct launches, the kernels the default benchmark frame does not reach, so that
ncu can capture them (profiles/: "an ncu capture for every kernel"):
  * config 2 (boids, maxNeighbors 1500): k_neighbors_wide, k_system_flock, k_rows_gather, k_pack / k_unpack
  * two slabs of a 2 M-entity config-4 scene in one process (weed_group_*): k_slab_pack, k_slab_headers,
    k_slab_wait, k_slab_drop, k_slab_unpack, k_slab_finish
  * the settled bed (config 3, frame 300): k_beyond_cap_dense, the F_XPOOL / F_XOVER paths of k_sweep
  * the TMA-tiled sweep: k_sweep_tile
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from multithreadedgameengine_b200 import binding as B, scenes  # noqa: E402
from multithreadedgameengine_b200.engine import GameEngine  # noqa: E402
from multithreadedgameengine_b200.slabs import SlabGroup  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"

if what in ("all", "boids"):
    cfg, cols = scenes.boids()
    cols["T.entityType"] = np.zeros(cfg["entityCount"], np.uint8)
    cols["T.entityType"][1:10001] = 1
    cols["T.entityType"][10001:] = 2
    eng = GameEngine(cfg, flags=B.FLAG_NO_GRAPH)
    eng.load_columns({k: v for k, v in cols.items() if k != "T.entityType"})
    eng.column("T.entityType")[:] = cols["T.entityType"]
    eng.upload(B.COL["T.entityType"])
    for _ in range(3):
        eng.step(1.0, 0, B.COLS_OUTPUT_ALL)
        eng.system_flock(scenes.PREDATORS_DEMO_CLASSES, 1.0)
    eng.fetch_neighbors()
    print("boids", eng.stats()["neighborsTotal"])
    eng.close()

if what in ("all", "slabs"):
    cfg, cols = scenes.scaled("config4", 2_000_000)
    grp = SlabGroup(cfg, cols, 2, flags=B.FLAG_NO_GRAPH)
    for _ in range(4):
        grp.step(1.0)
    print("slabs", [s.status()["owned"] for s in grp.slabs])
    grp.close()

if what in ("all", "bed"):
    cfg, cols = scenes.config3()
    eng = GameEngine(cfg, host_neighbor_rows=False)
    eng.load_columns(cols)
    eng.run(300)
    eng.sync()
    eng.close()
    # same state reached with direct launches for the last frames is too slow to replay under ncu from frame 0:
    # run 300 graph frames, then 2 direct ones in a second context fed with the downloaded state
    eng = GameEngine(cfg, host_neighbor_rows=False)
    eng.load_columns(cols)
    eng.run(300)
    eng.download(B.COLS_INPUT_ALL)
    state = {k: eng.col[k].copy() for k in cols}
    eng.close()
    eng = GameEngine(cfg, flags=B.FLAG_NO_GRAPH, host_neighbor_rows=False)
    eng.load_columns(state)
    eng.run(2)
    st = eng.stats()
    print("bed", st["cappedRows"], st["ms"][9])
    eng.close()

if what in ("all", "tile"):
    cfg, cols = scenes.scaled("config4", 2_000_000)
    eng = GameEngine(cfg, flags=B.FLAG_NO_GRAPH | B.FLAG_K6_TILE, host_neighbor_rows=False)
    eng.load_columns(cols)
    eng.run(4)
    print("tile", eng.stats()["collisionPairs"])
    eng.close()
