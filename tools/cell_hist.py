#!/usr/bin/env python
"""Cell occupancy of a scene after F frames (where does a settled bed put its entities?).
  python tools/cell_hist.py --frames 300"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multithreadedgameengine_b200 import binding as B
from multithreadedgameengine_b200.engine import GameEngine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config4")
ap.add_argument("--entities", type=int, default=None)
ap.add_argument("--frames", type=int, default=300)
a = ap.parse_args()
cfg, cols = bench.workload(a.workload, a.entities)
e = GameEngine(cfg, host_neighbor_rows=False)
e.load_columns(cols)
e.run(a.frames)
e.download(B.COLS_INPUT_ALL)
cs = cfg["spatial"]["cellSize"]
ncols = int(np.ceil(cfg["worldWidth"] / cs)); nrows = int(np.ceil(cfg["worldHeight"] / cs))
x, y = e.col["T.x"][1:], e.col["T.y"][1:]
cx = np.clip((x / cs).astype(np.int64), 0, ncols - 1); cy = np.clip((y / cs).astype(np.int64), 0, nrows - 1)
occ = np.bincount(cy * ncols + cx, minlength=ncols * nrows)
nz = occ[occ > 0]
srt = np.sort(occ)[::-1]
per_entity = occ[cy * ncols + cx]            # occupancy of each entity's own cell
edges = [1, 8, 16, 32, 64, 128, 256, 512, 1024, 4096, 1 << 30]
hist = {f"<{hi}": int(((per_entity >= lo) & (per_entity < hi)).sum()) for lo, hi in zip(edges[:-1], edges[1:])}
top = np.argsort(occ)[::-1][:12]
print(json.dumps({"frames": a.frames, "cells_nonempty": int(len(nz)), "max": int(srt[0]), "top12": [[int(t % ncols), int(t // ncols), int(occ[t])] for t in top],
                  "entities_by_own_cell_occupancy": hist, "sum_occ_sq": float((occ.astype(np.float64) ** 2).sum()),
                  "sum_occ_sq_cells_over_256": float((occ[occ > 256].astype(np.float64) ** 2).sum())}))
e.close()
