#!/usr/bin/env python
"""Run a few frames of one workload with direct kernel launches (for ncu / compute-sanitizer)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multithreadedgameengine_b200 import binding as B  # noqa: E402
from multithreadedgameengine_b200.engine import GameEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config3")
ap.add_argument("--entities", type=int, default=None)
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
cfg, cols = bench.workload(a.workload, a.entities)
eng = GameEngine(cfg, flags=0 if a.graph else B.FLAG_NO_GRAPH, host_neighbor_rows=False)
eng.load_columns(cols)
eng.run(a.frames)
eng.sync()
print(eng.stats())
eng.close()
