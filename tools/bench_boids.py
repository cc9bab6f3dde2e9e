#!/usr/bin/env python
"""Config 2 style boids scene at a chosen size: frame (spatial+physics) + device-side tick()
(weed_system_boids) versus what the host-tick path has to move over PCIe (the neighbor rows)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multithreadedgameengine_b200 import binding as B, scenes
from multithreadedgameengine_b200.engine import GameEngine

ap = argparse.ArgumentParser()
ap.add_argument("--prey", type=int, default=200_000)
ap.add_argument("--pred", type=int, default=10_000)
ap.add_argument("--max-neighbors", type=int, default=128)
ap.add_argument("--frames", type=int, default=30)
a = ap.parse_args()
cfg, cols = scenes.boids(a.prey, a.pred)
scale = ((a.prey + a.pred) / 10500.0) ** 0.5              # keep the demo's density
cfg["worldWidth"], cfg["worldHeight"] = round(5000.0 * scale / 128) * 128.0, round(2000.0 * scale / 128) * 128.0
for k, f in (("T.x", cfg["worldWidth"] / 5000), ("RB.px", cfg["worldWidth"] / 5000), ("T.y", cfg["worldHeight"] / 2000), ("RB.py", cfg["worldHeight"] / 2000)):
    cols[k] = (cols[k] * np.float32(f)).astype(np.float32)
cfg["spatial"]["maxNeighbors"] = a.max_neighbors
N = cfg["entityCount"]
etype = np.zeros(N, dtype=np.uint8)
etype[1:1 + a.prey] = 1
etype[1 + a.prey:] = 2
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    eng = GameEngine(cfg, stream=stream.cuda_stream, host_neighbor_rows=False)
    eng.load_columns(cols)
    eng.Transform.entityType[:] = etype
    eng.upload(eng.mask("T.entityType"))
    for _ in range(5):
        eng.run(1); eng.system_boids(1.0)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = ts = 0.0
    for _ in range(a.frames):
        e[0].record(stream); eng.run(1); e[1].record(stream); eng.system_boids(1.0); e[2].record(stream)
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); ts += e[1].elapsed_time(e[2])
    st = eng.stats()
    row_bytes = N * (1 + a.max_neighbors) * 8
    print(json.dumps({"workload": f"boids {a.prey} prey + {a.pred} predators, world {cfg['worldWidth']:.0f}x{cfg['worldHeight']:.0f}, "
                                  f"maxNeighbors {a.max_neighbors}, S=1", "entities": N,
                      "frame_ms": tf / a.frames, "device_tick_ms": ts / a.frames,
                      "kbar": st["neighborsTotal"] / max(1, st["activeInGrid"]),
                      "entity_substeps_per_s_with_device_tick": N * a.frames / ((tf + ts) * 1e-3),
                      "neighbor_row_bytes_a_host_tick_would_fetch_per_frame": row_bytes,
                      "pcie_ms_for_those_rows_at_50GBps": row_bytes / 50e9 * 1e3}))
    eng.close()
