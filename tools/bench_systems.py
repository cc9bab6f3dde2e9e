#!/usr/bin/env python
"""Device-side consumers timed beside their CPU restatements.

Scene: config-4 style clustered balls at --entities, maxCollisionPairs large enough to hold the
whole pair list, so the Enter/Stay/Exit diff runs over every colliding pair.  Device times are
CUDA events on the context's stream around the C-ABI call with NULL host outputs (no D2H);
"e2e" adds the D2H of the results.  The CPU figures are the C oracle (one thread) on the same
inputs — the reference does this work in JavaScript with Sets of Cantor keys."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import numpy as np
import torch

from multithreadedgameengine_b200 import binding as B, scenes
from multithreadedgameengine_b200.engine import GameEngine
from oracle.oracle_c import CollisionEventsC, screen_visibility_c

ap = argparse.ArgumentParser()
ap.add_argument("--entities", type=int, default=2_000_000)
ap.add_argument("--frames", type=int, default=10)
a = ap.parse_args()
scale = (a.entities / 16_000_000) ** 0.5
W, H = round(65536 * scale / 16) * 16.0, round(32768 * scale / 16) * 16.0
cfg, cols = scenes.balls_synthetic(a.entities, (W, H), 16.0, 64, 2, (2.0, 6.0), 16.0, 1234, clusters=max(1, int(256 * scale * scale)),
                                   cluster_sigma=400.0, cluster_edge="reflect")
cfg["physics"]["maxCollisionPairs"] = int(2.5 * a.entities)
N = cfg["entityCount"]
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    eng = GameEngine(cfg, stream=stream.cuda_stream, host_neighbor_rows=False)
    eng.load_columns(cols)
    L = B.lib()
    cam = B.Camera(1.0, W * 0.25, H * 0.25, W * 0.5, H * 0.5)
    cnt = B.CollisionEventCounts()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_ev = t_vis = 0.0
    pairs = entered = exited = 0
    for f in range(a.frames + 2):
        eng.run(1)
        ev[0].record(stream)
        B.check(eng.ctx, L.weed_system_collision_events(eng.ctx, 0, C.byref(cnt), None, None))
        ev[1].record(stream)
        B.check(eng.ctx, L.weed_system_screen_visibility(eng.ctx, C.byref(cam), None, None, None))
        ev[2].record(stream)
        torch.cuda.synchronize()
        if f >= 2:
            t_ev += ev[0].elapsed_time(ev[1]); t_vis += ev[1].elapsed_time(ev[2])
            pairs += cnt.pairs; entered += cnt.entered; exited += cnt.exited
    # end to end (results to host) and the CPU restatement on the same frames
    ora = CollisionEventsC()
    e2e_ev = e2e_vis = cpu_ev = cpu_vis = 0.0
    sx = np.zeros(N, np.float32); sy = np.zeros(N, np.float32); on = np.zeros(N, np.uint8)
    eng.collision_events(forget_previous=True)
    ora.process(eng.collisionData)
    for f in range(3):
        eng.run(1)
        t0 = time.perf_counter(); e = eng.collision_events(); t1 = time.perf_counter()
        eng.screen_visibility(1.0, W * 0.25, H * 0.25, W * 0.5, H * 0.5); t2 = time.perf_counter()
        e2e_ev += t1 - t0; e2e_vis += t2 - t1
        eng.download(eng.mask("T.x", "T.y"))
        t0 = time.perf_counter(); calls = ora.process(eng.collisionData); t1 = time.perf_counter()
        screen_visibility_c(eng.col["T.active"], eng.col["T.x"], eng.col["T.y"], 1.0, W * 0.25, H * 0.25, W * 0.5, H * 0.5, sx, sy, on)
        t2 = time.perf_counter()
        cpu_ev += t1 - t0; cpu_vis += t2 - t1
        assert eng.collision_callbacks(e) == calls
        assert np.array_equal(on, eng._onScreen)
    print(json.dumps({
        "workload": f"clustered balls, {N} entities, world {W:.0f}x{H:.0f}, S=2, maxCollisionPairs {cfg['physics']['maxCollisionPairs']}",
        "pairs_per_frame": pairs / a.frames, "entered_per_frame": entered / a.frames, "exited_per_frame": exited / a.frames,
        "collision_events_device_ms": t_ev / a.frames, "collision_events_e2e_ms": e2e_ev / 3 * 1e3,
        "collision_events_cpu_oracle_ms": cpu_ev / 3 * 1e3,
        "pairs_per_s_device": pairs / (t_ev * 1e-3),
        "screen_visibility_device_ms": t_vis / a.frames, "screen_visibility_e2e_ms": e2e_vis / 3 * 1e3,
        "screen_visibility_cpu_oracle_ms": cpu_vis / 3 * 1e3,
        "visibility_GBps_device": N * (16 + 1 + 9) / (t_vis / a.frames * 1e-3) / 1e9}))
    eng.close()
