#!/usr/bin/env python
"""Device time of the slab phases on ONE GPU: all slabs of a world in one process (SlabGroup in its
"nccl-buffers" mode: weed_slab_pack / weed_slab_apply with a device-to-device copy in between — the
staging path, whose pack and apply kernels are the ones the peer-to-peer transport runs too),
CUDA events on the shared stream around frame / pack / apply of slab 0.  Single process, so it
may run under ncu."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from multithreadedgameengine_b200.slabs import SlabGroup

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config4")
ap.add_argument("--entities", type=int, default=4_000_000)
ap.add_argument("--world", type=int, default=2)
ap.add_argument("--frames", type=int, default=5)
a = ap.parse_args()
cfg, cols = bench.workload(a.workload, a.entities)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    g = SlabGroup(cfg, cols, a.world, mode="nccl-buffers", stream=stream.cuda_stream)
    for _ in range(2):
        g.step()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t = [0.0, 0.0, 0.0]
    s0 = g.slabs[0]
    for _ in range(a.frames):
        for s in g.slabs[1:]:
            s.run()
        torch.cuda.synchronize()
        ev[0].record(stream); s0.run(); ev[1].record(stream); s0.pack(); ev[2].record(stream)
        for s in g.slabs[1:]:
            s.pack()
        torch.cuda.synchronize()
        for r, s in enumerate(g.slabs):
            if r > 0:
                s.recv_low.copy_(g.slabs[r - 1].send_high)
            if r + 1 < len(g.slabs):
                s.recv_high.copy_(g.slabs[r + 1].send_low)
        torch.cuda.synchronize()
        ev[2].record(stream); s0.apply(); ev[3].record(stream)
        for s in g.slabs[1:]:
            s.apply()
        torch.cuda.synchronize()
        t[0] += ev[0].elapsed_time(ev[1]); t[1] += ev[1].elapsed_time(ev[2]) if False else 0.0
        t[2] += ev[2].elapsed_time(ev[3])
    # pack timed separately (ev[2] was re-recorded above)
    tp = 0.0
    for _ in range(a.frames):
        g.step()
        torch.cuda.synchronize()
        s0.run(); torch.cuda.synchronize()
        ev[0].record(stream); s0.pack(); ev[1].record(stream); torch.cuda.synchronize()
        tp += ev[0].elapsed_time(ev[1])
        for s in g.slabs[1:]:
            s.run(); s.pack()
        torch.cuda.synchronize()
        for r, s in enumerate(g.slabs):
            if r > 0:
                s.recv_low.copy_(g.slabs[r - 1].send_high)
            if r + 1 < len(g.slabs):
                s.recv_high.copy_(g.slabs[r + 1].send_low)
        torch.cuda.synchronize()
        for s in g.slabs:
            s.apply()
        torch.cuda.synchronize()
    print(f"slab 0 of {a.world} ({a.entities} entities in the world): frame {t[0] / a.frames:.3f} ms, pack {tp / a.frames:.3f} ms, "
          f"apply {t[2] / a.frames:.3f} ms; {s0.status()}")
    g.close()
