#!/usr/bin/env python
"""Wall-clock breakdown of one slab frame (run under torchrun): frame kernels, pack, count
exchange, record exchange, apply."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from multithreadedgameengine_b200.slabs import SlabEngine, exchange_fixed, plan_slabs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg, cols = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "config4", int(sys.argv[2]) if len(sys.argv) > 2 else None)
plan = plan_slabs(cfg, cols, world)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    sl = SlabEngine(cfg, cols, rank, world, device=local, stream=stream.cuda_stream, plan=plan)
    for _ in range(3):
        sl.step_dist()
    T = np.zeros(5)
    frames = 10
    for _ in range(frames):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        sl.run(); sl.eng.sync()
        t1 = time.perf_counter()
        sl.pack(); sl.eng.sync()
        t2 = time.perf_counter()
        exchange_fixed(torch, rank, world, sl.send_low, sl.send_high, sl.recv_low, sl.recv_high)
        torch.cuda.current_stream().synchronize()
        t3 = time.perf_counter()
        sl.apply(); sl.eng.sync()
        t4 = time.perf_counter()
        T += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0]
    T *= 1e3 / frames
    st = sl.status()
    print(f"rank {rank}: frame {T[0]:.3f} ms, pack {T[1]:.3f}, exchange {T[2]:.3f}, apply {T[3]:.3f}, total {T[4]:.3f} "
          f"(each phase synchronised for the measurement); quota {sl.quota}, {st}", flush=True)
    sl.close()
dist.destroy_process_group()
