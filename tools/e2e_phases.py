"""Where the end-to-end frame goes on N ranks: wall time of eng.step (upload + frame + download)
and of the slab exchange, per rank (the GPU boxes of this pool are single-NUMA-node VMs: nothing to bind).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/e2e_phases.py"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--workload", default="config4")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from multithreadedgameengine_b200.engine import GameEngine
    from multithreadedgameengine_b200.slabs import SlabEngine, plan_slabs
    cfg, cols = bench.workload(args.workload, None)
    stream = torch.cuda.Stream()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # PCIe with every rank copying at once: is the link per GPU or shared?
    nb = 128 << 20
    dbuf = torch.empty(nb, dtype=torch.uint8, device="cuda")
    hbuf = torch.empty(nb, dtype=torch.uint8).pin_memory()
    hbuf2 = torch.empty(nb, dtype=torch.uint8).pin_memory()
    dbuf2 = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s2 = torch.cuda.Stream()
    pcie = {}
    for name in ("d2h", "h2d", "both"):
        for rep in range(3):
            sync_all()
            t0 = time.perf_counter()
            if name in ("d2h", "both"):
                hbuf.copy_(dbuf, non_blocking=True)
            if name in ("h2d", "both"):
                with torch.cuda.stream(s2):
                    dbuf2.copy_(hbuf2, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        pcie[name] = round(nb * (2 if name == "both" else 1) / dt / 1e9, 1)
    del dbuf, hbuf, hbuf2, dbuf2
    with torch.cuda.stream(stream):
        if world > 1:
            sl = SlabEngine(cfg, cols, rank, world, device=local, stream=stream.cuda_stream, plan=plan_slabs(cfg, cols, world))
            eng = sl.eng
        else:
            eng = GameEngine(cfg, device=local, stream=stream.cuda_stream, host_neighbor_rows=False)
            eng.load_columns(cols)
            sl = None
        up = eng.mask("RB.ax", "RB.ay")
        down = eng.mask("T.x", "T.y", "RB.vx", "RB.vy", "RB.velocityAngle", "RB.speed")
        rows = []
        for f in range(3 + args.frames):
            if world > 1:
                torch.cuda.synchronize()
                dist.barrier()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.step(1.0, up, 0)
            t1 = time.perf_counter()
            eng.step(1.0, 0, down)
            t2 = time.perf_counter()
            eng.step(1.0, up, down)
            t3 = time.perf_counter()
            if sl:
                sl.exchange_dist()
            t4 = time.perf_counter()
            torch.cuda.synchronize()
            t5 = time.perf_counter()
            eng.run(1)
            torch.cuda.synchronize()
            t6 = time.perf_counter()
            if sl:
                sl.exchange_dist()
                torch.cuda.synchronize()
            if f >= 3:
                rows.append([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5])
        r = np.array(rows).mean(0) * 1e3
        # the bench's e2e loop: free running, ranks meet only through the exchange
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.frames):
            eng.step(1.0, up, down)
            if sl:
                sl.exchange_dist()
        sync_all()
        free_ms = (time.perf_counter() - t0) * 1e3 / args.frames
        n_host = eng.totalEntityCount if world == 1 else sl.status()["top"]
        out = {"rank": rank, "entities_moved": int(n_host), "pcie_GBps_all_ranks_at_once": pcie, "e2e_loop_ms": round(free_ms, 3),
               "ms": {"step(up only)": round(r[0], 3), "step(down only)": round(r[1], 3), "step(up+down)": round(r[2], 3),
                      "exchange call": round(r[3], 3), "exchange wait": round(r[4], 3), "run(1) no copies": round(r[5], 3)},
               "GBps_down": round(24 * n_host / (r[1] - r[5]) / 1e6, 1) if r[1] > r[5] else None}
        print(json.dumps(out), flush=True)
        (sl.close if sl else eng.close)()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
