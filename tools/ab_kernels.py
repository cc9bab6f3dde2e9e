#!/usr/bin/env python
"""A/B measurement of kernel generations on one GPU (diagnostic; not part of bench.py).

For every variant (WEED_FLAG_K6_TILE or not; experimental builds through --lib) the same seeded scene is
run for --warmup + --frames frames with per-span CUDA-event timing (WEED_FLAG_KERNEL_TIMING);
the mean span times over the timed frames are printed, together with a hash of the final state,
of collisionData and of a sample of API rows, so that a variant that is faster but different is
caught in the same run.

  python tools/ab_kernels.py --workload config4 [--entities N] [--variants 22,12,21,11]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SPANS = ["k_cell_key", "k_cell_scan", "k_scatter+rank", "k_build_slots+prep", "K4", "K4b+c", "K6(all sweeps)", "WB+K7"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config4")
    ap.add_argument("--entities", type=int, default=None)
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--variants", default="22", help="comma list of <K4 form><K6 form>: 2 = current, T = TMA tiles (K6 only), e.g. 22,2T (the round-1 forms were retired with the cap-path redesign; their numbers are in profiles/README.md)")
    ap.add_argument("--rows", type=int, default=200_000, help="API rows hashed per sample block (3 blocks)")
    ap.add_argument("--no-rows", action="store_true", help="WEED_FLAG_NO_NEIGHBOR_ROWS: what the scan costs without the API rows")
    ap.add_argument("--lib", default=None, help="tag of an experimental build (tools/build_variant.sh) to load instead of the product library")
    args = ap.parse_args()

    from multithreadedgameengine_b200 import binding as B
    if args.lib:          # before anything loads the library
        B.LIB_PATH = os.path.join(os.path.dirname(B.LIB_PATH), "exp", f"libweedgpu_{args.lib}.so")
    import __graft_entry__ as entry
    entry.build()
    from bench import workload
    from multithreadedgameengine_b200.engine import GameEngine

    cfg, cols = workload(args.workload, args.entities)
    N = cfg["entityCount"]
    S = cfg["physics"]["subStepCount"]
    results = []
    for v in args.variants.split(","):
        flags = B.FLAG_KERNEL_TIMING
        flags |= {"2": 0}[v[0]]
        flags |= {"2": 0, "T": B.FLAG_K6_TILE}[v[1]]
        if args.no_rows:
            flags |= B.FLAG_NO_NEIGHBOR_ROWS
        eng = GameEngine(cfg, flags=flags, host_neighbor_rows=False)
        eng.load_columns(cols)
        eng.run(args.warmup)
        acc = np.zeros(8)
        dev = 0.0
        for _ in range(args.frames):
            eng.run(1)
            st = eng.stats()
            acc += np.array(st["ms"][:8])
            dev += st["ms"][8]
        ms = acc / args.frames
        eng.download(B.COLS_INPUT_ALL | B.COL_COLLISIONS)
        h = hashlib.sha256()
        for k in ("T.x", "T.y", "RB.px", "RB.py", "RB.vx", "RB.vy", "RB.speed", "RB.collisionCount", "RB.velocityAngle"):
            h.update(np.ascontiguousarray(eng.col[k]).view(np.uint8).tobytes())
        npairs = int(eng.collisionData[0])
        h.update(eng.collisionData[:1 + 2 * npairs].tobytes())
        state_hash = h.hexdigest()[:16]
        # sampled API rows: three blocks (start, middle, end), only the 1 + count words of each row
        stride = 1 + eng.maxNeighbors
        hr = hashlib.sha256()
        nrows = min(args.rows, N)
        act = eng.col["T.active"]
        for first in ([] if args.no_rows else sorted({0, max(0, N // 2 - nrows // 2), max(0, N - nrows)})):
            nd = np.empty(nrows * stride, np.int32)
            dd = np.empty(nrows * stride, np.float32)
            B.check(eng.ctx, B.lib().weed_fetch_neighbors_to(eng.ctx, first, nrows, nd.ctypes.data, dd.ctypes.data))
            nd = nd.reshape(nrows, stride)
            dd = dd.reshape(nrows, stride)
            live = act[first:first + nrows] != 0
            cnt = np.where(live, nd[:, 0], 0)
            keep = np.arange(stride)[None, :] <= cnt[:, None]
            keep &= live[:, None]
            hr.update(nd[keep].tobytes())
            hr.update(dd[keep].tobytes())
        st = eng.stats()
        res = {"lib": args.lib or "product", "variant": f"K4 v{v[0]} / K6 v{v[1]}", "entities": N, "frames": f"{args.warmup}..{args.warmup + args.frames}",
               "span_ms": {n: round(float(x), 4) for n, x in zip(SPANS, ms)},
               "K6_ms_per_sweep": round(float(ms[6]) / S, 4), "sum_ms": round(float(ms.sum()), 4),
               "device_frame_ms": round(dev / args.frames, 4),
               "kbar": st["neighborsTotal"] / max(1, st["activeInGrid"]), "capped_rows": st["cappedRows"],
               "explicit_pairs": st["explicitPairs"], "xover_rows": int(st["ms"][9]), "collision_pairs": st["collisionPairs"],
               "state_hash": state_hash, "rows_hash": hr.hexdigest()[:16]}
        print(json.dumps(res), flush=True)
        results.append(res)
        eng.close()
    same = len({(r["state_hash"], r["rows_hash"]) for r in results}) == 1
    print(json.dumps({"all_variants_bit_identical": same}))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
