#!/usr/bin/env python
"""bench.py — headline benchmark of the spatial+physics hot path (BASELINE.json metric:
entity-substep updates/sec, grid + neighbors + Verlet + collide, at 16M entities).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" is one lockstep frame (grid rebuild, neighbor query, integration, S constraint
substeps, write-back) over the whole scene.  Prints ONE JSON line (rank 0).

  value     device-resident throughput: A * S * K / t, t = CUDA-event time of K frames, inputs
            already in HBM (A = active entities, S = subStepCount)
  e2e       the same frames through GameEngine.step() with HOST component buffers: every step
            uploads the columns tick() writes (ax, ay) and downloads the columns the
            renderer/logic read (x, y, vx, vy, velocityAngle, speed)
  roofline  dominant kernel, algorithmic bytes / its CUDA-event time, against the
            measured HBM copy peak of MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the CPU restatement of the reference's two JS workers
            (oracle/, "port": no JS engine exists in this image) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "entity_substep_updates_per_sec"
UNIT = "entity-substeps/s"

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
# (profiles/), config4 16M on one B200; None = not captured for this kernel.
TRAFFIC_FROM_NCU = {          # bytes per launch, profiles/r2_ncu_config4_16M_final_frame_summary.md (frame 6 of the scene)
    "k_neighbors2": 4.742e9,                      # 1.695 read + 3.047 written
    "k_sweep": 2.32e9,                            # first sweep 2.05 + 0.41, last sweep 1.66 + 0.52
    "k_build_slots+k_slot_prep": 3.226e9,         # 0.80 + 1.01 and 0.51 + 0.91
}

KERNEL_NAMES = ["k_cell_key", "k_cell_scan", "k_scatter_ids+k_sort_big_cells+k_slot_rank", "k_build_slots+k_slot_prep", "k_neighbors2",
                "k_beyond_cap+k_back_alloc+k_back_write+k_back_sort+k_sort_lists", "k_sweep", "k_writeback+k_pair_scan+k_pair_emit"]
# ("k_sweep": one launch of k_sweep with its k_sweep_heavy branch beside it)
# algorithmic bytes per active entity of each timed span (SURVEY 8 d); spans without compulsory
# traffic of their own (the cap path, the pair log) can never be the "dominant kernel" of the roofline line
def span_bytes(kbar):
    return [13.0, 0.0, 8.0, 82.0, 24.0 + 8.0 * (1.0 + kbar), 0.0, 34.0 + 4.0 * (1.0 + kbar), 0.0]


def workload(name, n_override=None):
    from multithreadedgameengine_b200 import scenes
    full = {"config3": 1_000_000, "config4": 16_000_000, "config5": 128_000_000}
    if name == "config5":                 # 128M: build the columns frugally (float32 straight away)
        n_override = n_override or full[name]
    if name in full:
        n = n_override or full[name]
        if n == full[name]:
            return getattr(scenes, name)()
        return scenes.scaled(name, n)
    if name == "config4_reflect":      # diagnostic: same scene, clusters reflected at the walls instead of clipped
        return scenes.config4(cluster_edge="reflect")
    if name == "config1":
        return scenes.balls_readme()
    if name == "config1b":
        return scenes.balls_demo()
    if name == "config2":
        return scenes.boids()
    raise SystemExit(f"unknown workload {name}")


def full_config(name, n_override=None):
    """The config block of the FULL workload (what the GPU arm runs), without building the scene."""
    from multithreadedgameengine_b200 import scenes
    if name not in scenes._FULL:
        return None
    n, (W, H), cs, M, S, _, _, _ = scenes._FULL[name]
    if n_override and n_override != n:
        return None
    return dict(entityCount=n + 1, worldWidth=W, worldHeight=H, spatial=dict(cellSize=cs, maxNeighbors=M),
                physics=dict(subStepCount=S))


def describe(name, cfg):
    return {"workload": f"{name}: synthetic balls, {cfg['entityCount'] - 1} entities + Mouse, world "
                        f"{cfg['worldWidth']:.0f}x{cfg['worldHeight']:.0f}, cellSize {cfg['spatial']['cellSize']:g}, "
                        f"maxNeighbors {cfg['spatial']['maxNeighbors']}, subStepCount {cfg['physics']['subStepCount']}",
            "entities": cfg["entityCount"], "subStepCount": cfg["physics"]["subStepCount"]}


def algorithmic_bytes(kbar, S):
    """SURVEY §8(d) / DESIGN.md: compulsory bytes per active entity per frame."""
    per_kernel = {
        "k_cell_key": 13.0, "k_cell_scan+k_scatter_ids": 8.0,
        "k_neighbors": 24.0 + 8.0 * (1.0 + kbar),
        "k_build_slots (integrate+derived)": 64.0 + 18.0,
        "k_substep": S * (34.0 + 4.0 * (1.0 + kbar)),
    }
    return 127.0 + 8.0 * (1.0 + kbar) + S * (34.0 + 4.0 * (1.0 + kbar)), per_kernel


def state_checksum(gids, vals):
    """Order-independent 64-bit checksum of (gid, x, y, px, py, collisionCount) over a set of entities."""
    M = np.uint64
    with np.errstate(over="ignore"):
        h = gids.astype(M) * M(0x9E3779B97F4A7C15)
        for k, c in zip(("T.x", "T.y", "RB.px", "RB.py"), (0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x27D4EB2F165667C5, 0x85EBCA77C2B2AE63)):
            h = (h ^ (np.ascontiguousarray(vals[k]).view(np.uint32).astype(M) * M(c))) * M(0xFF51AFD7ED558CCD)
            h ^= h >> M(33)
        h = (h ^ vals["RB.collisionCount"].astype(M)) * M(0xC4CEB9FE1A85EC53)
        h ^= h >> M(29)
        return int(h.sum(dtype=M))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        mx = [float(r[2]) for r in self.rows if len(r) >= 8]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": reasons, "samples": len(sm)}


def oracle_for(cfg, cols):
    from multithreadedgameengine_b200.engine import PHYSICS_DEFAULTS
    from oracle.oracle_c import OracleC
    p = dict(PHYSICS_DEFAULTS)
    p.update({k: v for k, v in cfg["physics"].items() if k not in ("gravity", "maxCollisionPairs")})
    g = cfg["physics"].get("gravity", {"x": 0, "y": 0})
    o = OracleC(cfg["entityCount"], cfg["worldWidth"], cfg["worldHeight"], cfg["spatial"]["cellSize"],
                cfg["spatial"]["maxNeighbors"], int(cfg["physics"].get("maxCollisionPairs") or 10000), cfg.get("seed", 1.0),
                dict(subStepCount=p["subStepCount"], boundaryElasticity=p["boundaryElasticity"],
                     collisionResponseStrength=p["collisionResponseStrength"], verletDamping=p["verletDamping"],
                     minSpeedForRotation=p["minSpeedForRotation"], gravityX=g["x"], gravityY=g["y"]))
    o.load(cols)
    return o


def cpu_reference(name, steps, warmup, sample_entities, lockstep_frames=None, budget_s=None):
    """The reference's CPU structure: ONE spatial worker thread + ONE physics worker thread,
    free-running on shared buffers (gameEngine.js:978-996, AbstractWorker.js:114-146),
    restated in C (oracle/weed_oracle.c).  sample_entities=None runs the workload at its full
    size; a number runs a scaled instance at the same entity density (and says so)."""
    cfg, cols = workload(name, sample_entities)
    o = oracle_for(cfg, cols)
    S = cfg["physics"]["subStepCount"]
    active = int(cols["T.active"].sum())
    asked = (steps, warmup)
    if budget_s and warmup + steps > 1:
        # a full-size frame costs seconds on the host: time the first warm-up frame and, if the whole run would not
        # fit the budget, shorten the warm-up first (to one frame), then the timed frames (to two at least)
        ts1, tp1 = o.bench(1, 1.0, freerun=True)
        per1 = max(ts1, tp1)
        fit = max(3, int(budget_s / max(per1, 1e-9)))          # frames the budget pays for, the first one included
        if 1 + max(0, warmup - 1) + steps > fit:
            warmup = 1
            steps = max(2, min(steps, fit - 1))
        if warmup > 1:
            o.bench(warmup - 1, 1.0, freerun=True)
    elif warmup:
        o.bench(warmup, 1.0, freerun=True)
    ts, tp = o.bench(steps, 1.0, freerun=True)
    t = max(ts, tp)
    nl = max(1, steps // 2) if lockstep_frames is None else lockstep_frames
    lock_note = ""
    if nl:
        ls, lp = o.bench(nl, 1.0, freerun=False)
        lock = (ls + lp) / nl
        lock_note = f"; single-thread lockstep {lock * 1e3:.1f} ms/frame = {active * S / lock:.3e} {UNIT}"
    sampled = sample_entities is not None and full_config(name, sample_entities) is None
    what = (f"{name} at the same density scaled to {cfg['entityCount'] - 1} entities" if sampled
            else f"{name} at full size ({cfg['entityCount'] - 1} entities)")
    return {
        "value": active * S * steps / t, "unit": UNIT, "cores": 2, "kind": "port",
        "sample": f"{what}, {steps} frames after {warmup} warm-up frames"
                  + (f" (asked for {asked[0]} after {asked[1]}: shortened to fit --ref-budget-s)" if (steps, warmup) != asked else "") + ", "
                  f"2 free-running threads (spatial {ts / steps * 1e3:.1f} ms/frame, physics {tp / steps * 1e3:.1f} ms/frame)" + lock_note,
        "sampled": sampled, "sample_entities": cfg["entityCount"], "steps_run": steps, "warmup_run": warmup,
        "host_cores_available": len(os.sched_getaffinity(0)),
    }, cfg, t / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    # the reference arm runs the SAME workload as the GPU arm (config4 at 16M: ~6 s per frame on two host
    # threads); --cpu-sample N runs a scaled instance instead and is then labelled as such
    sample = args.cpu_sample if args.cpu_sample else args.entities
    base, cfg, per = cpu_reference(args.workload, args.steps, args.warmup, sample, lockstep_frames=0, budget_s=args.ref_budget_s)
    conf = describe(args.workload, cfg)
    if base["sampled"]:
        conf["workload"] += " [scaled sample of the workload: --cpu-sample]"
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": conf,
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0, "steps_run": base["steps_run"], "warmup_run": base["warmup_run"]}
    line["config"]["note"] = ("CPU restatement (C port) of the reference's JS spatial+physics workers; no JavaScript engine "
                              "exists in this image, so the original cannot be executed")
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from multithreadedgameengine_b200 import binding as B
    from multithreadedgameengine_b200.engine import GameEngine
    from multithreadedgameengine_b200.slabs import SlabEngine, plan_slabs, replan_from_times, row_costs

    name = args.workload
    cfg, cols = workload(name, args.entities)      # every rank builds the same seeded scene
    S = cfg["physics"]["subStepCount"]
    N = cfg["entityCount"]
    active = int(cols["T.active"].sum())
    plan = plan_slabs(cfg, cols, world) if world > 1 else None
    stream = torch.cuda.Stream()
    balance_note = "cost model of the start scene"
    if world > 1 and args.autobalance:
        # measured-feedback balancing (outside the timed region, like an autotuning pass): run a
        # few frames from the start scene, gather every slab's kernel time, move the cuts, restart
        row_weight = row_costs(cfg, cols)[0]
        for it in range(args.autobalance):
            with torch.cuda.stream(stream):
                sl = SlabEngine(cfg, cols, rank, world, device=local, stream=stream.cuda_stream, plan=plan, transport=args.transport)
                for _ in range(args.warmup + args.steps // 2):     # the scene evolves: balance for the frames that get timed
                    sl.step_dist()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                tsum = 0.0
                for _ in range(4):
                    torch.cuda.synchronize()
                    a0.record(stream)
                    sl.run()
                    a1.record(stream)
                    torch.cuda.synchronize()
                    tsum += a0.elapsed_time(a1)
                    sl.exchange_dist()
                sl.close()
            tt = torch.tensor([tsum / 4], device="cuda", dtype=torch.float64)
            allt = [torch.zeros_like(tt) for _ in range(world)]
            dist.all_gather(allt, tt)
            times = [float(t.item()) for t in allt]
            plan = (replan_from_times(plan[0], times, row_weight), plan[1])
        balance_note = (f"cost model + {args.autobalance} measured-feedback re-plans before the timed run, each measured "
                        f"{args.warmup + args.steps // 2} frames into the scene"
                        + (f"; during the run every cut follows the measured load by up to {args.balance_rows} rows per frame (weed_slab_balance)"
                           if args.balance_rows else " (static cuts during the run)"))

    def make(flags=0):
        if world == 1:
            e = GameEngine(cfg, device=local, flags=flags, stream=stream.cuda_stream, host_neighbor_rows=False)
            e.load_columns(cols)
            return e, e
        sl = SlabEngine(cfg, cols, rank, world, device=local, flags=flags, stream=stream.cuda_stream, plan=plan,
                        balance_rows=args.balance_rows, transport=args.transport)
        return sl, sl.eng

    def frames(obj, k):
        if world == 1:
            obj.run(k)
        else:
            for _ in range(k):
                obj.step_dist()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    with torch.cuda.stream(stream):
        obj, eng = make()
        # ---- device-resident timing ---------------------------------------------------------
        frames(obj, args.warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            t_wall = time.perf_counter()
            e0.record(stream)
            frames(obj, args.steps)
            e1.record(stream)
            barrier()
            wall = time.perf_counter() - t_wall
        # one rank alone is pure device work (events); with slabs the exchange has host-side
        # phases, so the wall clock of the same region is the honest figure
        ms = max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3) if world > 1 else e0.elapsed_time(e1))
        st = eng.stats()
        slab_st = obj.status() if world > 1 else None      # synchronises; raises on quota/table overflow
        owned_total = active if world == 1 else int(round(sum_over_ranks(slab_st["owned"])))
        kbar = sum_over_ranks(st["neighborsTotal"]) / max(1.0, sum_over_ranks(st["activeInGrid"]))
        value = owned_total * S * args.steps / (ms * 1e-3)
        halo_frac = 0.0 if world == 1 else sum_over_ranks(st["activeInGrid"]) / max(1, owned_total) - 1.0
        xbytes = 0 if world == 1 else int(sum_over_ranks(obj.exchange_bytes_per_frame))

        # ---- per-kernel CUDA-event timing (direct launches; same frames, same state) ----------
        if args.quick:
            kms = np.full(8, np.nan)       # per-kernel split not measured in --quick runs
            local_active, kbar_t = st["activeInGrid"], st["neighborsTotal"] / max(1, st["activeInGrid"])
        else:
            obj_t, eng_t = make(B.FLAG_KERNEL_TIMING)
            frames(obj_t, args.warmup)         # the SAME frame window as the headline: frames warmup .. warmup + steps of the scene
            acc = np.zeros(8)
            for _ in range(args.steps):
                frames(obj_t, 1)
                acc += np.array(eng_t.stats()["ms"][:8])
            kms = acc / args.steps
            kms[6] /= S                    # per LAUNCH of k_substep (the frame runs it S times)
            st_t = eng_t.stats()
            local_active = st_t["activeInGrid"]
            kbar_t = st_t["neighborsTotal"] / max(1, local_active)
            (obj_t.close if world > 1 else eng_t.close)()
        slab_rows = None
        if world > 1:
            t = torch.tensor([slab_st["rowBegin"], slab_st["rowEnd"], slab_st["cutMoves"]], device="cuda", dtype=torch.int64)
            allr = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allr, t)
            slab_rows = [[int(v) for v in x.tolist()] for x in allr]      # [rowBegin, rowEnd, cutMoves] per rank after the timed run
        # every rank's frame kernels (sum of the spans, k_substep counted S times): the slab balance
        kernel_ms_per_rank = None
        if world > 1 and not args.quick:
            mine = float(np.nansum(kms) + kms[6] * (S - 1))
            t = torch.tensor([mine], device="cuda", dtype=torch.float64)
            allk = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allk, t)
            kernel_ms_per_rank = [round(float(x.item()), 3) for x in allk]
        F, per_kernel = algorithmic_bytes(kbar_t, S)
        sb = span_bytes(kbar_t)
        # dominant kernel = the longest launch among the spans that move compulsory bytes
        top = 4 if args.quick else int(np.argmax(np.where(np.array(sb) > 0, np.nan_to_num(kms), -1.0)))
        alg = sb[top]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg * local_active / (kms[top] * 1e-3) / 1e9     # this rank's kernel over the entities it processes
        if args.quick:                 # only the whole-frame figure is meaningful
            achieved = F * owned_total * args.steps / (ms * 1e-3) / 1e9 / world
        frame_gbps = F * owned_total * args.steps / (ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "whole frame (per-kernel timing skipped: --quick)" if args.quick else KERNEL_NAMES[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak,
                    "traffic": TRAFFIC_FROM_NCU.get(KERNEL_NAMES[top]) if (world == 1 and name == "config4" and not args.entities) else None,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                    "algorithmic_bytes_per_entity": alg, "kernel_ms": None if args.quick else float(kms[top]), "entities_per_launch": int(local_active),
                    "whole_frame": {"bytes_per_entity_frame": F, "achieved_GBps_all_gpus": frame_gbps,
                                    "frac_of_n_gpus_peak": frame_gbps / (peak * world)}}

        # ---- end to end through the host-facing API -----------------------------------------------
        up = eng.mask("RB.ax", "RB.ay")
        down = eng.mask("T.x", "T.y", "RB.vx", "RB.vy", "RB.velocityAngle", "RB.speed")
        e2e_steps = max(1, args.steps // 4) if args.quick else args.steps
        e2e_balance = None
        if world > 1 and args.e2e_balance and not args.quick:
            # The end-to-end frame is bound by each GPU's host link, and the links of one box are not equal
            # (tools/e2e_phases.py).  Re-plan the slabs for THIS leg on the measured end-to-end step time of every
            # rank (outside the timed region, like the kernel-time re-plans of the device-resident leg), then time
            # the same frame window with static cuts.
            from multithreadedgameengine_b200.slabs import halo_rows, replan_by_rank_speed
            obj.close()
            e_plan, row_weight = plan, row_costs(cfg, cols)[0]
            rank_ms = []
            for it in range(args.e2e_balance + 1):
                obj = SlabEngine(cfg, cols, rank, world, device=local, stream=stream.cuda_stream, plan=e_plan,
                                 balance_rows=0, transport=args.transport)
                eng = obj.eng
                frames(obj, args.warmup + args.steps)            # where the device-resident leg left the scene
                if it == args.e2e_balance:
                    break
                ts = []
                for _ in range(6):
                    barrier()
                    t0 = time.perf_counter()
                    eng.step(1.0, up, down)
                    ts.append((time.perf_counter() - t0) * 1e3)
                    obj.exchange_dist()
                barrier()
                obj.close()
                t = torch.tensor([float(np.median(ts[1:]))], device="cuda", dtype=torch.float64)
                allt = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(allt, t)
                rank_ms = [round(float(x.item()), 3) for x in allt]
                blocks = replan_by_rank_speed(e_plan[0], rank_ms, row_weight)
                e_plan = (blocks, halo_rows(cfg, cols, blocks))
            e2e_balance = {"passes": args.e2e_balance, "step_ms_per_rank_before_last_replan": rank_ms,
                           "slab_rows_begin_end": [list(b) for b in e_plan[0]], "halo_rows": e_plan[1],
                           "what": "slabs re-planned on every rank's measured GameEngine.step time (copies included): the host links of "
                                   "one box differ, so kernel-time balance leaves the ranks on slow links late; static cuts while timed"}

        def e2e_frame():
            eng.step(1.0, up, down)
            if world > 1:
                obj.exchange_dist()

        for _ in range(1 if args.quick else min(3, args.warmup)):
            e2e_frame()
        barrier()
        t_wall = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_frame()
        barrier()
        ms_e2e = max_over_ranks((time.perf_counter() - t_wall) * 1e3)
        n_host = eng.totalEntityCount if world == 1 else obj.status()["top"]     # a slab's weed_step moves its table up to `top`
        e2e = {"value": owned_total * S * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(sum_over_ranks(8 * n_host)), "d2h_bytes_per_step": int(sum_over_ranks(24 * n_host)),
               "ms_per_step": ms_e2e / e2e_steps,
               "api": "GameEngine.step(dtRatio, upload=ax|ay, download=x|y|vx|vy|velocityAngle|speed) -> weed_step"
                      + ("; + SlabEngine.exchange_dist per frame" if world > 1 else ""),
               "slab_balance": e2e_balance}
        launches = st["kernelLaunchesPerStep"] * args.steps + (6 * args.steps if world > 1 else 0)
        (obj.close if world > 1 else eng.close)()

        # ---- steady state: the same measurement far into the scene (settled beds, not the opening explosion) ----
        steady = None
        if args.steady_frame and not args.quick:
            obj_s, eng_s = make()
            frames(obj_s, args.steady_frame)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            frames(obj_s, args.steps)
            s1.record(stream)
            barrier()
            ms_s = max_over_ranks(s0.elapsed_time(s1))
            st_s = eng_s.stats()
            steady = {"frame_window": [args.steady_frame, args.steady_frame + args.steps], "ms_per_step": ms_s / args.steps,
                      "value": owned_total * S * args.steps / (ms_s * 1e-3),
                      "kbar": sum_over_ranks(st_s["neighborsTotal"]) / max(1.0, sum_over_ranks(st_s["activeInGrid"])),
                      "capped_rows": int(sum_over_ranks(st_s["cappedRows"])),
                      "collision_pairs_last_substep": int(sum_over_ranks(st_s["collisionPairs"]))}
            (obj_s.close if world > 1 else eng_s.close)()

        # ---- verification of the partitioned run against ONE context (N > 1) ---------------------------------
        verified = None
        if world > 1 and args.verify and N > 40_000_000:
            verified = {"ok": None, "skipped": "one context cannot hold this scene (the comparison needs the whole world on rank 0)"}
        elif world > 1 and args.verify:
            vf = args.verify
            obj_v, eng_v = make()
            frames(obj_v, vf)
            g, vals, _ = obj_v.owned_state(("T.x", "T.y", "RB.px", "RB.py", "RB.collisionCount"))
            obj_v.status()
            cs_mine = state_checksum(g, vals)
            t = torch.tensor([len(g), cs_mine & 0xFFFFFFFF, cs_mine >> 32], device="cuda", dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            obj_v.close()
            n_all = int(t[0].item())
            cs_all = (int(t[1].item()) + (int(t[2].item()) << 32)) & 0xFFFFFFFFFFFFFFFF
            if rank == 0:
                ref = GameEngine(cfg, device=local, stream=stream.cuda_stream, host_neighbor_rows=False)
                ref.load_columns(cols)
                ref.run(vf)
                ref.download(B.COLS_INPUT_ALL)
                act = np.nonzero(ref.col["T.active"] != 0)[0].astype(np.uint32)
                cs_ref = state_checksum(act, {k: ref.col[k][act] for k in ("T.x", "T.y", "RB.px", "RB.py", "RB.collisionCount")})
                ref.close()
                verified = {"ok": bool(n_all == len(act) and cs_all == cs_ref), "frames": vf, "entities_owned_once": n_all == len(act),
                            "checksum_slabs": f"{cs_all:016x}", "checksum_single_context": f"{cs_ref:016x}",
                            "what": "order-independent 64-bit checksum of (gid, x, y, px, py, collisionCount) bits over the owned entities "
                                    "of all slabs, all-reduced, against one context that ran the same frames on rank 0"}
            dist.barrier()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu, _, _ = cpu_reference(name, args.cpu_steps, 1, args.cpu_sample or 400_000)
        conf = describe(name, cfg)
        conf.update({"parallelism": "1 GPU" if world == 1 else
                     f"{world} row slabs (1 per GPU, one process each), halo {plan[1]} rows recomputed redundantly, one neighbour exchange per frame: "
                     + ("pack kernels write straight into the neighbours' receive buffers over NVLink (CUDA IPC peer mappings), device-side arrival flags; "
                        "no host work and no library call inside a frame (NCCL: setup, barriers, reductions of the timings)" if args.transport == "p2p"
                        else "one fixed-size NCCL send/recv per neighbour (torch.distributed)"),
                     "frame_window": [args.warmup, args.warmup + args.steps], "steady_state": steady, "verified": verified,
                     "kbar": kbar, "active": active, "l2_policy": "working set >> 126 MB L2 (inputs larger than L2)"
                     if N / world > 2_000_000 else "per-GPU working set comparable to L2; frames run back-to-back on evolving state",
                     "kernel_ms_rank0_per_launch": None if args.quick else {n: float(v) for n, v in zip(KERNEL_NAMES, kms)},
                     "explicit_pairs": st["explicitPairs"], "capped_rows": st["cappedRows"],
                     "collision_pairs_last_substep": st["collisionPairs"],
                     "halo_replica_fraction": halo_frac, "exchange_bytes_per_frame": xbytes,
                     "frame_kernels_ms_per_rank": kernel_ms_per_rank, "slab_rows_begin_end_moves_per_rank": slab_rows,
                     "slab_balance": balance_note if world > 1 else None})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak" if name == "config5" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": conf,
                "verified": None if verified is None else verified.get("ok"),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clk.summary()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config4")
    ap.add_argument("--entities", type=int, default=None, help="override the entity count (same density)")
    ap.add_argument("--cpu-sample", type=int, default=None,
                    help="entities of a scaled CPU sample (default: 400000 for the cpu_baseline leg of the GPU arm; "
                         "--impl reference runs the full workload unless this is given)")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--ref-budget-s", type=float, default=270.0,
                    help="--impl reference: seconds of host time the oracle frames may take; a run that would not fit is shortened "
                         "(warm-up first, then timed frames) and says so (0 = run exactly --steps after --warmup)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the per-kernel timing pass, shorten the e2e pass (huge scenes)")
    ap.add_argument("--balance-rows", type=int, default=2, help="N>1: rows a cut may move per frame toward the slower slab (weed_slab_balance; 0 = static cuts)")
    ap.add_argument("--autobalance", type=int, default=3, help="measured-feedback slab re-plans before timing (N>1)")
    ap.add_argument("--e2e-balance", type=int, default=2, help="N>1: re-plans of the slabs on the measured end-to-end step time before the e2e leg (0 = keep the device leg's slabs)")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="N>1 neighbour exchange: peer-to-peer writes over NVLink (default) or the NCCL send/recv fallback")
    ap.add_argument("--verify", type=int, default=8, help="N>1: frames of the checksum comparison against one context (0 = skip)")
    ap.add_argument("--steady-frame", type=int, default=None,
                    help="also time --steps frames starting this far into the scene (default: 300 for config3/config4 on one GPU, else off)")
    args = ap.parse_args()
    if args.steady_frame is None:
        args.steady_frame = 300 if (args.workload in ("config3", "config4") and args.gpus == 1 and not args.entities) else 0
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
