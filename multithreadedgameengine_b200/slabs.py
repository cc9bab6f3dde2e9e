"""World-space row slabs: one libweedgpu context per GPU (SURVEY §8 e).

* Partition: contiguous blocks of grid rows; cell index = row * gridCols + col
  (spatial_worker.js:161) makes a row block a contiguous cell range, so the reference's
  row-major scan order is preserved inside a slab.  Cuts are chosen from the row histogram of
  the start scene so slabs hold equal entity counts.
* Ownership follows position: a slab owns the entities whose cell row lies in its block.
* Halo: H = (S+1) * ceil(max visualRange / cellSize) rows beyond each cut are replicated.
  With H that deep every substep of every owned entity — including the pair-membership
  inference that looks at a partner's row — only depends on data inside the slab + halo, so
  the replicas are simply recomputed redundantly and ONE exchange per frame (halo refresh +
  migration, 64-byte records to the two adjacent slabs) is enough.
* The exchange moves plain device buffers: `torch.distributed` send/recv (NCCL over NVLink)
  between processes (`SlabEngine.exchange_dist`), or device-to-device copies between the
  contexts of one process (`SlabGroup`, used by the single-GPU tests).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import binding as B
from .engine import GameEngine


def cell_rows(cfg, y):
    """Clamped grid row of each y (spatial_worker.js:158,160), float64 like the reference."""
    cs = float(cfg["spatial"]["cellSize"])
    rows = math.ceil(cfg["worldHeight"] / cs)
    inv = 1.0 / cs
    with np.errstate(invalid="ignore"):
        r = np.trunc(y.astype(np.float64) * inv)
    r = np.where(np.isfinite(r), r, 0.0)
    return np.clip(r, 0, rows - 1).astype(np.int64), rows


def halo_rows(cfg, cols):
    """Replicated rows beyond each cut.

    Entities that can move or push (non-trigger colliders) propagate position dependencies one
    visual range per substep, and the pair-membership inference looks one more range out (a
    partner's row must be complete): (S+1) * h_dyn.  Observers (triggers such as the Mouse,
    src/core/Mouse.js:139-145, and non-colliders) never move anybody, so they do not extend the
    chain; but an owned entity's collisionCount needs the observer's row to be complete when it
    is capped: 2 * h_obs."""
    act = cols["T.active"] != 0
    vr = np.where(np.isfinite(cols["C.visualRange"]), cols["C.visualRange"], 0).astype(np.float64)
    observer = (cols["C.isTrigger"] != 0) | (cols["C.active"] == 0)
    cs = float(cfg["spatial"]["cellSize"])
    S = int(cfg["physics"].get("subStepCount", 4))
    h_dyn = math.ceil(float(vr[act & ~observer].max(initial=0.0)) / cs)
    h_obs = math.ceil(float(vr[act & observer].max(initial=0.0)) / cs)
    return max(1, (S + 1) * h_dyn, 2 * h_obs)


def row_costs(cfg, cols):
    """Estimated per-row work of the start scene.  The kernels' cost per entity grows with the
    number of candidates its scan visits (the 3x3 cell neighbourhood): measured on B200,
    cost ~ 1 + 0.06 * (entities in the 3x3 block), in units of ~0.27 us."""
    cs = float(cfg["spatial"]["cellSize"])
    ncols = math.ceil(cfg["worldWidth"] / cs)
    act = (cols["T.active"] != 0) & np.isfinite(cols["T.x"]) & np.isfinite(cols["T.y"])
    row, rows = cell_rows(cfg, cols["T.y"])
    with np.errstate(invalid="ignore"):
        c = np.trunc(cols["T.x"].astype(np.float64) * (1.0 / cs))
    col = np.clip(np.where(np.isfinite(c), c, 0.0), 0, ncols - 1).astype(np.int64)
    m = np.bincount((row * ncols + col)[act], minlength=rows * ncols).reshape(rows, ncols).astype(np.float32)
    box = np.zeros_like(m)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            src = m[max(0, dr):rows + min(0, dr), max(0, dc):ncols + min(0, dc)]
            box[max(0, -dr):rows + min(0, -dr), max(0, -dc):ncols + min(0, -dc)] += src
    return (m * (1.0 + 0.06 * box)).sum(axis=1).astype(np.float64), m.sum(axis=1)


def plan_slabs(cfg, cols, world, balance="cost"):
    """-> list of (rowBegin, rowEnd) per rank and the halo depth.  Cuts equalise the estimated
    work (balance="cost", default) or the entity counts (balance="count") of the start scene."""
    cost, count = row_costs(cfg, cols)
    rows = len(cost)
    cum = np.cumsum(cost if balance == "cost" else count)
    total = float(cum[-1])
    cuts = [0]
    for k in range(1, world):
        target = total * k / world
        r = int(np.searchsorted(cum, target)) + 1
        cuts.append(min(max(r, cuts[-1] + 1), rows - (world - k)))
    cuts.append(rows)
    return [(cuts[k], cuts[k + 1]) for k in range(world)], halo_rows(cfg, cols)


def replan_from_times(blocks, times_ms, row_weight=None):
    """Measured-feedback balancing: given the frame time of every slab under the current cuts,
    spread each slab's time over its rows — in proportion to `row_weight` (the cost model of
    row_costs, so that dense rows inside a slab keep their share) or uniformly — and cut the
    cumulative curve into equal parts.  Returns the new (rowBegin, rowEnd) list (same halo)."""
    rows = blocks[-1][1]
    world = len(blocks)
    per_row = np.zeros(rows, dtype=np.float64)
    for (a, b), t in zip(blocks, times_ms):
        w = None if row_weight is None else np.asarray(row_weight[a:b], dtype=np.float64)
        if w is not None and w.sum() > 0:
            per_row[a:b] = float(t) * w / w.sum()
        else:
            per_row[a:b] = float(t) / max(1, b - a)
    cum = np.cumsum(per_row)
    total = float(cum[-1])
    cuts = [0]
    for k in range(1, world):
        r = int(np.searchsorted(cum, total * k / world)) + 1
        cuts.append(min(max(r, cuts[-1] + 1), rows - (world - k)))
    cuts.append(rows)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def exchange_fixed(torch, rank, world, send_low, send_high, recv_low, recv_high):
    """Neighbour exchange of one frame: ONE fixed-size message per adjacent rank and direction
    (rank-1 = low, rank+1 = high).  Every buffer holds quota + 1 records of 64 bytes; record 0 is
    a header carrying the number of valid records, so no size negotiation and no host
    synchronisation are needed.  Works on any backend/device pair torch.distributed supports
    (NCCL + CUDA tensors on the GPUs, gloo + CPU tensors in the CPU tests)."""
    dist = torch.distributed
    lo, hi = rank - 1, rank + 1
    ops = []
    if lo >= 0:
        ops += [dist.P2POp(dist.isend, send_low, lo), dist.P2POp(dist.irecv, recv_low, lo)]
    if hi < world:
        ops += [dist.P2POp(dist.isend, send_high, hi), dist.P2POp(dist.irecv, recv_high, hi)]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()            # CUDA tensors: orders the current stream after the transfer, does not block the host


def header_count(buf):
    """Record count stored in the header (first 4 bytes) of an exchange buffer (host-side helper)."""
    return int(buf[:4].cpu().numpy().view(np.uint32)[0])


class SlabEngine:
    """One slab = one GameEngine over a LOCAL entity table + the exchange buffers."""

    def __init__(self, cfg, cols, rank, world, device=0, flags=0, stream=None, plan=None,
                 capacity_factor=1.35, host_neighbor_rows=False, balance_rows=0, balance_hysteresis=3):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.blocks, self.H = plan if plan is not None else plan_slabs(cfg, cols, world)
        self.rb, self.re = self.blocks[rank]
        row, self.rows = cell_rows(cfg, cols["T.y"])
        act = cols["T.active"] != 0
        fin = np.isfinite(cols["T.x"]) & np.isfinite(cols["T.y"])
        inside = act & fin & (row >= self.rb - self.H) & (row < self.re + self.H)
        if rank == 0:   # active entities that never enter the grid (NaN position) live on slab 0
            inside |= act & ~fin
        sel = np.nonzero(inside)[0].astype(np.uint32)
        self.capacity = int(len(sel) * capacity_factor) + 4096
        if balance_rows:        # moving cuts even the slabs out: size every table for an even share (+ halo) as well
            self.capacity = max(self.capacity, int((int(act.sum()) / world) * (capacity_factor + 0.25)) + 4096)
        lcfg = dict(cfg)
        lcfg["entityCount"] = self.capacity
        self.cfg = cfg
        self.S = int(cfg["physics"].get("subStepCount", 4))
        self.eng = GameEngine(lcfg, device=device, flags=flags, stream=stream, host_neighbor_rows=host_neighbor_rows,
                              slab=(self.rb, self.re, self.H))
        for k, v in cols.items():
            self.eng.column(k)[:len(sel)] = v[sel]
        self.eng.upload(B.COLS_INPUT_ALL)
        B.check(self.eng.ctx, B.lib().weed_slab_set_gids(self.eng.ctx, sel.ctypes.data, len(sel)))
        # exchange quota: 1.5x the start population of the widest boundary band of ANY cut (+ slack),
        # the same on every rank, because the two sides of a cut must agree on the message size;
        # every frame moves exactly (quota + 1) records per neighbour and direction
        hist = np.bincount(row[act & fin], minlength=self.rows)
        band = 0
        if balance_rows:
            # cuts move: size the messages for the most crowded (halo + shift)-row window of the
            # start scene wherever it lies (the mean band x 1.5 would overflow inside a cluster)
            w = self.H + int(balance_rows)
            csum = np.concatenate([[0], np.cumsum(hist)])
            band = int((csum[w:] - csum[:-w]).max()) if len(hist) > w else int(hist.sum())
            self.quota = band + band // 4 + 8192
        else:
            for (_, cut) in self.blocks[:-1]:
                band = max(band, int(hist[max(0, cut - self.H):cut].sum()), int(hist[cut:cut + self.H].sum()))
            self.quota = band + band // 2 + 8192
        if balance_rows:        # cuts follow the measured load (weed_slab_balance)
            B.check(self.eng.ctx, B.lib().weed_slab_balance(self.eng.ctx, int(balance_rows), int(balance_hysteresis)))
        dev = torch.device("cuda", device)
        mk = lambda: torch.zeros((self.quota + 1) * B.SLAB_RECORD_BYTES, dtype=torch.uint8, device=dev)
        self.send_low, self.send_high, self.recv_low, self.recv_high = mk(), mk(), mk(), mk()

    # ---- one frame of this slab: everything below is asynchronous on the context's stream ---------
    def run(self, dtRatio=1.0):
        self.eng.run(1, dtRatio)

    def pack(self):
        B.check(self.eng.ctx, B.lib().weed_slab_pack(self.eng.ctx, self.send_low.data_ptr(), self.send_high.data_ptr(), self.quota))

    def apply(self):
        lo = self.recv_low.data_ptr() if self.rank > 0 else None
        hi = self.recv_high.data_ptr() if self.rank + 1 < self.world else None
        B.check(self.eng.ctx, B.lib().weed_slab_apply(self.eng.ctx, lo, hi, self.quota))

    def exchange_dist(self):
        """pack -> one fixed-size NCCL send/recv per adjacent rank -> apply; no host synchronisation."""
        self.pack()
        exchange_fixed(self.torch, self.rank, self.world, self.send_low, self.send_high, self.recv_low, self.recv_high)
        self.apply()

    def step_dist(self, dtRatio=1.0):
        self.run(dtRatio)
        self.exchange_dist()

    def status(self):
        """Synchronises; raises if a quota or the entity table overflowed at any time."""
        st = B.SlabStats()
        B.check(self.eng.ctx, B.lib().weed_slab_status(self.eng.ctx, C.byref(st)))
        return {n: getattr(st, n) for n, _ in st._fields_}

    @property
    def exchange_bytes_per_frame(self):
        n = (self.rank > 0) + (self.rank + 1 < self.world)
        return n * (self.quota + 1) * B.SLAB_RECORD_BYTES

    # ---- results -----------------------------------------------------------------------------------
    def gids(self):
        g = np.empty(self.capacity, dtype=np.uint32)
        top = C.c_uint32()
        B.check(self.eng.ctx, B.lib().weed_slab_get_gids(self.eng.ctx, g.ctypes.data, C.byref(top)))
        return g, top.value

    def owned_state(self, keys=("T.x", "T.y", "RB.px", "RB.py", "RB.vx", "RB.vy", "RB.speed", "RB.velocityAngle",
                                "RB.collisionCount", "RB.ax", "RB.ay")):
        """(gids, {column: values}) of the entities this slab owns NOW (by current position)."""
        self.eng.download(B.COLS_INPUT_ALL)
        g, top = self.gids()
        c = self.eng.col
        act = (c["T.active"][:top] != 0)
        row, _ = cell_rows(self.cfg, c["T.y"][:top])
        fin = np.isfinite(c["T.x"][:top]) & np.isfinite(c["T.y"][:top])
        st = self.status()                      # the cuts may have moved (weed_slab_balance)
        rb, re = st["rowBegin"], st["rowEnd"]
        own = act & ((fin & (row >= rb) & (row < re)) | (~fin & (self.rank == 0)))
        idx = np.nonzero(own)[0]
        return g[idx], {k: c[k][idx].copy() for k in keys}, idx

    def close(self):
        self.eng.close()


class SlabGroup:
    """All slabs of a world inside ONE process (contexts may share a device): the exchange is a
    device-to-device copy.  Used by the single-GPU tests; the multi-process path is
    SlabEngine.step_dist."""

    def __init__(self, cfg, cols, world, devices=None, plan=None, **kw):
        plan = plan or plan_slabs(cfg, cols, world)
        devices = devices or [0] * world
        self.slabs = [SlabEngine(cfg, cols, r, world, device=devices[r], plan=plan, **kw) for r in range(world)]

    def step(self, dtRatio=1.0):
        for s in self.slabs:
            s.run(dtRatio)
        for s in self.slabs:
            s.pack()
        self.slabs[0].torch.cuda.synchronize()
        for r, s in enumerate(self.slabs):
            if r > 0:
                s.recv_low.copy_(self.slabs[r - 1].send_high)
            if r + 1 < len(self.slabs):
                s.recv_high.copy_(self.slabs[r + 1].send_low)
        self.slabs[0].torch.cuda.synchronize()
        for s in self.slabs:
            s.apply()
        for s in self.slabs:
            s.status()

    def gather(self, N, keys=("T.x", "T.y", "RB.px", "RB.py", "RB.vx", "RB.vy", "RB.speed", "RB.collisionCount")):
        out = {k: None for k in keys}
        seen = np.zeros(N, dtype=np.int32)
        for s in self.slabs:
            g, vals, _ = s.owned_state(keys)
            seen[g] += 1
            for k in keys:
                if out[k] is None:
                    out[k] = np.zeros(N, dtype=vals[k].dtype)
                out[k][g] = vals[k]
        return out, seen

    def close(self):
        for s in self.slabs:
            s.close()
