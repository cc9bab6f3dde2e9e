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


GUARD_ROWS = 8      # an observer this close to its reach of a cut counts as "near" when the slabs are planned


def halo_rows(cfg, cols, blocks=None):
    """Replicated rows beyond each cut.

    Entities that can move or push (non-trigger colliders) propagate position dependencies one
    visual range per substep, and the pair-membership inference looks one more range out (a
    partner's row must be complete): reach (S+1) * h, wherever they are — they move.
    Observers (triggers such as the Mouse, src/core/Mouse.js:139-145, and non-colliders) never move
    anybody, so they do not extend the chain; but when one lies within h rows of a cut, the slabs
    on both sides need its capped row complete: reach 2 * h.  The Mouse of the large scenes
    (visualRange 150, ten cells) parked in a corner therefore costs nothing: with `blocks` (the
    planned cuts) only the observers near a cut count.  The library checks the same rule on the
    device every frame (k_slab_pack): an entity whose reach exceeds the halo and that comes near a
    cut makes weed_slab_status fail instead of letting the partition diverge."""
    act = cols["T.active"] != 0
    vr = cols["C.visualRange"].astype(np.float64)
    cs = float(cfg["spatial"]["cellSize"])
    rows = math.ceil(cfg["worldHeight"] / cs)
    with np.errstate(invalid="ignore"):
        cr = np.ceil(vr * (1.0 / cs))
    cr = np.where(np.isnan(cr), 0.0, np.clip(cr, 0, rows))
    observer = (cols["C.isTrigger"] != 0) | (cols["C.active"] == 0)
    S = int(cfg["physics"].get("subStepCount", 4))
    need = np.where(observer, 2 * cr, (S + 1) * cr)
    H = int(need[act & ~observer].max(initial=1.0))
    obs = act & observer
    if obs.any():
        if blocks is None:
            H = max(H, int(need[obs].max()))
        else:
            row, _ = cell_rows(cfg, cols["T.y"])
            cuts = np.array([b for (_, b) in blocks[:-1]], dtype=np.int64)
            if len(cuts):
                dist = np.abs(row[obs][:, None] - cuts[None, :]).min(axis=1)
                near = dist <= cr[obs] + GUARD_ROWS
                H = max(H, int(need[obs][near].max(initial=0)))
    return max(1, H)


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
    blocks = [(cuts[k], cuts[k + 1]) for k in range(world)]
    return blocks, halo_rows(cfg, cols, blocks)


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


def replan_by_rank_speed(blocks, times_ms, row_weight):
    """Balancing when the time belongs to the RANK rather than to its rows — the end-to-end frame,
    whose host<->device copies run at the rate of each GPU's own host link (measured on the 8-GPU
    boxes of this pool with every rank copying: 11.8 GB/s on four links, 18.6 GB/s on the other four).
    Each rank's speed is the row weight it carried per millisecond; the new cuts give every rank a
    share of the total weight in proportion to its speed.  Returns the new (rowBegin, rowEnd) list."""
    w = np.asarray(row_weight, dtype=np.float64)
    rows, world = blocks[-1][1], len(blocks)
    speed = np.array([max(w[a:b].sum(), 1e-12) / max(float(t), 1e-9) for (a, b), t in zip(blocks, times_ms)])
    share = np.cumsum(speed) / speed.sum()
    cum = np.cumsum(w)
    cuts = [0]
    for k in range(1, world):
        r = int(np.searchsorted(cum, cum[-1] * share[k - 1])) + 1
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
    """One slab = one GameEngine over a LOCAL entity table + its share of the exchange.

    transport "p2p" (default): the library's peer-to-peer exchange (weed_slab_exchange_*): the pack
    kernel writes into the neighbour's receive buffer over NVLink and the frame needs no host work at
    all — `step_dist` is ONE call, weed_slab_frame.  Between processes the buffers are shared through
    CUDA IPC handles that travel once, at setup, over torch.distributed; inside one process
    (SlabGroup) through plain device pointers.  transport "nccl": the round-1 path, one fixed-size
    send/recv per neighbour through torch.distributed, kept as the fallback for boxes without peer
    access."""

    def __init__(self, cfg, cols, rank, world, device=0, flags=0, stream=None, plan=None,
                 capacity_factor=1.35, host_neighbor_rows=False, balance_rows=0, balance_hysteresis=3,
                 transport="p2p", connect=True, adopt_ctx=None, capacity=None, quota=None):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.blocks, self.H = plan if plan is not None else plan_slabs(cfg, cols, world)
        self.rb, self.re = self.blocks[rank]
        self.cfg = cfg
        self.S = int(cfg["physics"].get("subStepCount", 4))
        sel = self.select(cfg, cols, rank, self.blocks, self.H)
        self.capacity = capacity or self.plan_capacity(cfg, cols, world, len(sel), capacity_factor, balance_rows)
        self.quota = quota or self.plan_quota(cfg, cols, self.blocks, self.H, balance_rows)
        lcfg = dict(cfg)
        lcfg["entityCount"] = self.capacity
        if transport == "nccl" and stream is None:
            # the NCCL calls of torch.distributed are ordered against torch's CURRENT stream only: the
            # context must run on a stream torch knows, and the exchange under `with torch.cuda.stream`
            self._tstream = torch.cuda.Stream(device=device)
            stream = self._tstream.cuda_stream
        else:
            self._tstream = None
        self._ext_stream = stream
        self.eng = GameEngine(lcfg, device=device, flags=flags, stream=stream, host_neighbor_rows=host_neighbor_rows,
                              slab=(self.rb, self.re, self.H), adopt_ctx=adopt_ctx)
        for k, v in cols.items():
            self.eng.column(k)[:len(sel)] = v[sel]
        self.eng.upload(B.COLS_INPUT_ALL)
        B.check(self.eng.ctx, B.lib().weed_slab_set_gids(self.eng.ctx, sel.ctypes.data, len(sel)))
        if balance_rows:        # cuts follow the measured load (weed_slab_balance)
            B.check(self.eng.ctx, B.lib().weed_slab_balance(self.eng.ctx, int(balance_rows), int(balance_hysteresis)))
        self.transport = transport
        self.device = device
        if transport == "p2p":
            if adopt_ctx is None:
                B.check(self.eng.ctx, B.lib().weed_slab_exchange_create(self.eng.ctx, self.quota))
                if connect and world > 1:
                    self.connect_dist()
        else:
            dev = torch.device("cuda", device)
            mk = lambda: torch.zeros((self.quota + 1) * B.SLAB_RECORD_BYTES, dtype=torch.uint8, device=dev)
            self.send_low, self.send_high, self.recv_low, self.recv_high = mk(), mk(), mk(), mk()

    # ---- planning helpers (identical on every rank: the two sides of a cut must agree) --------------
    @staticmethod
    def select(cfg, cols, rank, blocks, H):
        rb, re = blocks[rank]
        row, _ = cell_rows(cfg, cols["T.y"])
        act = cols["T.active"] != 0
        fin = np.isfinite(cols["T.x"]) & np.isfinite(cols["T.y"])
        inside = act & fin & (row >= rb - H) & (row < re + H)
        if rank == 0:   # active entities that never enter the grid (NaN position) live on slab 0
            inside |= act & ~fin
        return np.nonzero(inside)[0].astype(np.uint32)

    @staticmethod
    def plan_capacity(cfg, cols, world, n_local, capacity_factor=1.35, balance_rows=0):
        cap = int(n_local * capacity_factor) + 4096
        if balance_rows:        # moving cuts even the slabs out: size every table for an even share (+ halo) as well
            act = int((cols["T.active"] != 0).sum())
            cap = max(cap, int((act / world) * (capacity_factor + 0.25)) + 4096)
        return cap

    @staticmethod
    def plan_quota(cfg, cols, blocks, H, balance_rows=0):
        """Records one message can hold: the same on every rank.  The peer-to-peer transport only
        moves the records that exist (the quota is the size of the receive buffers: memory, not
        traffic); the NCCL fallback sends whole buffers."""
        row, rows = cell_rows(cfg, cols["T.y"])
        act = cols["T.active"] != 0
        fin = np.isfinite(cols["T.x"]) & np.isfinite(cols["T.y"])
        hist = np.bincount(row[act & fin], minlength=rows)
        if balance_rows:
            # cuts move: size for the most crowded (halo + shift)-row window of the start scene wherever it lies
            w = H + int(balance_rows)
            csum = np.concatenate([[0], np.cumsum(hist)])
            band = int((csum[w:] - csum[:-w]).max()) if len(hist) > w else int(hist.sum())
            return band + band // 4 + 8192
        band = 0
        for (_, cut) in blocks[:-1]:
            band = max(band, int(hist[max(0, cut - H):cut].sum()), int(hist[cut:cut + H].sum()))
        return band + band // 2 + 8192

    # ---- peer-to-peer setup -----------------------------------------------------------------------
    def export(self):
        """(64-byte IPC handle, device pointer) of this slab's receive buffers."""
        h = (C.c_ubyte * 64)()
        base = C.c_void_p()
        B.check(self.eng.ctx, B.lib().weed_slab_exchange_export(self.eng.ctx, h, C.byref(base)))
        return bytes(h), base.value

    def connect(self, side, handle=None, base=None):
        hb = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        B.check(self.eng.ctx, B.lib().weed_slab_exchange_connect(self.eng.ctx, side, hb, base))

    def connect_dist(self):
        """One process per GPU: the IPC handles travel once over torch.distributed."""
        dist = self.torch.distributed
        handle, _ = self.export()
        handles = [None] * self.world
        dist.all_gather_object(handles, handle)
        if self.rank > 0:
            self.connect(0, handle=handles[self.rank - 1])
        if self.rank + 1 < self.world:
            self.connect(1, handle=handles[self.rank + 1])
        self._connected_dist = True
        dist.barrier()

    # ---- one frame of this slab: everything below is asynchronous on the context's stream ---------
    def run(self, dtRatio=1.0):
        self.eng.run(1, dtRatio)

    def pack(self):
        B.check(self.eng.ctx, B.lib().weed_slab_pack(self.eng.ctx, self.send_low.data_ptr(), self.send_high.data_ptr(), self.quota))

    def apply(self):
        lo = self.recv_low.data_ptr() if self.rank > 0 else None
        hi = self.recv_high.data_ptr() if self.rank + 1 < self.world else None
        B.check(self.eng.ctx, B.lib().weed_slab_apply(self.eng.ctx, lo, hi, self.quota))

    def _nccl_stream(self):
        t = self.torch
        if self._tstream is not None:
            return t.cuda.stream(self._tstream)
        return t.cuda.stream(t.cuda.ExternalStream(self._ext_stream, device=self.device))

    def exchange_dist(self):
        """The exchange alone (after a frame run by `run` or `eng.step`)."""
        if self.transport == "p2p":
            B.check(self.eng.ctx, B.lib().weed_slab_exchange(self.eng.ctx))
            return
        with self._nccl_stream():        # pack -> one fixed-size send/recv per adjacent rank -> apply, all on ONE stream
            self.pack()
            exchange_fixed(self.torch, self.rank, self.world, self.send_low, self.send_high, self.recv_low, self.recv_high)
            self.apply()

    def step_dist(self, dtRatio=1.0):
        if self.transport == "p2p":      # frame kernels + pack into the neighbours' buffers + wait + apply: one call, no host work
            B.check(self.eng.ctx, B.lib().weed_slab_frame(self.eng.ctx, float(dtRatio)))
            return
        self.run(dtRatio)
        self.exchange_dist()

    def frame_begin(self, dtRatio=1.0):
        B.check(self.eng.ctx, B.lib().weed_slab_frame_begin(self.eng.ctx, float(dtRatio)))

    def frame_end(self):
        B.check(self.eng.ctx, B.lib().weed_slab_frame_end(self.eng.ctx))

    def status(self):
        """Synchronises; raises if a quota, the entity table or the halo reach was exceeded at any time."""
        st = B.SlabStats()
        B.check(self.eng.ctx, B.lib().weed_slab_status(self.eng.ctx, C.byref(st)))
        return {n: getattr(st, n) for n, _ in st._fields_}

    @property
    def exchange_bytes_per_frame(self):
        """Bytes this slab sent in its last exchange (synchronises).  p2p: the records that exist plus
        one header per neighbour; nccl: whole buffers."""
        n = (self.rank > 0) + (self.rank + 1 < self.world)
        if self.transport == "p2p":
            st = self.status()
            return (st["sentLow"] + st["sentHigh"] + n) * B.SLAB_RECORD_BYTES
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
        if getattr(self, "_connected_dist", False) and self.eng.ctx:
            # unmap the neighbours' buffers, wait for everybody, only then free mine
            B.lib().weed_slab_exchange_disconnect(self.eng.ctx)
            self.torch.distributed.barrier()
            self._connected_dist = False
        self.eng.close()


class SlabGroup:
    """All slabs of a world inside ONE process (contexts may share a device).

    mode "c" (default): the library's weed_group_* entry points — what a single Node engine would call
    — create the contexts and step them; mode "py": SlabEngines created here and connected through
    device pointers (weed_slab_exchange_connect); mode "nccl-buffers": the round-1 staging path
    (weed_slab_pack / weed_slab_apply with a device-to-device copy in between)."""

    def __init__(self, cfg, cols, world, devices=None, plan=None, mode="c", **kw):
        plan = plan or plan_slabs(cfg, cols, world)
        devices = devices or [0] * world
        self.mode = mode
        self.group = None
        balance_rows = kw.get("balance_rows", 0)
        if mode == "c":
            blocks, H = plan
            sels = [SlabEngine.select(cfg, cols, r, blocks, H) for r in range(world)]
            caps = [SlabEngine.plan_capacity(cfg, cols, world, len(s), kw.get("capacity_factor", 1.35), balance_rows) for s in sels]
            quota = SlabEngine.plan_quota(cfg, cols, blocks, H, balance_rows)
            probe = GameEngine.config_struct(cfg, flags=kw.get("flags", 0))
            cuts = (C.c_uint32 * (world + 1))(*([b[0] for b in blocks] + [blocks[-1][1]]))
            devs = (C.c_int32 * world)(*devices)
            capa = (C.c_uint32 * world)(*caps)
            g = C.c_void_p()
            rc = B.lib().weed_group_create(C.byref(probe), world, devs, cuts, H, capa, quota, C.byref(g))
            if rc != B.WEED_OK:
                msg = B.lib().weed_group_last_error(None)
                raise B.WeedError(rc, msg.decode() if msg else "")
            self.group = g
            self.slabs = [SlabEngine(cfg, cols, r, world, device=devices[r], plan=plan, transport="p2p",
                                     adopt_ctx=B.lib().weed_group_slab(g, r), capacity=caps[r], quota=quota, **kw)
                          for r in range(world)]
        elif mode == "py":
            self.slabs = [SlabEngine(cfg, cols, r, world, device=devices[r], plan=plan, transport="p2p", connect=False, **kw)
                          for r in range(world)]
            bases = [s.export()[1] for s in self.slabs]
            for r, s in enumerate(self.slabs):
                if r > 0:
                    s.connect(0, base=bases[r - 1])
                if r + 1 < world:
                    s.connect(1, base=bases[r + 1])
        else:
            self.slabs = [SlabEngine(cfg, cols, r, world, device=devices[r], plan=plan, transport="nccl", **kw) for r in range(world)]

    def step(self, dtRatio=1.0, check=True):
        if self.mode == "c":
            rc = B.lib().weed_group_step(self.group, float(dtRatio))
            if rc != B.WEED_OK:
                raise B.WeedError(rc, B.lib().weed_group_last_error(self.group).decode())
            if check:
                rc = B.lib().weed_group_sync(self.group)
                if rc != B.WEED_OK:
                    raise B.WeedError(rc, B.lib().weed_group_last_error(self.group).decode())
            return
        if self.mode == "py":
            for s in self.slabs:
                s.frame_begin(dtRatio)
            for s in self.slabs:
                s.frame_end()
        else:
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
        if check:
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
        if self.group is not None:
            B.lib().weed_group_destroy(self.group)
            self.group = None
