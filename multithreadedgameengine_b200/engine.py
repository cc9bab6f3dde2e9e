"""Host-side mirror of the reference's engine/worker interface for the spatial+physics path.

Mirrors (same names, argument meaning and defaults):
  * ``GameEngine(config)``            src/core/gameEngine.js:22-110  (config.spatial{cellSize,
    maxNeighbors}, config.physics{subStepCount, gravity, verletDamping, ...}, worldWidth/Height)
  * ``createSharedBuffers()``         src/core/gameEngine.js:534-777 (one buffer per component,
    dense; neighborData / distanceData / collisionData)
  * ``SpatialWorker.update()``        src/workers/spatial_worker.js:283-294
  * ``PhysicsWorker.update(dt, r)``   src/workers/physics_worker.js:103-108
  * ``updatePhysicsConfig(partial)``  src/core/gameEngine.js:1304-1325 -> physics_worker.js:114-129
  * ``GameObject.updateNeighbors``    src/core/gameObject.js:700-729  (``neighbors_of(i)``)

The two workers are replaced by one libweedgpu context (include/weedgpu.h).  The buffers
created here play the role of the SharedArrayBuffers: ``tick()``-style host code reads and
writes the component columns (``engine.Transform.x[i]``, ``engine.RigidBody.ax[i] = ...``) and
``step()`` moves the columns named in the masks across PCIe.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import binding as B
from .components import COLUMN_KEYS, new_component_classes

PHYSICS_DEFAULTS = dict(subStepCount=4, boundaryElasticity=0.8, collisionResponseStrength=0.5,
                        verletDamping=0.995, minSpeedForRotation=0.1)  # gameEngine.js:39-45


def _pinned_empty(nbytes):
    """Page-aligned host buffer (stand-in for a SharedArrayBuffer)."""
    raw = np.empty(nbytes + 4096, dtype=np.uint8)
    off = (-raw.ctypes.data) % 4096
    buf = raw[off:off + nbytes]
    buf[:] = 0
    return buf


class _Worker:
    def __init__(self, engine):
        self.engine = engine


class SpatialWorker(_Worker):
    def update(self, deltaTime=16.67, dtRatio=1.0, resuming=False):
        """spatial_worker.js:283-294: rebuildGrid() + findAllNeighbors() on the device."""
        B.check(self.engine.ctx, B.lib().weed_spatial(self.engine.ctx))


class PhysicsWorker(_Worker):
    def update(self, deltaTime=16.67, dtRatio=1.0, resuming=False):
        """physics_worker.js:103-108: updateVerlet(deltaTime, dtRatio) on the device."""
        B.check(self.engine.ctx, B.lib().weed_physics(self.engine.ctx, float(dtRatio)))


class GameEngine:
    def __init__(self, config, device=0, flags=0, stream=None, host_neighbor_rows=True, slab=None, adopt_ctx=None):
        L = B.lib()
        self.config = dict(config)
        # gameEngine.js:34-49 default merge
        phys = dict(PHYSICS_DEFAULTS)
        phys.update({k: v for k, v in (config.get("physics") or {}).items() if k not in ("gravity", "noLimitFPS")})
        gravity = (config.get("physics") or {}).get("gravity") or config.get("gravity") or {"x": 0, "y": 0}
        phys["gravity"] = dict(gravity)
        self.config["physics"] = phys
        spatial = config.get("spatial") or {}
        self.cellSize = float(spatial.get("cellSize") or config.get("cellSize"))            # spatial_worker.js:80
        self.maxNeighbors = int(spatial.get("maxNeighbors") or config.get("maxNeighbors") or 100)  # gameEngine.js:552
        self.maxCollisionPairs = int(phys.get("maxCollisionPairs") or config.get("maxCollisionPairs") or 10000)  # :689-693
        self.totalEntityCount = int(config["entityCount"])
        N = self.totalEntityCount

        cfg = B.Config()
        L.weed_default_config(C.byref(cfg))
        cfg.entityCount = N
        cfg.worldWidth = float(config["worldWidth"])
        cfg.worldHeight = float(config["worldHeight"])
        cfg.cellSize = self.cellSize
        cfg.maxNeighbors = self.maxNeighbors
        cfg.maxCollisionPairs = self.maxCollisionPairs
        cfg.seed = float(config.get("seed") or 1.0)
        cfg.physics = self._physics_struct(phys)
        cfg.device = device
        cfg.flags = flags
        cfg.stream = stream
        self._cfg_struct = cfg
        if slab is not None:      # (rowBegin, rowEnd, haloRows): this context is one slab of the world
            cfg.slabRowBegin, cfg.slabRowEnd, cfg.slabHaloRows = (int(v) for v in slab)
        self.slab = slab
        self._owns_ctx = adopt_ctx is None
        if adopt_ctx is not None:      # a context somebody else created and will destroy (a slab of a weed_group)
            self.ctx = C.c_void_p(adopt_ctx)
        else:
            self.ctx = C.c_void_p()
            rc = L.weed_create(C.byref(cfg), C.byref(self.ctx))
            if rc != B.WEED_OK:
                msg = L.weed_last_error(None)
                raise B.WeedError(rc, msg.decode() if msg else "")
        self.spatial = SpatialWorker(self)
        self.physics_worker = PhysicsWorker(self)
        self.buffers = {}
        self.createSharedBuffers(host_neighbor_rows)

    @staticmethod
    def config_struct(config, device=0, flags=0):
        """weed_config for `config` with the engine's defaults applied (gameEngine.js:34-49), without creating
        a context: the template weed_group_create takes."""
        L = B.lib()
        phys = dict(PHYSICS_DEFAULTS)
        phys.update({k: v for k, v in (config.get("physics") or {}).items() if k not in ("gravity", "noLimitFPS")})
        phys["gravity"] = dict((config.get("physics") or {}).get("gravity") or config.get("gravity") or {"x": 0, "y": 0})
        spatial = config.get("spatial") or {}
        cfg = B.Config()
        L.weed_default_config(C.byref(cfg))
        cfg.entityCount = int(config["entityCount"])
        cfg.worldWidth = float(config["worldWidth"])
        cfg.worldHeight = float(config["worldHeight"])
        cfg.cellSize = float(spatial.get("cellSize") or config.get("cellSize"))
        cfg.maxNeighbors = int(spatial.get("maxNeighbors") or config.get("maxNeighbors") or 100)
        cfg.maxCollisionPairs = int(phys.get("maxCollisionPairs") or config.get("maxCollisionPairs") or 10000)
        cfg.seed = float(config.get("seed") or 1.0)
        cfg.physics = GameEngine._physics_struct(phys)
        cfg.device = device
        cfg.flags = flags
        return cfg

    @staticmethod
    def _physics_struct(phys):
        p = B.PhysicsConfig()
        p.subStepCount = int(phys.get("subStepCount", 4))
        p.boundaryElasticity = float(phys.get("boundaryElasticity", 0.8))
        p.collisionResponseStrength = float(phys.get("collisionResponseStrength", 0.5))
        p.verletDamping = float(phys.get("verletDamping", 0.995))
        p.minSpeedForRotation = float(phys.get("minSpeedForRotation", 0.1))
        g = phys.get("gravity") or {}
        p.gravityX = float(g.get("x", 0.0))
        p.gravityY = float(g.get("y", 0.0))
        return p

    # ---- gameEngine.js:534-777 ---------------------------------------------------------------
    def createSharedBuffers(self, host_neighbor_rows=True):
        L = B.lib()
        N, M, P = self.totalEntityCount, self.maxNeighbors, self.maxCollisionPairs
        self.Transform, self.RigidBody, self.Collider = new_component_classes()
        for bid, cls, name in ((B.BUF_TRANSFORM, self.Transform, "Transform"),
                               (B.BUF_RIGIDBODY, self.RigidBody, "RigidBody"),
                               (B.BUF_COLLIDER, self.Collider, "Collider")):
            size = cls.getBufferSize(N)
            assert size == L.weed_buffer_bytes(bid, N, M, P)
            buf = _pinned_empty(size)
            self.buffers[name] = buf
            cls.initializeArrays(buf, N)
            B.check(self.ctx, L.weed_bind(self.ctx, bid, buf.ctypes.data, size))
        if host_neighbor_rows:
            self.neighborData = _pinned_empty(N * (1 + M) * 4).view(np.int32)
            self.distanceData = _pinned_empty(N * (1 + M) * 4).view(np.float32)
            B.check(self.ctx, L.weed_bind(self.ctx, B.BUF_NEIGHBOR, self.neighborData.ctypes.data, self.neighborData.nbytes))
            B.check(self.ctx, L.weed_bind(self.ctx, B.BUF_DISTANCE, self.distanceData.ctypes.data, self.distanceData.nbytes))
        else:
            self.neighborData = self.distanceData = None
        self.collisionData = _pinned_empty((1 + 2 * P) * 4).view(np.int32)
        B.check(self.ctx, L.weed_bind(self.ctx, B.BUF_COLLISION, self.collisionData.ctypes.data, self.collisionData.nbytes))

    def column(self, key):
        comp, name = COLUMN_KEYS[key]
        return getattr((self.Transform, self.RigidBody, self.Collider)[comp], name)

    @property
    def col(self):
        return {k: self.column(k) for k in COLUMN_KEYS}

    def load_columns(self, columns, upload=True):
        for k, v in columns.items():
            self.column(k)[:] = v
        if upload:
            self.upload(B.COLS_INPUT_ALL)

    # ---- data movement -----------------------------------------------------------------------
    @staticmethod
    def mask(*keys):
        m = 0
        for k in keys:
            m |= B.COL[k] if isinstance(k, str) else int(k)
        return m

    def upload(self, mask=B.COLS_INPUT_ALL):
        B.check(self.ctx, B.lib().weed_upload(self.ctx, mask))

    def download(self, mask=B.COLS_OUTPUT_ALL):
        B.check(self.ctx, B.lib().weed_download(self.ctx, mask))

    def fetch_neighbors(self, first=0, count=None):
        count = self.totalEntityCount - first if count is None else count
        B.check(self.ctx, B.lib().weed_fetch_neighbors(self.ctx, first, count))

    def neighbors_of(self, i, fetch=True):
        """GameObject.updateNeighbors (gameObject.js:700-729): (ids, squared distances) of row i."""
        if self.neighborData is None:      # no host mirror of the rows: fetch this row into a scratch pair
            stride = 1 + self.maxNeighbors
            nd = np.empty(stride, dtype=np.int32)
            dd = np.empty(stride, dtype=np.float32)
            B.check(self.ctx, B.lib().weed_fetch_neighbors_to(self.ctx, i, 1, nd.ctypes.data, dd.ctypes.data))
            n = int(nd[0])
            return nd[1:1 + n], dd[1:1 + n]
        if fetch:
            self.fetch_neighbors(i, 1)
        off = i * (1 + self.maxNeighbors)
        n = int(self.neighborData[off])
        return self.neighborData[off + 1:off + 1 + n], self.distanceData[off + 1:off + 1 + n]

    # ---- frames ------------------------------------------------------------------------------
    def step(self, dtRatio=1.0, upload=0, download=B.COLS_OUTPUT_ALL):
        """One lockstep frame: upload(mask); spatial; physics(dtRatio); download(mask)."""
        B.check(self.ctx, B.lib().weed_step(self.ctx, float(dtRatio), upload, download))

    def run(self, frames, dtRatio=1.0):
        """`frames` frames back to back on the device, no host interaction (asynchronous)."""
        B.check(self.ctx, B.lib().weed_run(self.ctx, float(dtRatio), frames))

    def sync(self):
        B.check(self.ctx, B.lib().weed_sync(self.ctx))

    def system_boids(self, dtRatio=1.0, protectedRange=None, centeringFactor=0.001, avoidFactor=0.3,
                     matchingFactor=0.1, turnFactor=0.01, margin=20.0, mouseEntityType=0):
        """Device-side tick() of the boids demo (demos/predators/boid.js:115-240, :318-341): reads the
        neighbor rows on the device and accumulates RigidBody.ax/ay there (weed_system_boids)."""
        p = B.BoidsParams(centeringFactor, avoidFactor, matchingFactor, turnFactor, margin, mouseEntityType, 0)
        ptr = None
        if protectedRange is not None:
            self._prot = np.ascontiguousarray(protectedRange, dtype=np.float32)
            ptr = self._prot.ctypes.data
        B.check(self.ctx, B.lib().weed_system_boids(self.ctx, C.byref(p), ptr, float(dtRatio)))

    def system_flock(self, classes, dtRatio=1.0, mouseDown=False, mouseEntityType=0, protectedRange=None):
        """The predators demo's tick() on the device (weed_system_flock): `classes` is a list of
        dicts with entityType, role ("boid" | "prey" | "predator"), otherEntityType and the
        Flocking numbers each constructor sets (boid.js:64-69, prey.js:37,55-60,
        predator.js:43,57-62); see scenes.PREDATORS_DEMO_CLASSES."""
        roles = {"boid": B.FLOCK_BOID, "prey": B.FLOCK_PREY, "predator": B.FLOCK_PREDATOR}
        arr = (B.FlockClass * len(classes))()
        for k, c in enumerate(classes):
            arr[k] = B.FlockClass(int(c["entityType"]), roles[c.get("role", "boid")], int(c.get("otherEntityType", 0)), 0,
                                  float(c.get("protectedRangeScale", 2.0)), float(c.get("centeringFactor", 0.001)),
                                  float(c.get("avoidFactor", 0.3)), float(c.get("matchingFactor", 0.1)),
                                  float(c.get("turnFactor", 0.01)), float(c.get("margin", 20.0)), float(c.get("roleFactor", 0.0)))
        fp = B.FlockParams(int(mouseEntityType), 1 if mouseDown else 0, float(dtRatio))
        ptr = None
        if protectedRange is not None:
            self._prot = np.ascontiguousarray(protectedRange, dtype=np.float32)
            ptr = self._prot.ctypes.data
        B.check(self.ctx, B.lib().weed_system_flock(self.ctx, arr, len(classes), C.byref(fp), ptr))

    # ---- device-side consumers of collisionData / positions / rows ----------
    def collision_events(self, forget_previous=False):
        """Enter / Stay / Exit diff of this frame's collisionData against the previous call's
        (logic_worker.js:429-526), computed on the device (weed_system_collision_events).
        Returns {"pairs": int32[n,2], "state": uint8[n] (1 Enter, 2 Stay), "exits": int32[m,2]}."""
        maxPairs = self.maxCollisionPairs
        if not hasattr(self, "_ev_state"):
            self._ev_state = np.zeros(max(1, maxPairs), np.uint8)
            self._ev_exit = np.zeros(1 + 2 * maxPairs, np.int32)
        c = B.CollisionEventCounts()
        B.check(self.ctx, B.lib().weed_system_collision_events(
            self.ctx, B.EVENTS_FORGET_PREVIOUS if forget_previous else 0, C.byref(c),
            self._ev_state.ctypes.data, self._ev_exit.ctypes.data))
        self.download(B.COL_COLLISIONS)
        n = int(self.collisionData[0])
        assert n == c.pairs
        return {"pairs": self.collisionData[1:1 + 2 * n].reshape(n, 2).copy(),
                "state": self._ev_state[:n].copy(),
                "exits": self._ev_exit[1:1 + 2 * c.exited].reshape(c.exited, 2).copy(),
                "entered": c.entered, "stayed": c.stayed, "exited": c.exited}

    @staticmethod
    def collision_callbacks(ev):
        """The reference's callback sequence for one frame as (type, self, other) triples:
        both objects of every current pair in list order (logic_worker.js:471-489), then the
        Exit calls — which the reference issues once per Cantor key of an ended pair, i.e.
        (A,B),(B,A) for keyAB and (B,A),(A,B) for keyBA (:493-516)."""
        calls = []
        for (a, b), st in zip(ev["pairs"].tolist(), ev["state"].tolist()):
            calls.append((st, a, b))
            calls.append((st, b, a))
        for a, b in ev["exits"].tolist():
            calls += [(B.COLLISION_EXIT, a, b), (B.COLLISION_EXIT, b, a), (B.COLLISION_EXIT, b, a), (B.COLLISION_EXIT, a, b)]
        return calls

    def screen_visibility(self, zoom, cameraX, cameraY, canvasWidth, canvasHeight, download=True):
        """particle_worker.js:1012-1062 on the device (weed_system_screen_visibility).  Returns
        (screenX, screenY, isItOnScreen) host arrays (persistent across calls, like the SAB)."""
        cam = B.Camera(float(zoom), float(cameraX), float(cameraY), float(canvasWidth), float(canvasHeight))
        if not hasattr(self, "_screenX"):
            N = self.totalEntityCount
            self._screenX = np.zeros(N, np.float32)
            self._screenY = np.zeros(N, np.float32)
            self._onScreen = np.zeros(N, np.uint8)
        ptrs = (self._screenX.ctypes.data, self._screenY.ctypes.data, self._onScreen.ctypes.data) if download else (None, None, None)
        B.check(self.ctx, B.lib().weed_system_screen_visibility(self.ctx, C.byref(cam), *ptrs))
        return self._screenX, self._screenY, self._onScreen

    def shadows_upload(self, lightActive, lightIntensity, casterActive, shadowRadius, height):
        """LightEmitter / ShadowCaster columns the shadow system reads (weed_system_shadows_upload)."""
        N = self.totalEntityCount
        self._sh_cols = [np.ascontiguousarray(lightActive, np.uint8), np.ascontiguousarray(lightIntensity, np.float32),
                         np.ascontiguousarray(casterActive, np.uint8), np.ascontiguousarray(shadowRadius, np.float32),
                         np.ascontiguousarray(height, np.float32)]
        assert all(a.shape == (N,) for a in self._sh_cols)
        cols = B.ShadowColumns(*[a.ctypes.data for a in self._sh_cols])
        B.check(self.ctx, B.lib().weed_system_shadows_upload(self.ctx, C.byref(cols)))

    def shadows(self, maxShadowCastingLights=20, maxShadowsPerLight=15, maxShadowSprites=None):
        """particle_worker.js:861-1003 on the device (weed_system_shadows).  Defaults as
        gameEngine.js:148-151.  Returns a dict of the eight shadow-sprite arrays + "count"."""
        if maxShadowSprites is None:
            maxShadowSprites = maxShadowCastingLights * maxShadowsPerLight
        S = int(maxShadowSprites)
        out = {"active": np.zeros(S, np.uint8)}
        for k in ("radius", "x", "y", "rotation", "scaleX", "scaleY", "alpha"):
            out[k] = np.zeros(S, np.float32)
        sp = B.ShadowSprites(*[out[k].ctypes.data for k in ("active", "radius", "x", "y", "rotation", "scaleX", "scaleY", "alpha")])
        n = C.c_uint32(0)
        B.check(self.ctx, B.lib().weed_system_shadows(self.ctx, int(maxShadowCastingLights), int(maxShadowsPerLight), S,
                                                      C.byref(sp), C.byref(n)))
        out["count"] = int(n.value)
        return out

    # ---- spawn / despawn pools ---------------------------------------------------
    def create_pool(self, startIndex, totalCount, rigidBody=True, collider=True):
        """One entity class = one pool (GameObject.initializeFreeList, gameObject.js:794-833)."""
        pid = C.c_uint32(0)
        comps = (B.POOL_HAS_RIGIDBODY if rigidBody else 0) | (B.POOL_HAS_COLLIDER if collider else 0)
        B.check(self.ctx, B.lib().weed_pool_create(self.ctx, int(startIndex), int(totalCount), comps, C.byref(pid)))
        return int(pid.value)

    def spawn(self, pool, records):
        """GameObject.spawn (:840-951) for a batch: records = float32[n,4] (x, y, vx, vy).  Returns
        the entity indices (-1 where the pool was exhausted).  Host columns are stale until
        downloaded."""
        rec = np.ascontiguousarray(records, dtype=np.float32).reshape(-1, 4)
        out = np.full(len(rec), -1, np.int32)
        B.check(self.ctx, B.lib().weed_pool_spawn(self.ctx, int(pool), rec.ctypes.data, len(rec), out.ctypes.data))
        return out

    def despawn(self, pool, indices):
        """GameObject.despawn (:668-690) for a batch, in order.  Returns how many were active."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        n = C.c_uint32(0)
        B.check(self.ctx, B.lib().weed_pool_despawn(self.ctx, int(pool), idx.ctypes.data, len(idx), C.byref(n)))
        return int(n.value)

    def despawn_all(self, pool):
        """GameObject.despawnAll (:1001-1034)."""
        n = C.c_uint32(0)
        B.check(self.ctx, B.lib().weed_pool_despawn_all(self.ctx, int(pool), C.byref(n)))
        return int(n.value)

    def pool_stats(self, pool):
        t, a = C.c_uint32(0), C.c_uint32(0)
        B.check(self.ctx, B.lib().weed_pool_stats(self.ctx, int(pool), C.byref(t), C.byref(a)))
        return {"total": int(t.value), "available": int(a.value), "active": int(t.value) - int(a.value)}

    def updatePhysicsConfig(self, partial):
        """gameEngine.js:1304-1325 -> applyPhysicsConfig + validatePhysicsConfig."""
        phys = self.config["physics"]
        for k, v in partial.items():
            if k == "gravity":
                phys["gravity"] = {**phys.get("gravity", {}), **v}
            else:
                phys[k] = v
        p = self._physics_struct(phys)
        B.check(self.ctx, B.lib().weed_set_physics(self.ctx, C.byref(p)))

    def physics_settings(self):
        p = B.PhysicsConfig()
        B.check(self.ctx, B.lib().weed_get_physics(self.ctx, C.byref(p)))
        return {n: getattr(p, n) for n, _ in p._fields_ if not n.startswith("_")}

    def stats(self):
        s = B.Stats()
        B.check(self.ctx, B.lib().weed_get_stats(self.ctx, C.byref(s)))
        d = {n: getattr(s, n) for n, _ in s._fields_ if n not in ("ms", "_pad")}
        d["ms"] = list(s.ms)
        return d

    def device_ptr(self, which):
        p, n = C.c_void_p(), C.c_size_t()
        B.check(self.ctx, B.lib().weed_device_ptr(self.ctx, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def close(self):
        if getattr(self, "ctx", None):
            if self._owns_ctx:
                B.lib().weed_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
