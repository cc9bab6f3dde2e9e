"""ctypes wrapper around oracle/libweedoracle.so (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

COLS = {  # key -> (component id, schema name, dtype)
    "T.active": (0, "active", np.uint8), "T.x": (0, "x", np.float32), "T.y": (0, "y", np.float32),
    "RB.active": (1, "active", np.uint8), "RB.static": (1, "static", np.uint8),
    "RB.vx": (1, "vx", np.float32), "RB.vy": (1, "vy", np.float32),
    "RB.ax": (1, "ax", np.float32), "RB.ay": (1, "ay", np.float32),
    "RB.px": (1, "px", np.float32), "RB.py": (1, "py", np.float32),
    "RB.maxVel": (1, "maxVel", np.float32), "RB.velocityAngle": (1, "velocityAngle", np.float32),
    "RB.speed": (1, "speed", np.float32), "RB.collisionCount": (1, "collisionCount", np.uint8),
    "C.active": (2, "active", np.uint8), "C.radius": (2, "radius", np.float32),
    "C.isTrigger": (2, "isTrigger", np.uint8), "C.visualRange": (2, "visualRange", np.float32),
}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libweedoracle.so")
    src = os.path.join(_HERE, "weed_oracle.c")
    src2 = os.path.join(_HERE, "weed_oracle_systems.c")
    hdr = os.path.join(_HERE, "..", "include", "weed_nudge.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in (src, src2, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libweedoracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.wo_create.restype = C.c_void_p
        L.wo_create.argtypes = [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_double]
        L.wo_destroy.argtypes = [C.c_void_p]
        L.wo_set_physics.argtypes = [C.c_void_p, C.c_int32] + [C.c_double] * 6
        L.wo_bind.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.wo_spatial.argtypes = [C.c_void_p]
        L.wo_physics.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.wo_step.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.wo_grid_export.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.wo_grid_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.wo_column_offset.restype = C.c_int64
        L.wo_column_offset.argtypes = [C.c_int, C.c_char_p, C.c_int64]
        L.wo_buffer_size.restype = C.c_int64
        L.wo_buffer_size.argtypes = [C.c_int, C.c_int64]
        L.wo_bench_freerun.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.wo_bench_lockstep.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.wo_js_toint32.restype = C.c_int32
        L.wo_js_toint32.argtypes = [C.c_double]
        L.wo_seeded_random.restype = C.c_double
        L.wo_seeded_random.argtypes = [C.c_double, C.c_int]
        L.wo_nudge_dir.argtypes = [C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.wo_nudge_hash.restype = C.c_uint32
        L.wo_nudge_hash.argtypes = [C.c_uint32] * 5
        L.wo_events_create.restype = C.c_void_p
        L.wo_events_destroy.argtypes = [C.c_void_p]
        L.wo_events_process.restype = C.c_int64
        L.wo_events_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.wo_screen_visibility.argtypes = [C.c_int32] + [C.c_void_p] * 3 + [C.c_double] * 5 + [C.c_void_p] * 3
        L.wo_shadow_sprites.restype = C.c_int32
        L.wo_shadow_sprites.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 11 + [C.c_int32] * 3 + [C.c_void_p] * 8
        L.wo_pool_create.restype = C.c_void_p
        L.wo_pool_create.argtypes = [C.c_int32, C.c_int32, C.c_int, C.c_int]
        L.wo_pool_destroy.argtypes = [C.c_void_p]
        L.wo_pool_available.restype = C.c_int32
        L.wo_pool_available.argtypes = [C.c_void_p]
        L.wo_pool_spawn.restype = C.c_int32
        L.wo_pool_spawn.argtypes = [C.c_void_p, C.c_void_p] + [C.c_float] * 4
        L.wo_pool_despawn.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.wo_pool_despawn_all.restype = C.c_int32
        L.wo_pool_despawn_all.argtypes = [C.c_void_p, C.c_void_p]
        L.wo_pool_free_list.argtypes = [C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


class PoolC:
    """GameObject.initializeFreeList / spawn / despawn / despawnAll (gameObject.js:794-951, 668-690,
    1001-1034) over a dict of numpy columns keyed like OracleC.col."""
    _KEYS = ["T.active", "RB.active", "C.active", "T.x", "T.y", "RB.vx", "RB.vy", "RB.ax", "RB.ay", "RB.px", "RB.py",
             "RB.speed", "RB.velocityAngle"]

    def __init__(self, col, startIndex, totalCount, rigidBody=True, collider=True):
        self.col = col
        self.total = totalCount
        self.h = lib().wo_pool_create(int(startIndex), int(totalCount), int(rigidBody), int(collider))
        self._cols = (C.c_void_p * len(self._KEYS))(*[col[k].ctypes.data for k in self._KEYS])

    def spawn(self, records):
        rec = np.asarray(records, np.float32).reshape(-1, 4)
        return np.array([lib().wo_pool_spawn(self.h, self._cols, *[float(v) for v in r]) for r in rec], np.int32)

    def despawn(self, indices):
        return sum(lib().wo_pool_despawn(self.h, self._cols, int(i)) for i in indices)

    def despawn_all(self):
        return int(lib().wo_pool_despawn_all(self.h, self._cols))

    def available(self):
        return int(lib().wo_pool_available(self.h))

    def free_list(self):
        out = np.zeros(self.total, np.int32)
        lib().wo_pool_free_list(self.h, out.ctypes.data)
        return out[:self.available()]

    def __del__(self):
        if getattr(self, "h", None):
            lib().wo_pool_destroy(self.h)
            self.h = None


class CollisionEventsC:
    """LogicWorker.processCollisionCallbacks (logic_worker.js:429-526), one logic worker."""

    def __init__(self):
        self.h = lib().wo_events_create()

    def process(self, collisionData):
        cd = np.ascontiguousarray(collisionData, np.int32)
        cap = 8 * int(cd[0]) + 8 * getattr(self, "_last", 0) + 16
        calls = np.zeros(3 * cap, np.int32)
        n = lib().wo_events_process(self.h, cd.ctypes.data, calls.ctypes.data, cap)
        assert n <= cap
        self._last = int(cd[0])
        return [tuple(t) for t in calls[:3 * n].reshape(n, 3).tolist()]

    def __del__(self):
        if getattr(self, "h", None):
            lib().wo_events_destroy(self.h)
            self.h = None


def screen_visibility_c(active, x, y, zoom, cameraX, cameraY, canvasWidth, canvasHeight, screenX, screenY, isItOnScreen):
    """particle_worker.js:1012-1062; the three output arrays are updated in place."""
    active = np.ascontiguousarray(active, np.uint8); x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    lib().wo_screen_visibility(len(active), active.ctypes.data, x.ctypes.data, y.ctypes.data, float(zoom), float(cameraX),
                               float(cameraY), float(canvasWidth), float(canvasHeight), screenX.ctypes.data,
                               screenY.ctypes.data, isItOnScreen.ctypes.data)


def shadow_sprites_c(maxNeighbors, neighborData, distanceData, transformActive, worldX, worldY, lightEnabled, lightIntensity,
                     shadowCasterActive, shadowRadius, shadowHeight, isOnScreen, maxLights=20, perLight=15, maxSprites=None):
    """particle_worker.js:861-1003.  Returns the dict of shadow-sprite arrays + "count"."""
    if maxSprites is None:
        maxSprites = maxLights * perLight
    ins = [np.ascontiguousarray(neighborData, np.int32), np.ascontiguousarray(distanceData, np.float32),
           np.ascontiguousarray(transformActive, np.uint8), np.ascontiguousarray(worldX, np.float32),
           np.ascontiguousarray(worldY, np.float32), np.ascontiguousarray(lightEnabled, np.uint8),
           np.ascontiguousarray(lightIntensity, np.float32), np.ascontiguousarray(shadowCasterActive, np.uint8),
           np.ascontiguousarray(shadowRadius, np.float32), np.ascontiguousarray(shadowHeight, np.float32),
           np.ascontiguousarray(isOnScreen, np.uint8)]
    out = {"active": np.zeros(maxSprites, np.uint8)}
    for k in ("radius", "x", "y", "rotation", "scaleX", "scaleY", "alpha"):
        out[k] = np.zeros(maxSprites, np.float32)
    n = lib().wo_shadow_sprites(len(ins[2]), int(maxNeighbors), *[a.ctypes.data for a in ins], int(maxLights), int(perLight),
                                int(maxSprites), *[out[k].ctypes.data for k in ("active", "radius", "x", "y", "rotation", "scaleX", "scaleY", "alpha")])
    out["count"] = int(n)
    return out


class OracleC:
    """The C oracle over its own SAB-layout buffers; .col[...] are numpy views into them."""

    def __init__(self, N, worldWidth, worldHeight, cellSize, maxNeighbors, maxPairs=10000,
                 seed=1.0, physics=None):
        L = lib()
        self.N, self.M, self.maxPairs = N, maxNeighbors, maxPairs
        self.bufs = [np.zeros(L.wo_buffer_size(k, N), dtype=np.uint8) for k in range(3)]
        self.neighborData = np.zeros(N * (1 + maxNeighbors), dtype=np.int32)
        self.distanceData = np.zeros(N * (1 + maxNeighbors), dtype=np.float32)
        self.collisionData = np.zeros(1 + 2 * maxPairs, dtype=np.int32)
        self.h = L.wo_create(N, worldWidth, worldHeight, cellSize, maxNeighbors, maxPairs, float(seed))
        for k in range(3):
            L.wo_bind(self.h, k, self.bufs[k].ctypes.data)
        L.wo_bind(self.h, 3, self.neighborData.ctypes.data)
        L.wo_bind(self.h, 4, self.distanceData.ctypes.data)
        L.wo_bind(self.h, 5, self.collisionData.ctypes.data)
        self.col = {}
        for key, (comp, name, dt) in COLS.items():
            off = L.wo_column_offset(comp, name.encode(), N)
            self.col[key] = self.bufs[comp][off:off + N * np.dtype(dt).itemsize].view(dt)
        cols, rows = C.c_int32(), C.c_int32()
        L.wo_grid_dims(self.h, C.byref(cols), C.byref(rows))
        self.gridCols, self.gridRows = cols.value, rows.value
        if physics:
            self.set_physics(**physics)

    def set_physics(self, subStepCount=4, boundaryElasticity=0.8, collisionResponseStrength=0.5,
                    verletDamping=0.995, minSpeedForRotation=0.1, gravityX=0.0, gravityY=0.0):
        lib().wo_set_physics(self.h, subStepCount, boundaryElasticity, collisionResponseStrength,
                             verletDamping, minSpeedForRotation, gravityX, gravityY)

    def load(self, columns):
        for k, v in columns.items():
            self.col[k][:] = v

    def spatial(self):
        lib().wo_spatial(self.h)

    def physics(self, dtRatio=1.0, order=0):
        lib().wo_physics(self.h, float(dtRatio), order)

    def step(self, dtRatio=1.0, order=0):
        lib().wo_step(self.h, float(dtRatio), order)

    def grid_csr(self):
        Cn = self.gridCols * self.gridRows
        cellOf = np.zeros(self.N, dtype=np.int32)
        start = np.zeros(Cn + 1, dtype=np.int32)
        idx = np.zeros(self.N, dtype=np.int32)
        lib().wo_grid_export(self.h, cellOf.ctypes.data, start.ctypes.data, idx.ctypes.data)
        return cellOf, start, idx[: start[Cn]]

    def bench(self, frames, dtRatio=1.0, freerun=True):
        a, b = C.c_double(), C.c_double()
        f = lib().wo_bench_freerun if freerun else lib().wo_bench_lockstep
        f(self.h, frames, float(dtRatio), C.byref(a), C.byref(b))
        return a.value, b.value

    def close(self):
        if self.h:
            lib().wo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
