"""ctypes binding of the C ABI in include/weedgpu.h (libweedgpu.so, built in-tree by
__graft_entry__.build()).  This is the binding the tests and bench drive; a Node engine would
bind the same symbols through the N-API shim in addon/weed_napi.cc.

No fallback: if the shared library is missing, ``lib()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libweedgpu.so")

WEED_OK, WEED_E_INVALID, WEED_E_CUDA, WEED_E_NOT_BOUND, WEED_E_SIZE, WEED_E_OVERFLOW, WEED_E_STATE, WEED_E_NOMEM = (
    0, -1, -2, -3, -4, -5, -6, -7)
ERROR_NAMES = {0: "WEED_OK", -1: "WEED_E_INVALID", -2: "WEED_E_CUDA", -3: "WEED_E_NOT_BOUND", -4: "WEED_E_SIZE",
               -5: "WEED_E_OVERFLOW", -6: "WEED_E_STATE", -7: "WEED_E_NOMEM"}

BUF_TRANSFORM, BUF_RIGIDBODY, BUF_COLLIDER, BUF_NEIGHBOR, BUF_DISTANCE, BUF_COLLISION = range(6)

# WEED_COL_* bits, in the order of include/weedgpu.h
COL = {
    "T.active": 1 << 0, "T.x": 1 << 1, "T.y": 1 << 2,
    "RB.active": 1 << 3, "RB.static": 1 << 4, "RB.vx": 1 << 5, "RB.vy": 1 << 6, "RB.ax": 1 << 7,
    "RB.ay": 1 << 8, "RB.px": 1 << 9, "RB.py": 1 << 10, "RB.maxVel": 1 << 11,
    "RB.velocityAngle": 1 << 12, "RB.speed": 1 << 13, "RB.collisionCount": 1 << 14,
    "C.active": 1 << 15, "C.radius": 1 << 16, "C.isTrigger": 1 << 17, "C.visualRange": 1 << 18,
    "T.entityType": 1 << 19,
}
COLS_INPUT_ALL = 0x000FFFFF
COLS_OUTPUT_ALL = (COL["T.x"] | COL["T.y"] | COL["RB.vx"] | COL["RB.vy"] | COL["RB.ax"] | COL["RB.ay"]
                   | COL["RB.px"] | COL["RB.py"] | COL["RB.velocityAngle"] | COL["RB.speed"]
                   | COL["RB.collisionCount"])
COL_NEIGHBORS = 1 << 24
COL_COLLISIONS = 1 << 25

FLAG_NO_GRAPH = 1 << 0
FLAG_KERNEL_TIMING = 1 << 1
FLAG_NO_NEIGHBOR_ROWS = 1 << 2
FLAG_K6_TILE = 1 << 10
FLAG_K4_WIDE = 1 << 12
FLAG_K4_THREAD = 1 << 13

DEV_NEIGHBOR, DEV_DISTANCE, DEV_COLLISION, DEV_STATE, DEV_ATTR, DEV_VEL, DEV_NEIGHBOR_COUNT, DEV_SLOT_OF = range(8)


class PhysicsConfig(C.Structure):
    _fields_ = [("subStepCount", C.c_int32), ("_pad0", C.c_int32),
                ("boundaryElasticity", C.c_double), ("collisionResponseStrength", C.c_double),
                ("verletDamping", C.c_double), ("minSpeedForRotation", C.c_double),
                ("gravityX", C.c_double), ("gravityY", C.c_double)]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("entityCount", C.c_uint32),
                ("worldWidth", C.c_double), ("worldHeight", C.c_double), ("cellSize", C.c_double),
                ("maxNeighbors", C.c_uint32), ("maxCollisionPairs", C.c_uint32),
                ("seed", C.c_double), ("physics", PhysicsConfig),
                ("device", C.c_int32), ("flags", C.c_uint32), ("stream", C.c_void_p),
                ("slabRowBegin", C.c_uint32), ("slabRowEnd", C.c_uint32),
                ("slabHaloRows", C.c_uint32), ("_pad1", C.c_uint32)]


class BoidsParams(C.Structure):
    _fields_ = [("centeringFactor", C.c_double), ("avoidFactor", C.c_double), ("matchingFactor", C.c_double),
                ("turnFactor", C.c_double), ("margin", C.c_double), ("mouseEntityType", C.c_uint32), ("_pad", C.c_uint32)]


class FlockClass(C.Structure):
    _fields_ = [("entityType", C.c_uint32), ("role", C.c_uint32), ("otherEntityType", C.c_uint32), ("_pad", C.c_uint32),
                ("protectedRangeScale", C.c_double), ("centeringFactor", C.c_double), ("avoidFactor", C.c_double),
                ("matchingFactor", C.c_double), ("turnFactor", C.c_double), ("margin", C.c_double), ("roleFactor", C.c_double)]


class FlockParams(C.Structure):
    _fields_ = [("mouseEntityType", C.c_uint32), ("mouseDown", C.c_uint32), ("dtRatio", C.c_double)]


class CollisionEventCounts(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("pairs", "entered", "stayed", "exited")]


class Camera(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("zoom", "cameraX", "cameraY", "canvasWidth", "canvasHeight")]


class ShadowColumns(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("lightActive", "lightIntensity", "casterActive", "shadowRadius", "height")]


class ShadowSprites(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("active", "radius", "x", "y", "rotation", "scaleX", "scaleY", "alpha")]


class SlabStats(C.Structure):
    _fields_ = ([(n, C.c_uint32) for n in ("top", "capacity", "owned", "sentLow", "sentHigh", "receivedLow",
                                            "receivedHigh", "overflow")]
                + [("rowBegin", C.c_int32), ("rowEnd", C.c_int32), ("cutMoves", C.c_uint32), ("loadNs", C.c_uint32)])


class Stats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("gridCols", C.c_uint32), ("gridRows", C.c_uint32),
                ("activeInGrid", C.c_uint32), ("maxCellOccupancy", C.c_uint32),
                ("neighborsTotal", C.c_uint64), ("cappedRows", C.c_uint32),
                ("explicitPairs", C.c_uint32), ("collisionPairs", C.c_uint32),
                ("kernelLaunchesPerStep", C.c_uint32), ("ms", C.c_float * 12)]


# every symbol include/weedgpu.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "weed_buffer_bytes": (C.c_size_t, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]),
    "weed_column_offset": (C.c_size_t, [C.c_int, C.c_uint32, C.c_uint32]),
    "weed_column_count": (C.c_uint32, [C.c_int]),
    "weed_column_name": (C.c_char_p, [C.c_int, C.c_uint32]),
    "weed_default_config": (None, [C.POINTER(Config)]),
    "weed_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "weed_destroy": (None, [C.c_void_p]),
    "weed_bind": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "weed_upload": (C.c_int, [C.c_void_p, C.c_uint32]),
    "weed_download": (C.c_int, [C.c_void_p, C.c_uint32]),
    "weed_spatial": (C.c_int, [C.c_void_p]),
    "weed_physics": (C.c_int, [C.c_void_p, C.c_double]),
    "weed_step": (C.c_int, [C.c_void_p, C.c_double, C.c_uint32, C.c_uint32]),
    "weed_run": (C.c_int, [C.c_void_p, C.c_double, C.c_uint32]),
    "weed_set_physics": (C.c_int, [C.c_void_p, C.POINTER(PhysicsConfig)]),
    "weed_get_physics": (C.c_int, [C.c_void_p, C.POINTER(PhysicsConfig)]),
    "weed_fetch_neighbors": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "weed_fetch_neighbors_to": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "weed_sync": (C.c_int, [C.c_void_p]),
    "weed_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "weed_last_error": (C.c_char_p, [C.c_void_p]),
    "weed_device_ptr": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "weed_row_pitch": (C.c_uint32, [C.c_void_p]),
    "weed_entity_count": (C.c_uint32, [C.c_void_p]),
    "weed_system_boids": (C.c_int, [C.c_void_p, C.POINTER(BoidsParams), C.c_void_p, C.c_double]),
    "weed_system_flock": (C.c_int, [C.c_void_p, C.POINTER(FlockClass), C.c_uint32, C.POINTER(FlockParams), C.c_void_p]),
    "weed_system_collision_events": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(CollisionEventCounts), C.c_void_p, C.c_void_p]),
    "weed_system_screen_visibility": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.c_void_p, C.c_void_p, C.c_void_p]),
    "weed_system_shadows_upload": (C.c_int, [C.c_void_p, C.POINTER(ShadowColumns)]),
    "weed_system_shadows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(ShadowSprites), C.POINTER(C.c_uint32)]),
    "weed_pool_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    "weed_pool_spawn": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]),
    "weed_pool_despawn": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "weed_pool_despawn_all": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "weed_pool_stats": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "weed_slab_set_gids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "weed_slab_get_gids": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]),
    "weed_slab_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "weed_slab_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "weed_slab_balance": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "weed_slab_status": (C.c_int, [C.c_void_p, C.POINTER(SlabStats)]),
    "weed_slab_exchange_create": (C.c_int, [C.c_void_p, C.c_uint32]),
    "weed_slab_exchange_export": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "weed_slab_exchange_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "weed_slab_frame": (C.c_int, [C.c_void_p, C.c_double]),
    "weed_slab_frame_begin": (C.c_int, [C.c_void_p, C.c_double]),
    "weed_slab_frame_end": (C.c_int, [C.c_void_p]),
    "weed_slab_exchange": (C.c_int, [C.c_void_p]),
    "weed_slab_exchange_disconnect": (C.c_int, [C.c_void_p]),
    "weed_group_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "weed_group_size": (C.c_uint32, [C.c_void_p]),
    "weed_group_slab": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "weed_group_step": (C.c_int, [C.c_void_p, C.c_double]),
    "weed_group_sync": (C.c_int, [C.c_void_p]),
    "weed_group_destroy": (None, [C.c_void_p]),
    "weed_group_last_error": (C.c_char_p, [C.c_void_p]),
}
SLAB_RECORD_BYTES = 64
FLOCK_BOID, FLOCK_PREY, FLOCK_PREDATOR, FLOCK_ANY_TYPE = 0, 1, 2, 0xFFFFFFFF
POOL_HAS_RIGIDBODY, POOL_HAS_COLLIDER = 1, 2
EVENTS_FORGET_PREVIOUS = 1
COLLISION_ENTER, COLLISION_STAY, COLLISION_EXIT = 1, 2, 3

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for the spatial+physics path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


class WeedError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


def check(ctx, rc):
    if rc != WEED_OK:
        msg = lib().weed_last_error(ctx)
        raise WeedError(rc, msg.decode() if msg else "")
    return rc
