"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/weedgpu.h declares, reproduces the Component.js layout rule, and fails loudly (no
CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from multithreadedgameengine_b200 import binding as B
from multithreadedgameengine_b200.components import Collider, RigidBody, Transform
from oracle.oracle_np import layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    entry.build()
    return B.lib()


def test_every_declared_symbol_is_exported_and_bound(L):
    hdr = open(os.path.join(ROOT, "include", "weedgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(weed_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(B.SYMBOLS), declared ^ set(B.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_struct_sizes_match_header(L):
    cfg = B.Config()
    L.weed_default_config(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(B.Config)
    # engine defaults: gameEngine.js:39-49, :553, :689-693
    assert (cfg.maxNeighbors, cfg.maxCollisionPairs) == (100, 10000)
    p = cfg.physics
    assert (p.subStepCount, p.boundaryElasticity, p.collisionResponseStrength, p.verletDamping,
            p.minSpeedForRotation, p.gravityX, p.gravityY) == (4, 0.8, 0.5, 0.995, 0.1, 0.0, 0.0)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof of every public struct as a C compiler sees include/weedgpu.h == the ctypes mirror."""
    import subprocess
    pairs = [("weed_config", B.Config), ("weed_physics_config", B.PhysicsConfig), ("weed_stats", B.Stats),
             ("weed_boids_params", B.BoidsParams), ("weed_flock_class", B.FlockClass), ("weed_flock_params", B.FlockParams),
             ("weed_collision_event_counts", B.CollisionEventCounts), ("weed_camera", B.Camera),
             ("weed_shadow_columns", B.ShadowColumns), ("weed_shadow_sprites", B.ShadowSprites),
             ("weed_slab_stats", B.SlabStats)]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "weedgpu.h"\nint main(void){' +
                   "".join(f'printf("%zu\\n", sizeof({n}));' for n, _ in pairs) + "return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(t) for _, t in pairs], list(zip([n for n, _ in pairs], got, [C.sizeof(t) for _, t in pairs]))


@pytest.mark.parametrize("N", [1, 2, 3, 5, 1001, 4099, 65537])
def test_layout_rule_matches_component_js(L, N):
    for bid, cls, name in ((0, Transform, "Transform"), (1, RigidBody, "RigidBody"), (2, Collider, "Collider")):
        ref, size = layout(name, N)              # oracle restatement of Component.js:20-42
        assert L.weed_buffer_bytes(bid, N, 7, 9) == size == cls.getBufferSize(N)
        assert L.weed_column_count(bid) == len(cls.ARRAY_SCHEMA)
        for k, col in enumerate(cls.ARRAY_SCHEMA):
            assert L.weed_column_name(bid, k).decode() == col
            assert L.weed_column_offset(bid, k, N) == ref[col] == cls.columnOffset(col, N)
        assert L.weed_column_offset(bid, len(cls.ARRAY_SCHEMA), N) == C.c_size_t(-1).value
    assert L.weed_buffer_bytes(B.BUF_NEIGHBOR, N, 7, 9) == N * 8 * 4
    assert L.weed_buffer_bytes(B.BUF_DISTANCE, N, 7, 9) == N * 8 * 4
    assert L.weed_buffer_bytes(B.BUF_COLLISION, N, 7, 9) == (1 + 18) * 4


def test_survey_a1_offsets(L):
    # SURVEY §8 a1, N = 1001
    assert L.weed_column_offset(0, 2, 1001) == 2004 and L.weed_buffer_bytes(0, 1001, 0, 0) == 14016
    assert L.weed_column_offset(1, 22, 1001) == 82084 and L.weed_buffer_bytes(1, 1001, 0, 0) == 83085
    assert L.weed_column_offset(2, 15, 1001) == 47052 and L.weed_buffer_bytes(2, 1001, 0, 0) == 51056


def test_component_views_alias_one_buffer():
    T = type("T", (Transform,), {})
    buf = np.zeros(T.getBufferSize(10), dtype=np.uint8)
    T.initializeArrays(buf, 10)
    T.x[3] = 1.5
    T.active[9] = 1
    assert buf[T.columnOffset("x", 10) + 12:T.columnOffset("x", 10) + 16].view(np.float32)[0] == 1.5
    assert buf[9] == 1
    with pytest.raises(ValueError):
        T.initializeArrays(np.zeros(5, dtype=np.uint8), 10)


def test_invalid_configs_are_rejected_without_touching_cuda(L):
    ctx = C.c_void_p()
    cfg = B.Config()
    L.weed_default_config(C.byref(cfg))
    cfg.entityCount, cfg.worldWidth, cfg.worldHeight, cfg.cellSize = 10, 100.0, 100.0, 10.0
    bad = B.Config.from_buffer_copy(cfg); bad.struct_size = 8
    assert L.weed_create(C.byref(bad), C.byref(ctx)) == B.WEED_E_INVALID
    bad = B.Config.from_buffer_copy(cfg); bad.entityCount = 0
    assert L.weed_create(C.byref(bad), C.byref(ctx)) == B.WEED_E_INVALID
    bad = B.Config.from_buffer_copy(cfg); bad.cellSize = 0.0
    assert L.weed_create(C.byref(bad), C.byref(ctx)) == B.WEED_E_INVALID
    assert b"positive" in L.weed_last_error(None)
    assert L.weed_create(None, C.byref(ctx)) == B.WEED_E_INVALID


def test_no_cpu_fallback(L):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    ctx = C.c_void_p()
    cfg = B.Config()
    L.weed_default_config(C.byref(cfg))
    cfg.entityCount, cfg.worldWidth, cfg.worldHeight, cfg.cellSize = 10, 100.0, 100.0, 10.0
    assert L.weed_create(C.byref(cfg), C.byref(ctx)) == B.WEED_E_CUDA
    assert b"no CPU fallback" in L.weed_last_error(None)
    from multithreadedgameengine_b200.engine import GameEngine
    with pytest.raises(B.WeedError):
        GameEngine(dict(entityCount=10, worldWidth=100.0, worldHeight=100.0, spatial=dict(cellSize=10.0, maxNeighbors=4)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multithreadedgameengine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|oracle[/.]|libweedoracle|weed_oracle", src), \
                    f"{f} references the oracle"


def test_library_holds_sm_100a_code_for_every_frame_kernel(L):
    """The built library carries native sm_100a code (not PTX to be JIT-compiled for something else) for
    every kernel of the frame, the settled-bed forms included, and the sweep keeps its register budget
    (40 registers: 12 blocks of 128 threads per SM — the occupancy the measurements rest on)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    so = os.path.join(ROOT, "multithreadedgameengine_b200", "libweedgpu.so")
    out = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in out
    for k in ("k_cell_key", "k_cell_scan", "k_scatter_ids", "k_sort_big_cells", "k_slot_rank", "k_build_slots", "k_slot_prep",
              "k_neighbors2", "k_neighbors_wide", "k_beyond_cap", "k_back_alloc", "k_back_write", "k_back_sort", "k_sort_lists",
              "k_sweep", "k_sweep_heavy", "k_sweep_tile", "k_writeback", "k_pair_scan", "k_pair_emit",
              "k_slab_pack", "k_slab_wait", "k_slab_unpack"):
        assert re.search(r"Function _ZN4weed\d+%s[A-Z]" % k, out), k
    sweep = re.findall(r"Function _ZN4weed7k_sweepILb[01]ELb[01]E\S*:\s*\n\s*REG:(\d+)", out)
    assert len(sweep) == 4 and all(int(r) <= 40 for r in sweep), sweep
