"""addon/weed_napi.cc — the N-API shim a Node engine loads (reference seam: AbstractWorker.js:298-330,
gameEngine.js:1049-1125) — compiled against tests/mock/node_api.h (this image has neither Node nor its
headers) and driven through a fake napi_env (tests/mock/fake_napi.cc):

  * without a GPU: the addon compiles, links against libweedgpu.so, and `create` surfaces the library's
    "no CPU fallback" error as a JavaScript exception;
  * on the GPU box: create -> bind x6 -> step x3 -> fetchNeighbors through the addon produce, bit for
    bit, the columns, rows and collisionData the ctypes binding produces for the same scene; bound
    buffers are referenced while the context lives and released by its finalizer; a typed array that is
    too short is refused.
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multithreadedgameengine_b200")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    import __graft_entry__ as entry
    entry.build()
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("addon") / "addon_harness")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "mock"), "-I" + os.path.join(ROOT, "include"),
           "-o", exe, os.path.join(ROOT, "tests", "mock", "fake_napi.cc"), os.path.join(ROOT, "addon", "weed_napi.cc"),
           "-L" + PKG, "-lweedgpu", "-Wl,-rpath," + PKG]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_addon_compiles_and_reports_missing_gpu_as_exception(harness, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the full run is test_addon_matches_ctypes_binding")
    r = subprocess.run([harness, str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


def scene():
    """The scene tests/mock/fake_napi.cc builds (same LCG, same float32 arithmetic)."""
    from multithreadedgameengine_b200 import scenes
    N, W, H = 2001, np.float32(1600), np.float32(800)
    c = scenes._blank(N)
    s = 12345

    def unit():
        nonlocal s
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        return np.float32(s >> 8) * np.float32(1.0 / 16777216.0)
    c["T.active"][0] = 1; c["C.active"][0] = 1; c["C.isTrigger"][0] = 1; c["C.visualRange"][0] = 150
    for i in range(1, N):
        c["T.active"][i] = c["RB.active"][i] = c["C.active"][i] = 1
        c["T.x"][i] = unit() * W
        c["T.y"][i] = unit() * H
        c["RB.px"][i] = c["T.x"][i]; c["RB.py"][i] = c["T.y"][i]
        c["C.radius"][i] = np.float32(10) + np.float32(20) * unit()
        c["C.visualRange"][i] = np.float32(66.5); c["RB.maxVel"][i] = 50
    cfg = dict(entityCount=N, worldWidth=1600.0, worldHeight=800.0, seed=1234,
               spatial=dict(cellSize=50.0, maxNeighbors=24),
               physics=dict(subStepCount=2, gravity=dict(x=0.0, y=0.5), verletDamping=0.99, maxCollisionPairs=5000))
    return cfg, c


@pytest.mark.gpu
def test_addon_matches_ctypes_binding(harness, tmp_path):
    from multithreadedgameengine_b200 import binding as B
    from multithreadedgameengine_b200.engine import GameEngine
    out = str(tmp_path / "out.bin")
    r = subprocess.run([harness, out, "3"], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    raw = np.fromfile(out, dtype=np.uint32)
    N, M = int(raw[0]), int(raw[1])
    x, y = raw[2:2 + N].view(np.float32), raw[2 + N:2 + 2 * N].view(np.float32)
    nd = raw[2 + 2 * N:2 + 2 * N + N * (1 + M)].view(np.int32)
    coll = raw[2 + 2 * N + N * (1 + M):].view(np.int32)
    cfg, cols = scene()
    eng = GameEngine(cfg)
    eng.load_columns(cols)
    for _ in range(3):
        eng.step(1.0, 0, B.COLS_INPUT_ALL | B.COL_NEIGHBORS | B.COL_COLLISIONS)
    assert np.array_equal(x.view(np.uint32), eng.col["T.x"].view(np.uint32))
    assert np.array_equal(y.view(np.uint32), eng.col["T.y"].view(np.uint32))
    stride = 1 + M
    for i in range(N):
        n = int(eng.neighborData[i * stride])
        assert int(nd[i * stride]) == n and np.array_equal(nd[i * stride + 1:i * stride + 1 + n], eng.neighborData[i * stride + 1:i * stride + 1 + n])
    n = int(eng.collisionData[0])
    assert n > 0 and int(coll[0]) == n and np.array_equal(coll[1:1 + 2 * n], eng.collisionData[1:1 + 2 * n])
    eng.close()
