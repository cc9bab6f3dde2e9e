"""Drive baseline/node_harness.mjs (the reference's real JS workers under Node) from Python.

Returns None when `node` or the reference tree is not available — which is the case in this
image; see SURVEY.md §0.  Used by tests/test_node_reference.py and, when present, reported by
bench.py --impl reference as a separate figure."""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def available(ref_root="/root/reference"):
    return shutil.which("node") is not None and os.path.isdir(os.path.join(ref_root, "src", "workers"))


def run(cfg, cols, frames, ref_root="/root/reference"):
    """-> dict(columns..., neighborData, distanceData, collisionData, seconds) or None."""
    if not available(ref_root):
        return None
    from oracle.oracle_np import SCHEMAS, layout
    N = cfg["entityCount"]
    M = cfg["spatial"]["maxNeighbors"]
    P = int((cfg.get("physics") or {}).get("maxCollisionPairs") or 10000)
    names = {"T": "Transform", "RB": "RigidBody", "C": "Collider"}
    bufs, lays = {}, {}
    for short, name in names.items():
        lay, size = layout(name, N)
        b = np.zeros(size, dtype=np.uint8)
        types = dict(SCHEMAS[name])
        for key, v in cols.items():
            pre, colname = key.split(".")
            if pre != short:
                continue
            dt = {1: np.uint8, 2: np.uint16, 4: np.float32}[types[colname]]
            b[lay[colname]:lay[colname] + N * types[colname]].view(dt)[:] = v
        bufs[name], lays[name] = b, (lay, types)
    with tempfile.TemporaryDirectory() as tmp:
        sj, sb, ob = (os.path.join(tmp, f) for f in ("scene.json", "scene.bin", "out.bin"))
        config = dict(cfg)
        config.setdefault("seed", 1)
        json.dump({"entityCount": N, "config": config,
                   "sizes": {"Transform": int(bufs["Transform"].size), "RigidBody": int(bufs["RigidBody"].size),
                             "Collider": int(bufs["Collider"].size), "neighbor": N * (1 + M) * 4,
                             "collision": (1 + 2 * P) * 4}}, open(sj, "w"))
        np.concatenate([bufs[n] for n in ("Transform", "RigidBody", "Collider")]).tofile(sb)
        out = subprocess.run(["node", os.path.join(HERE, "node_harness.mjs"), ref_root, sj, sb, ob, str(frames)],
                             capture_output=True, text=True, timeout=3600)
        if out.returncode != 0:
            raise RuntimeError("node harness failed:\n" + out.stderr[-2000:])
        info = json.loads(out.stdout.strip().splitlines()[-1])
        raw = np.fromfile(ob, dtype=np.uint8)
    res, off = {"seconds": info["seconds"], "node": info["node"]}, 0
    for short, name in names.items():
        lay, types = lays[name]
        size = bufs[name].size
        blk = raw[off:off + size]
        off += size
        for key in cols:
            pre, colname = key.split(".")
            if pre == short:
                dt = {1: np.uint8, 2: np.uint16, 4: np.float32}[types[colname]]
                res[key] = blk[lay[colname]:lay[colname] + N * types[colname]].view(dt).copy()
    nb = N * (1 + M) * 4
    res["neighborData"] = raw[off:off + nb].view(np.int32).copy(); off += nb
    res["distanceData"] = raw[off:off + nb].view(np.float32).copy(); off += nb
    res["collisionData"] = raw[off:off + (1 + 2 * P) * 4].view(np.int32).copy()
    return res
