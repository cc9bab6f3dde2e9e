"""Pin the oracle against the reference's REAL JavaScript workers when a Node runtime and the
reference tree are available (neither is in this image: the test then skips, and parity stays
"unpinned" as DESIGN.md §2 says)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline"))
import run_node_harness as node  # noqa: E402

from helpers import active_rows_equal, bits, make_oracle
from multithreadedgameengine_b200 import scenes
from oracle.oracle_c import OracleC


@pytest.mark.skipif(not node.available(), reason="no `node` on PATH or no reference tree: the reference JS cannot be executed here")
def test_oracle_matches_reference_javascript():
    cfg, cols = scenes.balls_readme(n_balls=600, seed=7, world=(1200.0, 700.0))
    frames = 5
    js = node.run(cfg, cols, frames)
    ora = make_oracle(OracleC, cfg, cols)
    for _ in range(frames):
        ora.step(1.0, 0)                      # order 0: the reference's own sequential sweep
    for k in ("T.x", "T.y", "RB.px", "RB.py", "RB.vx", "RB.vy", "RB.speed", "RB.collisionCount", "RB.ax", "RB.ay"):
        assert np.array_equal(bits(js[k]), bits(ora.col[k])), k
    cellOf, _, _ = ora.grid_csr()
    active_rows_equal(js["neighborData"], js["distanceData"], ora.neighborData, ora.distanceData,
                      cfg["entityCount"], cfg["spatial"]["maxNeighbors"], np.nonzero(cellOf >= 0)[0])
    n = int(ora.collisionData[0])
    assert int(js["collisionData"][0]) == n
    assert np.array_equal(js["collisionData"][:1 + 2 * n], ora.collisionData[:1 + 2 * n])
