"""CPU coverage of the multi-GPU (N > 1) host logic: slab planning, halo depth, and the
neighbour exchange protocol over torch.distributed with the gloo backend (world_size 2 and 3).
The kernels themselves are covered on the GPU by tests/test_gpu_slabs.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multithreadedgameengine_b200 import binding as B, scenes
from multithreadedgameengine_b200.slabs import cell_rows, exchange_fixed, halo_rows, header_count, plan_slabs, replan_by_rank_speed, replan_from_times


def test_cell_rows_follow_the_reference_key():
    cfg = dict(worldHeight=100.0, spatial=dict(cellSize=30.0))
    y = np.array([0.0, 29.999998, 30.0, 59.9, 99.9, 150.0, -5.0, np.nan, np.inf], dtype=np.float32)
    rows, n = cell_rows(cfg, y)
    assert n == 4
    # 29.999998f * (1/30) truncates to 0 in binary64 (fp32 would say 1); clamps; NaN/Inf -> 0
    assert rows.tolist() == [0, 0, 1, 1, 3, 3, 0, 0, 0]


def test_halo_depth_rules():
    cfg, cols = scenes.scaled("config4", 2000)       # S = 2, cell 16, balls vr 16, Mouse vr 150
    assert halo_rows(cfg, cols) == 20                # 2 * ceil(150/16) beats (2+1) * 1
    cols["T.active"][0] = 0                          # without the Mouse
    assert halo_rows(cfg, cols) == 3
    cfg5, cols5 = scenes.scaled("config5", 2000)     # S = 4, cell 8, vr 4
    cols5["T.active"][0] = 0
    assert halo_rows(cfg5, cols5) == 5
    b, c = scenes.boids(300, 20)                     # predators (vr 250, cell 128) are dynamic: (1+1) * 2
    c["T.active"][0] = 0
    assert halo_rows(b, c) == 4


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_is_a_balanced_partition(world):
    cfg, cols = scenes.scaled("config4", 40000)
    blocks, H = plan_slabs(cfg, cols, world, balance="count")
    rows, n = cell_rows(cfg, cols["T.y"])
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    assert all(a[1] == b[0] and a[0] < a[1] for a, b in zip(blocks, blocks[1:]))
    counts = [int(((rows >= a) & (rows < b)).sum()) for a, b in blocks]
    assert sum(counts) == cfg["entityCount"]
    assert max(counts) < 1.35 * cfg["entityCount"] / world + 200    # clustered scene, whole rows


def test_cost_balanced_plan_equalises_estimated_work():
    from multithreadedgameengine_b200.slabs import row_costs
    cfg, cols = scenes.scaled("config4", 60000)
    cost, count = row_costs(cfg, cols)
    assert int(count.sum()) == cfg["entityCount"]
    blocks, _ = plan_slabs(cfg, cols, 4)
    per = [cost[a:b].sum() for a, b in blocks]
    assert max(per) < 1.25 * sum(per) / 4
    by_count, _ = plan_slabs(cfg, cols, 4, balance="count")
    per_c = [cost[a:b].sum() for a, b in by_count]
    assert max(per) <= max(per_c) + 1e-6      # never worse than count balancing on its own objective


def test_replan_with_row_weights_keeps_dense_rows_heavy():
    # two slabs of 10 rows; inside the slow first one all the work sits in rows 6-9
    w = np.ones(20); w[:6] = 0.001; w[6:10] = 10.0
    new = replan_from_times([(0, 10), (10, 20)], [3.0, 1.0], w)
    # total 4, half = 2: rows 6, 7, 8 carry 0.75 each -> the cut falls after row 8
    assert new == [(0, 9), (9, 20)]
    flat = replan_from_times([(0, 10), (10, 20)], [3.0, 1.0])
    assert flat[0][1] < new[0][1]                        # uniform spreading would overshoot


def test_replan_from_times_moves_cuts_toward_the_slow_slab():
    blocks = [(0, 100), (100, 200), (200, 300), (300, 400)]
    new = replan_from_times(blocks, [1.0, 1.0, 1.0, 3.0])       # the last slab is 3x slower per row
    assert new[0][0] == 0 and new[-1][1] == 400 and all(a[1] == b[0] for a, b in zip(new, new[1:]))
    assert new[-1][1] - new[-1][0] < 100 and new[0][1] - new[0][0] > 100
    # equal times are a fixed point (up to one row of rounding)
    same = replan_from_times(blocks, [2.0] * 4)
    assert all(abs(a[0] - b[0]) <= 1 and abs(a[1] - b[1]) <= 1 for a, b in zip(same, blocks))


def test_replan_by_rank_speed_gives_slow_links_fewer_rows():
    # four ranks, uniform rows; ranks 0-1 sit on links half as fast: the same 100 rows took them twice as long
    blocks = [(0, 100), (100, 200), (200, 300), (300, 400)]
    new = replan_by_rank_speed(blocks, [2.0, 2.0, 1.0, 1.0], np.ones(400))
    sizes = [b - a for a, b in new]
    assert new[0][0] == 0 and new[-1][1] == 400 and all(a[1] == b[0] for a, b in zip(new, new[1:]))
    assert all(abs(n - e) <= 2 for n, e in zip(sizes, [67, 67, 133, 133]))
    # with the speeds unchanged, the new cuts equalise the predicted times
    pred = [n / sp for n, sp in zip(sizes, [50, 50, 100, 100])]
    assert max(pred) / min(pred) < 1.05
    # equal times are a fixed point
    same = replan_by_rank_speed(blocks, [1.5] * 4, np.ones(400))
    assert all(abs(a[0] - b[0]) <= 1 and abs(a[1] - b[1]) <= 1 for a, b in zip(same, blocks))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _payload(frame, rank, side, quota, R):
    """Deterministic (count, bytes) a rank puts in its low (0) / high (1) send buffer."""
    g = np.random.default_rng(10007 * frame + 31 * rank + side)
    n = int(g.integers(0, quota + 1))
    return n, g.integers(0, 255, n * R, dtype=np.uint8)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    R = B.SLAB_RECORD_BYTES
    quota = 48
    mk = lambda: torch.zeros((quota + 1) * R, dtype=torch.uint8)
    ok = True
    for frame in range(3):
        bufs = {}
        for side in (0, 1):
            n, data = _payload(frame, rank, side, quota, R)
            b_ = mk()
            b_[:4] = torch.from_numpy(np.array([n], dtype=np.uint32).view(np.uint8))     # header: record count
            b_[R:R + n * R] = torch.from_numpy(data)
            bufs[side] = b_
        recv_low, recv_high = mk(), mk()
        exchange_fixed(torch, rank, world, bufs[0], bufs[1], recv_low, recv_high)
        if rank > 0:      # the low neighbour's HIGH buffer arrives on my LOW side
            n, data = _payload(frame, rank - 1, 1, quota, R)
            ok &= header_count(recv_low) == n and np.array_equal(recv_low[R:R + n * R].numpy(), data)
        else:
            ok &= header_count(recv_low) == 0
        if rank + 1 < world:
            n, data = _payload(frame, rank + 1, 0, quota, R)
            ok &= header_count(recv_high) == n and np.array_equal(recv_high[R:R + n * R].numpy(), data)
        else:
            ok &= header_count(recv_high) == 0
    open(os.path.join(out_dir, f"r{rank}"), "w").write("ok" if ok else "bad")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_protocol_gloo(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"r{r}").read() for r in range(world)] == ["ok"] * world
