"""CPU, world_size 2 over gloo: the N>1 host logic -- slice assignment, max-over-ranks timing,
cost gather -- with the oracle standing in for the GPU on each rank's slice."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
import workloads as wl
from roboken_fmskf_robot_controller_b200 import _cabi, layout, sharding


def test_shard_range_partitions():
    for n, w in [(16, 2), (17, 4), (1 << 24, 8), (5, 8), (0, 2)]:
        parts = [sharding.shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        for a, b in zip(parts, parts[1:]):
            assert a[1] == b[0]
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def _rollout_cost(lo, hi, steps, seed):
    n = hi - lo
    inp = wl.plant_inputs(n, steps, seed=seed, first=lo)
    goal = np.zeros((n, 2), dtype=np.float32)
    st = np.zeros(layout.VS_WORDS * n, dtype=np.uint32)
    ro = ol.HostRollout(n, steps, _cabi.RK_SENSOR_PLANT, inp["cmd"], inp["seg_len"], inp["yaw"], inp["yaw_period"], goal=goal)
    ol.run_port(st, n, ro)
    return ro.cost.copy(), st


def _worker(rank, world, port, q, total=64):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, lr, w = sharding.init("gloo")
    assert (r, w) == (rank, world)
    lo, hi = sharding.shard_range(total, r, w)
    cost, _ = _rollout_cost(lo, hi, 300, seed=31)
    sharding.barrier()
    t = sharding.max_over_ranks(1.0 + rank)
    tot = sharding.sum_over_ranks(hi - lo)
    allc = sharding.gather_costs(torch.from_numpy(cost))
    if rank == 0:
        q.put((t, tot, allc.numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 65])  # 65: unequal slices (33 + 32) through the padded gather
def test_two_rank_slices_equal_single_run(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + (1 if total == 65 else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, total)) for r in range(2)]
    for p in procs:
        p.start()
    t, tot, allc = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert t == 2.0 and tot == total
    ref, _ = _rollout_cost(0, total, 300, seed=31)
    np.testing.assert_array_equal(allc, ref)  # slice results == single-process results, bit for bit
