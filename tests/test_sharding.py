"""Host-side walker sharding and the energy / gradient reductions, on CPU with gloo (world_size 2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import philox as PH
from psiformer_torch_b200 import sharding


def test_shard_bounds_partition_all_walkers():
    for total in (1, 7, 4096, 16384, 32768 + 3):
        for ws in (1, 2, 3, 4, 8):
            spans = [sharding.WalkerShard(r, ws, total).bounds for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_philox_streams_do_not_depend_on_the_partition():
    full = PH.mh_normals(99, 0, 5, 64, 4)
    parts = [PH.mh_normals(99, lo, 5, hi - lo, 4) for lo, hi in
             (sharding.WalkerShard(r, 4, 64).bounds for r in range(4))]
    assert np.array_equal(full, np.concatenate(parts))
    assert not np.array_equal(PH.mh_normals(99, 0, 5, 8, 4), PH.mh_normals(99, 0, 6, 8, 4))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        e_all = torch.randn(10, generator=g, dtype=torch.float64) - 14.6
        lo, hi = sharding.current_shard(10).bounds
        e = e_all[lo:hi]
        acc = torch.stack([e.sum(), (e * e).sum(), torch.tensor(float(e.numel()), dtype=torch.float64)])
        sharding.allreduce_energy_stats(acc)
        mean, var = sharding.energy_mean_and_variance(acc)
        lin = torch.nn.Linear(3, 2)
        with torch.no_grad():
            lin.weight.fill_(1.0); lin.bias.fill_(0.0)
        lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
        lin.bias.grad = torch.full_like(lin.bias, float(10 * (rank + 1)))
        sharding.allreduce_mean_gradients(lin.parameters())
        other = torch.nn.Linear(3, 2)
        with torch.no_grad():
            other.weight.fill_(float(rank + 5)); other.bias.fill_(float(rank - 3))
        sharding.broadcast_parameters(other.parameters())          # replicas start from rank 0's parameters
        # a rank without valid samples still joins the statistics all-reduce with zeros: the "no samples anywhere"
        # decision is collective (train.py:131-135 under sharding) and nobody is left waiting
        empty = torch.zeros(3, dtype=torch.float64) if rank == 1 else torch.tensor([2.0, 4.0, 1.0], dtype=torch.float64)
        sharding.allreduce_energy_stats(empty)
        ret[rank] = (mean.item(), var.item(), e_all.mean().item(), e_all.var(unbiased=False).item(),
                     lin.weight.grad[0, 0].item(), lin.bias.grad[0].item(), sharding.current_shard(10).walker_id0,
                     other.weight[0, 0].item(), other.bias[0].item(), empty.tolist())
    finally:
        dist.destroy_process_group()


def test_energy_and_gradient_allreduce_world2():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    for r in range(2):
        mean, var, ref_mean, ref_var, gw, gb, w0, bw, bb, stats = ret[r]
        assert abs(mean - ref_mean) < 1e-12 and abs(var - ref_var) < 1e-10
        assert gw == 1.5 and gb == 15.0
        assert w0 == 5 * r
        assert bw == 5.0 and bb == -3.0
        assert stats == [2.0, 4.0, 1.0]


def _loop_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the shape of bench.py's timed loops: every iteration holds an all-reduce, the ranks' own clocks want different
        # iteration counts (3 and 5 here); the count that is run must be the agreed maximum on both
        want = 3 if rank == 1 else 5
        done = 0
        acc = torch.zeros(3, dtype=torch.float64)
        while sharding.agree_max(1.0 if done < want else 0.0) > 0.0:
            step = torch.tensor([1.0, float(rank), 2.0], dtype=torch.float64)
            sharding.allreduce_energy_stats(step)
            acc += step
            done += 1
        n = int(sharding.agree_max(7 + rank))                 # a loop bound derived from a local timing
        for _ in range(n):
            sharding.allreduce_energy_stats(torch.ones(3, dtype=torch.float64))
        ret[rank] = (done, n, acc.tolist())
    finally:
        dist.destroy_process_group()


def test_loop_bounds_with_collectives_are_agreed_world2():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_loop_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == ret[1] == (5, 8, [10.0, 5.0, 20.0])
    assert sharding.agree_max(3.5) == 3.5                     # single process: the value itself
