"""N>1 host logic on CPU (gloo, world_size 2): shard arithmetic, id transport, and the reduction
semantics the CUDA step relies on -- per-rank gradients of the shards, each scaled by the GLOBAL batch
where the loss is a mean, sum to the single-process gradient."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _util as U
from freeimpala_b200 import dp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_the_batch():
    for g in (1, 5, 64, 1024, 1027):
        for w in (1, 2, 3, 8):
            parts = [dp.shard_range(g, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == g
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(8, 2, 2)


def _worker(rank, world, port, out_dir):
    for p in (U.ROOT, os.path.join(U.ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import pyoracle as po
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # id transport: what rank 0 creates arrives bit-identical everywhere
        ids = dp.exchange_ids(3, rank, world, lambda n: bytes(range(128)) * n)
        assert ids == bytes(range(128)) * 3
        o = po.Oracle()
        # (1) V-trace actor-critic: sum losses -> plain sum of shard gradients
        m, t = 6, 9
        batch = U.vtrace_batch(5, m, t)
        lo, hi = dp.shard_range(m, rank, world)
        ac = o.actor_critic(U.ac_params(2))
        losses = ac.loss_grad(*[a[lo:hi] for a in batch])
        g = torch.from_numpy(ac.grads())
        l = torch.from_numpy(np.asarray(losses))
        dist.all_reduce(g)
        dist.all_reduce(l)
        # (2) FarmerLstm / MSE: the mean is over the GLOBAL batch (loss_denom), then a plain sum
        b, tt = 5, 8
        z, x, tg = U.farmer_batch(7, b, tt)
        flo, fhi = dp.shard_range(b, rank, world)
        fm = o.farmer(U.farmer_params(3))
        floss = fm.loss_grad(z[flo:fhi], x[flo:fhi], tg[flo:fhi], loss_denom=b)
        fg = torch.from_numpy(fm.grads())
        fl = torch.tensor([floss], dtype=torch.float64)
        dist.all_reduce(fg)
        dist.all_reduce(fl)
        if rank == 0:
            np.savez(os.path.join(out_dir, "dp.npz"), g=g.numpy(), l=l.numpy(), fg=fg.numpy(), fl=fl.numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_sum_to_the_full_batch_gradient(oracle, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "dp.npz")
    ac = oracle.actor_critic(U.ac_params(2))
    want_l = ac.loss_grad(*U.vtrace_batch(5, 6, 9))
    assert U.rel_l2(got["g"], ac.grads()) < 1e-12
    np.testing.assert_allclose(got["l"], want_l, rtol=1e-12)
    fm = oracle.farmer(U.farmer_params(3))
    z, x, tg = U.farmer_batch(7, 5, 8)
    want_fl = fm.loss_grad(z, x, tg)
    assert U.rel_l2(got["fg"], fm.grads()) < 1e-12
    assert abs(got["fl"][0] - want_fl) < 1e-12 * abs(want_fl)
