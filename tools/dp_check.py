#!/usr/bin/env python
"""Data-parallel parity on real GPUs (run under torchrun, one rank per GPU):
every rank steps on its shard of the global batch; after fi_learner_step the all-reduced gradient arena must
equal the float64 oracle's full-batch gradient (1e-5), the all-reduced losses the full-batch losses, and the
parameters must be bit-identical on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import _util as U
import freeimpala_b200 as fi
from freeimpala_b200 import dp
from oracle import pyoracle as po

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for model, mode in (("mlp_actor_critic", "simt"), ("mlp_actor_critic", "auto"), ("farmer_lstm", "simt")):
    gm, t = 4 * world, 9
    lo, hi = dp.shard_range(gm, rank, world)
    L = fi.Learner(1, max(hi - lo, 2), t, hi - lo, model=model, device=local, gemm_mode=mode)
    dp.init_learner_dp(L, rank, world)
    o = po.Oracle()
    if model == "mlp_actor_critic":
        params = U.ac_params(4)
        full = U.vtrace_batch(21, gm, t)
        slots = po.pack_vtrace_slots(*[a[lo:hi] for a in full])
        O = o.actor_critic(params)
        want = O.loss_grad(*full)
    else:
        params = U.farmer_params(4)
        z, x, tg = U.farmer_batch(22, gm, t)
        slots = po.pack_farmer_slots(z[lo:hi], x[lo:hi], tg[lo:hi])
        O = o.farmer(params)
        want = np.array([O.loss_grad(z, x, tg), 0, 0, 0])
    L.set_params(0, params)
    L.trainModel(0, L.stage_batch(0, slots))
    got_l, got_g, got_p = L.last_losses(0), L.get_grads(0), L.get_params(0)
    e_g = U.rel_l2(got_g, O.grads())
    e_l = np.abs(got_l - want).max() / np.abs(want).max()
    O.opt_step()
    e_p = U.rel_l2(got_p, O.params())
    p_all = [torch.zeros(got_p.size, device="cuda") for _ in range(world)]
    dist.all_gather(p_all, torch.from_numpy(got_p).cuda())
    same = all(torch.equal(p_all[0], q) for q in p_all)
    good = e_g < 1e-5 and e_l < 1e-5 and same
    ok &= good
    if rank == 0:
        print(f"dp_check world={world} {model}/{mode}: grads rel_l2 {e_g:.2e} losses {e_l:.2e} params(after Adam) {e_p:.2e} "
              f"replicas bit-identical {same} -> {'OK' if good else 'FAIL'}", flush=True)
    L.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
