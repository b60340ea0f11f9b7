"""Parity report (GPU): the UNTRIMMED, FREE-RUNNING parameter error of the CUDA learner step after N steps, next to the same
metric for the float64 oracle against the reference's fp32 libtorch step, so that "within 1e-5 of the reference" is a set of
measured numbers. No assertion here: tests/test_gpu_learner.py holds the thresholds.

    python tools/parity_report.py > gpurun_out/r2_parity.md

FarmerLstm cases: tests/golden/farmer_step.npz (the reference's own train_step, cmd/libtorch_bench/main.cpp:117-135).
  cuda~ref   rel L2 of the CUDA parameters vs the reference's after N steps (every 997th parameter is stored), all elements
  orc~ref    the float64 oracle (free-running) vs the reference, same elements
  cuda~orc   CUDA vs the float64 oracle, all parameters
  max        the largest element-wise difference cuda~ref relative to the largest |parameter|
Actor-critic V-trace cases have no reference counterpart: cuda~orc only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util as U  # noqa: E402
import freeimpala_b200 as fi  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def farmer(oracle, g, ci, mode):
    stride = int(g["stride"][0])
    b, t, steps, ps, ys, bs = (int(v) for v in g[f"c{ci}_meta"])
    loss, opt = (str(v) for v in g[f"c{ci}_kind"])
    lr = float(g[f"c{ci}_lr"][0])
    params = U.farmer_params(ps)
    L = fi.Learner(1, max(b, 2), t, b, 0, 0, "", "", 0, model="farmer_lstm", loss=loss, optimizer=opt, lr=lr, gemm_mode=mode)
    L.set_params(0, params)
    F = oracle.farmer(params, opt=opt, lr=lr, loss=loss)
    loss_err, grad_ref, grad_orc = 0.0, 0.0, 0.0
    for s in range(steps):
        z, x, tg = U.farmer_batch(bs + s, b, t)
        L.forward_backward(0, L.stage_batch(0, po.pack_farmer_slots(z, x, tg)))
        want = float(g[f"c{ci}_losses"][s])
        loss_err = max(loss_err, abs(L.last_losses(0)[0] - want) / abs(want))
        F.loss_grad(z, x, tg)
        gr = L.get_grads(0)
        grad_ref = max(grad_ref, U.rel_l2(gr[::stride], g[f"c{ci}_grads"][s]))
        grad_orc = max(grad_orc, U.rel_l2(gr, F.grads()))   # free-running from step 2 on
        L.apply_update(0)
        F.opt_step()
    p = L.get_params(0)
    ref = g[f"c{ci}_params"]
    row = (f"| farmer c{ci} {b}x{t} {loss}/{opt} x{steps} | {mode} | {loss_err:.2e} | {grad_ref:.2e} | {grad_orc:.2e} | "
           f"{U.rel_l2(p[::stride], ref):.2e} | {U.rel_l2(F.params()[::stride], ref):.2e} | {U.rel_l2(p, F.params()):.2e} | "
           f"{U.rel_max(p[::stride], ref):.2e} |")
    L.close()
    return row


def actor_critic(oracle, m, t, steps, mode):
    params = U.ac_params(11)
    L = fi.Learner(1, max(m, 2), t, m, 0, 0, "", "", 0, model="mlp_actor_critic", gemm_mode=mode)
    L.set_params(0, params)
    F = oracle.actor_critic(params, lr=5e-4)
    loss_err, grad_orc = 0.0, 0.0
    for s in range(steps):
        obs, mu, act, rew, disc, boot = U.vtrace_batch(100 + s, m, t, done_p=0.03)
        want = F.loss_grad(obs, mu, act, rew, disc, boot)
        L.forward_backward(0, L.stage_batch(0, po.pack_vtrace_slots(obs, mu, act, rew, disc, boot)))
        got = L.last_losses(0)
        loss_err = max(loss_err, abs(got[0] - want[0]) / abs(want[0]))
        grad_orc = max(grad_orc, U.rel_l2(L.get_grads(0), F.grads()))
        L.apply_update(0)
        F.opt_step()
    p = L.get_params(0)
    row = (f"| vtrace {m}x{t} x{steps} | {mode} | {loss_err:.2e} | - | {grad_orc:.2e} | - | - | {U.rel_l2(p, F.params()):.2e} | "
           f"{U.rel_max(p, F.params()):.2e} |")
    L.close()
    return row


def main():
    oracle = po.Oracle()
    g = np.load(os.path.join(U.GOLDEN, "farmer_step.npz"))
    print("| case | gemm_mode | loss rel err (max over steps) | grads cuda~ref | grads cuda~orc | params cuda~ref | params orc~ref | "
          "params cuda~orc | params max |")
    print("|---|---|---|---|---|---|---|---|---|")
    for ci in range(3, int(g["ncases"][0])):
        for mode in ("simt", "auto", "tcgen05_f16"):
            print(farmer(oracle, g, ci, mode), flush=True)
    for m, t, steps in ((4, 7, 3), (64, 100, 3), (9, 33, 2)):
        for mode in ("simt", "tcgen05", "tcgen05_f16"):
            print(actor_critic(oracle, m, t, steps, mode), flush=True)


if __name__ == "__main__":
    main()
