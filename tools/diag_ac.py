"""Diagnostic (GPU): per-tensor gradient error of the actor-critic step vs a float64 oracle rebuilt from the
CUDA learner's own parameters at every step."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _util as U
import freeimpala_b200 as fi
from oracle import pyoracle as po

m, t, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "simt"
o = po.Oracle()
L = fi.Learner(1, max(m, 2), t, m, model="mlp_actor_critic", gemm_mode=mode)
L.set_params(0, U.ac_params(11))
tab = L.tensor_table()
for s in range(steps):
    b = U.vtrace_batch(100 + s, m, t, done_p=0.03)
    p = L.get_params(0)
    O = o.actor_critic(p, lr=5e-4)
    want = O.loss_grad(*b)
    L.forward_backward(0, L.stage_batch(0, po.pack_vtrace_slots(*b)))
    got = L.last_losses(0)
    g, og = L.get_grads(0), O.grads()
    print(f"step {s}: loss rel {np.abs(got - want) / np.abs(want)} grads rel_l2 {U.rel_l2(g, og):.3e}")
    for i, (off, n, r, c) in enumerate(tab):
        d = g[off:off + n] - og[off:off + n]
        print(f"   tensor {i:2d} [{r}x{c}] rel_l2 {np.linalg.norm(d) / max(np.linalg.norm(og[off:off+n]), 1e-300):.3e} "
              f"max|d| {np.abs(d).max():.3e} |g|max {np.abs(og[off:off+n]).max():.3e} nbad(>1e-4 rel max) {(np.abs(d) > 1e-4 * np.abs(og[off:off+n]).max()).sum()}")
    L.apply_update(0)
L.close()
