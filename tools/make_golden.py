#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref.

Runs only in the build container (needs /root/reference -> `make -C oracle ref`). The
fixtures are what pins the oracle (the reference holds no tests or
golden vectors of its own, so they are produced by running the reference itself here):

  farmer_step.npz   cmd/libtorch_bench/main.cpp train_step (:117-135) on seeded inputs:
                    forward y, loss per step, gradient + parameter samples after N steps,
                    for mse/adam (README shape), mae/sgd and huber/adamw.
  ring_trace.npz    SharedBuffer (data_structures.h:191-307) write/try_write/readBatch/
                    setDraining trace with every returned batch.
  model_ckpt.npz    ModelManager checkpoint bytes (data_structures.h:87-113,388-423).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po  # noqa: E402
import _util as U  # noqa: E402

SAMPLE_STRIDE = 997  # every 997th parameter/gradient is stored


def farmer_fixture():
    out = {}
    cases = [("mse", "adam", 5e-4, 4, 6, 3), ("mae", "sgd", 1e-2, 3, 5, 2), ("huber", "adamw", 5e-4, 5, 4, 2),
             ("mse", "adam", 5e-4, 64, 100, 2),  # the README shape (configs[0])
             # T >= 8 so that the 1024-byte record layout can carry x (8 records x 64 words): GPU cases
             ("mae", "sgd", 1e-2, 3, 9, 2), ("huber", "adamw", 5e-4, 5, 8, 3), ("mse", "adam", 5e-4, 9, 33, 3),
             # the shape bench.py's reference arm and `--workload farmer` run (BASELINE.json configs[3] batch x seq)
             ("mse", "adam", 5e-4, 1024, 100, 2)]
    path = os.path.join(U.GOLDEN, "farmer_step.npz")
    old = dict(np.load(path)) if os.path.exists(path) and "--regen" not in sys.argv else {}
    for ci, (loss, opt, lr, b, t, steps) in enumerate(cases):
        if f"c{ci}_meta" in old:   # fixtures already committed stay byte-identical: only new cases are generated
            out.update({k: v for k, v in old.items() if k.startswith(f"c{ci}_")})
            continue
        r = po.RefNN(seed=1, opt=opt, lr=lr, loss=loss)
        p0 = U.farmer_params(100 + ci)
        r.set_params(p0)
        z0, x0, _ = U.farmer_batch(200 + ci, b, t)
        y0 = r.forward(z0, x0)
        losses, grads = [], []
        for s in range(steps):
            z, x, tg = U.farmer_batch(300 + 10 * ci + s, b, t)
            losses.append(r.loss(z, x, tg))      # loss of current weights == the loss train_step computes
            r.train_step(z, x, tg)
            grads.append(r.grads()[::SAMPLE_STRIDE].copy())
        p = r.params()
        out[f"c{ci}_meta"] = np.array([b, t, steps, 100 + ci, 200 + ci, 300 + 10 * ci], np.int64)
        out[f"c{ci}_lr"] = np.array([lr])
        out[f"c{ci}_kind"] = np.array([loss, opt])
        out[f"c{ci}_y0"] = y0
        out[f"c{ci}_losses"] = np.array(losses, np.float64)
        out[f"c{ci}_grads"] = np.stack(grads)
        out[f"c{ci}_params"] = p[::SAMPLE_STRIDE].copy()
        out[f"c{ci}_param_sum"] = np.array([p.astype(np.float64).sum(), np.abs(p.astype(np.float64)).sum()])
    out["ncases"] = np.array([len(cases)])
    out["stride"] = np.array([SAMPLE_STRIDE])
    np.savez_compressed(path, **out)


def ring_fixture():
    h = po.RefHost()
    rng = np.random.default_rng(5)
    entry, cap = 1, 5
    ring = h.ring(entry, cap)
    slot = entry * 1024
    ops, outs = [], []
    count = 0
    for i in range(200):
        k = rng.integers(0, 10)
        if k < 5 and count < cap:       # write (sometimes short, sometimes oversize)
            n = int(rng.choice([slot, slot, slot // 2, 17, slot + 1]))
            data = rng.integers(0, 256, size=n, dtype=np.uint8)
            kind = 0 if k < 4 else 1
            ok = ring.write(data) if kind == 0 else ring.try_write(data)
            ops.append((kind, n, 0, ok))
            outs.append(data)
            count += int(ok and n <= slot)
        elif count > 0:
            m = int(rng.integers(1, count + 1))
            n, b = ring.read_batch(m)
            assert n == m
            ops.append((2, m, 0, n))
            outs.append(b.reshape(-1))
            count -= m
        assert ring.filled_count() == count
    ring.set_draining()
    n, b = ring.read_batch(cap)  # draining and count < M -> empty batch
    ops.append((3, cap, 0, int(n)))
    outs.append(np.zeros(0, np.uint8))
    if count:
        n, b = ring.read_batch(count)  # draining but enough entries -> still served
        ops.append((2, count, 0, int(n)))
        outs.append(b.reshape(-1))
    lens = np.array([len(o) for o in outs], np.int64)
    np.savez_compressed(os.path.join(U.GOLDEN, "ring_trace.npz"), ops=np.array(ops, np.int64),
                        blob=np.concatenate(outs), lens=lens, entry=np.array([entry]), cap=np.array([cap]))


def ckpt_fixture():
    h = po.RefHost()
    with tempfile.TemporaryDirectory() as d:
        mm = h.lib.ref_mm_create(2, 4096, d.encode())
        rng = np.random.default_rng(9)
        blobs = []
        for v in range(3):
            b = rng.integers(0, 256, size=4096, dtype=np.uint8)
            h.lib.ref_mm_publish(mm, 1, b.ctypes.data, b.nbytes)
            blobs.append(b)
        ver = h.lib.ref_mm_latest_version(mm, 1)
        h.lib.ref_mm_save(mm, 1, 7)
        files = {}
        for f in sorted(os.listdir(d)):
            files[f] = np.fromfile(os.path.join(d, f), dtype=np.uint8)
        h.lib.ref_mm_destroy(mm)
    np.savez_compressed(os.path.join(U.GOLDEN, "model_ckpt.npz"), version=np.array([ver], np.uint64),
                        last_blob=blobs[-1], names=np.array(list(files.keys())),
                        **{"file_" + str(i): v for i, v in enumerate(files.values())})


if __name__ == "__main__":
    po.build(ref=True)
    os.makedirs(U.GOLDEN, exist_ok=True)
    farmer_fixture()
    if "--regen" in sys.argv or not os.path.exists(os.path.join(U.GOLDEN, "ring_trace.npz")):
        ring_fixture()
    if "--regen" in sys.argv or not os.path.exists(os.path.join(U.GOLDEN, "model_ckpt.npz")):
        ckpt_fixture()
    for f in sorted(os.listdir(U.GOLDEN)):
        print(f, os.path.getsize(os.path.join(U.GOLDEN, f)))
