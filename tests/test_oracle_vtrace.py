"""V-trace / actor-critic oracle (oracle/oracle_vtrace.c, oracle_nn.c).

The reference has no V-trace: parity against the reference is UNPINNED
here. The float64 recurrence is cross-checked two independent ways instead:
  - against the non-recursive closed form of Espeholt et al. eq. (1);
  - losses and gradients against torch float64 autograd of the same objective, written
    independently in Python (stop-gradient on vs and pg_adv as in the paper, section 4.2)."""
import numpy as np
import torch

import _util as U


def test_scan_matches_closed_form(oracle):
    rng = np.random.default_rng(0)
    for m, t, lam in ((3, 1, 1.0), (4, 7, 1.0), (5, 33, 0.9), (2, 100, 1.0)):
        log_rho = rng.standard_normal((m, t)) * 0.7
        disc = 0.99 * (rng.random((m, t)) > 0.1)
        rew = rng.standard_normal((m, t))
        val = rng.standard_normal((m, t))
        boot = rng.standard_normal(m)
        vs, adv = oracle.vtrace(log_rho, disc, rew, val, boot, rho_bar=1.3, c_bar=0.8, pg_rho_bar=1.1, lambda_=lam)
        cf = oracle.vtrace_closed_form(log_rho, disc, rew, val, boot, rho_bar=1.3, c_bar=0.8, lambda_=lam)
        np.testing.assert_allclose(vs, cf, rtol=1e-12, atol=1e-12)
        vs_next = np.concatenate([vs[:, 1:], boot[:, None]], axis=1)
        ref_adv = np.minimum(1.1, np.exp(log_rho)) * (rew + disc * vs_next - val)
        np.testing.assert_allclose(adv, ref_adv, rtol=1e-12, atol=1e-12)


def test_on_policy_reduces_to_nstep_return(oracle):
    # log_rho = 0, rho_bar = c_bar = 1 -> vs is the discounted n-step return (paper remark 1)
    rng = np.random.default_rng(1)
    m, t = 3, 20
    disc = np.full((m, t), 0.9)
    rew, val, boot = rng.standard_normal((m, t)), rng.standard_normal((m, t)), rng.standard_normal(m)
    vs, _ = oracle.vtrace(np.zeros((m, t)), disc, rew, val, boot)
    ret = np.empty((m, t))
    acc = boot.copy()
    for s in range(t - 1, -1, -1):
        acc = rew[:, s] + 0.9 * acc
        ret[:, s] = acc
    np.testing.assert_allclose(vs, ret, rtol=1e-12)


def _torch_losses(logits, value, mu, act, rew, disc, boot, cfg):
    logits = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    value = torch.tensor(value, dtype=torch.float64, requires_grad=True)
    mu, rew, disc, boot = (torch.tensor(np.asarray(v, np.float64)) for v in (mu, rew, disc, boot))
    act = torch.tensor(act, dtype=torch.int64)
    logp = torch.log_softmax(logits, -1)
    logmu = torch.log_softmax(mu, -1)
    lp_a = logp.gather(-1, act[..., None])[..., 0]
    log_rho = (lp_a - logmu.gather(-1, act[..., None])[..., 0]).detach()
    v = value.detach()
    is_w = log_rho.exp()
    rho = is_w.clamp(max=cfg["rho_bar"])
    c = cfg["lambda_"] * is_w.clamp(max=cfg["c_bar"])
    v_next = torch.cat([v[:, 1:], boot[:, None]], 1)
    delta = rho * (rew + disc * v_next - v)
    acc = torch.zeros_like(boot)
    vs = []
    for s in range(v.shape[1] - 1, -1, -1):
        acc = delta[:, s] + disc[:, s] * c[:, s] * acc
        vs.append(v[:, s] + acc)
    vs = torch.stack(vs[::-1], 1)
    vs_next = torch.cat([vs[:, 1:], boot[:, None]], 1)
    adv = is_w.clamp(max=cfg["pg_rho_bar"]) * (rew + disc * vs_next - v)
    pg = -(lp_a * adv).sum()
    bl = 0.5 * ((vs - value) ** 2).sum()
    ent = (logp.exp() * logp).sum()
    total = pg + cfg["baseline_cost"] * bl + cfg["entropy_cost"] * ent
    total.backward()
    return [float(x) for x in (total, pg, bl, ent)], logits.grad.numpy(), value.grad.numpy(), vs.numpy(), adv.numpy()


def test_losses_and_gradients_match_autograd(oracle):
    rng = np.random.default_rng(2)
    cfg = dict(rho_bar=1.0, c_bar=1.0, pg_rho_bar=1.0, lambda_=1.0, baseline_cost=0.5, entropy_cost=0.01)
    for m, t in ((2, 1), (3, 9), (4, 40)):
        _, mu, act, rew, disc, boot = U.vtrace_batch(10 + t, m, t, done_p=0.1)
        logits = rng.standard_normal((m, t, 16))
        value = rng.standard_normal((m, t))
        o = oracle.vtrace_losses(logits, value, mu, act, rew, disc, boot, **cfg)
        L, dl, dv, vs, adv = _torch_losses(logits, value, mu, act, rew, disc, boot, cfg)
        np.testing.assert_allclose(o["losses"], L, rtol=1e-11)
        np.testing.assert_allclose(o["vs"], vs, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(o["pg_adv"], adv, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(o["dlogits"], dl, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(o["dvalue"], dv, rtol=1e-10, atol=1e-12)


def test_actor_critic_grads_match_autograd(oracle):
    """Whole-model check of orc_ac_loss_grad against torch float64 autograd."""
    m, t = 3, 5
    p = U.ac_params(7)
    obs, mu, act, rew, disc, boot = U.vtrace_batch(8, m, t, done_p=0.2)
    ac = oracle.actor_critic(p)
    losses = ac.loss_grad(obs, mu, act, rew, disc, boot)
    g = ac.grads()
    off, num = oracle.ac_table()
    ws = [torch.tensor(p[o:o + n].astype(np.float64).reshape(s), requires_grad=True)
          for o, n, s in zip(off, num, U.AC_SHAPES)]
    h = torch.tensor(obs.reshape(m * t, 162).astype(np.float64))
    for l in range(5):
        h = torch.relu(h @ ws[2 * l].T + ws[2 * l + 1])
    head = h @ ws[10].T + ws[11]
    logits, value = head[:, :16].reshape(m, t, 16), head[:, 16].reshape(m, t)
    cfg = dict(rho_bar=1.0, c_bar=1.0, pg_rho_bar=1.0, lambda_=1.0, baseline_cost=0.5, entropy_cost=0.01)
    # reuse the functional form above through a differentiable path
    logp = torch.log_softmax(logits, -1)
    a = torch.tensor(act, dtype=torch.int64)
    lp_a = logp.gather(-1, a[..., None])[..., 0]
    L, _, _, vs, adv = _torch_losses(logits.detach().numpy(), value.detach().numpy(), mu, act, rew, disc, boot, cfg)
    vs, adv = torch.tensor(vs), torch.tensor(adv)
    total = -(lp_a * adv).sum() + 0.5 * 0.5 * ((vs - value) ** 2).sum() + 0.01 * (logp.exp() * logp).sum()
    total.backward()
    np.testing.assert_allclose(losses[0], float(total), rtol=1e-11)
    gt = np.concatenate([w.grad.numpy().ravel() for w in ws])
    assert U.rel_l2(g, gt) < 1e-11


def test_record_layout_roundtrip(oracle):
    from oracle import pyoracle as po
    obs, mu, act, rew, disc, boot = U.vtrace_batch(3, 4, 6)
    slots = po.pack_vtrace_slots(obs, mu, act, rew, disc, boot)
    assert slots.shape == (4, 6 * 1024)
    out = oracle.decode_vtrace(slots, 4, 6)
    for a, b in zip(out, (obs, mu, act, rew, disc, boot)):
        assert np.array_equal(a, b)
    z, x, tg = U.farmer_batch(4, 3, 10)
    slots = po.pack_farmer_slots(z, x, tg)
    z2, x2, t2 = oracle.decode_farmer(slots, 3, 10)
    assert np.array_equal(z, z2) and np.array_equal(x, x2) and np.array_equal(tg, t2)
