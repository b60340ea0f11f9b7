"""TEST INFRASTRUCTURE ONLY: ctypes bindings to the checker libraries.

  oracle/_build/liboracle.so      plain-C float64 restatement (oracle.h)       -> class Oracle
  oracle/_ref/libfi_ref_nn.so     the reference's libtorch step, unmodified     -> class RefNN
  oracle/_ref/libfi_ref_host.so   the reference's SharedBuffer/ModelManager     -> class RefHost

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package freeimpala_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_NN_SO = os.path.join(HERE, "_ref", "libfi_ref_nn.so")
REF_HOST_SO = os.path.join(HERE, "_ref", "libfi_ref_host.so")
DROPIN_BIN = os.path.join(HERE, "_ref", "agent_dropin")

ELEMENT_SIZE = 1024
REC_WORDS = 256
Z_DIM, X_DIM, NUM_ACTIONS = 162, 484, 16
W_MU, W_ACTION, W_REWARD, W_DISCOUNT, W_AUX, W_X, X_PER_REC = 162, 178, 179, 180, 181, 192, 64
FARMER_PARAMS, AC_PARAMS = 1514497, 1142801
OPT = {"adam": 0, "sgd": 1, "adamw": 2}
LOSS = {"mse": 0, "mae": 1, "huber": 2}


def build(ref: bool = True) -> None:
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE, "-s"] + targets, check=True)
    # drop-in proof (the reference's unmodified agent.h against fi_host.hpp): needs the reference and the built product library
    lib = os.path.join(os.path.dirname(HERE), "freeimpala_b200", "_build", "libfreeimpala_b200.so")
    if ref and os.path.exists(lib):
        subprocess.run(["make", "-C", HERE, "-s", "dropin"], check=True)


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class VtraceCfg(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("rho_bar", "c_bar", "pg_rho_bar", "lambda_", "baseline_cost", "entropy_cost")]


DEFAULT_VTRACE = dict(rho_bar=1.0, c_bar=1.0, pg_rho_bar=1.0, lambda_=1.0, baseline_cost=0.5,
                      entropy_cost=0.01)


# ------------------------------------------------------------------------------------------
class Oracle:
    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.orc_ring_create.restype = C.c_void_p
        L.orc_ring_create.argtypes = [C.c_size_t, C.c_size_t]
        L.orc_ring_destroy.argtypes = [C.c_void_p]
        L.orc_ring_write.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_ring_read_batch.restype = C.c_long
        L.orc_ring_read_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_ring_set_draining.argtypes = [C.c_void_p]
        L.orc_ring_filled_count.restype = C.c_size_t
        L.orc_ring_filled_count.argtypes = [C.c_void_p]
        L.orc_ring_slot_bytes.restype = C.c_size_t
        L.orc_ring_slot_bytes.argtypes = [C.c_void_p]
        L.orc_decode_farmer.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t] + [C.c_void_p] * 3
        L.orc_decode_vtrace.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t] + [C.c_void_p] * 6
        L.orc_opt_update.argtypes = [C.c_int, C.c_double, C.c_int64, C.c_size_t] + [C.c_void_p] * 4
        L.orc_opt_update_f32.argtypes = [C.c_int, C.c_double, C.c_int64, C.c_size_t] + [C.c_void_p] * 4
        L.orc_farmer_create.restype = C.c_void_p
        L.orc_farmer_create.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.orc_farmer_destroy.argtypes = [C.c_void_p]
        L.orc_farmer_tensor_table.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_farmer_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_farmer_loss_grad.restype = C.c_double
        L.orc_farmer_loss_grad.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int] * 3
        L.orc_farmer_opt_step.argtypes = [C.c_void_p]
        L.orc_farmer_train_step.restype = C.c_double
        L.orc_farmer_train_step.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int] * 2
        for n in ("get_params", "get_grads", "set_grads"):
            getattr(L, "orc_farmer_" + n).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, "orc_ac_" + n).argtypes = [C.c_void_p, C.c_void_p]
        L.orc_vtrace.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_double] * 4 + [C.c_void_p] * 2
        L.orc_vtrace_closed_form.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_double] * 3 + [C.c_void_p]
        L.orc_vtrace_losses.argtypes = [C.c_int] * 3 + [C.c_void_p] * 7 + [C.POINTER(VtraceCfg)] + [C.c_void_p] * 5
        L.orc_ac_create.restype = C.c_void_p
        L.orc_ac_create.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(VtraceCfg)]
        L.orc_ac_destroy.argtypes = [C.c_void_p]
        L.orc_ac_tensor_table.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_ac_loss_grad.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int] * 2 + [C.c_void_p]
        L.orc_ac_loss_grad_masked.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int] * 2 + [C.c_void_p] * 2 + [C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.orc_ac_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_ac_opt_step.argtypes = [C.c_void_p]

    @property
    def num_threads(self) -> int:
        return self.lib.orc_num_threads()

    # ---- ring ----
    def ring(self, entry_size: int, capacity: int) -> "OracleRing":
        return OracleRing(self, entry_size, capacity)

    # ---- decode ----
    def decode_farmer(self, batch: np.ndarray, m: int, s: int):
        batch = np.ascontiguousarray(batch, dtype=np.uint8)
        z = np.empty((m, s, Z_DIM), np.float32)
        x = np.empty((m, X_DIM), np.float32)
        t = np.empty((m,), np.float32)
        self.lib.orc_decode_farmer(_p(batch), m, s, _p(z), _p(x), _p(t))
        return z, x, t

    def decode_vtrace(self, batch: np.ndarray, m: int, s: int):
        batch = np.ascontiguousarray(batch, dtype=np.uint8)
        obs = np.empty((m, s, Z_DIM), np.float32)
        mu = np.empty((m, s, NUM_ACTIONS), np.float32)
        act = np.empty((m, s), np.int32)
        rew = np.empty((m, s), np.float32)
        disc = np.empty((m, s), np.float32)
        boot = np.empty((m,), np.float32)
        self.lib.orc_decode_vtrace(_p(batch), m, s, _p(obs), _p(mu), _p(act), _p(rew), _p(disc), _p(boot))
        return obs, mu, act, rew, disc, boot

    # ---- optimiser ----
    def opt_update(self, kind, lr, step, p, g, m, v):
        self.lib.orc_opt_update(OPT[kind], lr, step, p.size, _p(p), _p(g), _p(m), _p(v))

    def opt_update_f32(self, kind, lr, step, p, g, m, v):
        self.lib.orc_opt_update_f32(OPT[kind], lr, step, p.size, _p(p), _p(g), _p(m), _p(v))

    # ---- vtrace ----
    def vtrace(self, log_rho, discount, reward, value, bootstrap, rho_bar=1.0, c_bar=1.0,
               pg_rho_bar=1.0, lambda_=1.0):
        m, t = log_rho.shape
        a = [_f64(v) for v in (log_rho, discount, reward, value, bootstrap)]
        vs = np.empty((m, t), np.float64)
        adv = np.empty((m, t), np.float64)
        self.lib.orc_vtrace(m, t, *[_p(v) for v in a], rho_bar, c_bar, pg_rho_bar, lambda_, _p(vs), _p(adv))
        return vs, adv

    def vtrace_closed_form(self, log_rho, discount, reward, value, bootstrap, rho_bar=1.0,
                           c_bar=1.0, lambda_=1.0):
        m, t = log_rho.shape
        a = [_f64(v) for v in (log_rho, discount, reward, value, bootstrap)]
        vs = np.empty((m, t), np.float64)
        self.lib.orc_vtrace_closed_form(m, t, *[_p(v) for v in a], rho_bar, c_bar, lambda_, _p(vs))
        return vs

    def vtrace_losses(self, logits, value, mu_logits, action, reward, discount, bootstrap, **cfg):
        m, t, a = logits.shape
        c = VtraceCfg(**{**DEFAULT_VTRACE, **cfg})
        logits, value = _f64(logits), _f64(value)
        mu_logits, reward, discount, bootstrap = map(_f32, (mu_logits, reward, discount, bootstrap))
        action = np.ascontiguousarray(action, dtype=np.int32)
        losses = np.empty(4, np.float64)
        dlogits = np.empty((m, t, a), np.float64)
        dvalue = np.empty((m, t), np.float64)
        vs = np.empty((m, t), np.float64)
        adv = np.empty((m, t), np.float64)
        self.lib.orc_vtrace_losses(m, t, a, _p(logits), _p(value), _p(mu_logits), _p(action),
                                   _p(reward), _p(discount), _p(bootstrap), C.byref(c), _p(losses),
                                   _p(dlogits), _p(dvalue), _p(vs), _p(adv))
        return dict(losses=losses, dlogits=dlogits, dvalue=dvalue, vs=vs, pg_adv=adv)

    # ---- models ----
    def farmer(self, params, opt="adam", lr=5e-4, loss="mse") -> "OracleFarmer":
        return OracleFarmer(self, params, opt, lr, loss)

    def actor_critic(self, params, opt="adam", lr=5e-4, **cfg) -> "OracleAC":
        return OracleAC(self, params, opt, lr, cfg)

    def farmer_table(self):
        off = np.empty(16, np.int64)
        num = np.empty(16, np.int64)
        self.lib.orc_farmer_tensor_table(_p(off), _p(num))
        return off, num

    def ac_table(self):
        off = np.empty(12, np.int64)
        num = np.empty(12, np.int64)
        self.lib.orc_ac_tensor_table(_p(off), _p(num))
        return off, num


class OracleRing:
    def __init__(self, o: Oracle, entry_size: int, capacity: int):
        self.o, self.h = o, o.lib.orc_ring_create(entry_size, capacity)
        self.slot_bytes = entry_size * ELEMENT_SIZE

    def __del__(self):
        if getattr(self, "h", None):
            self.o.lib.orc_ring_destroy(self.h)
            self.h = None

    def write(self, data: bytes | np.ndarray) -> int:
        buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        buf = np.ascontiguousarray(buf)
        return self.o.lib.orc_ring_write(self.h, _p(buf), buf.nbytes)

    def read_batch(self, m: int):
        out = np.empty((m, self.slot_bytes), np.uint8)
        rc = self.o.lib.orc_ring_read_batch(self.h, m, _p(out))
        return rc, (out if rc > 0 else None)

    def set_draining(self):
        self.o.lib.orc_ring_set_draining(self.h)

    def filled_count(self) -> int:
        return self.o.lib.orc_ring_filled_count(self.h)


class OracleFarmer:
    def __init__(self, o, params, opt, lr, loss):
        params = _f32(params)
        assert params.size == FARMER_PARAMS
        self.o, self.h = o, o.lib.orc_farmer_create(_p(params), OPT[opt], lr, LOSS[loss])

    def __del__(self):
        if getattr(self, "h", None):
            self.o.lib.orc_farmer_destroy(self.h)
            self.h = None

    def forward(self, z, x):
        z, x = _f32(z), _f32(x)
        b, t = z.shape[0], z.shape[1]
        y = np.empty(b, np.float64)
        self.o.lib.orc_farmer_forward(self.h, _p(z), _p(x), b, t, _p(y))
        return y

    def loss_grad(self, z, x, target, loss_denom=None):
        z, x, target = _f32(z), _f32(x), _f32(target)
        b, t = z.shape[0], z.shape[1]
        return self.o.lib.orc_farmer_loss_grad(self.h, _p(z), _p(x), _p(target), b, t, loss_denom or b)

    def opt_step(self):
        self.o.lib.orc_farmer_opt_step(self.h)

    def train_step(self, z, x, target):
        z, x, target = _f32(z), _f32(x), _f32(target)
        return self.o.lib.orc_farmer_train_step(self.h, _p(z), _p(x), _p(target), z.shape[0], z.shape[1])

    def params(self):
        out = np.empty(FARMER_PARAMS, np.float64)
        self.o.lib.orc_farmer_get_params(self.h, _p(out))
        return out

    def grads(self):
        out = np.empty(FARMER_PARAMS, np.float64)
        self.o.lib.orc_farmer_get_grads(self.h, _p(out))
        return out

    def set_grads(self, g):
        g = _f64(g)
        self.o.lib.orc_farmer_set_grads(self.h, _p(g))


class OracleAC:
    def __init__(self, o, params, opt, lr, cfg):
        params = _f32(params)
        assert params.size == AC_PARAMS
        c = VtraceCfg(**{**DEFAULT_VTRACE, **cfg})
        self.o, self.h = o, o.lib.orc_ac_create(_p(params), OPT[opt], lr, C.byref(c))

    def __del__(self):
        if getattr(self, "h", None):
            self.o.lib.orc_ac_destroy(self.h)
            self.h = None

    def forward(self, obs):
        obs = _f32(obs)
        rows = obs.size // Z_DIM
        logits = np.empty((rows, NUM_ACTIONS), np.float64)
        value = np.empty(rows, np.float64)
        self.o.lib.orc_ac_forward(self.h, _p(obs), rows, _p(logits), _p(value))
        return logits, value

    def loss_grad(self, obs, mu, act, rew, disc, boot):
        obs, mu, rew, disc, boot = map(_f32, (obs, mu, rew, disc, boot))
        act = np.ascontiguousarray(act, dtype=np.int32)
        m, t = act.shape
        losses = np.empty(4, np.float64)
        self.o.lib.orc_ac_loss_grad(self.h, _p(obs), _p(mu), _p(act), _p(rew), _p(disc), _p(boot), m, t, _p(losses))
        return losses

    def loss_grad_masked(self, obs, mu, act, rew, disc, boot, relu_mask):
        """loss_grad with the ReLU decisions pinned to `relu_mask` ([5, m*t*512] uint8). Returns
        (losses, number of overridden units, largest |z|/rms among them)."""
        obs, mu, rew, disc, boot = map(_f32, (obs, mu, rew, disc, boot))
        act = np.ascontiguousarray(act, dtype=np.int32)
        relu_mask = np.ascontiguousarray(relu_mask, dtype=np.uint8)
        m, t = act.shape
        assert relu_mask.size == 5 * m * t * 512
        losses = np.empty(4, np.float64)
        n_over, max_over = C.c_int64(), C.c_double()
        self.o.lib.orc_ac_loss_grad_masked(self.h, _p(obs), _p(mu), _p(act), _p(rew), _p(disc), _p(boot), m, t,
                                           _p(losses), _p(relu_mask), C.byref(n_over), C.byref(max_over))
        return losses, n_over.value, max_over.value

    def opt_step(self):
        self.o.lib.orc_ac_opt_step(self.h)

    def params(self):
        out = np.empty(AC_PARAMS, np.float64)
        self.o.lib.orc_ac_get_params(self.h, _p(out))
        return out

    def grads(self):
        out = np.empty(AC_PARAMS, np.float64)
        self.o.lib.orc_ac_get_grads(self.h, _p(out))
        return out

    def set_grads(self, g):
        g = _f64(g)
        self.o.lib.orc_ac_set_grads(self.h, _p(g))


# ------------------------------------------------------------------------------------------
def ref_available() -> bool:
    return os.path.exists(REF_NN_SO) and os.path.exists(REF_HOST_SO)


class RefNN:
    """The reference's own libtorch learner step (cmd/libtorch_bench/main.cpp), CPU."""

    def __init__(self, seed=1234, opt="adam", lr=5e-4, loss="mse", path: str = REF_NN_SO):
        import torch  # noqa: F401  (loads libtorch's global deps; a bare dlopen of the shim segfaults)
        L = self.lib = C.CDLL(path)
        L.ref_nn_create.restype = C.c_void_p
        L.ref_nn_create.argtypes = [C.c_uint64, C.c_char_p, C.c_double, C.c_char_p]
        L.ref_nn_destroy.argtypes = [C.c_void_p]
        L.ref_nn_param_count.restype = C.c_int64
        L.ref_nn_param_count.argtypes = [C.c_void_p]
        L.ref_nn_num_tensors.argtypes = [C.c_void_p]
        L.ref_nn_tensor_numel.restype = C.c_int64
        L.ref_nn_tensor_numel.argtypes = [C.c_void_p, C.c_int]
        for n in ("get_params", "set_params", "get_grads"):
            getattr(L, "ref_nn_" + n).argtypes = [C.c_void_p, C.c_void_p]
        L.ref_nn_make_batch.argtypes = [C.c_uint64, C.c_int, C.c_int] + [C.c_void_p] * 3
        L.ref_nn_loss.restype = C.c_double
        L.ref_nn_loss.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int] * 2
        L.ref_nn_train_step.restype = C.c_double
        L.ref_nn_train_step.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int] * 2
        L.ref_nn_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_nn_bench.restype = C.c_double
        L.ref_nn_bench.argtypes = [C.c_void_p] + [C.c_int] * 4
        L.ref_nn_set_num_threads.argtypes = [C.c_int]
        self.h = L.ref_nn_create(seed, opt.encode(), lr, loss.encode())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_nn_destroy(self.h)
            self.h = None

    @property
    def num_threads(self):
        return self.lib.ref_nn_num_threads()

    def param_count(self):
        return self.lib.ref_nn_param_count(self.h)

    def tensor_numels(self):
        return [self.lib.ref_nn_tensor_numel(self.h, i) for i in range(self.lib.ref_nn_num_tensors(self.h))]

    def params(self):
        out = np.empty(self.param_count(), np.float32)
        self.lib.ref_nn_get_params(self.h, _p(out))
        return out

    def set_params(self, p):
        p = _f32(p)
        self.lib.ref_nn_set_params(self.h, _p(p))

    def grads(self):
        out = np.empty(self.param_count(), np.float32)
        self.lib.ref_nn_get_grads(self.h, _p(out))
        return out

    def make_batch(self, seed, b, t):
        z = np.empty((b, t, Z_DIM), np.float32)
        x = np.empty((b, X_DIM), np.float32)
        tg = np.empty((b, 1), np.float32)
        self.lib.ref_nn_make_batch(seed, b, t, _p(z), _p(x), _p(tg))
        return z, x, tg

    def loss(self, z, x, target):
        z, x, target = _f32(z), _f32(x), _f32(target)
        return self.lib.ref_nn_loss(self.h, _p(z), _p(x), _p(target), z.shape[0], z.shape[1])

    def train_step(self, z, x, target):
        z, x, target = _f32(z), _f32(x), _f32(target)
        return self.lib.ref_nn_train_step(self.h, _p(z), _p(x), _p(target), z.shape[0], z.shape[1])

    def forward(self, z, x):
        z, x = _f32(z), _f32(x)
        y = np.empty(z.shape[0], np.float32)
        self.lib.ref_nn_forward(self.h, _p(z), _p(x), z.shape[0], z.shape[1], _p(y))
        return y

    def bench(self, b, t, warmups, runs):
        return self.lib.ref_nn_bench(self.h, b, t, warmups, runs)


class RefHost:
    """The reference's own SharedBuffer / ModelManager / Learner (header-only C++)."""

    def __init__(self, path: str = REF_HOST_SO):
        L = self.lib = C.CDLL(path)
        L.ref_ring_create.restype = C.c_void_p
        L.ref_ring_create.argtypes = [C.c_size_t, C.c_size_t]
        L.ref_ring_destroy.argtypes = [C.c_void_p]
        L.ref_ring_write.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_ring_try_write.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_ring_read_batch.restype = C.c_size_t
        L.ref_ring_read_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_ring_set_draining.argtypes = [C.c_void_p]
        L.ref_ring_filled_count.restype = C.c_size_t
        L.ref_ring_filled_count.argtypes = [C.c_void_p]
        L.ref_ring_bench_read_batch.restype = C.c_double
        L.ref_ring_bench_read_batch.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        L.ref_mm_create.restype = C.c_void_p
        L.ref_mm_create.argtypes = [C.c_size_t, C.c_size_t, C.c_char_p]
        L.ref_mm_destroy.argtypes = [C.c_void_p]
        L.ref_mm_latest_version.restype = C.c_uint64
        L.ref_mm_latest_version.argtypes = [C.c_void_p, C.c_size_t]
        L.ref_mm_publish.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_mm_get.restype = C.c_uint64
        L.ref_mm_get.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_mm_save.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.ref_mm_load.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_learner_run.argtypes = [C.c_size_t] * 7 + [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]

    def ring(self, entry_size, capacity) -> "RefRing":
        return RefRing(self, entry_size, capacity)

    def bench_read_batch(self, entry_size, capacity, batch, iters) -> float:
        return self.lib.ref_ring_bench_read_batch(entry_size, capacity, batch, iters)

    def learner_run(self, players, capacity, entry_size, batch, train_ms, writers, per_writer, ckpt_dir):
        sec, upd = C.c_double(), C.c_uint64()
        self.lib.ref_learner_run(players, capacity, entry_size, batch, train_ms, writers, per_writer,
                                 ckpt_dir.encode(), C.byref(sec), C.byref(upd))
        return sec.value, upd.value


class RefRing:
    def __init__(self, host: RefHost, entry_size, capacity):
        self.host, self.h = host, host.lib.ref_ring_create(entry_size, capacity)
        self.slot_bytes = entry_size * ELEMENT_SIZE

    def __del__(self):
        if getattr(self, "h", None):
            self.host.lib.ref_ring_destroy(self.h)
            self.h = None

    def write(self, data) -> int:
        buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
        return self.host.lib.ref_ring_write(self.h, _p(buf), buf.nbytes)

    def try_write(self, data) -> int:
        buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
        return self.host.lib.ref_ring_try_write(self.h, _p(buf), buf.nbytes)

    def read_batch(self, m):
        out = np.empty((m, self.slot_bytes), np.uint8)
        n = self.host.lib.ref_ring_read_batch(self.h, m, _p(out))
        return n, (out if n > 0 else None)

    def set_draining(self):
        self.host.lib.ref_ring_set_draining(self.h)

    def filled_count(self):
        return self.host.lib.ref_ring_filled_count(self.h)


# ------------------------------------------------------------------------------------------
def pack_farmer_slots(z, x, target, s=None) -> np.ndarray:
    """Pack (z[m,t,162], x[m,484], target[m]) into m trajectory slots of s records
    (record layout: DESIGN.md / oracle.h). Unused words are zero."""
    z, x = _f32(z), _f32(x)
    target = _f32(target).reshape(-1)
    m, t = z.shape[0], z.shape[1]
    s = s or t
    w = np.zeros((m, s, REC_WORDS), np.float32)
    w[:, :t, :Z_DIM] = z
    xp = np.zeros((m, 8 * X_PER_REC), np.float32)
    xp[:, :X_DIM] = x
    nrec = min(8, s)
    w[:, :nrec, W_X:W_X + X_PER_REC] = xp.reshape(m, 8, X_PER_REC)[:, :nrec]
    w[:, 0, W_AUX] = target
    return w.reshape(m, s * REC_WORDS).view(np.uint8)


def pack_vtrace_slots(obs, mu_logits, action, reward, discount, bootstrap) -> np.ndarray:
    obs = _f32(obs)
    m, t = obs.shape[0], obs.shape[1]
    w = np.zeros((m, t, REC_WORDS), np.float32)
    w[:, :, :Z_DIM] = obs
    w[:, :, W_MU:W_MU + NUM_ACTIONS] = _f32(mu_logits)
    w[:, :, W_ACTION] = np.ascontiguousarray(action, dtype=np.int32).view(np.float32)
    w[:, :, W_REWARD] = _f32(reward)
    w[:, :, W_DISCOUNT] = _f32(discount)
    aux = w[:, t - 1, W_AUX]
    aux[...] = _f32(bootstrap)
    return w.reshape(m, t * REC_WORDS).view(np.uint8)
