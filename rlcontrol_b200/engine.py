"""Host-side wrapper of the C-ABI: torch tensors in, torch tensors out, device memory and streams
from PyTorch, all arithmetic in librlc.so.  One :class:`Engine` per process/GPU; one
:class:`Critic` per critic network (online / target)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (ACT_PER_STATE, ACT_SHARED, ADAM_TF, ADAM_TORCH, LAYOUT_IN_OUT, LAYOUT_OUT_IN,
                   PREC_AUTO, PREC_BY_NAME, PREC_FP32, TIN, TMID, RlcCritic, RlcMlp, check)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(x, device) -> torch.Tensor:
    """numpy / torch, float64 or float32 -> contiguous fp32 on device (the reference casts at the
    feed: torch.FloatTensor(x) forwardkl_network.py:125-129; TF placeholders are float32)."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).to(device)


_BOUNDS = {}


def _bounds(x, A: int, dev) -> torch.Tensor:
    """Action bounds as an fp32 device vector [A]; device tensors pass through, host values are uploaded once and
    cached (they are constants of the environment, and an H2D copy per call would also break graph capture)."""
    if isinstance(x, torch.Tensor) and x.device == dev:
        return x.to(torch.float32).reshape(-1).contiguous()
    arr = np.broadcast_to(np.asarray(x, np.float64).reshape(-1) if np.ndim(x) else np.asarray(x, np.float64), (A,))
    key = (str(dev), A, arr.tobytes())
    t = _BOUNDS.get(key)
    if t is None:
        t = _BOUNDS[key] = torch.as_tensor(np.ascontiguousarray(arr), dtype=torch.float32).to(dev)
    return t


class Engine:
    """Owns the rlc_handle (workspace + packed tensor-core operands) of one CUDA device."""

    def __init__(self, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise _lib.RlcError("rlcontrol_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        with torch.cuda.device(self.device_index):
            check(self.lib.rlc_create(C.byref(h), self.device_index))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.rlc_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.rlc_launch_count(self.h))

    def umma_error(self) -> int:
        return int(self.lib.rlc_umma_last_error(self.h, _stream()))

    # ------------------------------------------------------------------ reductions (K3)
    def topk(self, q: torch.Tensor, k: int, actions: Optional[torch.Tensor] = None):
        """``row.argsort()[::-1][:k]`` per state (ActorExpert.py:177). Returns (idx[B,k] int64,
        q_sel[B,k], elites[B,k,A] or None)."""
        B, N = q.shape
        idx = torch.empty((B, k), dtype=torch.int64, device=q.device)
        qs = torch.empty((B, k), dtype=torch.float32, device=q.device)
        elites, A, mode = None, 0, ACT_SHARED
        if actions is not None:
            A = actions.shape[-1]
            mode = ACT_PER_STATE if actions.dim() == 3 else ACT_SHARED
            elites = torch.empty((B, k, A), dtype=torch.float32, device=q.device)
        check(self.lib.rlc_reduce_topk(self.h, _ptr(q), B, N, k, _ptr(idx), _ptr(qs), _ptr(actions),
                                       A, mode, _ptr(elites), _stream()))
        return idx, qs, elites

    def stats(self, q: torch.Tensor):
        """(argmax int64 [B], max [B], mean [B]) over the sample axis."""
        B, N = q.shape
        am = torch.empty((B,), dtype=torch.int64, device=q.device)
        mx = torch.empty((B,), dtype=torch.float32, device=q.device)
        mean = torch.empty((B,), dtype=torch.float32, device=q.device)
        check(self.lib.rlc_reduce_stats(self.h, _ptr(q), B, N, _ptr(am), _ptr(mx), _ptr(mean), _stream()))
        return am, mx, mean

    def mean_into(self, x: torch.Tensor, out: torch.Tensor):
        """out[0] = mean(x) over a flat fp32 vector (the ``.mean(-1)`` over states of the policy losses,
        forwardkl_network.py:194) -- rlc_reduce_stats on a single row."""
        check(self.lib.rlc_reduce_stats(self.h, _ptr(x), 1, int(x.numel()), None, None, _ptr(out), _stream()))
        return out

    def soft_value(self, q: torch.Tensor, action_dim: int):
        """SQL ``logsumexp - log N + A log 2`` (sql_network.py:76-84)."""
        B, N = q.shape
        v = torch.empty((B,), dtype=torch.float32, device=q.device)
        check(self.lib.rlc_reduce_lse(self.h, _ptr(q), B, N, int(action_dim), _ptr(v), _stream()))
        return v

    def fkl(self, q, w, logp, entropy_scale: float, b_total: Optional[int] = None,
            want_boltz: bool = True, want_grad: bool = True):
        """ForwardKL grid reduction (forwardkl_network.py:165-194).
        Returns (loss_b [B], boltz [B,N] | None, dlogp [B,N] | None)."""
        B, N = q.shape
        loss_b = torch.empty((B,), dtype=torch.float32, device=q.device)
        boltz = torch.empty_like(q) if want_boltz else None
        dlogp = torch.empty_like(q) if want_grad else None
        check(self.lib.rlc_reduce_fkl(self.h, _ptr(q), _ptr(w), _ptr(logp), B, N, float(entropy_scale),
                                      int(b_total or B), _ptr(loss_b), _ptr(boltz), _ptr(dlogp),
                                      _stream()))
        return loss_b, boltz, dlogp

    def rkl(self, q, v, w, logp, entropy_scale: float, hard: bool = False,
            b_total: Optional[int] = None, want_grad: bool = True):
        """ReverseKL grid reduction (reversekl_network.py:181-190 / 197-203)."""
        B, N = q.shape
        loss_b = torch.empty((B,), dtype=torch.float32, device=q.device)
        dlogp = torch.empty_like(q) if want_grad else None
        check(self.lib.rlc_reduce_rkl(self.h, _ptr(q), _ptr(v), _ptr(w), _ptr(logp), B, N,
                                      float(entropy_scale), int(bool(hard)), int(b_total or B),
                                      _ptr(loss_b), _ptr(dlogp), _stream()))
        return loss_b, dlogp

    def fkl_policy(self, q, w, grid, action_scale: float, mean, log_std, entropy_scale: float,
                   b_total: Optional[int] = None, want_grad: bool = True, want_logp: bool = False, out=None):
        """ForwardKL grid reduction with ``PolicyNetwork.get_logprob`` evaluated in place
        (forwardkl_network.py:165-194 + :324-351).  Returns (loss_b [B], dmean [B,A] | None,
        dlog_std [B,A] | None, logp [B,N] | None)."""
        B, N = q.shape
        A = grid.shape[-1]
        if out is not None:                 # preallocated (loss_b [B], dmean [B,A], dlog_std [B,A]) device tensors
            loss_b, dm, ds = out
        else:
            loss_b = torch.empty((B,), dtype=torch.float32, device=q.device)
            dm = torch.empty((B, A), dtype=torch.float32, device=q.device) if want_grad else None
            ds = torch.empty((B, A), dtype=torch.float32, device=q.device) if want_grad else None
        lp = torch.empty_like(q) if want_logp else None
        check(self.lib.rlc_reduce_fkl_policy(self.h, _ptr(q), _ptr(w), _ptr(grid), A, float(action_scale),
                                             _ptr(mean), _ptr(log_std), B, N, float(entropy_scale),
                                             int(b_total or B), _ptr(loss_b), _ptr(dm), _ptr(ds), _ptr(lp),
                                             _stream()))
        return loss_b, dm, ds, lp

    def rkl_policy(self, q, v, w, grid, action_scale: float, mean, log_std, entropy_scale: float,
                   hard: bool = False, b_total: Optional[int] = None, want_grad: bool = True,
                   want_logp: bool = False, out=None):
        """ReverseKL counterpart (reversekl_network.py:181-203 + :346-374)."""
        B, N = q.shape
        A = grid.shape[-1]
        if out is not None:
            loss_b, dm, ds = out
        else:
            loss_b = torch.empty((B,), dtype=torch.float32, device=q.device)
            dm = torch.empty((B, A), dtype=torch.float32, device=q.device) if want_grad else None
            ds = torch.empty((B, A), dtype=torch.float32, device=q.device) if want_grad else None
        lp = torch.empty_like(q) if want_logp else None
        check(self.lib.rlc_reduce_rkl_policy(self.h, _ptr(q), _ptr(v), _ptr(w), _ptr(grid), A,
                                             float(action_scale), _ptr(mean), _ptr(log_std), B, N,
                                             float(entropy_scale), int(bool(hard)), int(b_total or B),
                                             _ptr(loss_b), _ptr(dm), _ptr(ds), _ptr(lp), _stream()))
        return loss_b, dm, ds, lp

    # ------------------------------------------------------------------ Actor-Expert actor side (N1)
    def mixture_sample(self, alpha, mean, sigma, comp_u, normal, a_min, a_max, equal_modal: bool = False,
                       uni_u=None, want_comp: bool = False):
        """``sample_action`` of the AE networks with the draws supplied (ae_network.py:461-496).
        alpha [B,M] (None with equal_modal), mean/sigma [B,M,A], comp_u [B,N], normal [B,N,A];
        uni_u [B,n_uniform,A] optional.  Returns actions [B,N,A] (and comp [B,N] int32)."""
        dev = self.device
        mean, sigma, comp_u, normal = (_f32(x, dev) for x in (mean, sigma, comp_u, normal))
        alpha = None if alpha is None else _f32(alpha, dev)
        B, M, A = mean.shape
        N = comp_u.shape[1]
        amin = _bounds(a_min, A, dev)
        amax = _bounds(a_max, A, dev)
        uni = None if uni_u is None else _f32(uni_u, dev)
        n_uni = 0 if uni is None else int(uni.shape[1])
        acts = torch.empty((B, N, A), dtype=torch.float32, device=dev)
        comp = torch.empty((B, N), dtype=torch.int32, device=dev) if want_comp else None
        check(self.lib.rlc_mixture_sample(self.h, _ptr(alpha), _ptr(mean), _ptr(sigma), B, M, A, N,
                                          int(bool(equal_modal)), _ptr(comp_u), _ptr(normal), _ptr(amin),
                                          _ptr(amax), n_uni, _ptr(uni), _ptr(acts), _ptr(comp), _stream()))
        return (acts, comp) if want_comp else acts

    def mixture_nll(self, alpha, mean, sigma, actions, equal_modal: bool = False, b_total: Optional[int] = None):
        """``get_lossfunc`` on the elites (ae_network.py:262-278).  actions [B,k,A].
        Returns (loss [1], nll [B,k], dalpha [B,M], dmean [B,M,A], dsigma [B,M,A])."""
        dev = self.device
        mean, sigma, actions = (_f32(x, dev) for x in (mean, sigma, actions))
        alpha = None if alpha is None else _f32(alpha, dev)
        B, M, A = mean.shape
        k = actions.shape[1]
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        nll = torch.empty((B, k), dtype=torch.float32, device=dev)
        da = torch.zeros((B, M), dtype=torch.float32, device=dev)
        dm, ds = torch.empty_like(mean), torch.empty_like(sigma)
        check(self.lib.rlc_mixture_nll(self.h, _ptr(alpha), _ptr(mean), _ptr(sigma), _ptr(actions), B, M, A, k,
                                       int(bool(equal_modal)), int(b_total or B), _ptr(loss), _ptr(nll), _ptr(da),
                                       _ptr(dm), _ptr(ds), _stream()))
        return loss, nll, da, dm, ds

    # ------------------------------------------------------------------ FKL / RKL B-row pieces
    def policy_evaluate(self, head, eps, action_scale: float, log_std_min: float = -20.0,
                        log_std_max: float = 2.0, out=None):
        """``PolicyNetwork.evaluate`` (forwardkl_network.py:303-322) on the raw head [B,2A] with the
        normal draws ``eps`` [B,A] (None = mean action).  Returns dict(action, logp, mean, mu_raw,
        log_std, z); ``out`` may carry preallocated tensors under the same keys."""
        B, A2 = head.shape
        A = A2 // 2
        o = dict(out or {})
        for k, sh in (("action", (B, A)), ("logp", (B,)), ("mean", (B, A)), ("mu_raw", (B, A)),
                      ("log_std", (B, A)), ("z", (B, A))):
            if k not in o:
                o[k] = torch.empty(sh, dtype=torch.float32, device=head.device)
        check(self.lib.rlc_policy_evaluate(self.h, _ptr(head), _ptr(eps), B, A, float(action_scale),
                                           float(log_std_min), float(log_std_max), _ptr(o["action"]),
                                           _ptr(o["logp"]), _ptr(o["mean"]), _ptr(o["mu_raw"]),
                                           _ptr(o["log_std"]), _ptr(o["z"]), _stream()))
        return o

    def kl_targets(self, r, gamma, v_next, q_new, logp, v, entropy_scale: float, sac: bool,
                   b_total: Optional[int] = None, out=None):
        """TD / soft-value targets (forwardkl_network.py:137-150). Returns (y_q [B], dv [B], v_loss [1])."""
        B = r.shape[0]
        if out is not None:
            y, dv, vl = out
        else:
            y = torch.empty((B,), dtype=torch.float32, device=r.device)
            dv = torch.empty((B,), dtype=torch.float32, device=r.device)
            vl = torch.empty((1,), dtype=torch.float32, device=r.device)
        check(self.lib.rlc_kl_targets(self.h, _ptr(r), _ptr(gamma), _ptr(v_next), _ptr(q_new), _ptr(logp),
                                      _ptr(v), B, int(b_total or B), float(entropy_scale), int(bool(sac)),
                                      _ptr(y), _ptr(dv), _ptr(vl), _stream()))
        return y, dv, vl

    def policy_head_grad(self, head, mode: int, dmean=None, dlog_std=None, z=None, logp=None, q_new=None,
                         v=None, entropy_scale: float = 0.0, b_total: Optional[int] = None,
                         log_std_min: float = -20.0, log_std_max: float = 2.0, out=None, loss_out=None):
        """Gradient wrt the raw policy head [B,2A] (include/rlc.h rlc_policy_head_grad)."""
        B, A2 = head.shape
        dhead = out if out is not None else torch.empty_like(head)
        check(self.lib.rlc_policy_head_grad(self.h, _ptr(head), B, A2 // 2, float(log_std_min),
                                            float(log_std_max), int(mode), _ptr(dmean), _ptr(dlog_std),
                                            _ptr(z), _ptr(logp), _ptr(q_new), _ptr(v), float(entropy_scale),
                                            int(b_total or B), _ptr(dhead), _ptr(loss_out), _stream()))
        return dhead

    def gmm_refit(self, X: torch.Tensor, num_modal: int, resp0: Optional[torch.Tensor] = None,
                  tol: float = 1e-2, max_iter: int = 100):
        """Bounded diagonal GMM refit (utils/boundedvar_gaussian_mixture.py). X [B,k,A]."""
        B, k, A = X.shape
        w = torch.empty((B, num_modal), dtype=torch.float32, device=X.device)
        mu = torch.empty((B, num_modal, A), dtype=torch.float32, device=X.device)
        var = torch.empty((B, num_modal, A), dtype=torch.float32, device=X.device)
        nit = torch.empty((B,), dtype=torch.int32, device=X.device)
        check(self.lib.rlc_gmm_refit(self.h, _ptr(X), B, k, A, num_modal, _ptr(resp0), float(tol),
                                     int(max_iter), _ptr(w), _ptr(mu), _ptr(var), _ptr(nit), _stream()))
        return w, mu, var, nit

    # ------------------------------------------------------------------ optimiser pieces
    def adam_step(self, theta, grad, m, v, step: int, lr: float, variant: int = ADAM_TORCH,
                  beta1=0.9, beta2=0.999, eps=1e-8, target=None, tau: float = 0.0):
        check(self.lib.rlc_adam_step(self.h, _ptr(theta), _ptr(grad), _ptr(m), _ptr(v), theta.numel(),
                                     int(step), float(lr), float(beta1), float(beta2), float(eps),
                                     int(variant), _ptr(target), float(tau), _stream()))

    def adam_step_dev(self, theta, grad, m, v, state_dev, lr: float, variant: int = ADAM_TORCH,
                      beta1=0.9, beta2=0.999, eps=1e-8, target=None, tau: float = 0.0):
        """Graph-safe Adam step: the step count lives in ``state_dev`` (int32[4] on the device)."""
        check(self.lib.rlc_adam_step_dev(self.h, _ptr(theta), _ptr(grad), _ptr(m), _ptr(v), theta.numel(),
                                         _ptr(state_dev), float(lr), float(beta1), float(beta2), float(eps),
                                         int(variant), _ptr(target), float(tau), _stream()))

    def soft_update(self, target, online, tau: float):
        check(self.lib.rlc_soft_update(self.h, _ptr(target), _ptr(online), target.numel(), float(tau),
                                       _stream()))


class Critic:
    """One critic network on the device: dims + canonical ``theta`` (see include/rlc.h)."""

    def __init__(self, engine: Engine, topology: int, S: int, A: int, H1: int, H2: int,
                 state_min: Optional[Sequence[float]] = None,
                 state_max: Optional[Sequence[float]] = None):
        self.eng = engine
        self.topology, self.S, self.A, self.H1, self.H2 = int(topology), int(S), int(A), int(H1), int(H2)
        n = engine.lib.rlc_theta_numel(self.topology, S, A, H1, H2)
        if n <= 0:
            raise ValueError("invalid critic dimensions")
        self.theta = torch.zeros((n,), dtype=torch.float32, device=engine.device)
        off = (C.c_int64 * 6)()
        check(engine.lib.rlc_theta_offsets(self.topology, S, A, H1, H2, off))
        self.offsets = list(off)
        self.smin = None if state_min is None else _f32(np.asarray(state_min, np.float64).reshape(-1), engine.device)
        self.smax = None if state_max is None else _f32(np.asarray(state_max, np.float64).reshape(-1), engine.device)
        if (self.smin is None) != (self.smax is None):
            raise ValueError("state_min and state_max must be given together")
        if self.smin is not None and (self.smin.numel() != S or self.smax.numel() != S):
            raise ValueError("state bounds must have S entries")
        self._desc = RlcCritic()
        self._refresh()

    def _refresh(self):
        d = self._desc
        d.topology, d.S, d.A, d.H1, d.H2 = self.topology, self.S, self.A, self.H1, self.H2
        d.theta = self.theta.data_ptr()
        d.smin = None if self.smin is None else self.smin.data_ptr()
        d.smax = None if self.smax is None else self.smax.data_ptr()

    @property
    def in1(self):
        return self.S + self.A if self.topology == TIN else self.S

    @property
    def in2(self):
        return self.H1 if self.topology == TIN else self.H1 + self.A

    # ------------------------------------------------------------------ parameters
    def load(self, W1, b1, W2, b2, W3, b3, layout: int):
        """Load the six tensors in the reference's own layout (torch ``[out,in]`` or TF
        ``[in,out]``)."""
        dev = self.eng.device
        ts = [_f32(x, dev) for x in (W1, b1, W2, b2, W3, b3)]
        exp1 = (self.H1, self.in1) if layout == LAYOUT_OUT_IN else (self.in1, self.H1)
        exp2 = (self.H2, self.in2) if layout == LAYOUT_OUT_IN else (self.in2, self.H2)
        if tuple(ts[0].shape) != exp1 or tuple(ts[2].shape) != exp2 or ts[1].numel() != self.H1 \
                or ts[3].numel() != self.H2 or ts[4].numel() != self.H2 or ts[5].numel() != 1:
            raise ValueError(f"weight shapes do not match critic dims: {[tuple(t.shape) for t in ts]}")
        check(self.eng.lib.rlc_pack_theta(self.topology, self.S, self.A, self.H1, self.H2, layout,
                                          *[_ptr(t) for t in ts], _ptr(self.theta), _stream()))
        self.invalidate()
        return self

    def export(self, layout: int):
        dev = self.eng.device
        s1 = (self.H1, self.in1) if layout == LAYOUT_OUT_IN else (self.in1, self.H1)
        s2 = (self.H2, self.in2) if layout == LAYOUT_OUT_IN else (self.in2, self.H2)
        s3 = (1, self.H2) if layout == LAYOUT_OUT_IN else (self.H2, 1)
        outs = [torch.empty(s, dtype=torch.float32, device=dev) for s in
                (s1, (self.H1,), s2, (self.H2,), s3, (1,))]
        check(self.eng.lib.rlc_unpack_theta(self.topology, self.S, self.A, self.H1, self.H2, layout,
                                            _ptr(self.theta), *[_ptr(t) for t in outs], _stream()))
        return outs

    def invalidate(self):
        check(self.eng.lib.rlc_invalidate_pack(self.eng.h, _ptr(self.theta)))

    def copy_from(self, other: "Critic"):
        self.theta.copy_(other.theta)
        self.invalidate()

    # ------------------------------------------------------------------ evaluation (K1/K2)
    def eval(self, s, a, precision="auto") -> torch.Tensor:
        """Q for every (state, action) pair: s [B,S]; a [N,A] (shared grid) or [B,N,A].
        Returns q [B,N] fp32 on the device."""
        dev = self.eng.device
        s = _f32(s, dev)
        a = _f32(a, dev)
        if s.dim() != 2 or s.shape[1] != self.S:
            raise ValueError(f"states must be [B,{self.S}], got {tuple(s.shape)}")
        B = s.shape[0]
        if a.dim() == 2:
            mode, N = ACT_SHARED, a.shape[0]
        elif a.dim() == 3 and a.shape[0] == B:
            mode, N = ACT_PER_STATE, a.shape[1]
        else:
            raise ValueError(f"actions must be [N,A] or [B,N,A], got {tuple(a.shape)}")
        if a.shape[-1] != self.A:
            raise ValueError(f"action dim {a.shape[-1]} != {self.A}")
        prec = PREC_BY_NAME[precision] if isinstance(precision, str) else int(precision)
        q = torch.empty((B, N), dtype=torch.float32, device=dev)
        if B * N == 0:
            return q
        check(self.eng.lib.rlc_critic_eval(self.eng.h, C.byref(self._desc), _ptr(s), B, _ptr(a), N,
                                           mode, prec, _ptr(q), _stream()))
        return q

    def eval_into(self, s: torch.Tensor, a: torch.Tensor, q_out: torch.Tensor, precision="auto") -> torch.Tensor:
        """Allocation-free variant of :meth:`eval` for steady-state loops: ``s``, ``a`` and
        ``q_out`` [B,N] are contiguous fp32 tensors already on this engine's device."""
        for x in (s, a, q_out):
            if x.device != self.eng.device or x.dtype != torch.float32 or not x.is_contiguous():
                raise ValueError("eval_into needs contiguous fp32 tensors on the engine's device")
        B = s.shape[0]
        mode, N = (ACT_SHARED, a.shape[0]) if a.dim() == 2 else (ACT_PER_STATE, a.shape[1])
        if s.shape[1] != self.S or a.shape[-1] != self.A or tuple(q_out.shape) != (B, N) or \
                (a.dim() == 3 and a.shape[0] != B):
            raise ValueError("eval_into: shape mismatch")
        prec = PREC_BY_NAME[precision] if isinstance(precision, str) else int(precision)
        if B * N:
            check(self.eng.lib.rlc_critic_eval(self.eng.h, C.byref(self._desc), _ptr(s), B, _ptr(a), N,
                                               mode, prec, _ptr(q_out), _stream()))
        return q_out

    def eval_reduce_policy(self, s, grid, w, action_scale: float, mean, log_std, entropy_scale: float, kind: str = "fkl",
                           v=None, hard: bool = False, b_total: Optional[int] = None, precision="auto",
                           want_q: bool = False, out=None, fuse: bool = False):
        """The whole sampled-action step in one call (``rlc_critic_eval_reduce_policy``): Q on the shared grid ``[N,A]`` and
        its per-state ForwardKL (``kind="fkl"``, forwardkl_network.py:160-194) or ReverseKL (``"rkl"``, needs ``v`` [B])
        policy reduction with ``get_logprob`` in place.  Returns (loss_b [B], dmean [B,A], dlog_std [B,A], q [B,N] | None).
        ``fuse=True`` with the split tensor precisions and B >= 8 x the SM count runs the reduction inside the evaluation
        kernel (``q`` is then never written unless ``want_q``); the default composes the two kernels, which is ~30 us
        faster at cfg4."""
        dev = self.eng.device
        s, grid, w = _f32(s, dev), _f32(grid, dev), _f32(w, dev).reshape(-1)
        mean, log_std = _f32(mean, dev), _f32(log_std, dev)
        B, N, A = s.shape[0], grid.shape[0], grid.shape[1]
        if s.shape[1] != self.S or A != self.A or tuple(mean.shape) != (B, A) or tuple(log_std.shape) != (B, A) or w.numel() != N:
            raise ValueError("eval_reduce_policy: shape mismatch")
        if kind not in ("fkl", "rkl") or (kind == "rkl" and v is None):
            raise ValueError("kind must be 'fkl' or 'rkl' (with v)")
        v = None if v is None else _f32(v, dev).reshape(-1)
        prec = PREC_BY_NAME[precision] if isinstance(precision, str) else int(precision)
        if out is not None:
            loss_b, dm, ds = out
        else:
            loss_b = torch.empty((B,), dtype=torch.float32, device=dev)
            dm = torch.empty((B, A), dtype=torch.float32, device=dev)
            ds = torch.empty((B, A), dtype=torch.float32, device=dev)
        q = torch.empty((B, N), dtype=torch.float32, device=dev) if want_q is True else (want_q if isinstance(want_q, torch.Tensor) else None)
        if B:
            check(self.eng.lib.rlc_critic_eval_reduce_policy(
                self.eng.h, C.byref(self._desc), _ptr(s), B, _ptr(grid), N, _ptr(w), float(action_scale), _ptr(mean),
                _ptr(log_std), _ptr(v), float(entropy_scale), 0 if kind == "fkl" else 1, int(bool(hard)), int(b_total or B),
                prec, int(bool(fuse)), _ptr(q), _ptr(loss_b), _ptr(dm), _ptr(ds), _stream()))
        return loss_b, dm, ds, q

    def tensor_arithmetic(self, shared_actions: bool, precision="fp16") -> str:
        """Which stated arithmetic the tensor path uses for this critic ("ss" | "folded" | "grid" | "grid3" |
        "unsupported"; include/rlc.h rlc_umma_mode_prec) -- the oracle restatement to compare against."""
        prec = PREC_BY_NAME[precision] if isinstance(precision, str) else int(precision)
        m = int(self.eng.lib.rlc_umma_mode_prec(C.byref(self._desc), ACT_SHARED if shared_actions else ACT_PER_STATE,
                                                prec))
        return {0: "ss", 1: "folded", 3: "grid", 4: "grid3", 5: "grid3c8"}.get(m, "unsupported")

    def eval_grad(self, s, a):
        """T-mid only: (q [B,N], dq/da [B,N,A]) without materialising the stack."""
        dev = self.eng.device
        s, a = _f32(s, dev), _f32(a, dev)
        B = s.shape[0]
        mode, N = (ACT_SHARED, a.shape[0]) if a.dim() == 2 else (ACT_PER_STATE, a.shape[1])
        q = torch.empty((B, N), dtype=torch.float32, device=dev)
        g = torch.empty((B, N, self.A), dtype=torch.float32, device=dev)
        check(self.eng.lib.rlc_tmid_eval_grad(self.eng.h, C.byref(self._desc), _ptr(s), B, _ptr(a), N,
                                              mode, _ptr(q), _ptr(g), _stream()))
        return q, g

    def grad_action(self, s_rows, a_rows):
        """``tf.gradients(q, action)`` on R stacked rows. Returns (dq/da [R,A], q [R])."""
        dev = self.eng.device
        s, a = _f32(s_rows, dev), _f32(a_rows, dev)
        R = s.shape[0]
        g = torch.empty((R, self.A), dtype=torch.float32, device=dev)
        q = torch.empty((R,), dtype=torch.float32, device=dev)
        check(self.eng.lib.rlc_critic_grad_action(self.eng.h, C.byref(self._desc), _ptr(s), _ptr(a), R,
                                                  _ptr(g), _ptr(q), _stream()))
        return g, q

    def grads(self, s, a, y, b_total: Optional[int] = None):
        """Gradient of mean((y-Q)^2) wrt theta. Returns (grad [numel], loss [1], q [B])."""
        dev = self.eng.device
        s, a, y = _f32(s, dev), _f32(a, dev), _f32(y, dev).reshape(-1)
        B = s.shape[0]
        grad = torch.empty_like(self.theta)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        q = torch.empty((B,), dtype=torch.float32, device=dev)
        check(self.eng.lib.rlc_critic_grads(self.eng.h, C.byref(self._desc), _ptr(s), _ptr(a), _ptr(y), B,
                                            int(b_total or B), _ptr(grad), _ptr(loss), _ptr(q), _stream()))
        return grad, loss, q

    def grads_into(self, s, a, y, grad_out, loss_out, q_out=None, b_total: Optional[int] = None):
        """Allocation-free :meth:`grads` (device fp32 tensors in, preallocated outputs)."""
        B = s.shape[0]
        check(self.eng.lib.rlc_critic_grads(self.eng.h, C.byref(self._desc), _ptr(s), _ptr(a), _ptr(y), B,
                                            int(b_total or B), _ptr(grad_out), _ptr(loss_out), _ptr(q_out),
                                            _stream()))
        return grad_out

    def ae_expert_step(self, s, k: int, alpha, mean, sigma, comp_u, normal, a_min, a_max,
                       equal_modal: bool = False, uni_u=None, want_actions: bool = False, want_q: bool = False):
        """The expert step of ``ActorExpert.update_network`` (ActorExpert.py:162-181) in one launch: sample N
        actions per state from the actor's mixture (draws supplied), Q through this (T-mid) critic,
        per-state top-k, elite gather.  Returns dict(idx [B,k] int64, q_sel [B,k], elites [B,k,A]
        [, actions [B,N,A]] [, q [B,N]])."""
        dev = self.eng.device
        s, mean, sigma, comp_u, normal = (_f32(x, dev) for x in (s, mean, sigma, comp_u, normal))
        alpha = None if alpha is None else _f32(alpha, dev)
        B, M, A = mean.shape
        N = comp_u.shape[1]
        amin = _bounds(a_min, A, dev)
        amax = _bounds(a_max, A, dev)
        uni = None if uni_u is None else _f32(uni_u, dev)
        n_uni = 0 if uni is None else int(uni.shape[1])
        out = dict(idx=torch.empty((B, k), dtype=torch.int64, device=dev),
                   q_sel=torch.empty((B, k), dtype=torch.float32, device=dev),
                   elites=torch.empty((B, k, A), dtype=torch.float32, device=dev))
        if want_actions:
            out["actions"] = torch.empty((B, N, A), dtype=torch.float32, device=dev)
        if want_q:
            out["q"] = torch.empty((B, N), dtype=torch.float32, device=dev)
        check(self.eng.lib.rlc_ae_expert_step(self.eng.h, C.byref(self._desc), _ptr(s), B, N, int(k), _ptr(alpha),
                                              _ptr(mean), _ptr(sigma), M, int(bool(equal_modal)), _ptr(comp_u),
                                              _ptr(normal), _ptr(amin), _ptr(amax), n_uni, _ptr(uni),
                                              _ptr(out.get("actions")), _ptr(out.get("q")), _ptr(out["idx"]),
                                              _ptr(out["q_sel"]), _ptr(out["elites"]), _stream()))
        return out

    def svgd_action_grads(self, s, fixed, updated, h_min: float = 1e-3, eps: float = 1e-6, want_aux: bool = False):
        """Soft-Q-learning SVGD direction (sql_network.py:96-117, utils/sql_kernel.py): fixed [B,Kf,A],
        updated [B,Ku,A] -> action_gradients [B,Ku,A] (and dict(q_fixed, dqda, kappa, h) with want_aux)."""
        dev = self.eng.device
        s, fixed, updated = _f32(s, dev), _f32(fixed, dev), _f32(updated, dev)
        B, Kf, A = fixed.shape
        Ku = updated.shape[1]
        dq = torch.empty((B, Kf, A), dtype=torch.float32, device=dev)
        g = torch.empty((B, Ku, A), dtype=torch.float32, device=dev)
        aux = dict(dqda=dq)
        if want_aux:
            aux.update(q_fixed=torch.empty((B, Kf), dtype=torch.float32, device=dev),
                       kappa=torch.empty((B, Kf, Ku), dtype=torch.float32, device=dev),
                       h=torch.empty((B,), dtype=torch.float32, device=dev))
        check(self.eng.lib.rlc_svgd_action_grads(self.eng.h, C.byref(self._desc), _ptr(s), B, _ptr(fixed), Kf,
                                                 _ptr(updated), Ku, float(h_min), float(eps), _ptr(dq), _ptr(g),
                                                 _ptr(aux.get("q_fixed")), _ptr(aux.get("kappa")), _ptr(aux.get("h")),
                                                 _stream()))
        return (g, aux) if want_aux else g

    # ------------------------------------------------------------------ CEM (K4)
    def cem(self, s, u0, noise, comp_u, top_m: int, num_modal: int, a_min, a_max,
            want_idx: bool = False):
        """QT-Opt CEM (qt_opt_network.py:132-175) with supplied draws. u0 [B,N,A];
        noise [iters-1,B,N,A] or None; comp_u [iters-1,B,N] or None."""
        dev = self.eng.device
        s, u0 = _f32(s, dev), _f32(u0, dev)
        B, N, A = u0.shape
        iters = 1 + (0 if noise is None else noise.shape[0])
        noise = None if noise is None else _f32(noise, dev)
        comp_u = None if comp_u is None else _f32(comp_u, dev)
        amin = _bounds(a_min, A, dev)
        amax = _bounds(a_max, A, dev)
        w = torch.empty((B, num_modal), dtype=torch.float32, device=dev)
        mu = torch.empty((B, num_modal, A), dtype=torch.float32, device=dev)
        var = torch.empty((B, num_modal, A), dtype=torch.float32, device=dev)
        best = torch.empty((B, A), dtype=torch.float32, device=dev)
        idx = torch.empty((iters, B, top_m), dtype=torch.int64, device=dev) if want_idx else None
        check(self.eng.lib.rlc_cem(self.eng.h, C.byref(self._desc), _ptr(s), B, N, iters, int(top_m),
                                   int(num_modal), _ptr(u0), _ptr(noise), _ptr(comp_u), _ptr(amin),
                                   _ptr(amax), _ptr(w), _ptr(mu), _ptr(var), _ptr(best), _ptr(idx),
                                   _stream()))
        return w, mu, var, best, idx


class Mlp:
    """Generic B-row MLP ``x -> relu(FC) -> relu(FC) -> FC(O)`` on the device (include/rlc.h rlc_mlp):
    the reference's ``ValueNetwork`` (O=1) and ``PolicyNetwork.forward`` (O=2A, [mean | log_std])."""

    def __init__(self, engine: Engine, inp: int, H1: int, H2: int, O: int):
        self.eng = engine
        self.inp, self.H1, self.H2, self.O = int(inp), int(H1), int(H2), int(O)
        n = engine.lib.rlc_mlp_numel(self.inp, self.H1, self.H2, self.O)
        if n <= 0:
            raise ValueError("invalid MLP dimensions")
        self.theta = torch.zeros((n,), dtype=torch.float32, device=engine.device)
        off = (C.c_int64 * 6)()
        check(engine.lib.rlc_mlp_offsets(self.inp, self.H1, self.H2, self.O, off))
        self.offsets = list(off)
        self._desc = RlcMlp()
        d = self._desc
        d.inp, d.H1, d.H2, d.O, d.theta = self.inp, self.H1, self.H2, self.O, self.theta.data_ptr()

    def _views(self):
        o, t = self.offsets, self.theta
        return (t[o[0]:o[1]].view(self.inp, self.H1), t[o[1]:o[2]], t[o[2]:o[3]].view(self.H1, self.H2),
                t[o[3]:o[4]], t[o[4]:o[5]].view(self.H2, self.O), t[o[5]:])

    def load_torch(self, W1, b1, W2, b2, W3, b3):
        """torch ``nn.Linear`` tensors (weight [out,in]); several output heads are passed as lists and
        laid side by side in the O columns (``[mean_linear, log_std_linear]``)."""
        dev = self.eng.device
        cat = lambda x, dim: torch.cat([_f32(t, dev) for t in x], dim) if isinstance(x, (list, tuple)) else _f32(x, dev)
        W3, b3 = cat(W3, 0), cat(b3, 0)
        v = self._views()
        v[0].copy_(_f32(W1, dev).t()); v[1].copy_(_f32(b1, dev).reshape(-1))
        v[2].copy_(_f32(W2, dev).t()); v[3].copy_(_f32(b2, dev).reshape(-1))
        v[4].copy_(W3.reshape(self.O, self.H2).t()); v[5].copy_(b3.reshape(-1))
        return self

    def export_torch(self):
        """[W1 [H1,in], b1, W2 [H2,H1], b2, W3 [O,H2], b3 [O]] in torch layout (device tensors)."""
        v = self._views()
        return [v[0].t().contiguous(), v[1].clone(), v[2].t().contiguous(), v[3].clone(),
                v[4].t().contiguous(), v[5].clone()]

    def copy_from(self, other: "Mlp"):
        self.theta.copy_(other.theta)

    def act_buffer(self, B: int) -> torch.Tensor:
        n = self.eng.lib.rlc_mlp_act_numel(self.H1, self.H2, int(B))
        return torch.empty((max(int(n), 1),), dtype=torch.float32, device=self.eng.device)

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, act: Optional[torch.Tensor] = None):
        B = x.shape[0]
        if x.dim() != 2 or x.shape[1] != self.inp:
            raise ValueError(f"input must be [B,{self.inp}], got {tuple(x.shape)}")
        if out is None:
            out = torch.empty((B, self.O), dtype=torch.float32, device=x.device)
        check(self.eng.lib.rlc_mlp_forward(self.eng.h, C.byref(self._desc), _ptr(x), B, _ptr(out), _ptr(act),
                                           _stream()))
        return out

    def grads(self, x: torch.Tensor, dout: torch.Tensor, act: Optional[torch.Tensor] = None,
              grad_out: Optional[torch.Tensor] = None, want_dx: bool = False):
        B = x.shape[0]
        g = grad_out if grad_out is not None else torch.empty_like(self.theta)
        dx = torch.empty_like(x) if want_dx else None
        check(self.eng.lib.rlc_mlp_grads(self.eng.h, C.byref(self._desc), _ptr(x), _ptr(act), _ptr(dout), B,
                                         _ptr(g), _ptr(dx), _stream()))
        return (g, dx) if want_dx else g


class CriticOptimizer:
    """Adam state for one critic + the optional data-parallel gradient all-reduce
    (critic regression step, rows a15/a16)."""

    def __init__(self, critic: Critic, lr: float, variant: int = ADAM_TORCH, target: Optional[Critic] = None,
                 tau: float = 0.0, process_group=None):
        self.critic, self.lr, self.variant = critic, float(lr), int(variant)
        self.m = torch.zeros_like(critic.theta)
        self.v = torch.zeros_like(critic.theta)
        self.t = 0
        self.state_dev = torch.zeros((4,), dtype=torch.int32, device=critic.theta.device)   # graph-safe step count
        self.target, self.tau = target, float(tau)
        self.pg = process_group

    def step_graph_safe(self, s, a, y, world_size: int = 1):
        """Same as :meth:`step` but with the Adam step count on the device (``rlc_adam_step_dev``), so the
        call can sit inside a captured CUDA graph and still advance on every replay.  Do not mix with
        :meth:`step` on the same optimiser (two counters)."""
        B = s.shape[0]
        grad, loss, q = self.critic.grads(s, a, y, b_total=B * world_size)
        if world_size > 1:
            from .parallel import allreduce_grad_
            allreduce_grad_(grad, self.pg)
        self.critic.eng.adam_step_dev(self.critic.theta, grad, self.m, self.v, self.state_dev, self.lr, self.variant,
                                      target=None if self.target is None else self.target.theta, tau=self.tau)
        return loss, q

    def step(self, s, a, y, world_size: int = 1):
        """One regression step on this rank's shard of the batch. With world_size>1 the per-rank
        gradients are already scaled by 1/B_total, so a SUM all-reduce reproduces the reference's
        batch mean."""
        B = s.shape[0]
        grad, loss, q = self.critic.grads(s, a, y, b_total=B * world_size)
        if world_size > 1:
            from .parallel import allreduce_grad_
            allreduce_grad_(grad, self.pg)
        self.t += 1
        self.critic.eng.adam_step(self.critic.theta, grad, self.m, self.v, self.t, self.lr, self.variant,
                                  target=None if self.target is None else self.target.theta, tau=self.tau)
        return loss, q
