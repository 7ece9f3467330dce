"""CPU oracle for the sampled-action critic path of RLControl (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing in ``rlcontrol_b200/`` may import this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` as the *checker* (or the timed CPU arm), never as the product path.

Every function cites the reference file:line (relative to the RLControl checkout) whose
arithmetic it restates.  Parity status:

* T-in critic (``SoftQNetwork``), FKL / RKL reductions, critic regression step (torch Adam):
  PINNED against the reference's own torch classes run in the build container
  (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
* T-mid critic topology, concat order and weight layout: PINNED against the five trained
  ``Bimodal1DEnv_trueQ_ckpt`` checkpoints + the env ``reward_func`` closed forms.
* Bounded diagonal GMM refit: EM pinned against scikit-learn's ``GaussianMixture`` run through a
  ``BoundedVarGaussianMixture``-equivalent subclass given the same initial responsibilities
  (k-means init consumes RNG and is version dependent: "parity unpinned" for the init only).
* TF Adam / TF graphs: TensorFlow 1.15 cannot be installed -> restated from the TF documentation
  of ``tf.train.AdamOptimizer``; "parity unpinned" for the TF optimizer step.
"""
from __future__ import annotations

import math
import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# Critic forward passes
# --------------------------------------------------------------------------------------


def stack_state_major(s: np.ndarray, n: int) -> np.ndarray:
    """``state.unsqueeze(1).repeat(1,N,1).reshape(-1,S)`` (forwardkl_network.py:161-162) ==
    ``np.repeat(state_batch, N, axis=0)`` (ActorExpert.py:168): row = b*N + n."""
    return np.repeat(np.asarray(s), n, axis=0)


def stack_actions(a: np.ndarray, b: int) -> np.ndarray:
    """Shared grid ``[N,A]`` -> ``intgrl_actions.repeat(B,1,1).reshape(-1,A)``
    (forwardkl_network.py:104-105, reversekl_network.py:178-179); per-state ``[B,N,A]`` ->
    ``reshape(B*N, A)`` (ActorExpert.py:166)."""
    a = np.asarray(a)
    if a.ndim == 2:
        return np.tile(a, (b, 1))
    return a.reshape(-1, a.shape[-1])


def tin_forward(s, a, W1, b1, W2, b2, W3, b3, dtype=F32):
    """T-in critic, ``SoftQNetwork.forward`` (forwardkl_network.py:263-268,
    reversekl_network.py:272-282): ``x=cat([s,a],1)``; Linear+ReLU; Linear+ReLU; Linear.
    Weights are torch ``nn.Linear`` layout ``[out, in]`` (y = x W^T + b).
    ``s``: [R,S], ``a``: [R,A] (already stacked).  Returns q [R] in ``dtype``."""
    x = np.concatenate([np.asarray(s, dtype), np.asarray(a, dtype)], axis=1)
    h1 = np.maximum(x @ np.asarray(W1, dtype).T + np.asarray(b1, dtype), 0)
    h2 = np.maximum(h1 @ np.asarray(W2, dtype).T + np.asarray(b2, dtype), 0)
    q = h2 @ np.asarray(W3, dtype).reshape(-1) + np.asarray(b3, dtype).reshape(())
    return q.astype(dtype)


def tin_eval(s, a, params, dtype=F32):
    """B x N sampled-action evaluation exactly as the reference executes it: materialise the
    state-major stack (forwardkl_network.py:160-164) and run the critic on B*N rows.
    ``s`` [B,S]; ``a`` [N,A] (shared grid) or [B,N,A].  Returns q [B,N]."""
    s = np.asarray(s)
    a = np.asarray(a)
    B = s.shape[0]
    N = a.shape[0] if a.ndim == 2 else a.shape[1]
    q = tin_forward(stack_state_major(s, N), stack_actions(a, B), *params, dtype=dtype)
    return q.reshape(B, N)


def round_operand(x, kind):
    """Round-to-nearest-even of an fp64/fp32 array to the tensor-core operand type ``kind``
    ('fp16' saturating at +-65504, or 'bf16'), returned as float64.  Used to restate the
    *stated arithmetic* of the tcgen05 path (fp16/bf16 operands, fp32 accumulate)."""
    x = np.asarray(x, np.float32)
    if kind == "fp16":
        return np.clip(x, -65504.0, 65504.0).astype(np.float16).astype(np.float64)
    if kind == "bf16":
        u = x.view(np.uint32).astype(np.uint64)
        u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
        return u.astype(np.uint32).view(np.float32).astype(np.float64)
    raise ValueError(kind)


def round_fp8(x, kind):
    """Round-to-nearest-even to an 8-bit float (``cvt.rn.satfinite``), returned as float64: 'e4m3' (3 mantissa bits,
    min normal 2^-6, max 448) or 'e5m2' (2 mantissa bits, min normal 2^-14, max 57344); subnormals kept, overflow
    saturates.  Operand types of the FP8 correction terms of RLC_PREC_FP16C8."""
    mb, emin, mx = {"e4m3": (3, -6, 448.0), "e5m2": (2, -14, 57344.0)}[kind]
    x = np.asarray(x, np.float64)
    ax = np.abs(x)
    e = np.maximum(np.floor(np.log2(np.where(ax > 0, ax, 1.0))), emin)
    step = np.exp2(e - mb)
    return np.sign(x) * np.minimum(np.round(ax / step) * step, mx)


def tin_eval_rounded(s, a, params, kind="fp16", head="ss"):
    """T-in evaluation with the operand rounding of the tensor-core path made explicit (fp64
    accumulate): x=[s;a], W1, b1 (folded as a ones column) and the ReLU'd layer-1 activations are
    rounded to ``kind``.  Two stated arithmetics for layer 2 + head (rlc_umma_mode()):

    ``head="ss"``     W2 rounded; b2, W3, b3 and all accumulation wide:
                      q = b3 + sum_j w3_j relu(h1 . r(W2[j,:]) + b2_j)
    ``head="folded"`` the output head is folded into layer 2's operands (TS kernel):
                      q = b3 + 2^-k sum_j sign(w3_j) relu(h1 . r(2^k |w3_j| W2[j,:]) + r(2^k |w3_j| b2_j))
                      with 2^k max|w3| in [1,2); the products 2^k|w3_j|*x are formed in fp32 (one
                      rounding, as the pack kernel does) before the rounding to ``kind``.

    ``head="grid"``   shared action grids only (K1-grid kernel): layer 1 is hoisted,
                      h1 = relu(r(r(b1 + W1s s_b) + r(W1a a_n)))  -- the two partial pre-activations
                      are formed in fp32 and rounded to ``kind``, their sum is rounded once more
                      (one packed fma) -- followed by the folded head.

    ``head="grid3"``  the strict split mode (RLC_PREC_FP16X3, shared grids, ``kind`` is ignored: fp16):
                      h = relu(fl32(PS_b + PA_n)) with fp32 tables (same fp32 FMA chains), h = h_hi + h_lo with
                      h_hi = r(h), h_lo = r(h - h_hi); W' = fl32(2^k |w3_j| W2[j,:]) = W_hi + W_lo likewise, 2^k the
                      head-folding scale times 2^e2 with 2^e2 max(|W2|,|b2|) in [2^8, 2^9);
                      z = h_hi.W_hi + h_lo.W_hi + h_hi.W_lo (the bias rides on an always-one feature), folded head.
    ``head="grid3c8"`` RLC_PREC_FP16C8: same h, W' and hi parts, the two corrections with 8-bit operands:
                      z = h_hi.W_hi + e4m3(2^9 (h - h_hi)).e4m3(2^-9 W_hi) + e5m2(h_hi).e4m3(W' - W_hi),
                      e5m2(h_hi) = upper byte of the fp16 bits of h_hi + 0x80 (round half up).

    The CUDA kernel must match THIS to ~1e-5 rms (it does exactly this arithmetic with fp32
    accumulators); the distance from ``tin_eval(..., float64)`` is the cost of the operand type."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, np.float64) for p in params]
    s = np.asarray(s)
    a = np.asarray(a)
    B = s.shape[0]
    N = a.shape[0] if a.ndim == 2 else a.shape[1]
    x = np.concatenate([stack_state_major(s, N), stack_actions(a, B)], axis=1)
    r = lambda z: round_operand(z, kind)
    if head == "grid3":
        return _tin_eval_grid3(s, a, params)
    if head == "grid3c8":
        return _tin_eval_grid3(s, a, params, c8=True)
    if head == "grid":
        if a.ndim != 2:
            raise ValueError("head='grid' is the shared-grid arithmetic")
        S = s.shape[1]
        f32 = np.float32

        def fma_chain(xs, Wt, init):
            """acc = init; for k: acc = fl32(acc + x_k * W_k)  -- the pre-pass kernel's fp32 FMA
            order (the product of two fp32 values is exact in fp64)."""
            acc = np.broadcast_to(np.asarray(init, f32), (xs.shape[0], Wt.shape[1])).astype(f32)
            for k in range(xs.shape[1]):
                acc = (acc.astype(np.float64) + xs[:, k:k + 1].astype(np.float64) * Wt[k].astype(np.float64)).astype(f32)
            return acc

        W1f = np.asarray(params[0], f32)                       # [H1, S+A]
        ps = r(fma_chain(np.asarray(s, f32), W1f[:, :S].T.copy(), np.asarray(params[1], f32)))
        pa = r(fma_chain(np.asarray(a, f32), W1f[:, S:].T.copy(), f32(0)))
        h1 = r(np.maximum(ps[:, None, :] + pa[None, :, :], 0)).reshape(B * N, -1)
        head = "folded"
    else:
        h1 = r(np.maximum(r(x) @ r(W1).T + r(b1), 0))
    w3 = W3.reshape(-1)
    if head == "ss":
        h2 = np.maximum(h1 @ r(W2).T + b2, 0)
        q = h2 @ w3 + b3.reshape(())
    elif head == "folded":
        mx = np.float32(np.abs(w3).max())
        scale = np.float32(1.0)
        if mx > 0 and np.isfinite(mx):
            scale = np.float32(np.ldexp(1.0, 1 - int(np.frexp(mx)[1])))
        sw = scale * np.abs(w3).astype(np.float32)                                   # exact (power of two)
        W2s = r(sw[:, None] * np.asarray(params[2], np.float32))                     # fp32 product, then kind
        bs = r(sw * np.asarray(params[3], np.float32).reshape(-1))
        z = np.maximum(h1 @ W2s.T + bs, 0)
        sign = np.where(w3 < 0, -1.0, 1.0)
        q = (z @ sign) / np.float64(scale) + b3.reshape(())
    else:
        raise ValueError(head)
    return q.reshape(B, N)


def _tin_eval_grid3(s, a, params, c8=False, sa=9, sb=0):
    """Stated arithmetic of the split tensor modes (csrc/critic_umma_grid3.cuh), fp64 accumulate."""
    if np.asarray(a).ndim != 2:
        raise ValueError("head='grid3' is a shared-grid arithmetic")
    f32, f64 = np.float32, np.float64
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, f32) for p in params]
    s = np.asarray(s, f32)
    a = np.asarray(a, f32)
    B, N, S = s.shape[0], a.shape[0], s.shape[1]
    r16 = lambda z: round_operand(z, "fp16")

    def fma_chain(xs, Wt, init):
        acc = np.broadcast_to(np.asarray(init, f32), (xs.shape[0], Wt.shape[1])).astype(f32)
        for k in range(xs.shape[1]):
            acc = (acc.astype(f64) + xs[:, k:k + 1].astype(f64) * Wt[k].astype(f64)).astype(f32)
        return acc

    ps = fma_chain(s, W1[:, :S].T.copy(), b1)                        # [B,H1] fp32
    pa = fma_chain(a, W1[:, S:].T.copy(), f32(0))                    # [N,H1] fp32
    h = np.maximum((ps[:, None, :] + pa[None, :, :]).astype(f32), f32(0)).reshape(B * N, -1)   # one fp32 add
    h_hi = r16(h)
    h_lo = r16((h.astype(f64) - h_hi).astype(f32))                   # exact difference
    w3 = W3.reshape(-1)
    mx = f32(np.abs(w3).max())
    scale = f32(1.0)
    if mx > 0 and np.isfinite(mx):
        scale = f32(np.ldexp(1.0, 1 - int(np.frexp(mx)[1])))
    mx2 = f32(max(np.abs(W2).max(), np.abs(b2).max()))
    if mx2 > 0 and np.isfinite(mx2):
        scale = f32(scale * f32(np.ldexp(1.0, 9 - int(np.frexp(mx2)[1]))))
    sw = (scale * np.abs(w3)).astype(f32)                            # exact (power of two)
    Wp = np.clip((sw[:, None] * W2).astype(f32), -65504, 65504)      # [H2,H1], one fp32 rounding
    bp = np.clip((sw * b2.reshape(-1)).astype(f32), -65504, 65504)
    W_hi, b_hi = r16(Wp), r16(bp)
    if c8:
        h_lo8 = round_fp8((h.astype(f64) - h_hi).astype(f32).astype(f64) * 2.0 ** sa, "e4m3")
        # e5m2 copy of h_hi: the upper byte of the fp16 value after adding half an e5m2 ulp (round half UP in magnitude)
        u = np.asarray(h_hi, np.float32).astype(np.float16).view(np.uint16).astype(np.uint32)
        h_hi8 = (((u + 0x80) & 0xFF00).astype(np.uint16).view(np.float16).astype(f64)) * 2.0 ** -sb
        W_hi8 = round_fp8(W_hi * 2.0 ** -sa, "e4m3")
        W_lo8 = round_fp8((Wp.astype(f64) - W_hi).astype(f32).astype(f64) * 2.0 ** sb, "e4m3")
        b_lo8 = round_fp8((bp.astype(f64) - b_hi).astype(f32).astype(f64) * 2.0 ** sb, "e4m3")
        # the bias feature is h = 1: hi part 1, lo part 0, e5m2(2^-sb) exact
        z = h_hi @ W_hi.T + h_lo8 @ W_hi8.T + h_hi8 @ W_lo8.T + b_hi + (2.0 ** -sb) * b_lo8
    else:
        W_lo, b_lo = r16((Wp.astype(f64) - W_hi).astype(f32)), r16((bp.astype(f64) - b_hi).astype(f32))
        z = h_hi @ W_hi.T + h_lo @ W_hi.T + h_hi @ W_lo.T + (b_hi + b_lo)
    sign = np.where(w3 < 0, -1.0, 1.0)
    q = (np.maximum(z, 0) @ sign) / f64(scale) + f64(b3.reshape(()))
    return q.reshape(B, N)


def clip_state(s, smin, smax, dtype=F32):
    """Input "normalisation" of the TF graphs: ``RunningMeanStd`` mean/var are baked in as the
    constants 0/1 (utils/running_mean_std.py:4-7,31-32; base_network_manager.py:37) so the only
    live op is ``tf.clip_by_value(x, state_min, state_max)`` (critic_network.py:71,
    qt_opt_network.py:79)."""
    if smin is None:
        return np.asarray(s, dtype)
    return np.clip(np.asarray(s, dtype), np.asarray(smin, dtype), np.asarray(smax, dtype))


def tmid_forward(s, a, W1, b1, W2, b2, W3, b3, smin=None, smax=None, dtype=F32):
    """T-mid critic ``network()`` (critic_network.py:77-99, qt_opt_network.py:83-105,
    ae_network.py:213-227): ``h1=ReLU(clip(s) W1 + b1)``; ``h2=ReLU(concat([h1,a]) W2 + b2)``;
    ``q=h2 W3 + b3``.  TF ``fully_connected`` kernels are ``[in, out]`` (y = x W + b);
    rows 0..H1-1 of W2 multiply h1, rows H1.. multiply the action (verified on the trueQ
    checkpoint).  ``s`` [R,S], ``a`` [R,A] stacked.  Returns q [R]."""
    x = clip_state(s, smin, smax, dtype)
    h1 = np.maximum(x @ np.asarray(W1, dtype) + np.asarray(b1, dtype), 0)
    z = np.concatenate([h1, np.asarray(a, dtype)], axis=1)
    h2 = np.maximum(z @ np.asarray(W2, dtype) + np.asarray(b2, dtype), 0)
    q = h2 @ np.asarray(W3, dtype).reshape(-1) + np.asarray(b3, dtype).reshape(())
    return q.astype(dtype)


def tmid_eval(s, a, params, smin=None, smax=None, dtype=F32):
    """``predict_q(stacked_states, stacked_actions, phase)`` on the un-hoisted stack, as TF runs
    it (ae_network.py:377-387, qt_opt_network.py:107-117).  Returns q [B,N]."""
    s = np.asarray(s)
    a = np.asarray(a)
    B = s.shape[0]
    N = a.shape[0] if a.ndim == 2 else a.shape[1]
    q = tmid_forward(stack_state_major(s, N), stack_actions(a, B), *params,
                     smin=smin, smax=smax, dtype=dtype)
    return q.reshape(B, N)


def tmid_eval_hoisted(s, a, params, smin=None, smax=None, dtype=F32):
    """Algebraically identical T-mid evaluation with the state-only terms hoisted
    (SURVEY 0.4): p_b = h1_b W2[:H1] + b2; q = ReLU(p_b + a W2[H1:]) . w3 + b3."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, dtype) for p in params]
    H1 = W1.shape[1]
    x = clip_state(s, smin, smax, dtype)
    h1 = np.maximum(x @ W1 + b1, 0)
    p = h1 @ W2[:H1] + b2                                   # [B,H2]
    a = np.asarray(a, dtype)
    if a.ndim == 2:
        pa = (a @ W2[H1:])[None, :, :]                      # [1,N,H2]
    else:
        pa = a @ W2[H1:]                                    # [B,N,H2]
    h2 = np.maximum(p[:, None, :] + pa, 0)
    return (h2 @ W3.reshape(-1) + b3.reshape(())).astype(dtype)


# --------------------------------------------------------------------------------------
# Per-state reductions
# --------------------------------------------------------------------------------------


def topk_desc(q, k):
    """Per-state elites: ``row.argsort()[::-1][:k]`` (ActorExpert.py:177, qt_opt_network.py:166).
    numpy's default introsort does not define the order of ties, so the oracle fixes it as the
    reverse of a *stable* ascending sort (ties: larger index first), which is what numpy
    produces for the insertion-sort sized rows and what the CUDA kernel implements."""
    q = np.asarray(q)
    idx = np.argsort(q, axis=1, kind="stable")[:, ::-1][:, :k]
    return np.ascontiguousarray(idx.astype(np.int64))


def gather_elites(actions, idx):
    """``[actions[idxs] for actions, idxs in zip(action_batch, selected_idxs)]``
    (ActorExpert.py:178, qt_opt_network.py:168-170).  actions [B,N,A], idx [B,k]."""
    return np.take_along_axis(np.asarray(actions), np.asarray(idx)[:, :, None], axis=1)


def argmax_max_mean(q):
    """``np.argmax / np.max / np.mean(axis=1)`` (optimal_q_network.py:156-159,
    ActorCritic.py:139,158,210,223).  np.argmax returns the first maximal index."""
    q = np.asarray(q)
    return np.argmax(q, axis=1).astype(np.int64), np.max(q, axis=1), np.mean(q, axis=1, dtype=q.dtype)


def sql_soft_value(q_targ, action_dim):
    """SQL soft value (sql_network.py:76-84): ``logsumexp_n Q - log N + A log 2``."""
    q = np.asarray(q_targ, F32)
    m = np.max(q, axis=1, keepdims=True)
    lse = (m[:, 0] + np.log(np.sum(np.exp(q - m), axis=1, dtype=F32))).astype(F32)
    return (lse - F32(np.log(F32(q.shape[1]))) + F32(action_dim * np.log(2))).astype(F32)


def fkl_reduce(q, w, logp, entropy_scale, dtype=F32):
    """ForwardKL policy loss over the quadrature grid (forwardkl_network.py:165-194):
    ``t=q/alpha``; ``m_b=max_n t``; ``e=exp(t-m)``; ``z_b=sum_n w_n e``; ``p=e/z`` (detached);
    ``loss = -mean_b sum_n w_n p logpi``.
    Returns (loss scalar, per-state loss [B], boltzmann p [B,N], dloss/dlogp [B,N])."""
    q = np.asarray(q, dtype)
    w = np.asarray(w, dtype)
    logp = np.asarray(logp, dtype)
    t = q / dtype(entropy_scale)
    m = np.max(t, axis=1, keepdims=True)
    e = np.exp(t - m)
    z = np.sum(e * w[None, :], axis=1, keepdims=True)
    p = e / z
    per_state = -np.sum(p * logp * w[None, :], axis=1)
    loss = np.mean(per_state)
    dlogp = -(p * w[None, :]) / dtype(q.shape[0])
    return dtype(loss), per_state.astype(dtype), p.astype(dtype), dlogp.astype(dtype)


def rkl_reduce(q, v, w, logp, entropy_scale, hard=False, dtype=F32):
    """ReverseKL policy loss (reversekl_network.py:181-190; hard variant :197-203):
    integrand ``-exp(logpi) * ((q - v).detach() - alpha*logpi)`` (hard: no entropy term),
    times ``w_n``, summed over n, mean over b.
    Returns (loss, per-state loss [B], dloss/dlogp [B,N])."""
    q = np.asarray(q, dtype)
    v = np.asarray(v, dtype).reshape(-1, 1)
    w = np.asarray(w, dtype)
    logp = np.asarray(logp, dtype)
    al = dtype(0.0 if hard else entropy_scale)
    pe = np.exp(logp)
    adv = q - v
    integrand = -pe * (adv - al * logp)
    per_state = np.sum(integrand * w[None, :], axis=1)
    loss = np.mean(per_state)
    # d/dlogp [-e^l (adv - al*l)] = -e^l (adv - al*l) + al*e^l = -e^l (adv - al*l - al)
    dlogp = (-pe * (adv - al * logp - al)) * w[None, :] / dtype(q.shape[0])
    return dtype(loss), per_state.astype(dtype), dlogp.astype(dtype)


# --------------------------------------------------------------------------------------
# Quadrature grid (third-party: quadpy, unpinned in requirements.txt:4)
# --------------------------------------------------------------------------------------


def clenshaw_curtis(n):
    """``quadpy.c1.clenshaw_curtis(n)`` (call sites forwardkl_network.py:64,
    reversekl_network.py:69): n nodes x_j = -cos(j*pi/(n-1)) on [-1,1] and the classical
    Clenshaw-Curtis weights (published algorithm; quadpy is not vendored in the reference).
    The reference then drops both endpoints (``points[1:-1]``)."""
    if n < 2:
        raise ValueError("clenshaw_curtis needs n >= 2")
    N = n - 1
    j = np.arange(n)
    theta = np.pi * j / N
    x = -np.cos(theta)
    w = np.zeros(n)
    for i in range(n):
        s = 0.0
        for k in range(1, N // 2 + 1):
            bk = 1.0 if 2 * k == N else 2.0
            s += bk / (4.0 * k * k - 1.0) * math.cos(2.0 * k * theta[i])
        c = 1.0 if i in (0, N) else 2.0
        w[i] = c / N * (1.0 - s)
    return x, w


def clenshaw_curtis_fast(n):
    """Vectorised version of :func:`clenshaw_curtis` (same arithmetic, numpy)."""
    N = n - 1
    theta = np.pi * np.arange(n) / N
    x = -np.cos(theta)
    k = np.arange(1, N // 2 + 1)
    bk = np.where(2 * k == N, 1.0, 2.0)
    s = (bk / (4.0 * k * k - 1.0))[None, :] * np.cos(2.0 * k[None, :] * theta[:, None])
    c = np.full(n, 2.0)
    c[0] = c[-1] = 1.0
    w = c / N * (1.0 - s.sum(axis=1))
    return x, w


def intg_grid_1d(n_param, action_max):
    """A=1 integration grid of FKL/RKL (forwardkl_network.py:60-71): interior CC nodes scaled
    by action_max, interior weights unscaled, both fp32."""
    x, w = clenshaw_curtis_fast(int(n_param))
    acts = (x[1:-1].astype(F32)[:, None] * F32(action_max)).astype(F32)
    return acts, w[1:-1].astype(F32)


# --------------------------------------------------------------------------------------
# tanh-Gaussian log-prob on the grid (actor side; an *input* to the hot path)
# --------------------------------------------------------------------------------------


def tanh_gauss_logprob_1d(mean, log_std, actions, action_scale, eps=1e-6):
    """``PolicyNetwork.get_logprob`` for action_dim==1 (forwardkl_network.py:324-344):
    Normal(mean,std).log_prob(atanh(a/scale)) - log(1-(a/scale)^2+eps).
    mean/log_std [B,1], actions [N,1] -> [B,N]."""
    mean = np.asarray(mean, F32).reshape(-1, 1)
    log_std = np.asarray(log_std, F32).reshape(-1, 1)
    na = (np.asarray(actions, F32).reshape(1, -1) / F32(action_scale)).astype(F32)
    at = ((np.log(1 + na) - np.log(1 - na)) / 2).astype(F32)
    std = np.exp(log_std)
    lp = -((at - mean) ** 2) / (2 * std * std) - log_std - F32(math.log(math.sqrt(2 * math.pi)))
    lp = lp - np.log(1 - na * na + F32(eps))
    return lp.astype(F32)


def tanh_gauss_logprob(mean, log_std, actions, action_scale, eps=1e-6, dtype=np.float64):
    """``PolicyNetwork.get_logprob`` for any action_dim (forwardkl_network.py:324-351,
    reversekl_network.py:346-374).  u = atanh(a/scale) = (log(1+x) - log(1-x))/2 (:353-354).
      A == 1: Normal(mean, std).log_prob(u)                = -(u-m)^2/(2 std^2) - log std - log sqrt(2 pi)
      A  > 1: MultivariateNormal(mean, diag_embed(std))    -- std is passed as the COVARIANCE (:350), so
              log_prob = -1/2 sum_k (u_k-m_k)^2/std_k - 1/2 sum_k log std_k - A/2 log(2 pi)
    minus sum_k log(1 - x_k^2 + eps).  mean/log_std [B,A], actions [N,A] -> (logp [B,N],
    dlogp/dmean [B,N,A], dlogp/dlog_std [B,N,A])."""
    mean = np.asarray(mean, dtype)
    log_std = np.asarray(log_std, dtype)
    B, A = mean.shape
    x = np.asarray(actions, dtype) / dtype(action_scale)                 # [N,A]
    u = (np.log(1 + x) - np.log(1 - x)) / 2
    jac = np.log(1 - x * x + dtype(eps)).sum(axis=1)                     # [N]
    d = u[None, :, :] - mean[:, None, :]                                 # [B,N,A]
    std = np.exp(log_std)[:, None, :]
    if A == 1:
        lp = (-(d * d) / (2 * std * std) - log_std[:, None, :] - dtype(0.5 * math.log(2 * math.pi))).sum(-1)
        dmean = d / (std * std)
        dlstd = d * d / (std * std) - 1
    else:
        lp = (-(d * d) / (2 * std)).sum(-1) - 0.5 * log_std.sum(-1)[:, None] - dtype(0.5 * A * math.log(2 * math.pi))
        dmean = d / std
        dlstd = 0.5 * d * d / std - 0.5
    return (lp - jac[None, :]).astype(dtype), dmean.astype(dtype), dlstd.astype(dtype)


def fkl_policy_reduce(q, w, actions, mean, log_std, action_scale, entropy_scale, b_total=None, dtype=np.float64):
    """ForwardKL grid reduction with the policy log-density evaluated in place (a3 + a5):
    returns (per-state loss [B], dL/dmean [B,A], dL/dlog_std [B,A], logp [B,N]) for
    L = mean_b loss_b; Boltzmann weights are constants for the gradient (``.detach()``,
    forwardkl_network.py:176-184)."""
    lp, dm, ds = tanh_gauss_logprob(mean, log_std, actions, action_scale, dtype=dtype)
    _, per_state, _, dlogp = fkl_reduce(q, w, lp, entropy_scale, dtype=dtype)
    if b_total is not None:
        dlogp = dlogp * dtype(np.asarray(q).shape[0]) / dtype(b_total)
    return per_state, (dlogp[:, :, None] * dm).sum(1), (dlogp[:, :, None] * ds).sum(1), lp


def rkl_policy_reduce(q, v, w, actions, mean, log_std, action_scale, entropy_scale, hard=False, b_total=None,
                      dtype=np.float64):
    """ReverseKL counterpart of :func:`fkl_policy_reduce` (reversekl_network.py:181-203)."""
    lp, dm, ds = tanh_gauss_logprob(mean, log_std, actions, action_scale, dtype=dtype)
    _, per_state, dlogp = rkl_reduce(q, v, w, lp, entropy_scale, hard=hard, dtype=dtype)
    if b_total is not None:
        dlogp = dlogp * dtype(np.asarray(q).shape[0]) / dtype(b_total)
    return per_state, (dlogp[:, :, None] * dm).sum(1), (dlogp[:, :, None] * ds).sum(1), lp


# --------------------------------------------------------------------------------------
# CEM: bounded diagonal GMM refit + loop
# --------------------------------------------------------------------------------------

ACTION_BOUND = 2.0                  # utils/boundedvar_gaussian_mixture.py:7
SIGMA_BOUND = 1.0                   # utils/boundedvar_gaussian_mixture.py:8
VAR_LO = math.exp(-2 * SIGMA_BOUND)
VAR_HI = math.exp(2 * SIGMA_BOUND)
REG_COVAR = 1e-6                    # sklearn GaussianMixture default


def _estimate_diag(X, resp, reg_covar=REG_COVAR):
    """sklearn ``_estimate_gaussian_parameters(..., 'diag')``: nk = sum resp + 10 eps;
    means = resp^T X / nk; cov = resp^T X^2/nk - 2 means*(resp^T X)/nk + means^2 + reg."""
    nk = resp.sum(axis=0) + 10 * np.finfo(resp.dtype).eps
    means = resp.T @ X / nk[:, None]
    avg_X2 = resp.T @ (X * X) / nk[:, None]
    avg_means2 = means ** 2
    avg_X_means = means * (resp.T @ X) / nk[:, None]
    cov = avg_X2 - 2 * avg_X_means + avg_means2 + reg_covar
    return nk, means, cov


def _bound(means, cov):
    """The two ``np.clip`` lines of boundedvar_gaussian_mixture.py:27-28 and :68-69."""
    return (np.clip(means, -ACTION_BOUND, ACTION_BOUND), np.clip(cov, VAR_LO, VAR_HI))


def _e_step_diag(X, weights, means, cov):
    """sklearn ``_estimate_log_gaussian_prob`` (diag) + ``_estimate_log_weights`` +
    logsumexp normalisation.  Returns (mean log-likelihood, log_resp)."""
    n, d = X.shape
    prec_chol = 1.0 / np.sqrt(cov)                         # _compute_precision_cholesky('diag')
    log_det = np.sum(np.log(prec_chol), axis=1)
    prec = prec_chol ** 2
    log_prob = (np.sum(means ** 2 * prec, axis=1)
                - 2.0 * (X @ (means * prec).T)
                + (X ** 2) @ prec.T)
    log_gauss = -0.5 * (d * np.log(2 * np.pi) + log_prob) + log_det
    weighted = log_gauss + np.log(weights)
    m = np.max(weighted, axis=1, keepdims=True)
    norm = m[:, 0] + np.log(np.sum(np.exp(weighted - m), axis=1))
    log_resp = weighted - norm[:, None]
    return float(np.mean(norm)), log_resp


def gmm_fit_bounded(X, resp0, tol=1e-2, max_iter=100, reg_covar=REG_COVAR):
    """``BoundedVarGaussianMixture(n_components=M, covariance_type='diag', tol=1e-2).fit(X)``
    (qt_opt_network.py:173; utils/boundedvar_gaussian_mixture.py:10-75) given the initial
    responsibilities ``resp0`` [k,M] that sklearn derives from its k-means init (one-hot).
    sklearn ``BaseMixture.fit_predict`` loop: initialise -> repeat {E, M, lower-bound change <
    tol -> stop}.  Returns (weights [M], means [M,A], covariances [M,A], n_iter)."""
    X = np.asarray(X, np.float64)
    resp0 = np.asarray(resp0, np.float64)
    n = X.shape[0]
    nk, means, cov = _estimate_diag(X, resp0, reg_covar)       # _initialize :22-23
    means, cov = _bound(means, cov)                            # :27-28
    weights = nk / n                                           # :30
    lower = -np.inf
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        prev = lower
        ll, log_resp = _e_step_diag(X, weights, means, cov)
        nk, means, cov = _estimate_diag(X, np.exp(log_resp), reg_covar)   # _m_step :62-64
        means, cov = _bound(means, cov)                                   # :68-69
        weights = nk / n                                                  # :72
        lower = ll
        if abs(lower - prev) < tol:
            break
    return weights, means, cov, n_iter


def gmm_fit_1comp(X, reg_covar=REG_COVAR):
    """Closed form for ``n_components=1``: resp==1 so one M-step is the fixed
    point: mu=clip(mean), var=clip(E[x^2]-mu0^2+1e-6) with mu0 the *unclipped* mean."""
    X = np.asarray(X, np.float64)
    resp = np.ones((X.shape[0], 1))
    nk, means, cov = _estimate_diag(X, resp, reg_covar)
    means, cov = _bound(means, cov)
    return np.ones(1), means, cov


def farthest_point_labels(X):
    """Deterministic 2-way initial partition used by the device CEM in place of sklearn's
    RNG-consuming k-means init ("parity unpinned" for the init): seed c0 = the
    first elite (best Q), c1 = the elite farthest from c0 (lowest index on ties); label each
    elite by the nearer seed (c0 on ties)."""
    X = np.asarray(X, np.float64)
    d0 = np.sum((X - X[0]) ** 2, axis=1)
    j = int(np.argmax(d0))
    d1 = np.sum((X - X[j]) ** 2, axis=1)
    return (d1 < d0).astype(np.int64)


def labels_to_resp(labels, M):
    r = np.zeros((len(labels), M))
    r[np.arange(len(labels)), labels] = 1.0
    return r


def gmm_sample_with_noise(weights, means, cov, comp_u, normal_noise):
    """Restatement of ``GaussianMixture.sample`` with the RNG draws supplied as tensors:
    component pick by inverse CDF of ``comp_u`` [N] over ``weights``; x = mu_c + sqrt(var_c)*z.
    (sklearn draws a multinomial count and stacks per-component blocks; the *distribution* is the
    same, the ordering is not -- the CEM only consumes the samples as a set per state.)"""
    cdf = np.cumsum(weights)
    cdf[-1] = 1.0 + 1e-12
    c = np.searchsorted(cdf, comp_u, side="right")
    c = np.minimum(c, len(weights) - 1)
    return means[c] + np.sqrt(cov[c]) * normal_noise


def cem_iterate(q_fn, s, u0, noise, comp_u, top_m, num_modal, a_min, a_max, tol=1e-2):
    """``iterate_cem_multidim`` (qt_opt_network.py:132-175) with the random draws supplied:
    iter 0 actions = a_min + u0*(a_max-a_min), u0 [B,N,A] in [0,1); later iters sample the
    refit mixture (samples are *not* clipped to the action box, QT_OPT.py:41 clips only the
    final action).  ``q_fn(s, actions[B,N,A]) -> q [B,N]``.
    Returns (weights [B,M], means [B,M,A], covs [B,M,A], elites_idx per iter [iters,B,top_m])."""
    s = np.asarray(s)
    B, N, A = u0.shape
    iters = 1 + (0 if noise is None else noise.shape[0])
    a_min = np.asarray(a_min, np.float64)
    a_max = np.asarray(a_max, np.float64)
    actions = a_min + u0.astype(np.float64) * (a_max - a_min)
    Wt = np.zeros((B, num_modal))
    Mu = np.zeros((B, num_modal, A))
    Cv = np.zeros((B, num_modal, A))
    all_idx = np.zeros((iters, B, top_m), np.int64)
    for it in range(iters):
        if it > 0:
            actions = np.stack([
                gmm_sample_with_noise(Wt[b], Mu[b], Cv[b], comp_u[it - 1, b], noise[it - 1, b])
                for b in range(B)])
        q = q_fn(s, actions.astype(F32))
        idx = topk_desc(q, top_m)
        all_idx[it] = idx
        el = gather_elites(actions.astype(F32), idx).astype(np.float64)
        for b in range(B):
            if num_modal == 1:
                w, mu, cv = gmm_fit_1comp(el[b])
            else:
                resp0 = labels_to_resp(farthest_point_labels(el[b]), num_modal)
                w, mu, cv, _ = gmm_fit_bounded(el[b], resp0, tol=tol)
            Wt[b], Mu[b], Cv[b] = w, mu, cv
    return Wt, Mu, Cv, all_idx


def cem_iterate_forced(q_fn, s, u0, noise, comp_u, top_m, num_modal, a_min, a_max, forced_idx, tol=1e-2):
    """Teacher-forced replay of :func:`cem_iterate` for parity checks: every iteration's candidates are sampled from the
    mixture refit on ``forced_idx[it-1]`` (the elites the implementation under test picked), so each iteration is a pure
    function of identical inputs and a near-tie in one iteration cannot hide what later iterations do.
    Returns (q per iteration [iters,B,N] as q_fn returns it, this oracle's own top-m per iteration [iters,B,top_m],
    weights, means, covs after the last forced refit)."""
    s = np.asarray(s)
    B, N, A = u0.shape
    iters = 1 + (0 if noise is None else noise.shape[0])
    a_min = np.asarray(a_min, np.float64)
    a_max = np.asarray(a_max, np.float64)
    actions = a_min + u0.astype(np.float64) * (a_max - a_min)
    Wt = np.zeros((B, num_modal))
    Mu = np.zeros((B, num_modal, A))
    Cv = np.zeros((B, num_modal, A))
    qs, own = [], np.zeros((iters, B, top_m), np.int64)
    for it in range(iters):
        if it > 0:
            actions = np.stack([
                gmm_sample_with_noise(Wt[b], Mu[b], Cv[b], comp_u[it - 1, b], noise[it - 1, b])
                for b in range(B)])
        q = np.asarray(q_fn(s, actions.astype(F32)))
        qs.append(q)
        own[it] = topk_desc(q, top_m)
        el = gather_elites(actions.astype(F32), np.asarray(forced_idx[it])).astype(np.float64)
        for b in range(B):
            if num_modal == 1:
                w, mu, cv = gmm_fit_1comp(el[b])
            else:
                resp0 = labels_to_resp(farthest_point_labels(el[b]), num_modal)
                w, mu, cv, _ = gmm_fit_bounded(el[b], resp0, tol=tol)
            Wt[b], Mu[b], Cv[b] = w, mu, cv
    return np.stack(qs), own, Wt, Mu, Cv


def cem_final_action(weights, means):
    """``gmm.means_[np.argmax(gmm.weights_)]`` (qt_opt_network.py:180)."""
    k = np.argmax(weights, axis=1)
    return means[np.arange(means.shape[0]), k]


# --------------------------------------------------------------------------------------
# Backward: dQ/da and the critic regression step
# --------------------------------------------------------------------------------------


def tin_dq_da(s, a, params, dtype=np.float64):
    """``tf.gradients(q, action)`` for the T-in critic (sql_network.py:101-107):
    dq/da = W1[:, S:]^T ( relu'(z1) * (W2^T (relu'(z2) * w3)) ).  s [R,S], a [R,A]."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, dtype) for p in params]
    S = np.asarray(s).shape[1]
    x = np.concatenate([np.asarray(s, dtype), np.asarray(a, dtype)], axis=1)
    z1 = x @ W1.T + b1
    h1 = np.maximum(z1, 0)
    z2 = h1 @ W2.T + b2
    g2 = (z2 > 0) * W3.reshape(1, -1)
    g1 = (g2 @ W2) * (z1 > 0)
    return g1 @ W1[:, S:]


def tmid_dq_da(s, a, params, smin=None, smax=None, dtype=np.float64):
    """``self.action_grads = tf.gradients(self.q_prediction, self.action_input)``
    (ae_network.py:117, critic_network.py:58): dq/da = (relu'(z2) * w3) W2[H1:]^T."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, dtype) for p in params]
    H1 = W1.shape[1]
    x = clip_state(s, smin, smax, dtype)
    h1 = np.maximum(x @ W1 + b1, 0)
    z2 = np.concatenate([h1, np.asarray(a, dtype)], axis=1) @ W2 + b2
    g2 = (z2 > 0) * W3.reshape(1, -1)
    return g2 @ W2[H1:].T


def q_gradient_ascent(dq_da_fn, s, a0, lr, a_min, a_max, max_steps=10, stop_tol=1e-3):
    """AE+ action refinement (ae_plus_network.py:310-343): up to ``max_steps`` steps
    ``a += flag*lr*grad``, clipped to the box; a row stops (flag=0) once
    mean(|delta a|)/a_max <= stop_tol."""
    a = np.array(a0, np.float64)
    flag = np.ones((a.shape[0], 1))
    amax0 = float(np.asarray(a_max).reshape(-1)[0])
    for _ in range(max_steps):
        g = dq_da_fn(s, a)
        new = np.clip(a + flag * lr * g, a_min, a_max)
        delta = np.mean(np.abs(new - a), axis=1, keepdims=True) / amax0
        a = new
        flag = flag * (delta > stop_tol)
        if not flag.any():
            break
    return a


def tin_mse_grads(s, a, y, params, dtype=np.float64):
    """Gradients of ``nn.MSELoss()(q_net(s,a), y)`` (forwardkl_network.py:133-140) wrt the six
    T-in parameter tensors (torch layout).  Returns (loss, [gW1,gb1,gW2,gb2,gW3,gb3])."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, dtype) for p in params]
    x = np.concatenate([np.asarray(s, dtype), np.asarray(a, dtype)], axis=1)
    B = x.shape[0]
    z1 = x @ W1.T + b1
    h1 = np.maximum(z1, 0)
    z2 = h1 @ W2.T + b2
    h2 = np.maximum(z2, 0)
    q = h2 @ W3.reshape(-1) + b3.reshape(())
    d = q - np.asarray(y, dtype).reshape(-1)
    loss = np.mean(d * d)
    dq = (2.0 / B) * d                                      # [B]
    gW3 = (dq[:, None] * h2).sum(0).reshape(W3.shape)
    gb3 = np.array([dq.sum()]).reshape(b3.shape)
    g2 = dq[:, None] * W3.reshape(1, -1) * (z2 > 0)
    gW2 = g2.T @ h1
    gb2 = g2.sum(0)
    g1 = (g2 @ W2) * (z1 > 0)
    gW1 = g1.T @ x
    gb1 = g1.sum(0)
    return loss, [gW1, gb1, gW2, gb2, gW3, gb3]


def tmid_mse_grads(s, a, y, params, smin=None, smax=None, dtype=np.float64):
    """Gradients of ``tf.reduce_mean(tf.squared_difference(y, q))`` (critic_network.py:54-55,
    qt_opt_network.py:65-66) wrt the T-mid parameters (TF ``[in,out]`` layout)."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, dtype) for p in params]
    x = clip_state(s, smin, smax, dtype)
    B = x.shape[0]
    z1 = x @ W1 + b1
    h1 = np.maximum(z1, 0)
    zc = np.concatenate([h1, np.asarray(a, dtype)], axis=1)
    z2 = zc @ W2 + b2
    h2 = np.maximum(z2, 0)
    q = h2 @ W3.reshape(-1) + b3.reshape(())
    d = q - np.asarray(y, dtype).reshape(-1)
    loss = np.mean(d * d)
    dq = (2.0 / B) * d
    gW3 = (h2 * dq[:, None]).sum(0).reshape(W3.shape)
    gb3 = np.array([dq.sum()]).reshape(b3.shape)
    g2 = dq[:, None] * W3.reshape(1, -1) * (z2 > 0)
    gW2 = zc.T @ g2
    gb2 = g2.sum(0)
    H1 = W1.shape[1]
    g1 = (g2 @ W2[:H1].T) * (z1 > 0)
    gW1 = x.T @ g1
    gb1 = g1.sum(0)
    return loss, [gW1, gb1, gW2, gb2, gW3, gb3]


def adam_step_torch(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """``torch.optim.Adam`` (torch 1.7.1 default, no amsgrad / weight decay; optimizer built at
    forwardkl_network.py:54-56): denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= lr/(1-b1^t) * m/denom.
    ``t`` is the 1-based step count.  Returns (p, m, v)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


def adam_step_tf(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """``tf.train.AdamOptimizer`` (critic_network.py:55): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    p -= lr_t * m / (sqrt(v) + eps)  (epsilon outside the bias correction)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v


def soft_update(target, online, tau):
    """theta' += tau (theta - theta') (critic_network.py:29; torch form
    forwardkl_network.py:211-215 ``t*(1-tau) + p*tau``)."""
    return target + tau * (online - target)


# --------------------------------------------------------------------------------------
# Replay buffer sampling / gather
# --------------------------------------------------------------------------------------


def sample_n_k(rng: np.random.RandomState, n: int, k: int):
    """``RandomAccessQueue.sample_n_k`` (utils/custom_collections.py:107-131): k distinct
    uniform indices from range(n); ``choice(replace=False)`` when 3k >= n, else a rejection
    scheme over 2k draws."""
    if not 0 <= k <= n:
        raise ValueError("Sample larger than population or is negative")
    if k == 0:
        return np.empty((0,), dtype=np.int64)
    if 3 * k >= n:
        return rng.choice(n, k, replace=False)
    result = rng.choice(n, 2 * k)
    selected = set()
    j = k
    for i in range(k):
        x = result[i]
        while x in selected:
            x = result[i] = result[j]
            j += 1
            if j == 2 * k:
                result[k:] = rng.choice(n, k)
                j = k
        selected.add(x)
    return result[:k]


def replay_gather(store, idx):
    """``map(np.array, zip(*batch))`` (utils/replaybuffer.py:32-37) over a struct-of-arrays
    store ``{'state','action','reward','next_state','gamma'}`` in logical FIFO order."""
    idx = np.asarray(idx)
    return tuple(store[k][idx] for k in ("state", "action", "reward", "next_state", "gamma"))


# --------------------------------------------------------------------------------------
# TF checkpoint decoder for the trueQ fixture
# --------------------------------------------------------------------------------------

TRUEQ_OFFSETS = {                      # byte offsets inside every *.data-00000-of-00001
    "b1": (8, (200,)),
    "W1": (2408, (1, 200)),
    "b2": (4808, (200,)),
    "W2": (7208, (201, 200)),
    "b3": (489608, (1,)),
    "W3": (489620, (200, 1)),
}


def decode_trueq_checkpoint(path):
    """Read ``main/qf`` tensors of a ``Bimodal1DEnv_*_trueQ_learned.data-00000-of-00001`` bundle
    as raw little-endian fp32 at the offsets decoded from the ``.index`` table."""
    raw = open(path, "rb").read()
    out = {}
    for name, (off, shape) in TRUEQ_OFFSETS.items():
        n = int(np.prod(shape))
        out[name] = np.frombuffer(raw, dtype="<f4", count=n, offset=off).reshape(shape).copy()
    return out


def bimodal_reward(action, maxima=(-0.6, 0.6), stddev=(0.2, 0.2), height=(1.0, 1.0)):
    """Closed-form ``reward_func`` of the Bimodal1D bandits (environments/environments.py:573-587
    for eq_var1; the variants differ in maxima/stddev/height)."""
    a = np.asarray(action, np.float64)
    return sum(h * np.exp(-0.5 * ((a - mx) / sd) ** 2) for mx, sd, h in zip(maxima, stddev, height))


# --------------------------------------------------------------------------------------
# Actor-Expert actor side: mixture sampling and the mixture NLL on the elites
# --------------------------------------------------------------------------------------

def mixture_pick(alpha, comp_u, equal_modal=False):
    """Component indices of ``rng.choice(M, N, p=alpha_b)`` (ae_network.py:483-486) given the uniforms numpy
    draws for it: numpy's legacy ``choice`` computes cdf = cumsum(p); cdf /= cdf[-1] and returns
    ``cdf.searchsorted(random_sample(N), side='right')`` (float64).  ``equal_modal``: ``rng.choice(M, N)``
    is ``randint`` in numpy (a different stream); with supplied uniforms it is restated as floor(u*M).
    alpha [B,M], comp_u [B,N] -> idx [B,N] int64."""
    alpha = np.asarray(alpha, np.float64)
    u = np.asarray(comp_u, np.float64)
    B, M = alpha.shape
    if equal_modal:
        return np.minimum((u * M).astype(np.int64), M - 1)
    idx = np.empty(u.shape, np.int64)
    for b in range(B):
        cdf = np.cumsum(alpha[b])
        cdf /= cdf[-1]
        idx[b] = np.minimum(cdf.searchsorted(u[b], side="right"), M - 1)
    return idx


def mixture_sample(alpha, mean, sigma, comp_u, normal, a_min, a_max, equal_modal=False, uni_u=None):
    """``sample_action`` of the Actor-Expert networks (ae_network.py:461-496; ae_actor_network.py:310-341)
    with the draws supplied: ``np.clip(rng.normal(m[idx], s[idx]), action_min, action_max)`` -- numpy's
    ``normal(loc, scale)`` is loc + scale * N(0,1) in float64 -- then the first ``uni_u.shape[1]`` samples of
    every state replaced by ``rng.uniform(action_min, action_max)`` = low + (high-low)*u (:489-493).
    mean/sigma [B,M,A] (float32, as TF returns them), normal [B,N,A] -> (actions [B,N,A] float64, idx)."""
    idx = mixture_pick(alpha, comp_u, equal_modal)
    mean, sigma = np.asarray(mean, np.float64), np.asarray(sigma, np.float64)
    B = mean.shape[0]
    rows = np.arange(B)[:, None]
    act = mean[rows, idx] + sigma[rows, idx] * np.asarray(normal, np.float64)
    lo, hi = np.asarray(a_min, np.float64), np.asarray(a_max, np.float64)
    act = np.clip(act, lo, hi)
    if uni_u is not None and np.asarray(uni_u).shape[1] > 0:
        u = np.asarray(uni_u, np.float64)
        act[:, :u.shape[1]] = lo + (hi - lo) * u
    return act, idx


def mixture_nll(alpha, mean, sigma, actions, equal_modal=False, b_total=None):
    """``get_lossfunc`` (ae_network.py:262-278, tf_normal :230-243) on the elite actions [B,k,A] (the reference
    repeats each state k times, ActorExpert.py:179-182): density_m = prod_a sqrt(1/(2 pi s^2)) exp(-(y-m)^2/(2 s^2)),
    mix = sum_m w_m density_m (w = alpha or 1/M), loss = mean over the B*k rows of -log(clip(mix, 1e-30, 1e30)).
    Returns (loss, nll [B,k], dalpha [B,M], dmean [B,M,A], dsigma [B,M,A]) -- gradients of the loss wrt the
    mixture parameters the actor network outputs (float64)."""
    alpha, mean, sigma = (np.asarray(x, np.float64) for x in (alpha, mean, sigma))
    y = np.asarray(actions, np.float64)
    B, k, A = y.shape
    M = alpha.shape[1]
    R = float((b_total or B) * k)
    d = y[:, :, None, :] - mean[:, None, :, :]                               # [B,k,M,A]
    s = sigma[:, None, :, :]
    dens = np.prod(np.sqrt(1.0 / (2 * np.pi * s * s)) * np.exp(-(d * d) / (2 * s * s)), axis=3)   # [B,k,M]
    w = np.full((B, 1, M), 1.0 / M) if equal_modal else alpha[:, None, :]
    mix = (w * dens).sum(2)                                                  # [B,k]
    inside = (mix >= 1e-30) & (mix <= 1e30)
    nll = -np.log(np.clip(mix, 1e-30, 1e30))
    loss = nll.sum() / R
    g = np.where(inside, -1.0 / np.where(inside, mix, 1.0), 0.0) / R         # dloss/dmix  [B,k]
    dalpha = np.zeros_like(alpha) if equal_modal else (g[:, :, None] * dens).sum(1)
    gd = g[:, :, None] * w * dens                                            # dloss/ddens * dens  [B,k,M]
    dmean = (gd[..., None] * d / (s * s)).sum(1)
    dsigma = (gd[..., None] * (d * d / (s ** 3) - 1.0 / s)).sum(1)
    return loss, nll, dalpha, dmean, dsigma


# --------------------------------------------------------------------------------------
# Soft-Q-learning SVGD.  TF is not installable: parity UNPINNED, restated
# from the explicit formulas of utils/sql_kernel.py:7-69 and sql_network.py:96-117.
# --------------------------------------------------------------------------------------

def adaptive_isotropic_gaussian_kernel(xs, ys, h_min=1e-3, dtype=np.float64):
    """utils/sql_kernel.py:7-69: xs [B,Kx,D], ys [B,Ky,D] -> (kappa [B,Kx,Ky], dkappa/dxs [B,Kx,Ky,D], h [B]).
    median = last of top_k(dist_sq, k = Kx*Ky//2 + 1) (the (k)-th largest); h = max(median/log(Kx), h_min)."""
    xs, ys = np.asarray(xs, dtype), np.asarray(ys, dtype)
    B, Kx, D = xs.shape
    Ky = ys.shape[1]
    diff = xs[:, :, None, :] - ys[:, None, :, :]
    dist = (diff ** 2).sum(-1)
    flat = np.sort(dist.reshape(B, Kx * Ky), axis=1)[:, ::-1]
    med = flat[:, Kx * Ky // 2]
    h = np.maximum(med / dtype(np.log(Kx)), dtype(h_min))
    kappa = np.exp(-dist / h[:, None, None])
    grad = -2 * diff / h[:, None, None, None] * kappa[..., None]
    return kappa, grad, h


def svgd_action_gradients(dqda, fixed, updated, h_min=1e-3, eps=1e-6, dtype=np.float64):
    """sql_network.py:100-114: log_p = Q + sum log(1 - a^2 + EPS) on the fixed particles,
    action_gradients = mean_i(kappa * grad_log_p + dkappa) -> [B,Ku,A]."""
    fixed = np.asarray(fixed, dtype)
    glp = np.asarray(dqda, dtype) + (-2 * fixed) / (1 - fixed ** 2 + dtype(eps))
    kappa, kgrad, h = adaptive_isotropic_gaussian_kernel(fixed, updated, h_min, dtype)
    return (kappa[..., None] * glp[:, :, None, :] + kgrad).mean(axis=1), kappa, h
