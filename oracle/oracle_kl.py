"""ORACLE (test infrastructure, never imported by the product): numpy float64 restatement of one whole
ForwardKL / ReverseKL ``update_network`` + ``update_target_network``
(/root/reference/agents/network/forwardkl_network.py:123-215, reversekl_network.py:130-224) with the
policy's N(0,1) draws fed in.  Parity: PINNED on the reference itself -- tests/golden/full_*.npz are
recorded from the unmodified reference classes by oracle/make_golden.py (all three networks before and
after two consecutive updates, six optim_type / q_update_type variants); tests/test_oracle_kl.py checks
this restatement against them, and the GPU tests check librlc against both.

Parameter lists are in torch layout (weights [out,in]):
    q, v, tv = [W1, b1, W2, b2, W3, b3]          pi = [W1, b1, W2, b2, Wm, bm, Ws, bs]
"""
import math

import numpy as np

from . import oracle_np as onp

F64 = np.float64
LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0            # PolicyNetwork.__init__, forwardkl_network.py:294


def mlp_forward(x, W1, b1, W2, b2, W3, b3):
    """ValueNetwork.forward / PolicyNetwork trunk (forwardkl_network.py:283-287,310-312). Returns
    (out, cache)."""
    x = np.asarray(x, F64)
    z1 = x @ np.asarray(W1, F64).T + np.asarray(b1, F64)
    h1 = np.maximum(z1, 0)
    z2 = h1 @ np.asarray(W2, F64).T + np.asarray(b2, F64)
    h2 = np.maximum(z2, 0)
    out = h2 @ np.asarray(W3, F64).T + np.asarray(b3, F64)
    return out, (x, z1, h1, z2, h2)


def mlp_grads(cache, dout, W2, W3):
    """Back-propagation of dLoss/dout [B,O] through :func:`mlp_forward` -> [gW1,gb1,gW2,gb2,gW3,gb3]."""
    x, z1, h1, z2, h2 = cache
    dout = np.asarray(dout, F64)
    gW3 = dout.T @ h2
    gb3 = dout.sum(0)
    g2 = (dout @ np.asarray(W3, F64)) * (z2 > 0)
    gW2 = g2.T @ h1
    gb2 = g2.sum(0)
    g1 = (g2 @ np.asarray(W2, F64)) * (z1 > 0)
    return [g1.T @ x, g1.sum(0), gW2, gb2, gW3, gb3]


def policy_forward(s, pi):
    """PolicyNetwork.forward (:310-317): head = [mean | log_std_raw]; log_std clamped."""
    W3 = np.concatenate([pi[4], pi[6]], 0)
    b3 = np.concatenate([pi[5], pi[7]], 0)
    head, cache = mlp_forward(s, pi[0], pi[1], pi[2], pi[3], W3, b3)
    A = np.asarray(pi[4]).shape[0]
    mean, raw = head[:, :A], head[:, A:]
    return mean, np.clip(raw, LOG_STD_MIN, LOG_STD_MAX), raw, cache, W3


def policy_evaluate(mean, log_std, eps, action_scale, epsilon=1e-6):
    """PolicyNetwork.evaluate (:319-338) given the normal draws: returns (action, logp [B], z, tanh-mean)."""
    mean, log_std, eps = np.asarray(mean, F64), np.asarray(log_std, F64), np.asarray(eps, F64)
    A = mean.shape[1]
    std = np.exp(log_std)
    if A == 1:                                   # Normal(mean, std)
        z = mean + std * eps
        lp = (-(z - mean) ** 2 / (2 * std * std) - log_std - 0.5 * math.log(2 * math.pi)).sum(1)
    else:                                        # MultivariateNormal(mean, covariance=diag_embed(std)) (:346-351)
        z = mean + np.sqrt(std) * eps
        lp = (-0.5 * (z - mean) ** 2 / std - 0.5 * log_std).sum(1) - 0.5 * A * math.log(2 * math.pi)
    act = np.tanh(z)
    lp = lp - np.log(1 - act * act + epsilon).sum(1)
    return act * action_scale, lp, z, np.tanh(mean) * action_scale


class Adam:
    """torch.optim.Adam over a list of tensors (oracle_np.adam_step_torch per tensor)."""

    def __init__(self, params, lr):
        self.lr, self.t = float(lr), 0
        self.m = [np.zeros_like(np.asarray(p, F64)) for p in params]
        self.v = [np.zeros_like(np.asarray(p, F64)) for p in params]

    def step(self, params, grads):
        self.t += 1
        out = []
        for i, (p, g) in enumerate(zip(params, grads)):
            p2, self.m[i], self.v[i] = onp.adam_step_torch(np.asarray(p, F64), np.asarray(g, F64).reshape(np.shape(p)),
                                                           self.m[i], self.v[i], self.t, self.lr)
            out.append(p2)
        return out


class KLAgent:
    """State of one ForwardKLNetwork / ReverseKLNetwork: the four parameter lists and three Adam states."""

    def __init__(self, kind, q, v, tv, pi, grid_a, grid_w, action_scale, entropy_scale, pi_lr, qf_vf_lr, tau,
                 optim_type="intg", q_update_type="non_sac"):
        assert kind in ("fkl", "rkl")
        f = lambda ps: [np.asarray(p, F64).copy() for p in ps]
        self.kind, self.q, self.v, self.tv, self.pi = kind, f(q), f(v), f(tv), f(pi)
        self.grid_a, self.grid_w = np.asarray(grid_a, F64), np.asarray(grid_w, F64)
        self.scale, self.alpha, self.tau = float(action_scale), float(entropy_scale), float(tau)
        self.optim_type, self.q_update_type = str(optim_type), str(q_update_type)
        self.q_opt, self.v_opt, self.pi_opt = Adam(self.q, qf_vf_lr), Adam(self.v, qf_vf_lr), Adam(self.pi, pi_lr)

    def update(self, s, a, s2, r, g, eps):
        """update_network (:123-209) followed by update_target_network (:211-215).
        Returns (q_loss, v_loss, pi_loss)."""
        s, a, s2 = np.asarray(s, F64), np.asarray(a, F64), np.asarray(s2, F64)
        r, g = np.asarray(r, F64).reshape(-1), np.asarray(g, F64).reshape(-1)
        B, A = a.shape
        alpha = self.alpha
        v_val, v_cache = mlp_forward(s, *self.v)
        v_val = v_val.reshape(-1)
        mean, log_std, raw, pi_cache, piW3 = policy_forward(s, self.pi)
        new_action, logp, z, _ = policy_evaluate(mean, log_std, eps, self.scale)
        v_next = mlp_forward(s2, *self.tv)[0].reshape(-1)
        y_q = r + g * v_next
        q_loss, q_grads = onp.tin_mse_grads(s, a, y_q, self.q, dtype=F64)
        new_q = onp.tin_forward(s, new_action, *self.q, dtype=F64).reshape(-1)
        if self.q_update_type == "sac":
            target_v = new_q - alpha * logp
        else:
            target_v = (r - alpha * logp) + g * v_next
        dv = v_val - target_v
        v_loss = np.mean(dv * dv)
        v_grads = mlp_grads(v_cache, (2.0 / B * dv)[:, None], self.v[2], self.v[4])
        # ---- policy loss and its gradient wrt forward()'s outputs
        if self.optim_type in ("intg", "hard_intg"):
            q_grid = onp.tin_eval(s, self.grid_a, self.q, dtype=F64)                    # [B,N], detached
            if self.kind == "fkl":
                loss_b, dmean, dls, _ = onp.fkl_policy_reduce(q_grid, self.grid_w, self.grid_a, mean, log_std,
                                                              self.scale, alpha)
            else:
                loss_b, dmean, dls, _ = onp.rkl_policy_reduce(q_grid, v_val, self.grid_w, self.grid_a, mean, log_std,
                                                              self.scale, alpha, hard=self.optim_type == "hard_intg")
            pi_loss = float(np.mean(loss_b))
        else:                                                                           # 'll' / 'hard_ll' (:161-169)
            c = (new_q - v_val - alpha * logp) if self.optim_type == "ll" else (new_q - v_val)
            pi_loss = float(np.mean(-logp * c))
            coef = (-c / B)[:, None]
            std = np.exp(log_std)
            t = z - mean                                                                # z is a detached sample
            if A == 1:
                dmean, dls = coef * t / (std * std), coef * (t * t / (std * std) - 1)
            else:
                dmean, dls = coef * t / std, coef * (0.5 * t * t / std - 0.5)
        dls = dls * ((raw >= LOG_STD_MIN) & (raw <= LOG_STD_MAX))                       # torch.clamp backward
        g_pi = mlp_grads(pi_cache, np.concatenate([dmean, dls], 1), self.pi[2], piW3)
        pi_grads = g_pi[:4] + [g_pi[4][:A], g_pi[5][:A], g_pi[4][A:], g_pi[5][A:]]
        # ---- three optimiser steps (:199-209), then the Polyak step of the target V net
        self.q = self.q_opt.step(self.q, q_grads)
        self.v = self.v_opt.step(self.v, v_grads)
        self.pi = self.pi_opt.step(self.pi, pi_grads)
        self.tv = [onp.soft_update(t_, p_, self.tau) for t_, p_ in zip(self.tv, self.v)]
        return float(q_loss), float(v_loss), pi_loss

    def sample_action(self, s, eps):
        mean, log_std, _, _, _ = policy_forward(s, self.pi)
        return policy_evaluate(mean, log_std, eps, self.scale)[0]

    def predict_action(self, s):
        mean, log_std, _, _, _ = policy_forward(s, self.pi)
        return np.tanh(mean) * self.scale
