"""Drop-in ``ForwardKLNetwork`` / ``ReverseKLNetwork`` (agents/network/forwardkl_network.py:17-235,
reversekl_network.py:17-245): same constructor ``(sess, input_norm, config)``, same config keys and the
same methods the managers call (``agents/ForwardKL.py``, ``agents/ReverseKL.py:31-90``):
``update_network(s, a, s', r, gamma)``, ``update_target_network()``, ``sample_action(s)``,
``predict_action(s)``, ``getQFunction(state)``, ``getPolicyFunction(state)``; numpy in, numpy out.

One ``update_network`` is ONE CUDA-graph launch: the minibatch goes up in one pinned copy, then (in the
reference's order -- every forward pass sees the pre-update parameters, :131-194)

    V(s), V_targ(s'), policy head(s)                      rlc_mlp_forward        (B rows)
    PolicyNetwork.evaluate with the step's N(0,1) draws   rlc_policy_evaluate
    Q(s, a_new)                                           rlc_critic_eval        (B rows)
    y_q, dV                                               rlc_kl_targets
    Q on the B x N integration grid                       rlc_critic_eval        <- the hot path
    FKL / RKL reduction with get_logprob fused            rlc_reduce_{fkl,rkl}_policy
    dLoss/d(policy head)                                  rlc_policy_head_grad
    three backward passes + three Adam steps              rlc_critic_grads / rlc_mlp_grads / rlc_adam_step_dev

and the three losses come back in one pinned copy.  Nothing runs on the CPU except drawing the normal
noise from torch's global CPU generator -- the stream ``normal.sample()`` consumes in the reference, so
``torch.manual_seed`` reproduces the reference's trajectory (pass ``eps=`` to supply the draws).
Initial weights are drawn exactly as the reference's constructors draw them (same modules, same order)."""
from __future__ import annotations

import numpy as np
import torch

import ctypes as C

from ._lib import (ADAM_TORCH, LAYOUT_OUT_IN, SB_MAX_B, SB_ROLE_PI, SB_ROLE_Q, SB_ROLE_V, TIN, RlcSbNet, RlcSbTrain,
                   check)
from .engine import Critic, Engine, Mlp, _f32, _ptr, _stream
from .quadrature import integration_grid

LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0          # PolicyNetwork defaults, forwardkl_network.py:294


def _flag(x):
    return x is True or x == "True"


def _linear_init(n_in, n_out, init_w=None):
    """The tensors ``nn.Linear(n_in, n_out)`` would hold, drawn from torch's global generator in the
    order the reference's constructors draw them (weight, bias; then the optional ``uniform_`` re-draws,
    forwardkl_network.py:258-259,302-308)."""
    lin = torch.nn.Linear(n_in, n_out)
    if init_w is not None:
        lin.weight.data.uniform_(-init_w, init_w)
        lin.bias.data.uniform_(-init_w, init_w)
    return lin.weight.detach().clone(), lin.bias.detach().clone()


class _Adam:
    """torch.optim.Adam state of one flat parameter vector, step count on the device (graph-safe)."""

    def __init__(self, eng: Engine, theta: torch.Tensor, lr: float, grad: torch.Tensor):
        self.eng, self.theta, self.lr = eng, theta, float(lr)
        self.m, self.v = torch.zeros_like(theta), torch.zeros_like(theta)
        self.state_dev = torch.zeros((4,), dtype=torch.int32, device=theta.device)
        self.grad = grad                               # a view into the agent's flat gradient buffer

    def step(self):
        self.eng.adam_step_dev(self.theta, self.grad, self.m, self.v, self.state_dev, self.lr, ADAM_TORCH)


class _QNet:
    """``self.q_net`` of the reference object: callable on stacked rows -> [R,1] (numpy or tensors in,
    device tensor out)."""

    def __init__(self, critic: Critic):
        self.critic = critic

    def __call__(self, state, action):
        dev = self.critic.eng.device
        s, a = _f32(state, dev), _f32(action, dev)
        return self.critic.eval(s, a[:, None, :], "fp32").reshape(-1, 1)

    forward = __call__

    def eval_grid(self, state_batch, grid_actions, precision="auto"):
        return self.critic.eval(state_batch, grid_actions, precision)


class _KLNetwork(object):
    KIND = None                                   # "fkl" | "rkl"
    OPTIM_TYPES = ()

    def __init__(self, sess, input_norm, config):
        self.sess, self.input_norm, self.config = sess, input_norm, config
        self.state_dim, self.action_dim = int(config.state_dim), int(config.action_dim)
        self.state_min, self.state_max = config.state_min, config.state_max
        self.action_min, self.action_max = config.action_min, config.action_max
        self.learning_rate = [float(config.pi_lr), float(config.qf_vf_lr)]
        self.tau, self.norm_type = float(config.tau), getattr(config, "norm_type", "none")
        self.optim_type = config.optim_type
        self.q_update_type = config.q_update_type
        self.use_true_q = _flag(getattr(config, "use_true_q", "False"))
        self.rng = np.random.RandomState(config.random_seed)
        self.entropy_scale = float(config.entropy_scale)
        if self.q_update_type not in ("sac", "non_sac"):
            raise ValueError("invalid config.q_update_type")
        if self.optim_type not in self.OPTIM_TYPES:
            # forwardkl_network.py:152-153 raises for 'll'; reversekl 'reparam' leaves policy_loss undefined (:171-173)
            raise NotImplementedError("optim_type %r" % (self.optim_type,))
        S, A = self.state_dim, self.action_dim
        self.action_scale = float(np.asarray(self.action_max, np.float64).reshape(-1)[0])
        a1, a2 = int(config.actor_l1_dim), int(config.actor_l2_dim)
        c1, c2 = int(config.critic_l1_dim), int(config.critic_l2_dim)
        eng = getattr(config, "engine", None)
        self.eng = eng if eng is not None else Engine()
        # one handle = one scratch arena and one operand-pack cache: an agent that captures CUDA graphs on it must own it
        # (a second agent's kernels would write the same scratch from another stream while the first one's graph replays)
        # Sequential reuse is safe (outgrown scratch blocks are retired, not freed: csrc/api.cu), concurrent use is not.
        import weakref
        prev = getattr(self.eng, "_graph_owner", None)
        if prev is not None and prev() is not None:
            import warnings
            warnings.warn("config.engine is shared with another live ForwardKLNetwork / ReverseKLNetwork: their updates "
                          "must not overlap in time (one scratch arena per handle); give concurrently running agents "
                          "their own rlcontrol_b200.Engine", RuntimeWarning, stacklevel=3)
        self.eng._graph_owner = weakref.ref(self)
        # 'auto' = parity-preserving: the split tensor mode (fp16 hi+lo operands, ~1e-5 of the fp32 reference) for large
        # shared-grid evaluations, the fp32 CUDA-core path otherwise; the single-rounding 'fp16'/'bf16' modes carry up to
        # 5e-3 of Q, which exp(q / entropy_scale) amplifies at the small entropy scales the reference sweeps (0.01, 0.001),
        # so they are opt-in only
        self.precision = getattr(config, "precision", "auto")
        dev = self.eng.device
        # The update's branches (Q regression | grid evaluation + policy | V) overlap on separate streams inside the
        # captured graph; a handle's scratch is per handle, so each branch gets its own (cheap: scratch only).
        self.eng_grid, self.eng_v, self.eng_pi = (Engine(self.eng.device_index) for _ in range(3))
        # ---- parameters, drawn in the reference's order: pi_net, q_net, v_net, target_v_net (:41-45)
        pW1, pb1 = _linear_init(S, a1)
        pW2, pb2 = _linear_init(a1, a2)
        mW, mb = _linear_init(a2, A, 3e-3)
        sW, sb = _linear_init(a2, A, 3e-3)
        qW1, qb1 = _linear_init(S + A, c1)
        qW2, qb2 = _linear_init(c1, c2)
        qW3, qb3 = _linear_init(c2, 1, 3e-3)
        vW1, vb1 = _linear_init(S, c1)
        vW2, vb2 = _linear_init(c1, c2)
        vW3, vb3 = _linear_init(c2, 1, 3e-3)
        for n_in, n_out, iw in ((S, c1, None), (c1, c2, None), (c2, 1, 3e-3)):     # target_v_net's own draws, then
            _linear_init(n_in, n_out, iw)                                          # overwritten by the copy (:47-49)
        self.pi = Mlp(self.eng_pi, S, a1, a2, 2 * A).load_torch(pW1, pb1, pW2, pb2, [mW, sW], [mb, sb])
        self.critic = Critic(self.eng, TIN, S, A, c1, c2).load(qW1, qb1, qW2, qb2, qW3, qb3, LAYOUT_OUT_IN)
        self.critic_grid = Critic(self.eng_grid, TIN, S, A, c1, c2)       # same theta, the grid branch's handle
        self.critic_grid.theta = self.critic.theta
        self.critic_grid._refresh()
        self.v = Mlp(self.eng_v, S, c1, c2, 1).load_torch(vW1, vb1, vW2, vb2, vW3, vb3)
        self.target_v = Mlp(self.eng_v, S, c1, c2, 1)
        self.target_v.copy_from(self.v)
        self.q_net = _QNet(self.critic)
        # one contiguous gradient buffer [theta_Q | theta_V | theta_pi]: the data-parallel exchange is ONE all-reduce
        nq, nv, npi = self.critic.theta.numel(), self.v.theta.numel(), self.pi.theta.numel()
        self.grad_flat = torch.zeros((nq + nv + npi,), dtype=torch.float32, device=dev)
        self.pi_opt = _Adam(self.eng_pi, self.pi.theta, self.learning_rate[0], self.grad_flat[nq + nv:])
        self.q_opt = _Adam(self.eng, self.critic.theta, self.learning_rate[1], self.grad_flat[:nq])
        self.v_opt = _Adam(self.eng_v, self.v.theta, self.learning_rate[1], self.grad_flat[nq:nq + nv])
        # data-parallel: every rank updates on its own shard of the minibatch; all per-rank
        # gradients are pre-scaled by 1/B_total inside the kernels, so one SUM all-reduce of grad_flat gives the
        # reference's batch means and every rank then applies identical Adam steps.
        self.process_group = getattr(config, "process_group", None)
        self.world_size = int(getattr(config, "world_size", 1))
        # global minibatch size of a data-parallel update (default: equal shards, world_size x local batch); with
        # parallel.shard_bounds' uneven shards (B % world != 0) the caller passes the true total
        self.global_batch = getattr(config, "global_batch_size", None)
        # ---- integration grid (:58-102)
        grid = getattr(config, "integration_grid", None)      # optional (actions [N,A], weights [N]) override
        acts, w = grid if grid is not None else integration_grid(A, self.action_scale, getattr(config, "N_param", 64),
                                                                 getattr(config, "l_param", 6))
        acts, w = np.asarray(acts, np.float32).reshape(-1, A), np.asarray(w, np.float32).reshape(-1)
        self.intgrl_actions, self.intgrl_weights = _f32(acts, dev), _f32(w, dev)
        self.intgrl_actions_len = int(acts.shape[0])
        self.device = dev
        self._steps = {}                      # batch size -> captured update
        self._act = {}                        # batch size -> action-selection buffers
        # The data-parallel update (world_size > 1) is TWO graphs around an eager all-reduce: everything up to the joined
        # gradients, then torch's NCCL all-reduce of [g_Q | g_V | g_pi] and of the loss shares on the update's stream,
        # then the three Adam steps and the download.  (One graph with the collective inside -- four stream branches +
        # NCCL in one capture -- hung on 2 GPUs in the warm-up/capture sequence; config.dp_cuda_graph=False: all eager.)
        self.use_graph = bool(getattr(config, "use_cuda_graph", True)) and (
            self.world_size == 1 or bool(getattr(config, "dp_cuda_graph", True)))
        # small minibatches (cfg1 / cfg5): all B-row forward passes in one launch, all backward passes + Adam in one
        # launch (csrc/small_batch.cu) instead of ~60 dependent small kernels
        self.fused_small = bool(getattr(config, "fused_small_batch", True))

    # ------------------------------------------------------------------ parameter access (tests, checkpoints)
    def load_reference_parameters(self, q, v, target_v, pi):
        """Lists in torch layout: q/v/target_v = [W1,b1,W2,b2,W3,b3]; pi = [W1,b1,W2,b2,Wm,bm,Ws,bs]."""
        self.critic.load(*q, LAYOUT_OUT_IN)
        self.critic_grid.invalidate()
        self.v.load_torch(*v)
        self.target_v.load_torch(*target_v)
        self.pi.load_torch(pi[0], pi[1], pi[2], pi[3], [pi[4], pi[6]], [pi[5], pi[7]])
        torch.cuda.synchronize(self.device)

    def export_parameters(self):
        """dict(q, v, tv, pi) of numpy lists in the same torch layouts."""
        A = self.action_dim
        n = lambda ts: [t.detach().cpu().numpy() for t in ts]
        p = self.pi.export_torch()
        pi = [p[0], p[1], p[2], p[3], p[4][:A], p[5][:A], p[4][A:], p[5][A:]]
        return dict(q=n(self.critic.export(LAYOUT_OUT_IN)), v=n(self.v.export_torch()),
                    tv=n(self.target_v.export_torch()), pi=n(pi))

    # ------------------------------------------------------------------ the update
    def _build_step(self, B):
        dev, S, A, N = self.device, self.state_dim, self.action_dim, self.intgrl_actions_len
        st = type("Step", (), {})()
        n_in = B * (2 * S + 2 * A + 2)
        st.in_host = torch.zeros((n_in,), dtype=torch.float32).pin_memory()
        st.in_dev = torch.zeros((n_in,), dtype=torch.float32, device=dev)

        def views(flat):
            out, off = {}, 0
            for k, sh in (("s", (B, S)), ("a", (B, A)), ("s2", (B, S)), ("r", (B,)), ("g", (B,)), ("eps", (B, A))):
                n = int(np.prod(sh))
                out[k] = flat[off:off + n].view(sh)
                off += n
            return out
        st.h, st.d = views(st.in_host), views(st.in_dev)
        st.h_np = {k: v.numpy() for k, v in st.h.items()}      # numpy views of the pinned staging buffer
        st.out_host = torch.zeros((4,), dtype=torch.float32).pin_memory()       # q_loss, v_loss, pi_loss, -
        st.out_dev = torch.zeros((4,), dtype=torch.float32, device=dev)
        f = lambda *sh: torch.zeros(sh, dtype=torch.float32, device=dev)
        st.v_out, st.vnext, st.head = f(B, 1), f(B, 1), f(B, 2 * A)
        st.act_v, st.act_pi = self.v.act_buffer(B), self.pi.act_buffer(B)
        st.ev = dict(action=f(B, A), logp=f(B), mean=f(B, A), mu_raw=f(B, A), log_std=f(B, A), z=f(B, A))
        st.q_new, st.y, st.dv = f(B, 1), f(B), f(B)
        st.q_grid = f(B, N)
        st.loss_b, st.dmean, st.dls, st.dhead = f(B), f(B, A), f(B, A), f(B, 2 * A)
        st.q_reg = f(B)
        st.stream = torch.cuda.Stream(device=dev)
        st.s_grid, st.s_v, st.s_pi = (torch.cuda.Stream(device=dev) for _ in range(3))
        st.graph = None
        return st

    def _small_ok(self, B):
        dims = (self.pi.H1, self.pi.H2, self.critic.H1, self.critic.H2)
        return (self.fused_small and self.world_size == 1 and B <= SB_MAX_B and self.optim_type in ("intg", "hard_intg")
                and max(dims) <= 400 and self.state_dim + self.action_dim <= 256 and 2 * self.action_dim <= 32)

    def _build_small(self, st, B):
        """Descriptors of the two fused launches (kept alive on ``st``; the C side copies them at launch)."""
        dev, S, A, d = self.device, self.state_dim, self.action_dim, st.d
        f = lambda *sh: torch.zeros(sh, dtype=torch.float32, device=dev)
        st.sb_buf = dict(h1v=f(B, self.v.H1), h2v=f(B, self.v.H2), h1p=f(B, self.pi.H1), h2p=f(B, self.pi.H2),
                         h1q=f(B, self.critic.H1), h2q=f(B, self.critic.H2), w3v=f(self.v.H2), w3p=f(self.pi.H2 * 2 * A),
                         w3q=f(self.critic.H2), h1t=f(B, self.v.H1), h2t=f(B, self.v.H2))
        bf = st.sb_buf
        p = lambda t: None if t is None else t.data_ptr()
        sac = self.q_update_type == "sac"
        fwd = (RlcSbNet * 5)()

        def net(n, theta, inp, H1, H2, O, x0, n0, x1, n1, h1, h2, out, w3, opt):
            n.theta, n.inp, n.H1, n.H2, n.O = p(theta), inp, H1, H2, O
            n.x0, n.n0, n.x1, n.n1 = p(x0), n0, p(x1), n1
            n.h1, n.h2, n.out, n.w3_snapshot = p(h1), p(h2), p(out), p(w3)
            if opt is not None:
                n.adam_state, n.lr, n.beta1, n.beta2, n.adam_variant = p(opt.state_dev), opt.lr, 0.9, 0.999, ADAM_TORCH
        net(fwd[0], self.v.theta, S, self.v.H1, self.v.H2, 1, d["s"], S, None, 0, bf["h1v"], bf["h2v"], st.v_out, bf["w3v"],
            self.v_opt)
        net(fwd[1], self.target_v.theta, S, self.v.H1, self.v.H2, 1, d["s2"], S, None, 0, None, None, st.vnext, None, None)
        net(fwd[2], self.pi.theta, S, self.pi.H1, self.pi.H2, 2 * A, d["s"], S, None, 0, bf["h1p"], bf["h2p"], st.head,
            bf["w3p"], self.pi_opt)
        n = fwd[2]
        n.policy, n.eps, n.action_scale, n.log_std_min, n.log_std_max = 1, p(d["eps"]), self.action_scale, LOG_STD_MIN, LOG_STD_MAX
        ev = st.ev
        n.action, n.logp, n.mean, n.mu_raw, n.log_std, n.z = (p(ev[k]) for k in ("action", "logp", "mean", "mu_raw", "log_std", "z"))
        net(fwd[3], self.critic.theta, S + A, self.critic.H1, self.critic.H2, 1, d["s"], S, d["a"], A, bf["h1q"], bf["h2q"],
            st.q_reg, bf["w3q"], self.q_opt)
        # the B x N grid evaluation with the PRE-update theta_Q rides in the same launch when it is small (cfg1: 1 984 rows);
        # larger stacks go to rlc_critic_eval (tensor path) on the side stream
        N = self.intgrl_actions_len
        st.sb_grid_fused = B * N <= 4096
        net(fwd[4], self.critic.theta, S + A, self.critic.H1, self.critic.H2, 1, d["s"], S, self.intgrl_actions, A, None, None,
            st.q_grid, None, None)
        fwd[4].rows, fwd[4].x0_div, fwd[4].x1_mod = B * N, N, N
        st.sb_fwd = fwd
        fwd2 = (RlcSbNet * 1)()                       # 'sac': Q(s, a_new) needs the policy's sample first (:143-146)
        net(fwd2[0], self.critic.theta, S + A, self.critic.H1, self.critic.H2, 1, d["s"], S, ev["action"], A, None, None,
            st.q_new, None, None)
        st.sb_fwd2 = fwd2
        upd = (RlcSbTrain * 3)()

        def tr(n, role, theta, opt, inp, H1, H2, O, x0, n0, x1, n1, h1, h2, out, w3, loss):
            n.theta, n.m, n.v, n.adam_state = p(theta), p(opt.m), p(opt.v), p(opt.state_dev)
            n.beta1, n.beta2, n.eps = 0.9, 0.999, 1e-8
            n.inp, n.H1, n.H2, n.O, n.x0, n.n0, n.x1, n.n1 = inp, H1, H2, O, p(x0), n0, p(x1), n1
            n.h1, n.h2, n.out, n.w3_snapshot, n.role, n.loss_out = p(h1), p(h2), p(out), p(w3), role, p(loss)
            n.r, n.gamma, n.v_next, n.logp, n.q_new = p(d["r"]), p(d["g"]), p(st.vnext), p(ev["logp"]), p(st.q_new)
            n.entropy_scale, n.sac = self.entropy_scale, int(sac)
            n.log_std_min, n.log_std_max = LOG_STD_MIN, LOG_STD_MAX
        o = st.out_dev
        tr(upd[0], SB_ROLE_V, self.v.theta, self.v_opt, S, self.v.H1, self.v.H2, 1, d["s"], S, None, 0, bf["h1v"], bf["h2v"],
           st.v_out, bf["w3v"], o[1:2])
        tr(upd[1], SB_ROLE_Q, self.critic.theta, self.q_opt, S + A, self.critic.H1, self.critic.H2, 1, d["s"], S, d["a"], A,
           bf["h1q"], bf["h2q"], st.q_reg, bf["w3q"], o[0:1])
        tr(upd[2], SB_ROLE_PI, self.pi.theta, self.pi_opt, S, self.pi.H1, self.pi.H2, 2 * A, d["s"], S, None, 0, bf["h1p"],
           bf["h2p"], st.head, bf["w3p"], o[2:3])
        upd[2].dmean, upd[2].dlog_std, upd[2].loss_b = p(st.dmean), p(st.dls), p(st.loss_b)
        st.sb_upd = upd

    def _enqueue_small(self, st, B, device_inputs=False):
        """The same update as :meth:`_enqueue` in four launches: all forward passes, the B x N grid evaluation among them
        -> FKL/RKL reduction with get_logprob fused (2) -> all backward passes + Adam."""
        if getattr(st, "sb_fwd", None) is None:
            self._build_small(st, B)
        lib, alpha = self.eng.lib, self.entropy_scale
        main, s_grid = st.stream, st.s_grid
        if not device_inputs:
            st.in_dev.copy_(st.in_host, non_blocking=True)
        if st.sb_grid_fused:
            check(lib.rlc_sb_forward(self.eng.h, st.sb_fwd, 5, B, _stream()))
        else:
            s_grid.wait_stream(main)
            with torch.cuda.stream(s_grid):
                self.critic_grid.eval_into(st.d["s"], self.intgrl_actions, st.q_grid, self.precision)   # pre-update theta_Q
            check(lib.rlc_sb_forward(self.eng.h, st.sb_fwd, 4, B, _stream()))
        if self.q_update_type == "sac":
            check(lib.rlc_sb_forward(self.eng.h, st.sb_fwd2, 1, B, _stream()))
        if not st.sb_grid_fused:
            main.wait_stream(s_grid)
        if self.KIND == "fkl":
            self.eng.fkl_policy(st.q_grid, self.intgrl_weights, self.intgrl_actions, self.action_scale, st.ev["mu_raw"],
                                st.ev["log_std"], alpha, b_total=B, out=(st.loss_b, st.dmean, st.dls))
        else:
            self.eng.rkl_policy(st.q_grid, st.v_out.view(-1), self.intgrl_weights, self.intgrl_actions, self.action_scale,
                                st.ev["mu_raw"], st.ev["log_std"], alpha, hard=self.optim_type == "hard_intg", b_total=B,
                                out=(st.loss_b, st.dmean, st.dls))
        check(lib.rlc_sb_update(self.eng.h, st.sb_upd, 3, B, B, _stream()))
        self.critic.invalidate()
        self.critic_grid.invalidate()
        if not device_inputs:
            st.out_host.copy_(st.out_dev, non_blocking=True)

    def _enqueue_dp_post(self, st, device_inputs=False):
        """Data-parallel update, after the all-reduce: identical Adam steps on every rank, then the download."""
        self.q_opt.step()
        self.v_opt.step()
        self.pi_opt.step()
        self.critic.invalidate()
        self.critic_grid.invalidate()
        if not device_inputs:
            st.out_host.copy_(st.out_dev, non_blocking=True)

    def _allreduce_dp(self, st):
        from .parallel import allreduce_grad_
        allreduce_grad_(self.grad_flat, self.process_group)
        allreduce_grad_(st.out_dev, self.process_group)     # every loss slot holds this rank's share of the global mean

    def _enqueue(self, st, B, device_inputs=False, stop_before_allreduce=False):
        """One update on four streams (fork after the upload, join before the download); the dependency edges
        are exactly the data dependencies of update_network, so the captured graph runs the independent
        branches side by side: small batches are latency-bound, and this is what shortens the critical path.

            main  : H2D -> [Q(s,a_new)] -> targets -> Q regression grads -> (grid done) Adam(theta_Q) -> D2H
            s_pi  : policy head(s) -> evaluate
            s_v   : V(s), V_targ(s')          ... -> V grads -> Adam(theta_V)
            s_grid: Q on the B x N grid with the PRE-update theta_Q -> reduction -> head grad -> pi grads -> Adam(theta_pi)
        """
        if self._small_ok(B):
            return self._enqueue_small(st, B, device_inputs)
        eng, d, o = self.eng, st.d, st.out_dev
        alpha, sac = self.entropy_scale, self.q_update_type == "sac"
        intg = self.optim_type in ("intg", "hard_intg")
        dp = self.world_size > 1
        bt = int(self.global_batch) if (dp and self.global_batch) else B * self.world_size
        need_q_new = sac or not intg
        main, s_grid, s_v, s_pi = st.stream, st.s_grid, st.s_v, st.s_pi
        if not device_inputs:                  # device_loop.py fills st.d[...] on the device (replay gather, staged draws)
            st.in_dev.copy_(st.in_host, non_blocking=True)
        for br in (s_grid, s_v, s_pi):
            br.wait_stream(main)
        # forward passes, all on pre-update parameters (:131-137)
        with torch.cuda.stream(s_pi):
            self.pi.forward(d["s"], out=st.head, act=st.act_pi)
            self.eng_pi.policy_evaluate(st.head, d["eps"], self.action_scale, LOG_STD_MIN, LOG_STD_MAX, out=st.ev)
        with torch.cuda.stream(s_v):
            self.v.forward(d["s"], out=st.v_out, act=st.act_v)
            self.target_v.forward(d["s2"], out=st.vnext)
        if intg:
            with torch.cuda.stream(s_grid):
                self.critic_grid.eval_into(d["s"], self.intgrl_actions, st.q_grid, self.precision)
                ev_grid = torch.cuda.Event()
                ev_grid.record(s_grid)
        # main: regression targets (:137-150)
        main.wait_stream(s_pi)
        if need_q_new:
            self.critic.eval_into(d["s"], st.ev["action"].view(B, 1, -1), st.q_new, "fp32")
        main.wait_stream(s_v)
        eng.kl_targets(d["r"], d["g"], st.vnext, st.q_new, st.ev["logp"], st.v_out, alpha, sac, b_total=bt,
                       out=(st.y, st.dv, o[1:2]))
        # V branch: backward + Adam
        s_v.wait_stream(main)
        with torch.cuda.stream(s_v):
            self.v.grads(d["s"], st.dv.view(B, 1), act=st.act_v, grad_out=self.v_opt.grad)
            if not dp:
                self.v_opt.step()
        # policy branch (:155-194 / reversekl :175-203), on the grid stream
        s_grid.wait_stream(main)               # covers s_pi and s_v (evaluate, V(s)) and, for 'll', Q(s,a_new)
        with torch.cuda.stream(s_grid):
            if intg:
                if self.KIND == "fkl":
                    self.eng_grid.fkl_policy(st.q_grid, self.intgrl_weights, self.intgrl_actions, self.action_scale,
                                             st.ev["mu_raw"], st.ev["log_std"], alpha, b_total=bt,
                                             out=(st.loss_b, st.dmean, st.dls))
                else:
                    self.eng_grid.rkl_policy(st.q_grid, st.v_out.view(-1), self.intgrl_weights, self.intgrl_actions,
                                             self.action_scale, st.ev["mu_raw"], st.ev["log_std"], alpha,
                                             hard=self.optim_type == "hard_intg", b_total=bt,
                                             out=(st.loss_b, st.dmean, st.dls))
                self.eng_grid.policy_head_grad(st.head, 0, dmean=st.dmean, dlog_std=st.dls, out=st.dhead)
                self.eng_grid.mean_into(st.loss_b, o[2:3])
                if dp:
                    o[2:3].mul_(float(B) / float(bt))      # this rank's share of the global mean over states
            else:
                self.eng_grid.policy_head_grad(st.head, 1 if self.optim_type == "ll" else 2, z=st.ev["z"],
                                               logp=st.ev["logp"], q_new=st.q_new, v=st.v_out, entropy_scale=alpha,
                                               b_total=bt, out=st.dhead, loss_out=o[2:3])
            self.pi.grads(d["s"], st.dhead, act=st.act_pi, grad_out=self.pi_opt.grad)   # eng_pi's scratch: s_pi is idle
            if not dp:
                self.pi_opt.step()
        # main: Q regression (:133-140,199-201); theta_Q may only change once the grid branch has read it
        self.critic.grads_into(d["s"], d["a"], st.y, self.q_opt.grad, o[0:1], st.q_reg, b_total=bt)
        if dp:
            # join, ONE all-reduce over [g_Q | g_V | g_pi] (and the three loss shares), identical Adam steps everywhere
            for br in (s_grid, s_v, s_pi):
                main.wait_stream(br)
            if stop_before_allreduce:          # first of the two captured graphs
                return
            self._allreduce_dp(st)
            self._enqueue_dp_post(st, device_inputs)
            return
        else:
            if intg:
                main.wait_event(ev_grid)
            self.q_opt.step()
        self.critic.invalidate()
        self.critic_grid.invalidate()
        for br in (s_grid, s_v, s_pi):
            main.wait_stream(br)
        if not device_inputs:
            st.out_host.copy_(o, non_blocking=True)

    def _snapshot(self):
        ts = [self.critic.theta, self.v.theta, self.pi.theta]
        for o in (self.q_opt, self.v_opt, self.pi_opt):
            ts += [o.m, o.v, o.state_dev]
        return ts, [t.clone() for t in ts]

    def _step_for(self, B):
        st = self._steps.get(B)
        if st is None:
            # a new batch size may grow a handle's scratch, which would leave older captured graphs with stale
            # pointers: keep one captured update (the reference's batch size is fixed, config.batch_size)
            self._steps.clear()
            st = self._build_step(B)
            if self.use_graph:
                # warm-up (workspace growth, function attributes) must not advance the agent: snapshot/restore
                ts, saved = self._snapshot()
                with torch.cuda.stream(st.stream):
                    self._enqueue(st, B)
                st.stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st.stream):
                    self._enqueue(st, B, stop_before_allreduce=self.world_size > 1)
                st.graph = g
                st.graph_post = None
                if self.world_size > 1:
                    g2 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g2, stream=st.stream):
                        self._enqueue_dp_post(st)
                    st.graph_post = g2
                for t, s_ in zip(ts, saved):
                    t.copy_(s_)
                self.critic.invalidate()
                self.critic_grid.invalidate()
                torch.cuda.synchronize(self.device)
            self._steps[B] = st
        return st

    def update_network_async(self, state_batch, action_batch, next_state_batch, reward_batch, gamma_batch, eps=None):
        """Stage the minibatch and launch the update without waiting for it (independent agents of a sweep
        overlap on one GPU this way: launch them all, then :meth:`wait` each)."""
        as_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        state_batch, action_batch, next_state_batch, reward_batch, gamma_batch = (
            as_np(x) for x in (state_batch, action_batch, next_state_batch, reward_batch, gamma_batch))
        s = np.asarray(state_batch, np.float32)
        B = s.shape[0]
        st = self._step_for(B)
        h = st.h_np
        if eps is None:
            # the draws behind normal.sample() in pi_net.evaluate (:305-309): loc + scale * N(0,1), global CPU generator
            eps = torch.randn(B, self.action_dim)
        if isinstance(eps, torch.Tensor):
            eps = eps.numpy()
        # float64 -> float32 at the boundary, like torch.FloatTensor(x) in the reference (:125-129)
        for k, x in (("s", s), ("a", action_batch), ("s2", next_state_batch), ("r", reward_batch), ("g", gamma_batch),
                     ("eps", eps)):
            np.copyto(h[k], np.asarray(x).reshape(h[k].shape), casting="same_kind")
        with torch.cuda.stream(st.stream):
            if st.graph is not None:
                st.graph.replay()
                if getattr(st, "graph_post", None) is not None:      # data-parallel: eager collective between the graphs
                    self._allreduce_dp(st)
                    st.graph_post.replay()
                # the replay moves theta_Q on the device; the host-side validity of the cached tensor-core operand packs
                # was only cleared at capture time, so clear it after every replay (an eager q_net.eval_grid(...) between
                # updates would otherwise reuse operands packed from the previous theta)
                self.critic.invalidate()
                self.critic_grid.invalidate()
            else:
                self._enqueue(st, B)
        self._pending = st

    def wait(self):
        """Block until the launched update has finished; returns (q_loss, v_loss, pi_loss)."""
        st = self._pending
        st.stream.synchronize()
        self.last_losses = st.out_host[:3].clone().numpy()
        return self.last_losses

    def update_network(self, state_batch, action_batch, next_state_batch, reward_batch, gamma_batch, eps=None):
        self.update_network_async(state_batch, action_batch, next_state_batch, reward_batch, gamma_batch, eps)
        self.wait()

    def update_target_network(self):
        # target_v <- (1 - tau) target_v + tau v   (:211-215); on the update's own stream so that it is ordered
        # after a launched update and before the next one without any device-wide synchronisation
        st = getattr(self, "_pending", None)
        if st is None:
            self.eng_v.soft_update(self.target_v.theta, self.v.theta, self.tau)
        else:
            with torch.cuda.stream(st.stream):
                self.eng_v.soft_update(self.target_v.theta, self.v.theta, self.tau)

    # ------------------------------------------------------------------ acting
    def _evaluate(self, state_batch, eps):
        dev = self.device
        s = _f32(np.asarray(state_batch, np.float32), dev)
        head = self.pi.forward(s)
        e = None if eps is None else _f32(eps, dev)
        return self.eng_pi.policy_evaluate(head, e, self.action_scale, LOG_STD_MIN, LOG_STD_MAX)

    def sample_action(self, state_batch, eps=None):
        """pi_net.evaluate(state)[0] (:109-113): one tanh-Gaussian sample per state."""
        B = np.asarray(state_batch).shape[0]
        if eps is None:
            eps = torch.randn(B, self.action_dim)
        return self._evaluate(state_batch, eps)["action"].cpu().numpy()

    def predict_action(self, state_batch):
        """tanh(mean) * action_scale (:115-121).  The reference also draws (and discards) a sample here, which
        advances torch's generator: reproduced so that seeded runs stay aligned."""
        torch.randn(np.asarray(state_batch).shape[0], self.action_dim)
        return self._evaluate(state_batch, None)["mean"].cpu().numpy()

    def getQFunction(self, state):
        s = np.asarray(state, np.float32).reshape(1, -1)
        return lambda action: self.q_net(s, np.asarray([action], np.float32).reshape(1, -1)).cpu().numpy()

    def getPolicyFunction(self, state):
        ev = self._evaluate(np.asarray(state, np.float32).reshape(1, -1), None)
        mean, std = ev["mean"].cpu().numpy(), np.exp(ev["log_std"].cpu().numpy())
        return lambda action: 1 / (std * np.sqrt(2 * np.pi)) * np.exp(-(action - mean) ** 2 / (2 * std ** 2))


class ForwardKLNetwork(_KLNetwork):
    """agents/network/forwardkl_network.py:17 -- optim_type 'intg' only ('ll' raises there, :152-153)."""
    KIND = "fkl"
    OPTIM_TYPES = ("intg",)


class ReverseKLNetwork(_KLNetwork):
    """agents/network/reversekl_network.py:17 -- 'intg', 'hard_intg', 'll', 'hard_ll' ('reparam' has no loss)."""
    KIND = "rkl"
    OPTIM_TYPES = ("intg", "hard_intg", "ll", "hard_ll")
