"""Drop-in mirrors of the reference's ``agents/network`` critic entry points.

Same class names, constructor signatures ``(sess, input_norm, config)`` and positional method
signatures as the reference, so a manager file (``agents/QT_OPT.py``, ``agents/ActorExpert*.py``,
``agents/ReverseKL.py`` ...) keeps calling ``predict_q / predict_q_target / train* /
q_action_gradients / q_gradient_ascent / iterate_cem_multidim / init_target_network /
update_target_network / getQFunction`` unchanged; numpy in, numpy out (float64 accepted, cast to
float32 at the boundary, outputs float32 ``[R,1]``), stacked rows state-major.  All arithmetic runs
in librlc.so (``engine.py``); there is no CPU path.

Only the CRITIC side of each reference network lives here -- actors/policies are per-state work
outside the hot path and keep their own implementation.

``config`` is the reference's ``utils/config.Config``: any object with the attributes the
reference reads (``state_dim, state_min, state_max, action_dim, action_min, action_max, tau,
norm_type, random_seed`` and the per-agent keys cited at each class).  Optional new keys:
``precision`` ("auto" | "fp32" | "fp16" | "bf16") and ``engine`` (a shared :class:`Engine`)."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import ADAM_TF, ADAM_TORCH, LAYOUT_IN_OUT, LAYOUT_OUT_IN, TIN, TMID
from .engine import Critic, CriticOptimizer, Engine

_ENGINE = None


def _engine(config) -> Engine:
    global _ENGINE
    eng = getattr(config, "engine", None)
    if eng is not None:
        return eng
    if _ENGINE is None:
        _ENGINE = Engine()
    return _ENGINE


def _np(x):
    """numpy fp32 view of what a manager passes in -- numpy (the reference contract) or a tensor on any device
    (ReplayBuffer.sample_batch(as_numpy=False) keeps minibatches on the GPU)."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x, np.float32)


class _GMM(object):
    """What the managers read from a fitted ``BoundedVarGaussianMixture``
    (qt_opt_network.py:155,180,186-189): ``weights_ / means_ / covariances_`` and ``sample``."""

    def __init__(self, weights, means, covs, rng):
        self.weights_, self.means_, self.covariances_, self._rng = weights, means, covs, rng

    def sample(self, n_samples=1):
        # sklearn GaussianMixture.sample (diag): multinomial component counts, then normals per component
        counts = self._rng.multinomial(n_samples, self.weights_ / self.weights_.sum())
        X = np.vstack([m + self._rng.randn(c, m.size) * np.sqrt(v)
                       for m, v, c in zip(self.means_, self.covariances_, counts)])
        y = np.concatenate([np.full(c, j, dtype=int) for j, c in enumerate(counts)])
        return X, y


class _TMidBase(object):
    """TF T-mid critic ``clip(s) -> FC(S,l1)+ReLU -> FC(l1+A,l2)+ReLU -> FC(l2,1)``
    (critic_network.py:61-99 and the identical bodies listed in SURVEY 8 row a6), online + target
    copy, MSE/Adam(TF) regression, Polyak target update."""

    def _build(self, sess, input_norm, config, l1, l2, lr):
        self.sess, self.input_norm, self.config = sess, input_norm, config
        self.state_dim, self.action_dim = int(config.state_dim), int(config.action_dim)
        self.state_min, self.state_max = np.asarray(config.state_min, np.float64), np.asarray(config.state_max, np.float64)
        self.action_min, self.action_max = np.asarray(config.action_min, np.float64), np.asarray(config.action_max, np.float64)
        self.learning_rate, self.tau = float(lr), float(config.tau)
        self.norm_type = getattr(config, "norm_type", "none")
        if self.norm_type not in ("none", "input_norm"):
            # base_network.py:53-65: 'batch' needs batch-norm statistics (the AE networks raise too, ae_network.py:93-94) and
            # 'layer' inserts tf.contrib.layers.layer_norm(center, scale) after both FC layers -- this critic has neither the
            # parameters nor the op, so accepting it would silently train a different network
            raise NotImplementedError("norm_type %r is not supported on the B200 path" % (self.norm_type,))
        self.l1, self.l2 = int(l1), int(l2)
        self.eng = _engine(config)
        self.precision = getattr(config, "precision", "fp32")
        # RunningMeanStd is baked in as identity: the only live op is the clip, and only when norm_type != 'none'
        clip = self.norm_type != "none"
        smin = np.broadcast_to(self.state_min, (self.state_dim,)) if clip else None
        smax = np.broadcast_to(self.state_max, (self.state_dim,)) if clip else None
        mk = lambda: Critic(self.eng, TMID, self.state_dim, self.action_dim, self.l1, self.l2, smin, smax)
        self.critic, self.target = mk(), mk()
        rng = np.random.RandomState(getattr(config, "random_seed", 0))
        S, A, H1, H2 = self.state_dim, self.action_dim, self.l1, self.l2
        u = lambda fan_in, *sh: rng.uniform(-np.sqrt(3.0 / fan_in), np.sqrt(3.0 / fan_in), sh).astype(np.float32)
        for c in (self.critic, self.target):          # independent inits, like two TF variable scopes
            c.load(u(S, S, H1), u(S, H1), u(H1 + A, H1 + A, H2), u(H1 + A, H2),
                   rng.uniform(-3e-3, 3e-3, (H2, 1)).astype(np.float32),
                   rng.uniform(-3e-3, 3e-3, 1).astype(np.float32), LAYOUT_IN_OUT)
        self.opt = CriticOptimizer(self.critic, lr=self.learning_rate, variant=ADAM_TF)

    # -- weights in the TF variable layout [in,out] (e.g. a decoded Bimodal1DEnv_trueQ_ckpt) --
    def set_weights(self, W1, b1, W2, b2, W3, b3, target=False):
        (self.target if target else self.critic).load(W1, b1, W2, b2, W3, b3, LAYOUT_IN_OUT)

    def load_tf_checkpoint(self, prefix, scope="main/qf", target=False):
        """Restore the critic from a TensorFlow V2 bundle (what ``tf.train.Saver.restore`` does for the
        ``main/qf`` variables, agents/SoftActorCritic.py:37-49), without TensorFlow (``tf_bundle.py``)."""
        from .tf_bundle import read_critic
        W1, b1, W2, b2, W3, b3 = read_critic(prefix, scope)
        self.set_weights(W1, b1, W2, b2, W3, b3, target=target)

    def save_tf_checkpoint(self, prefix, scope="main/qf", target=False):
        """Write the critic as a V2 bundle with the reference's variable names (byte-compatible container)."""
        from .tf_bundle import write_bundle
        names = ["fully_connected", "fully_connected_1", "fully_connected_2"]
        ws = self.get_weights(target=target)
        write_bundle(prefix, {"%s/%s/%s" % (scope, names[i // 2], "weights" if i % 2 == 0 else "biases"): np.asarray(w)
                              for i, w in enumerate(ws)})

    def get_weights(self, target=False):
        return [t.cpu().numpy() for t in (self.target if target else self.critic).export(LAYOUT_IN_OUT)]

    def _q(self, critic, inputs, action):
        inputs, action = _np(inputs), _np(action)
        if inputs.ndim != 2 or action.ndim != 2 or inputs.shape[0] != action.shape[0]:
            raise ValueError("inputs [R,S] and action [R,A] must have the same number of rows")
        R = inputs.shape[0]
        if R == 0:
            return np.zeros((0, 1), np.float32)
        # stacked rows: one "state" per row with a single action (N = 1) -- the B x N broadcast
        # entry points are eval()/cem() on the Critic itself
        q = critic.eval(inputs, action[:, None, :], self.precision)
        return q.reshape(R, 1).cpu().numpy()

    def _train(self, inputs, action, y):
        _, q = self.opt.step(torch.as_tensor(_np(inputs)), torch.as_tensor(_np(action)), _np(y).reshape(-1))
        return [q.reshape(-1, 1).cpu().numpy(), None]     # sess.run([outputs, optimize]) -> [q_pred, None]

    def _action_grads(self, critic, inputs, action):
        g, _ = critic.grad_action(_np(inputs), _np(action))
        return [g.cpu().numpy()]                          # sess.run(tf.gradients(q, action)) -> [grad]

    def _ascent(self, grad_fn, state, action_init, gd_alpha, gd_max_steps, gd_stop, is_training):
        # critic_network.py:126-146 / ae_network.py:321-350, verbatim control flow on the host
        action = np.copy(action_init)
        ascent_count = 0
        update_flag = np.ones([state.shape[0], self.action_dim])
        while np.any(update_flag > 0) and ascent_count < gd_max_steps:
            action_old = np.copy(action)
            gradients = grad_fn(state, action, is_training)[0]
            action += update_flag * gd_alpha * gradients
            action = np.clip(action, self.action_min, self.action_max)
            stop_idx = [idx for idx in range(len(action))
                        if np.mean(np.abs(action_old[idx] - action[idx]) / self.action_max) <= gd_stop]
            update_flag[stop_idx] = 0
            ascent_count += 1
        return action

    def init_target_network(self):
        self.target.copy_from(self.critic)

    def update_target_network(self):
        self.eng.soft_update(self.target.theta, self.critic.theta, self.tau)
        self.target.invalidate()

    def getQFunction(self, state):
        return lambda action: self._q(self.critic, np.expand_dims(state, 0), np.expand_dims([action], 0).reshape(1, -1))


class CriticNetwork(_TMidBase):
    """``agents/network/critic_network.py`` (DDPG-style critic; keys ``critic_lr, critic_l1_dim,
    critic_l2_dim``, :8-13)."""

    def __init__(self, sess, input_norm, config):
        self._build(sess, input_norm, config, config.critic_l1_dim, config.critic_l2_dim, config.critic_lr)

    def predict(self, *args):                 # (inputs, action, phase)   :101-111
        return self._q(self.critic, args[0], args[1])

    def predict_target(self, *args):          # :113-123
        return self._q(self.target, args[0], args[1])

    def train(self, *args):                   # (inputs, action, predicted_q_value)   :185-192
        return self._train(args[0], args[1], args[2])

    def action_gradients(self, inputs, action, is_training):            # :169-175
        return self._action_grads(self.critic, inputs, action)

    def action_gradients_target(self, inputs, action, is_training):     # :177-183
        return self._action_grads(self.target, inputs, action)

    def gradient_ascent(self, state, action_init, gd_alpha, gd_max_steps, gd_stop, is_training):   # :126-146
        return self._ascent(self.action_gradients, state, action_init, gd_alpha, gd_max_steps, gd_stop, is_training)

    def gradient_ascent_target(self, state, action_init, gd_alpha, gd_max_steps, gd_stop, is_training):
        return self._ascent(self.action_gradients_target, state, action_init, gd_alpha, gd_max_steps, gd_stop,
                            is_training)


class QTOPTNetwork(_TMidBase):
    """``agents/network/qt_opt_network.py`` (keys ``qnet_lr, qnet_l1_dim, qnet_l2_dim, num_iter,
    num_samples, top_m, num_modal, random_seed``, :10-24).  ``iterate_cem_multidim`` runs all CEM
    iterations for the whole batch in ONE kernel launch (rlc_cem); the random draws come from
    ``self.rng`` like the reference's (uniform proposals first, then mixture samples)."""

    def __init__(self, sess, input_norm, config):
        self._build(sess, input_norm, config, config.qnet_l1_dim, config.qnet_l2_dim, config.qnet_lr)
        self.rng = np.random.RandomState(config.random_seed)
        self.num_iter, self.num_samples = int(config.num_iter), int(config.num_samples)
        self.top_m, self.num_modal = int(config.top_m), int(config.num_modal)

    def predict_q(self, *args):               # :107-117
        return self._q(self.critic, args[0], args[1])

    def predict_q_target(self, *args):        # :119-129
        return self._q(self.target, args[0], args[1])

    def train(self, *args):                   # :193-200
        return self._train(args[0], args[1], args[2])

    def iterate_cem_multidim(self, state_batch):          # :132-175
        state_batch = _np(state_batch)
        B, N, A = len(state_batch), self.num_samples, self.action_dim
        u0 = self.rng.uniform(size=(B, N, A)).astype(np.float32)
        noise = self.rng.randn(self.num_iter - 1, B, N, A).astype(np.float32) if self.num_iter > 1 else None
        comp_u = self.rng.uniform(size=(self.num_iter - 1, B, N)).astype(np.float32) if self.num_iter > 1 else None
        amin = np.broadcast_to(self.action_min, (A,))
        amax = np.broadcast_to(self.action_max, (A,))
        w, mu, var, _, _ = self.critic.cem(state_batch, u0, noise, comp_u, self.top_m, self.num_modal, amin, amax)
        w, mu, var = w.cpu().numpy().astype(np.float64), mu.cpu().numpy().astype(np.float64), var.cpu().numpy().astype(np.float64)
        return [_GMM(w[b], mu[b], var[b], self.rng) for b in range(B)]

    def predict_action(self, state_batch):                # :177-181
        gmm_batch = self.iterate_cem_multidim(state_batch)
        return np.array([gmm.means_[np.argmax(gmm.weights_)] for gmm in gmm_batch])

    def sample_action(self, state_batch):                 # :183-191
        gmm_batch = self.iterate_cem_multidim(state_batch)
        final_action_samples_batch = np.array([gmm.sample(n_samples=1)[0] for gmm in gmm_batch])
        final_action_mean_batch = np.array([gmm.means_[np.argmax(gmm.weights_)] for gmm in gmm_batch])
        weight_mean_var_arr = [(gmm.weights_, gmm.means_, gmm.covariances_) for gmm in gmm_batch]
        return final_action_samples_batch, final_action_mean_batch, weight_mean_var_arr


class ActorExpertCritic(_TMidBase):
    """The Q ("expert") half of ``agents/network/ae_network.py`` / ``ae_expert_network.py`` /
    ``ae_plus_expert_network.py`` (keys ``expert_lr, shared_l1_dim | l1_dim, expert_l2_dim | l2_dim,
    better_q_gd_alpha, better_q_gd_max_steps, better_q_gd_stop``).  The actor half (mixture heads,
    NLL loss) stays in the reference network; the per-state elite selection the manager does after
    ``predict_q`` (ActorExpert.py:174-181) is ``select_elites``."""

    def __init__(self, sess, input_norm, config):
        l1 = getattr(config, "shared_l1_dim", getattr(config, "l1_dim", None))
        l2 = getattr(config, "expert_l2_dim", getattr(config, "l2_dim", None))
        lr = config.learning_rate[1] if hasattr(config, "learning_rate") and np.ndim(config.learning_rate) else config.expert_lr
        self._build(sess, input_norm, config, l1, l2, lr)
        self.rng = np.random.RandomState(config.random_seed)          # ae_network.py:20
        self.better_q_gd_alpha = getattr(config, "better_q_gd_alpha", 1e-2)
        self.better_q_gd_max_steps = getattr(config, "better_q_gd_max_steps", 10)
        self.better_q_gd_stop = getattr(config, "better_q_gd_stop", 1e-3)

    def predict_q(self, *args):               # ae_network.py:377-387
        return self._q(self.critic, args[0], args[1])

    def predict_q_target(self, *args):        # ae_network.py:389-399
        return self._q(self.target, args[0], args[1])

    def train_expert(self, *args):            # ae_network.py:360-367
        return self._train(args[0], args[1], args[2])

    def q_action_gradients(self, inputs, action, is_training):          # ae_network.py:352-358
        return self._action_grads(self.critic, inputs, action)

    def q_gradient_ascent(self, state, action_init, is_training, is_better_q_gd=False):   # ae_network.py:321-350
        assert is_better_q_gd
        return self._ascent(self.q_action_gradients, np.asarray(state), action_init, self.better_q_gd_alpha,
                            self.better_q_gd_max_steps, self.better_q_gd_stop, is_training)

    def select_elites(self, state_batch, action_batch, k):
        """ActorExpert.py:162-181 in one go: ``action_batch`` [B,N,A] samples per state ->
        (q [B,N], idx [B,k] = ``argsort()[::-1][:k]``, elites [B,k,A]); the stacked tensors are
        never materialised."""
        a = torch.as_tensor(_np(action_batch), device=self.eng.device)
        q = self.critic.eval(_np(state_batch), a, self.precision)
        idx, _, elites = self.eng.topk(q, int(k), a)
        return q.cpu().numpy(), idx.cpu().numpy(), elites.cpu().numpy()

    def sample_and_select_elites(self, state_batch, alpha, mean, sigma, k, num_samples, rng=None,
                                 equal_modal_selection=False, comp_u=None, normal=None):
        """The whole actor-update preamble of ``ActorExpert.update_network`` (ActorExpert.py:162-181:
        ``sample_action`` ae_network.py:461-496 -> ``predict_q`` -> per-state ``argsort()[::-1][:k]`` -> elite
        gather) as ONE launch, from the actor's mixture outputs ``alpha [B,M] | [B,M,1]``, ``mean, sigma [B,M,A]``.
        The draws come from ``rng`` exactly as the reference consumes them per state (``choice`` uniforms, then
        ``normal``), or are passed in.  Returns (elites [B,k,A], idx [B,k], sampled actions [B,N,A])."""
        mean = np.asarray(mean, np.float32)
        B, M, A = mean.shape
        N = int(num_samples)
        if comp_u is None:
            rng = self.rng if rng is None else rng
            comp_u, normal = np.empty((B, N)), np.empty((B, N, A))
            for b in range(B):                                  # same per-state order as the list comprehensions
                comp_u[b] = rng.random_sample(N)
            for b in range(B):
                normal[b] = rng.standard_normal((N, A))
        alpha = None if alpha is None else np.asarray(alpha, np.float32).reshape(B, M)
        out = self.critic.ae_expert_step(_np(state_batch), int(k), alpha, mean, _np(sigma), _np(comp_u), _np(normal),
                                         self.action_min, self.action_max, equal_modal=equal_modal_selection,
                                         want_actions=True)
        return out["elites"].cpu().numpy(), out["idx"].cpu().numpy(), out["actions"].cpu().numpy()


class SoftQNetwork(object):
    """torch ``SoftQNetwork`` (forwardkl_network.py:250-268, reversekl_network.py:259-282):
    ``q = net(state, action)`` on stacked rows -> ``[R,1]``; ``eval_grid(state_batch, grid)`` is the
    un-materialised B x N form the update loop needs (:160-164).  Tensors or arrays in; a torch
    tensor on the engine's device out (``.cpu().numpy()`` for the reference's numpy callers)."""

    def __init__(self, state_dim, action_dim, l1_dim, l2_dim, init_w=3e-3, engine=None, precision="auto", seed=None):
        self.eng = engine if engine is not None else _engine(None)
        self.critic = Critic(self.eng, TIN, int(state_dim), int(action_dim), int(l1_dim), int(l2_dim))
        self.precision = precision
        rng = np.random.RandomState(seed)
        S, A, H1, H2 = int(state_dim), int(action_dim), int(l1_dim), int(l2_dim)
        k1, k2 = 1 / np.sqrt(S + A), 1 / np.sqrt(H1)                          # nn.Linear default init
        u = lambda k, *sh: rng.uniform(-k, k, sh).astype(np.float32)
        self.critic.load(u(k1, H1, S + A), u(k1, H1), u(k2, H2, H1), u(k2, H2), u(init_w, 1, H2), u(init_w, 1),
                         LAYOUT_OUT_IN)

    def load_from_torch(self, module):
        """Copy ``linear1/2/3`` of a reference ``SoftQNetwork`` module."""
        g = lambda t: t.detach().cpu().numpy()
        self.critic.load(g(module.linear1.weight), g(module.linear1.bias), g(module.linear2.weight),
                         g(module.linear2.bias), g(module.linear3.weight), g(module.linear3.bias), LAYOUT_OUT_IN)
        return self

    def __call__(self, state, action):
        s = torch.as_tensor(state, dtype=torch.float32)
        a = torch.as_tensor(action, dtype=torch.float32)
        if s.dim() != 2 or a.dim() != 2 or s.shape[0] != a.shape[0]:
            raise ValueError("state [R,S] and action [R,A] must have the same number of rows")
        return self.critic.eval(s, a[:, None, :], "fp32" if s.shape[0] < 16384 else self.precision).reshape(-1, 1)

    forward = __call__

    def eval_grid(self, state_batch, grid_actions):
        return self.critic.eval(state_batch, grid_actions, self.precision)

    def optimizer(self, lr):
        return CriticOptimizer(self.critic, lr=lr, variant=ADAM_TORCH)
