"""Sweep INDEX -> hyper-parameter setting, and independent agents overlapped on one GPU.

The reference runs a sweep as one OS process per INDEX (``main_concurrent.py:64-81``, SLURM arrays): INDEX picks a
setting by mixed-radix decoding of the agent JSON's ``sweeps`` (first key fastest, wrapping for extra runs:
``utils/main_utils.py:90-98``), the run number and random seed are ``INDEX // total`` (``main.py:131-141``), and a
``Config`` is the defaults of ``utils/config.py:8-22`` with env, agent and CLI keys merged over them (``main.py:160-163``).
Runs share nothing, so on a B200 they are replicas only: each agent owns its handles, streams and captured update graph,
and the host launches one update per agent back to back and then waits -- the GPU executes them side by side."""
from __future__ import annotations

import json
from collections import OrderedDict
from types import SimpleNamespace
from typing import Iterable, Sequence


def sweep_setting(sweeps, index: int):
    """``get_sweep_parameters`` (utils/main_utils.py:90-98): returns (OrderedDict key -> value, number of settings)."""
    out, accum = OrderedDict(), 1
    for key in sweeps:
        num = len(sweeps[key])
        out[key] = sweeps[key][int(index / accum) % num]
        accum *= num
    return out, accum


def index_to_run(index: int, total: int):
    """(setting number, run number, random seed) of one INDEX (main.py:131-141: RANDOM_SEED = RUN_NUM)."""
    run = int(index / total)
    return index % total, run, run


DEFAULTS = dict(norm=None, exploration_policy=None, warmup_steps=0, batch_size=32, buffer_size=1e6, tau=0.01, gamma=0.99,
                ou_theta=0.15, ou_mu=0.0, ou_sigma=0.2)                     # utils/config.py:8-22


def make_config(env_params: dict, agent_params: dict, arg_params: dict | None = None, **extra):
    """``Config()`` + ``merge_config(env) / merge_config(agent) / merge_config(args)`` (main.py:160-163): every key
    becomes an attribute, later sources win.  ``extra`` carries this repo's optional keys (engine, precision ...)."""
    d = dict(DEFAULTS)
    for src in (env_params, agent_params, arg_params or {}, extra):
        d.update(src)
    return SimpleNamespace(**d)


def load_agent_json(path: str):
    with open(path) as f:
        j = json.load(f, object_pairs_hook=OrderedDict)                    # key order defines the radix order
    return j["agent"], j["sweeps"]


class SweepRunner:
    """K independent ForwardKL / ReverseKL agents (one per sweep INDEX) on one GPU.

    ``env_params``: what ``create_environment`` contributes to the config (state_dim, state_min/max, action_dim,
    action_min/max).  ``update_all(batches)`` takes one ``(s, a, s', r, gamma)`` tuple per agent, launches every agent's
    captured update (+ Polyak step) without waiting, then collects the losses."""

    def __init__(self, agent_name: str, sweeps, env_params: dict, indices: Iterable[int], device: int | None = None):
        import torch

        from . import kl_networks
        from .engine import Engine
        cls = {"ReverseKL": kl_networks.ReverseKLNetwork, "ForwardKL": kl_networks.ForwardKLNetwork}.get(agent_name)
        if cls is None:
            raise NotImplementedError("sweep runner covers the torch agents (ReverseKL, ForwardKL); got %r" % agent_name)
        self.indices, self.settings, self.agents = list(indices), [], []
        for idx in self.indices:
            params, total = sweep_setting(sweeps, idx)
            setting, run, seed = index_to_run(idx, total)
            cfg = make_config(env_params, params, dict(random_seed=seed), engine=Engine(device))
            torch.manual_seed(seed)                                         # main.py:141-146 seeds per run
            self.agents.append(cls(None, None, cfg))
            self.settings.append(dict(index=idx, setting=setting, run=run, agent_params=dict(params)))

    def update_all(self, batches: Sequence[tuple]):
        for ag, b in zip(self.agents, batches):
            ag.update_network_async(*b)
            ag.update_target_network()
        return [ag.wait() for ag in self.agents]
