"""Loader of the reference's OWN torch classes for the CPU arm of bench.py (cpu_baseline.kind == "reference").

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/oracle_np.py): only tests/, __graft_entry__ and bench.py's cpu_baseline /
--impl reference legs may import this.  ``__graft_entry__.build()`` copies the two unmodified reference files
(agents/network/forwardkl_network.py, reversekl_network.py) into ``oracle/_ref/`` when /root/reference exists (build
container); that directory is git-ignored (no reference source enters the history) but travels to the GPU box with the
snapshot.  The files import tensorflow-era siblings at module load (agents.network.base_network,
environments.environments, quadpy) that are absent here; they are stubbed exactly as oracle/make_golden.py does --
``SoftQNetwork`` / ``PolicyNetwork`` themselves are plain torch.nn modules and run unmodified."""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference_module(name="forwardkl_network"):
    """The unmodified reference module from oracle/_ref/, or None when the copy is not there."""
    path = os.path.join(REF_DIR, name + ".py")
    if not os.path.exists(path):
        return None
    _stub("agents")
    _stub("agents.network")
    _stub("agents.network.base_network", BaseNetwork=object)
    _stub("environments")
    _stub("environments.environments")
    _stub("quadpy")
    try:
        spec = importlib.util.spec_from_file_location("rlcontrol_ref_" + name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        return None


def reference_softq(params):
    """The reference's ``SoftQNetwork`` (forwardkl_network.py:250-268) holding ``params`` (torch [out,in] layout), or None."""
    import torch
    mod = load_reference_module("forwardkl_network")
    if mod is None:
        return None
    W1, b1, W2, b2, W3, b3 = [torch.as_tensor(p, dtype=torch.float32) for p in params]
    net = mod.SoftQNetwork(W1.shape[1] - 0, 0, W1.shape[0], W2.shape[0])      # Linear(state_dim + action_dim, l1): sum only
    with torch.no_grad():
        net.linear1.weight.copy_(W1); net.linear1.bias.copy_(b1.reshape(-1))
        net.linear2.weight.copy_(W2); net.linear2.bias.copy_(b2.reshape(-1))
        net.linear3.weight.copy_(W3.reshape(1, -1)); net.linear3.bias.copy_(b3.reshape(-1))
    for p in net.parameters():
        p.requires_grad_(False)
    return net
