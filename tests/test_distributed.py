"""World-size-2 gloo tests (CPU) of the multi-rank host logic: state sharding, the 1/B_total
gradient scaling + SUM all-reduce that reproduces the reference's batch mean, and the global
policy-loss mean.  The per-rank arithmetic is the oracle's (the CUDA kernels are covered by
test_gpu_parity.py::test_critic_grads_and_adam_variants on one GPU with the same b_total logic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_np as onp
from rlcontrol_b200.parallel import allreduce_grad_, global_mean_from_shards, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(0)                      # identical data on every rank
        S, A, H1, H2, B = 5, 2, 24, 16, 37                   # B not divisible by world
        k1, k2 = 1 / np.sqrt(S + A), 1 / np.sqrt(H1)
        p = [rng.uniform(-k1, k1, (H1, S + A)), rng.uniform(-k1, k1, H1), rng.uniform(-k2, k2, (H2, H1)),
             rng.uniform(-k2, k2, H2), rng.uniform(-.3, .3, (1, H2)), rng.uniform(-.3, .3, 1)]
        s, a, y = rng.randn(B, S), rng.randn(B, A), rng.randn(B)
        lo, hi = shard_bounds(B, rank, world)
        # per-rank gradient of sum_b (y-Q)^2 / B_total  ==  (local mean-gradient) * B_local / B_total
        loss_loc, g_loc = onp.tin_mse_grads(s[lo:hi], a[lo:hi], y[lo:hi], p)
        flat = np.concatenate([g.ravel() for g in g_loc]) * (hi - lo) / B
        g = torch.tensor(flat)
        allreduce_grad_(g)
        loss_full, g_full = onp.tin_mse_grads(s, a, y, p)
        ref = np.concatenate([x.ravel() for x in g_full])
        np.testing.assert_allclose(g.numpy(), ref, rtol=1e-9, atol=1e-12)
        # global mean of a per-state loss held shard-wise
        per_state = torch.tensor((y ** 2)[lo:hi])
        m = global_mean_from_shards(per_state, B)
        assert abs(float(m) - float((y ** 2).mean())) < 1e-12
        # every rank ends up with identical parameters after the same Adam step
        th = np.concatenate([x.ravel() for x in p])
        th2, _, _ = onp.adam_step_torch(th, g.numpy(), np.zeros_like(th), np.zeros_like(th), 1, 1e-3)
        gathered = [torch.zeros_like(torch.tensor(th2)) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor(th2))
        assert all(torch.equal(gathered[0], t) for t in gathered)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_gradient_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [out.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
