"""torchrun helper (not collected by pytest): data-parallel ForwardKL update on 2+ GPUs vs the single-GPU update on
the concatenated minibatch.  Usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1
--master-port 29511 tests/gpu_dp_update.py"""
import os, sys
from types import SimpleNamespace
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlcontrol_b200 as rb
from rlcontrol_b200 import kl_networks


def cfg(engine, B, **kw):
    d = dict(state_dim=3, state_min=[-10.0] * 3, state_max=[10.0] * 3, action_dim=1, action_min=[-2.0], action_max=[2.0],
             tau=0.01, norm_type="input_norm", random_seed=0, pi_lr=1e-3, qf_vf_lr=1e-3, optim_type="intg",
             q_update_type="sac", use_true_q="False", actor_l1_dim=64, actor_l2_dim=48, critic_l1_dim=64, critic_l2_dim=48,
             entropy_scale=0.2, N_param=64, l_param=6, batch_size=B, engine=engine, precision="fp32")
    d.update(kw)
    return SimpleNamespace(**d)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    Bl = 16
    B = Bl * world
    rng = np.random.RandomState(0)
    batches = [(rng.randn(B, 3), rng.uniform(-2, 2, (B, 1)), rng.randn(B, 3), rng.randn(B), np.full(B, 0.99),
                rng.randn(B, 1).astype(np.float32)) for _ in range(3)]
    for kind in (kl_networks.ForwardKLNetwork, kl_networks.ReverseKLNetwork):
        torch.manual_seed(5)
        dp = kind(None, None, cfg(rb.Engine(lr), Bl, world_size=world))
        torch.manual_seed(5)
        ref = kind(None, None, cfg(rb.Engine(lr), B))
        sl = slice(rank * Bl, (rank + 1) * Bl)
        for b in batches:
            dp.update_network(*[x[sl] for x in b[:5]], eps=b[5][sl]); dp.update_target_network()
            ref.update_network(*b[:5], eps=b[5]); ref.update_target_network()
            np.testing.assert_allclose(dp.last_losses, ref.last_losses, rtol=2e-4, atol=2e-5)
        a, r = dp.export_parameters(), ref.export_parameters()
        for k in a:
            for x, y in zip(a[k], r[k]):
                np.testing.assert_allclose(x, y, rtol=0, atol=2e-5)
        # all ranks hold identical parameters
        t = dp.critic.theta.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float((t - dp.critic.theta).abs().max()) == 0.0
        if rank == 0:
            print(kind.__name__, "data-parallel == single-GPU on the concatenated batch; losses", dp.last_losses, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
