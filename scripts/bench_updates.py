"""Full agent updates per second through the drop-in classes (rlcontrol_b200/kl_networks.py): one
``update_network`` + ``update_target_network`` = what the reference's manager runs per environment step
(agents/ReverseKL.py:81-90).  Host numpy arrays in, losses back on the host, one CUDA-graph launch per update.
One JSON line per configuration."""
import json, os, sys, time
from types import SimpleNamespace
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rlcontrol_b200 as rb
from rlcontrol_b200 import kl_networks


def cfg(eng, S, A, amax, B, n_param, l1, l2, optim="intg", qtype="non_sac", alpha=0.1, **kw):
    d = dict(state_dim=S, state_min=[-10.0] * S, state_max=[10.0] * S, action_dim=A, action_min=[-amax] * A,
             action_max=[amax] * A, tau=0.01, norm_type="input_norm", random_seed=0, pi_lr=1e-3, qf_vf_lr=1e-3,
             optim_type=optim, q_update_type=qtype, use_true_q="False", actor_l1_dim=l1, actor_l2_dim=l2,
             critic_l1_dim=l1, critic_l2_dim=l2, entropy_scale=alpha, N_param=n_param, l_param=6, batch_size=B, engine=eng)
    d.update(kw)
    return SimpleNamespace(**d)


def run(name, cls, c, reps):
    eng = c.engine
    net = cls(None, None, c)
    rng = np.random.RandomState(0)
    B, S, A = c.batch_size, c.state_dim, c.action_dim
    batches = [(rng.randn(B, S), rng.uniform(-1, 1, (B, A)) * c.action_max[0], rng.randn(B, S), rng.randn(B),
                np.full(B, 0.99)) for _ in range(8)]
    for i in range(5):
        net.update_network(*batches[i % 8]); net.update_target_network()
    torch.cuda.synchronize()
    l0 = eng.launches
    t0 = time.perf_counter()
    for i in range(reps):
        net.update_network(*batches[i % 8]); net.update_target_network()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    # device time of the captured update alone
    st = net._steps[B]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st.stream):
        e0.record()
        for _ in range(reps):
            st.graph.replay()
        e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"config": name, "ms_per_update_host": round(dt * 1e3, 4), "updates_per_sec": round(1 / dt, 1),
                      "ms_per_update_device": round(e0.elapsed_time(e1) / reps, 4),
                      "kernels_per_update": "graph (launch counter not advanced by replays: %d)" % (eng.launches - l0),
                      "N": net.intgrl_actions_len, "losses": [float(x) for x in net.last_losses]}), flush=True)


def main():
    eng = rb.Engine(0)
    run("cfg1 ReverseKL Pendulum: S=3 A=1 B=32 N_param=64 200-200 (README command), full update_network + target update",
        kl_networks.ReverseKLNetwork, cfg(eng, 3, 1, 2.0, 32, 64, 200, 200), 300)
    run("cfg1-shape ForwardKL", kl_networks.ForwardKLNetwork, cfg(eng, 3, 1, 2.0, 32, 64, 200, 200), 300)
    run("cfg4-exact ForwardKL: S=3 A=1 B=4096 N_param=1026 400-300, full update_network + target update",
        kl_networks.ForwardKLNetwork, cfg(eng, 3, 1, 2.0, 4096, 1026, 400, 300), 50)
    run("cfg4-exact ReverseKL", kl_networks.ReverseKLNetwork, cfg(eng, 3, 1, 2.0, 4096, 1026, 400, 300), 50)


if __name__ == "__main__":
    main()
