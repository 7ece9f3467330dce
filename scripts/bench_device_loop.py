"""cfg1 / cfg5 through the device-resident loop (rlcontrol_b200/device_loop.py): the README command's run
(Pendulum-v0 + reverse_kl.json, B=32, N=62, 200-200) as environment steps per second -- every step = env.step +
replay add + minibatch sample + full update_network + target update + sample_action, evaluation sessions (10 greedy
episodes every 500 steps) included -- for 1 run and for 8 runs interleaved on one GPU.
Usage: python scripts/bench_device_loop.py [steps_per_run] [n_runs]"""
import json
import os
import sys
import time
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlcontrol_b200 import device_loop as dl, kl_networks  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_runs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
env_json = {"environment": "Pendulum-v0", "TotalMilSteps": steps / 1e6, "EpisodeSteps": -1,
            "EvalIntervalMilSteps": 0.0005, "EvalEpisodes": 10}
spec = dl.EnvSpec(env_json)


def make(seed, entropy=0.1):
    cfg = SimpleNamespace(pi_lr=1e-3, qf_vf_lr=1e-3, tau=0.01, norm_type="input_norm", optim_type="intg",
                          q_update_type="non_sac", use_true_q="False", sample_for_eval="False", random_seed=seed,
                          entropy_scale=entropy, actor_l1_dim=200, actor_l2_dim=200, critic_l1_dim=200, critic_l2_dim=200,
                          N_param=64, l_param=6, batch_size=32, gamma=0.99, warmup_steps=0, buffer_size=1e6,
                          **spec.env_params())
    torch.manual_seed(seed)
    return dl.DeviceExperiment(kl_networks.ReverseKLNetwork(None, None, cfg), env_json, cfg)


def timed(exps):
    for e in exps:
        e._build()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = dl.run_interleaved(exps)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, res


dt1, res1 = timed([make(0)])
print(json.dumps({"config": "cfg1 device-resident run: Pendulum-v0 ReverseKL B=32 N=62 200-200, 1 run", "steps": steps,
                  "eval_sessions": len(res1[0][4]), "wall_s": round(dt1, 3), "env_steps_per_sec": round(steps / dt1, 1),
                  "note": "each step: env.step + replay add + sample 32 + full update_network + Polyak + sample_action; "
                          "eval sessions (10 x 200 greedy steps) inside the timed region",
                  "last_train_episode_rewards": [round(x, 1) for x in res1[0][0][-3:]]}))
dtn, resn = timed([make(s, (1, 0.1, 0.01, 0.001)[s % 4]) for s in range(n_runs)])
print(json.dumps({"config": f"cfg5 share of one GPU: {n_runs} independent runs interleaved (replicas only)", "steps_per_run": steps,
                  "wall_s": round(dtn, 3), "env_steps_per_sec_total": round(n_runs * steps / dtn, 1),
                  "env_steps_per_sec_per_run": round(steps / dtn, 1),
                  "mean_last_eval_return": [round(float(sum(r[1][-1]) / len(r[1][-1])), 1) for r in resn]}))
