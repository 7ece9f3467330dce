"""Device time of each launch of the small-minibatch update (cfg1 shape), each replayed alone in a CUDA graph of 50 copies."""
import os, sys, json
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlcontrol_b200 import device_loop as dl, kl_networks
from rlcontrol_b200._lib import check
from rlcontrol_b200.engine import _stream

env_json = {"environment": "Pendulum-v0", "TotalMilSteps": 0.001, "EpisodeSteps": -1, "EvalIntervalMilSteps": 0.0005, "EvalEpisodes": 10}
spec = dl.EnvSpec(env_json)
cfg = SimpleNamespace(pi_lr=1e-3, qf_vf_lr=1e-3, tau=0.01, norm_type="input_norm", optim_type="intg", q_update_type="non_sac",
                      use_true_q="False", sample_for_eval="False", random_seed=0, entropy_scale=0.1, actor_l1_dim=200,
                      actor_l2_dim=200, critic_l1_dim=200, critic_l2_dim=200, N_param=64, l_param=6, batch_size=32, gamma=0.99,
                      warmup_steps=0, buffer_size=1e6, **spec.env_params())
torch.manual_seed(0)
net = kl_networks.ReverseKLNetwork(None, None, cfg)
exp = dl.DeviceExperiment(net, env_json, cfg)
exp._build()
st, B, lib, h = exp.st, 32, net.eng.lib, net.eng.h
for k in ("s", "s2"):
    st.d[k].normal_()
st.d["a"].uniform_(-2, 2); st.d["r"].normal_(); st.d["g"].fill_(0.99); st.d["eps"].normal_()
exp.cur[1] = 1000


def timeit(name, fn, reps=50):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s); s.synchronize()
    print(json.dumps({"launch": name, "us": round(1e3 * e0.elapsed_time(e1) / (5 * reps), 2)}))


timeit("rlc_sb_forward x4 nets (V, Vtarg, pi+evaluate, Q), B=32", lambda: check(lib.rlc_sb_forward(h, st.sb_fwd, 4, B, _stream())))
timeit("rlc_sb_forward x5: the four B-row passes + the 32 x 62 grid rows", lambda: check(lib.rlc_sb_forward(h, st.sb_fwd, 5, B, _stream())))
timeit("rlc_sb_forward pi+evaluate, B=1 (sample_action)", lambda: check(lib.rlc_sb_forward(h, exp.sb_act, 1, 1, _stream())))
timeit("rlc_sb_forward pi, B=10 (evaluation step)", lambda: check(lib.rlc_sb_forward(h, exp.sb_ev, 1, 10, _stream())))
timeit("rlc_sb_update x3 nets (V, Q, pi) + Adam + Polyak, B=32", lambda: check(lib.rlc_sb_update(h, st.sb_upd, 3, B, B, _stream())))
timeit("grid evaluation 32 x 62 rows, 200-200 (rlc_critic_eval fp32)", lambda: net.critic_grid.eval_into(st.d["s"], net.intgrl_actions, st.q_grid, net.precision))
timeit("rlc_reduce_rkl_policy (k_grid_logterms + k_policy_reduce)", lambda: net.eng.rkl_policy(
    st.q_grid, st.v_out.view(-1), net.intgrl_weights, net.intgrl_actions, net.action_scale, st.ev["mu_raw"], st.ev["log_std"], 0.1,
    hard=False, b_total=B, out=(st.loss_b, st.dmean, st.dls)))
timeit("k_env_step_train", exp._env_step)
timeit("k_loop_stage", lambda: exp._stage(True))
timeit("whole training step (graph of 1)", lambda: exp._train_step(True), reps=20)

# %globaltimer trace of one rlc_sb_update launch (CTA 0, thread 0)
dbg = torch.zeros(128, dtype=torch.int64, device=net.device)
os.environ["RLC_SB_DEBUG"] = hex(dbg.data_ptr())
for _ in range(3):
    check(lib.rlc_sb_forward(h, st.sb_fwd, 5, B, _stream()))
    check(lib.rlc_sb_update(h, st.sb_upd, 3, B, B, _stream()))
torch.cuda.synchronize()
t = dbg.cpu().numpy()
names = {65: "prologue loads staged", 66: "dz2 done", 67: "bulk W2 rows + moments arrived", 68: "dh1/dz1 done", 72: "dW2 sums done (before Adam)", 69: "dW2 + Adam done",
         70: "barrier", 71: "small parameters + Adam done"}
print("rlc_sb_update trace (ns from kernel start):", {names[k]: int(t[k] - t[64]) for k in sorted(names)})
