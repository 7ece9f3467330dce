#!/usr/bin/env python
"""bench.py -- sampled-action critic evaluation throughput on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference-equivalent CPU path

Workload (config.workload = "cfg4"): ForwardKL large batch, per GPU B=4096 states x N=1024 grid
actions, S=17, A=6, 400-300 T-in critic (BASELINE.json configs[3], the shape the target is quoted
on).  One *step* = one pass of the hot path over one replay minibatch:
    K1  rlc_critic_eval   q[B,N] = Q(s_b, a_n)   (tcgen05 kernel, fp16 operands / fp32 accumulate)
    K3  rlc_reduce_fkl_policy  per-state Boltzmann weights + policy loss over the grid, with the
                          tanh-Gaussian log-density evaluated in place from mean/log_std [B,A]
metric = (s,a) Q-evaluations per second, whole job (all ranks).  States shard over ranks with no
data-path collective (weak scaling: every rank gets its own B=4096 minibatch).

value : inputs already resident in HBM, CUDA events on the launching stream around exactly K steps.
e2e   : the same step through the public API with HOST (pinned) inputs: H2D copy of the states and
        the policy head outputs (mean, log_std), D2H read of the per-state loss, in the timed region.
roofline : dominant kernel = K1; achieved = algorithmic flops / its mean launch time
        measured with CUDA events in this process; peak from MEASURED_PEAKS.json.
cpu_baseline / --impl reference : oracle/oracle_torch.py (the reference's torch-CPU arithmetic,
        materialised stacks) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOAD = dict(workload="cfg4", B_per_gpu=4096, N=1024, S=17, A=6, H1=400, H2=300, topology="T-in",
                reduction="forward_kl (Boltzmann weights + tanh-Gaussian log-density + policy loss and its gradient wrt the policy head)", action_layout="shared_grid[N,A]", entropy_scale=0.1)
RING = 10                     # rotating input/output sets: 10 x q[B,N] (16.8 MB each) = 168 MB > 126 MB L2
METRIC = "sampled_q_evals_per_sec"
UNIT = "Q-evals/s"


def algorithmic_flops(B, N, S, A, H1, H2):
    """F = B*2*S*H1 + B*N*2*(A*H1 + H1*H2 + H2)."""
    return B * 2 * S * H1 + B * N * 2 * (A * H1 + H1 * H2 + H2)


def make_params(rng, S, A, H1, H2):
    """torch nn.Linear default init + U(+-3e-3) last layer scaled x100 ("trained-like")."""
    k1, k2 = 1 / np.sqrt(S + A), 1 / np.sqrt(H1)
    u = lambda k, *sh: rng.uniform(-k, k, sh).astype(np.float32)
    return [u(k1, H1, S + A), u(k1, H1), u(k2, H2, H1), u(k2, H2), u(0.3, 1, H2), u(0.3, 1)]


def make_inputs(rng, B, N, S, A):
    from rlcontrol_b200 import quadrature          # grid weights only (Clenshaw-Curtis interior)
    s = np.clip(rng.randn(B, S), -10, 10).astype(np.float32)
    a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
    _, w = quadrature.grid_1d(N + 2, 1.0)
    mean = (rng.randn(B, A) * 0.5).astype(np.float32)          # policy head outputs (actor side, fed in)
    log_std = (rng.randn(B, A) * 0.3 - 0.5).astype(np.float32)
    return s, a, np.asarray(w, np.float32), (mean, log_std)


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------
_SAMPLER_SRC = r"""
import sys, time, json
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
mx = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
out = []
sys.stdout.write("ready\n"); sys.stdout.flush()
import select
while True:
    out.append((time.time(), int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get(h))))
    if select.select([sys.stdin], [], [], 0.0005)[0]:
        break
sys.stdout.write(json.dumps({"max": mx, "samples": out}) + "\n"); sys.stdout.flush()
"""


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled by a helper PROCESS (pynvml, ~1 kHz)
    so that the polling never takes this process's GIL away from the launch loop.  The helper runs from
    construction; only samples whose timestamp falls inside [__enter__, __exit__] are reported."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        import subprocess
        self.t0 = self.t1 = None
        self.proc = None
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(index)], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            if self.proc.stdout.readline().strip() != "ready":
                raise RuntimeError("sampler did not start")
        except Exception:
            self.proc = None

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *exc):
        self.t1 = time.time()
        self.data = None
        if self.proc is not None:                 # stop the helper right away: nothing polls NVML after the window
            try:
                out, _ = self.proc.communicate("stop\n", timeout=10)
                self.data = json.loads(out.strip().splitlines()[-1])
            except Exception:
                self.proc.kill()

    def summary(self):
        data = getattr(self, "data", None)
        if data is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        inside = [(mhz, mask) for (ts, mhz, mask) in data["samples"] if self.t0 <= ts <= self.t1]
        if not inside:      # region shorter than one sampling period: take the nearest samples around it
            inside = sorted(data["samples"], key=lambda x: abs(x[0] - 0.5 * (self.t0 + self.t1)))[:3]
            inside = [(mhz, mask) for (_, mhz, mask) in inside]
        reasons = sorted({k for k, bit in self.REASONS.items() for (_, mask) in inside if mask & bit})
        return {"sm_mhz": int(np.median([m for m, _ in inside])), "sm_max_mhz": data["max"], "reasons": reasons,
                "samples": len(inside), "sampler": "helper process, pynvml, ~1 kHz"}


# ---------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference's torch-CPU path)
# ---------------------------------------------------------------------------------------------
def cpu_arm(params, s, a, w, pol, entropy_scale, budget_s=12.0, b_sample=256, max_reps=40, warmup=1):
    """Times oracle_torch.fkl_sampled_step on the first b_sample states (a bounded sample of the
    workload; the full B=4096 stack would need 12 GB of fp32 activations on the host)."""
    from oracle import oracle_torch as ot
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = ot.SoftQNetworkPort(*params)
    ts, ta, tw = (torch.as_tensor(x) for x in (s[:b_sample], a, w))
    tl = (torch.as_tensor(pol[0][:b_sample]), torch.as_tensor(pol[1][:b_sample]))
    for _ in range(warmup):
        ot.fkl_sampled_step(net, ts, ta, tw, tl, entropy_scale)
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < max_reps and (time.perf_counter() < t_end or len(times) < 3):
        t0 = time.perf_counter()
        ot.fkl_sampled_step(net, ts, ta, tw, tl, entropy_scale)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    evals = b_sample * a.shape[0]
    return {"value": evals / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first {b_sample} of {s.shape[0]} states x N={a.shape[0]} ({evals} rows), "
                      f"median of {len(times)} reps, oracle/oracle_torch.py (reference torch-CPU arithmetic)",
            "ms_per_sample": med * 1e3}


def run_reference(args, rank, world):
    """--impl reference: the reference-equivalent CPU path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    W = WORKLOAD
    rng = np.random.RandomState(0)
    params = make_params(rng, W["S"], W["A"], W["H1"], W["H2"])
    s, a, w, pol = make_inputs(rng, W["B_per_gpu"], W["N"], W["S"], W["A"])
    from oracle import oracle_torch as ot
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = ot.SoftQNetworkPort(*params)
    b_sample = 256
    ts, ta, tw = (torch.as_tensor(x) for x in (s[:b_sample], a, w))
    tl = (torch.as_tensor(pol[0][:b_sample]), torch.as_tensor(pol[1][:b_sample]))
    for _ in range(max(args.warmup, 1)):
        ot.fkl_sampled_step(net, ts, ta, tw, tl, W["entropy_scale"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ot.fkl_sampled_step(net, ts, ta, tw, tl, W["entropy_scale"])
    dt = time.perf_counter() - t0
    evals = b_sample * W["N"]
    value = evals * args.steps / dt
    sample = (f"each step = first {b_sample} of {W['B_per_gpu']} states x N={W['N']} ({evals} rows) of the cfg4 "
              f"minibatch through oracle/oracle_torch.py (the reference's torch-CPU arithmetic, materialised stacks)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(W, sample_states_per_step=b_sample),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch.distributed as dist
    import rlcontrol_b200 as rb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = WORKLOAD
    B, N, S, A, H1, H2 = W["B_per_gpu"], W["N"], W["S"], W["A"], W["H1"], W["H2"]
    alpha = W["entropy_scale"]
    rng = np.random.RandomState(0)
    params = make_params(rng, S, A, H1, H2)                       # identical weights on every rank
    rng_in = np.random.RandomState(1000 + rank)                   # each rank: its own minibatch shard
    s_np, a_np, w_np, (mean_np, lstd_np) = make_inputs(rng_in, B, N, S, A)
    a_np, w_np = make_inputs(np.random.RandomState(1), 1, N, S, A)[1:3]   # the grid is shared by all ranks

    eng = rb.Engine(local_rank)
    critic = rb.Critic(eng, rb.TIN, S, A, H1, H2).load(*params, rb.LAYOUT_OUT_IN)
    prec = args.precision
    t = lambda x: torch.as_tensor(x, device=dev)
    a_d, w_d = t(a_np), t(w_np)
    # rotating sets so that consecutive steps never find their inputs/outputs in L2
    s_ring = [t(np.roll(s_np, i, axis=0).copy()) for i in range(RING)]
    mean_ring = [t(np.roll(mean_np, i, axis=0).copy()) for i in range(RING)]
    lstd_ring = [t(np.roll(lstd_np, i, axis=0).copy()) for i in range(RING)]
    ACTION_SCALE = 1.0
    q_ring = [torch.empty((B, N), dtype=torch.float32, device=dev) for _ in range(RING)]

    def step(i):
        j = i % RING
        q = critic.eval_into(s_ring[j], a_d, q_ring[j], prec)
        loss_b, dmean, dlstd, _ = eng.fkl_policy(q, w_d, a_d, ACTION_SCALE, mean_ring[j], lstd_ring[j], alpha)
        return loss_b

    # ---- parity gate, part of the cpu_baseline leg (rank 0 at N=1; the only place this arm touches oracle/, as the
    # checker): the oracle evaluates a sample of this rank's rows and the device results must match it ----
    loss0 = step(0)
    torch.cuda.synchronize()
    if eng.umma_error() != 0:
        raise SystemExit("bench.py: tcgen05 kernel raised its error flag")
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_np as onp
        rows = np.arange(0, B, B // 8)[:8]
        q_ref = onp.tin_eval(s_np[rows], a_np, params, dtype=np.float64)
        q_gpu = q_ring[0][torch.as_tensor(rows, device=dev)].cpu().numpy()
        den = np.maximum(np.abs(q_ref), np.sqrt((q_ref ** 2).mean(1, keepdims=True)))
        err = np.abs(q_gpu - q_ref) / den
        parity = {"rows_checked": int(q_ref.size), "rel_err_rms": float(np.sqrt((err ** 2).mean())),
                  "rel_err_max": float(err.max())}
        if prec in ("fp16", "bf16"):
            q_rnd = onp.tin_eval_rounded(s_np[rows], a_np, params, prec,
                                           head=critic.tensor_arithmetic(True))
            d = np.abs(q_gpu - q_rnd) / den
            parity["vs_stated_arithmetic_rms"] = float(np.sqrt((d ** 2).mean()))
            parity["vs_stated_arithmetic_max"] = float(d.max())
            ok = parity["vs_stated_arithmetic_rms"] < 3e-5 and parity["vs_stated_arithmetic_max"] < 2e-3
        else:
            ok = parity["rel_err_max"] < 2e-5
        per_state = onp.fkl_policy_reduce(q_gpu, w_np, a_np, mean_np[rows], lstd_np[rows], ACTION_SCALE, alpha)[0]
        ok = ok and np.allclose(loss0.cpu().numpy()[rows], per_state, rtol=1e-3, atol=1e-5)
        if not ok:
            raise SystemExit(f"bench.py: parity gate failed: {parity}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: exactly K steps between two events ----
    clocks = ClockSampler(local_rank)          # helper process starts sampling now; the window is marked below
    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks:
        ev0.record()
        for i in range(args.steps):
            step(args.warmup + i)
        ev1.record()
        barrier()
    launches = eng.launches - launches0
    ms_total = ev0.elapsed_time(ev1)

    # ---- dominant kernel alone (roofline): K1 launches only, CUDA events on the same stream ----
    k1_reps = max(args.steps, 10)
    for i in range(3):
        critic.eval_into(s_ring[i % RING], a_d, q_ring[i % RING], prec)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(k1_reps):
        critic.eval_into(s_ring[i % RING], a_d, q_ring[i % RING], prec)
    ev1.record()
    torch.cuda.synchronize()
    k1_ms = ev0.elapsed_time(ev1) / k1_reps

    # ---- end to end through the public API: host inputs in, host result out, every step ----
    # rlcontrol_b200.steps.ForwardKLGridStep: one CUDA-graph launch = H2D of the minibatch (states,
    # policy head outputs) from pinned memory, K1, K3, D2H of loss and policy-head gradients.
    from rlcontrol_b200.steps import ForwardKLGridStep
    fstep = ForwardKLGridStep(critic, a_d, w_d, ACTION_SCALE, alpha, B, precision=prec)
    host_in = [(torch.as_tensor(np.roll(s_np, i, axis=0).copy()), torch.as_tensor(np.roll(mean_np, i, axis=0).copy()),
                torch.as_tensor(np.roll(lstd_np, i, axis=0).copy())) for i in range(4)]

    def e2e_step(i):
        hs, hm, hl = host_in[i & 3]
        loss_host, dmean_host, _ = fstep(hs, hm, hl)       # copies into pinned staging, launches, synchronises
        return float(loss_host[0]) + float(dmean_host[0, 0])

    chk = e2e_step(0)
    # the end-to-end call must reproduce the device-resident step on the same inputs (ring slot 0 = host set 0); the
    # device-resident step itself is what the parity gate above checked against the oracle
    loss_dev0 = float(loss0[0])
    if not np.isfinite(chk) or abs(float(fstep.loss_host[0]) - loss_dev0) > 1e-5 * max(1.0, abs(loss_dev0)):
        raise SystemExit("bench.py: e2e step disagrees with the device-resident step")
    for i in range(args.warmup):
        e2e_step(i)
    # host-timed: exactly K steps per trial; 5 trials, the median is reported (host jitter on a shared box
    # occasionally doubles a single 20-step window), min/max kept in the line
    import gc
    gc.disable()
    trials = []
    for _ in range(5):
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        torch.cuda.synchronize()
        trials.append((time.perf_counter() - t0) * 1e3)
    gc.enable()
    e2e_blk_ms = float(np.median(trials))     # blocking call per step: copy in, launch, wait, copy out, strictly in sequence
    blk_trials = trials
    # the same steps through the pipelined API (steps.ForwardKLGridPipeline, two slots): every step still pays its own
    # H2D and D2H copies, but step i+1 is staged and uploaded while step i computes; this is the headline e2e number
    from rlcontrol_b200.steps import ForwardKLGridPipeline
    pipe = ForwardKLGridPipeline(critic, a_d, w_d, ACTION_SCALE, alpha, B, precision=prec, depth=2)

    def pipe_run(k_steps):
        acc = 0.0
        pipe.submit(*host_in[0])
        for i in range(1, k_steps):
            pipe.submit(*host_in[i & 3])
            r = pipe.result()
            acc += float(r[0][0]) + float(r[1][0, 0])
        r = pipe.result()
        return acc + float(r[0][0]) + float(r[1][0, 0])

    pipe.submit(*host_in[0])
    if abs(float(pipe.result()[0][0]) - loss_dev0) > 1e-5 * max(1.0, abs(loss_dev0)):
        raise SystemExit("bench.py: pipelined e2e step disagrees with the device-resident step")
    pipe_run(max(args.warmup, 3))
    gc.disable()
    trials = []
    for _ in range(5):
        barrier()
        t0 = time.perf_counter()
        pipe_run(args.steps)
        torch.cuda.synchronize()
        trials.append((time.perf_counter() - t0) * 1e3)
    gc.enable()
    e2e_wall_ms = float(np.median(trials))
    e2e_ms = e2e_wall_ms          # every result is read on the host inside the window: wall clock IS the end-to-end time
    barrier()
    h2d = (fstep.s_host.numel() + fstep.mean_host.numel() + fstep.log_std_host.numel()) * 4
    d2h = (fstep.loss_host.numel() + fstep.dmean_host.numel() + fstep.dlog_std_host.numel()) * 4

    # ---- secondary: critic regression update (a15/a16) incl. the NCCL grad all-reduce when N>1 ----
    a_reg = t(rng_in.uniform(-1, 1, (B, A)).astype(np.float32))
    y_reg = t(rng_in.randn(B).astype(np.float32))
    opt = rb.CriticOptimizer(rb.Critic(eng, rb.TIN, S, A, H1, H2).load(*params, rb.LAYOUT_OUT_IN), lr=1e-3)
    for _ in range(3):
        opt.step(s_ring[0], a_reg, y_reg, world_size=world)
    barrier()
    upd_reps = 20
    ev0.record()
    for _ in range(upd_reps):
        opt.step(s_ring[0], a_reg, y_reg, world_size=world)
    ev1.record()
    barrier()
    upd_ms = ev0.elapsed_time(ev1) / upd_reps

    # ---- secondary: cfg1 (the README command: Pendulum ReverseKL, B=32, N_param=64 -> N=62, S=3, A=1, 200-200).
    # One FULL agent update through the drop-in class: ReverseKLNetwork.update_network + update_target_network
    # (q, v and pi networks, three Adam steps, Polyak), numpy minibatch in, losses out, host-synchronous like
    # the reference's loop (agents/ReverseKL.py:81-90) -- one CUDA-graph launch per update.
    from types import SimpleNamespace
    from rlcontrol_b200 import kl_networks

    def kl_config(engine, S_, A_, amax, B_, n_param, l1_, l2_, **kw):
        d = dict(state_dim=S_, state_min=[-10.0] * S_, state_max=[10.0] * S_, action_dim=A_, action_min=[-amax] * A_,
                 action_max=[amax] * A_, tau=0.01, norm_type="input_norm", random_seed=0, pi_lr=1e-3, qf_vf_lr=1e-3,
                 optim_type="intg", q_update_type="non_sac", use_true_q="False", actor_l1_dim=l1_, actor_l2_dim=l2_,
                 critic_l1_dim=l1_, critic_l2_dim=l2_, entropy_scale=alpha, N_param=n_param, l_param=6,
                 batch_size=B_, engine=engine)
        d.update(kw)
        return SimpleNamespace(**d)

    rng1 = np.random.RandomState(5)
    B1 = 32
    batches1 = [(rng1.randn(B1, 3), rng1.uniform(-2, 2, (B1, 1)), rng1.randn(B1, 3), rng1.randn(B1), np.full(B1, 0.99))
                for _ in range(8)]
    torch.manual_seed(0)
    ag = kl_networks.ReverseKLNetwork(None, None, kl_config(rb.Engine(local_rank), 3, 1, 2.0, B1, 64, 200, 200))

    def update_cfg1(i):
        ag.update_network(*batches1[i % 8])
        ag.update_target_network()

    for i in range(20):
        update_cfg1(i)
    torch.cuda.synchronize()
    n_upd = 300
    t0 = time.perf_counter()
    for i in range(n_upd):
        update_cfg1(i)
    torch.cuda.synchronize()
    cfg1_ms = (time.perf_counter() - t0) * 1e3 / n_upd

    # cfg5 share of one GPU: 8 independent agents (sweep INDEX runs, main_concurrent.py:64-81), each with its own
    # handles / streams / graph, overlapping on the device; replicas only, no communication
    agents = [kl_networks.ReverseKLNetwork(None, None, kl_config(rb.Engine(local_rank), 3, 1, 2.0, B1, 64, 200, 200))
              for _ in range(8)]

    def sweep_step(i):
        for g in agents:
            g.update_network_async(*batches1[i % 8])
            g.update_target_network()
        return sum(float(g.wait()[0]) for g in agents)

    for i in range(10):
        sweep_step(i)
    torch.cuda.synchronize()
    n_sw = 100
    t0 = time.perf_counter()
    for i in range(n_sw):
        sweep_step(i)
    torch.cuda.synchronize()
    sweep_ms = (time.perf_counter() - t0) * 1e3 / n_sw

    # cfg1 / cfg5 as whole RUNS (environment + replay + agent on the device, rlcontrol_b200/device_loop.py): the README
    # command's loop -- env.step, replay add, minibatch sample, full update_network, Polyak, sample_action per step,
    # evaluation sessions (10 greedy episodes every 500 steps) included -- for 1 run and 8 interleaved runs per GPU
    from rlcontrol_b200 import device_loop as dl
    run_steps = 3000
    env_json = {"environment": "Pendulum-v0", "TotalMilSteps": run_steps / 1e6, "EpisodeSteps": -1,
                "EvalIntervalMilSteps": 0.0005, "EvalEpisodes": 10}
    spec_dl = dl.EnvSpec(env_json)

    def make_run(seed):
        c = kl_config(rb.Engine(local_rank), 3, 1, 2.0, B1, 64, 200, 200, random_seed=seed, gamma=0.99, warmup_steps=0,
                      buffer_size=1e6, sample_for_eval="False", **{k: v for k, v in spec_dl.env_params().items()
                                                                    if k in ("state_min", "state_max")})
        torch.manual_seed(seed)
        return dl.DeviceExperiment(kl_networks.ReverseKLNetwork(None, None, c), env_json, c)

    def timed_runs(exps):
        for e in exps:
            e._build()
        torch.cuda.synchronize()
        t0_ = time.perf_counter()
        dl.run_interleaved(exps)
        torch.cuda.synchronize()
        return time.perf_counter() - t0_
    run1_s = timed_runs([make_run(0)])
    run8_s = timed_runs([make_run(i) for i in range(8)])

    # cfg4, the full ForwardKL update_network on this rank's B=4096 minibatch with the synthetic [N,A] grid
    # (q/v/pi networks 400-300, three backward passes and Adam steps; tensor-core grid evaluation inside)
    torch.manual_seed(1)
    ag4 = kl_networks.ForwardKLNetwork(None, None, kl_config(rb.Engine(local_rank), S, A, 1.0, B, 64, H1, H2,
                                                               integration_grid=(a_np, w_np), precision=prec,
                                                               world_size=world))
    b4 = (s_np, rng_in.uniform(-1, 1, (B, A)).astype(np.float32), np.roll(s_np, 1, axis=0),
          rng_in.randn(B).astype(np.float32), np.full(B, 0.99, np.float32))
    for _ in range(3):
        ag4.update_network(*b4)
        ag4.update_target_network()
    torch.cuda.synchronize()
    n4 = 20
    t0 = time.perf_counter()
    for _ in range(n4):
        ag4.update_network(*b4)
        ag4.update_target_network()
    torch.cuda.synchronize()
    cfg4_full_ms = (time.perf_counter() - t0) * 1e3 / n4
    if world > 1:
        tt = torch.tensor([cfg4_full_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        cfg4_full_ms = float(tt.cpu())

    # ---- max over ranks ----
    tm = torch.tensor([ms_total, e2e_ms, k1_ms, upd_ms, e2e_wall_ms, e2e_blk_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, k1_ms, upd_ms, e2e_wall_ms, e2e_blk_ms = [float(x) for x in tm.cpu()]

    if rank == 0:
        peaks, peak_src = {}, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
            peak_src = "measured"
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        flops = algorithmic_flops(B, N, S, A, H1, H2)
        achieved = flops / (k1_ms * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        evals_total = world * B * N * args.steps
        line = {
            "metric": METRIC, "value": evals_total / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[prec],
            "dtype_detail": {"fp16": "f16 operands, f32 accumulate in TMEM (tcgen05 kind::f16)",
                             "bf16": "bf16 operands, f32 accumulate in TMEM (tcgen05 kind::f16)",
                             "fp32": "f32 CUDA cores"}[prec],
            "data": "synthetic",
            "config": dict(W, global_states=B * world, precision=prec,
                           l2="rotating %d input/output sets (%.0f MB > 126 MB L2), no flush kernels in the timed region"
                              % (RING, RING * 2 * B * N * 4 / 1e6),
                           parallelism=f"states sharded over {world} rank(s), no data-path collective"),
            "clocks": clocks.summary(),
            "e2e": {"value": evals_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "wall_ms_per_step": e2e_wall_ms / args.steps, "trials": 5,
                    "trial_ms_per_step_min_max": [min(trials) / args.steps, max(trials) / args.steps],
                    "blocking_value": evals_total / (e2e_blk_ms * 1e-3), "blocking_ms_per_step": e2e_blk_ms / args.steps,
                    "blocking_trial_ms_per_step_min_max": [min(blk_trials) / args.steps, max(blk_trials) / args.steps],
                    "api": "rlcontrol_b200.steps.ForwardKLGridPipeline: submit(states, mean, log_std) / result() -> (loss_b, dmean, "
                           "dlog_std), host arrays in, host arrays out, two slots in flight (step i+1 is staged and uploaded while "
                           "step i computes; every step pays its own H2D and D2H copy and its result is read on the host inside "
                           "the timed window); blocking_* = the same steps through ForwardKLGridStep.__call__ (copy, launch, wait, "
                           "copy in strict sequence); timed with the host clock"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "K1 fused T-in critic eval: k_critic_umma_grid (+ k_grid_parts pre-pass) [%s arithmetic]" % critic.tensor_arithmetic(True), "bound": "tensor",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "peak_source": f"{peak_src} bf16_tflops (burst; sustained "
                                        f"{peaks.get('bf16_tflops_sustained', 1400.0)})",
                         "flops_per_launch": flops, "ms_per_launch": k1_ms, "traffic": traffic},
            "parity": parity,
            "extra": {"critic_update_ms": upd_ms, "critic_updates_per_sec": 1e3 / upd_ms,
                      "agent_update_hot_path_ms": ms_total / args.steps + upd_ms,
                      "agent_updates_per_sec": world * 1e3 / (ms_total / args.steps + upd_ms),
                      "agent_update_definition": "cfg4 per rank: critic regression step (grads + all-reduce + Adam) "
                                                 "+ sampled-action evaluation + ForwardKL policy reduction",
                      "cfg4_full_update_ms": cfg4_full_ms, "cfg4_full_updates_per_sec": 1e3 / cfg4_full_ms,
                      "cfg4_full_update_definition": "kl_networks.ForwardKLNetwork.update_network + update_target_network, B=4096 "
                                                     "states per rank (q, v, pi networks 400-300, grid N=1024, three backward "
                                                     "passes + Adam steps), numpy minibatch in, losses out; " +
                                                     ("one CUDA graph" if world == 1 else
                                                      "global batch %d sharded over %d ranks, ONE NCCL sum all-reduce of the "
                                                      "[g_Q|g_V|g_pi] buffer per update, eager launches" % (B * world, world)),
                      "cfg1_update_ms": cfg1_ms, "cfg1_updates_per_sec": 1e3 / cfg1_ms,
                      "cfg1_definition": "README command shape (Pendulum ReverseKL: B=32 N=62 S=3 A=1 200-200): one FULL agent update "
                                         "through kl_networks.ReverseKLNetwork.update_network + update_target_network (q, v, pi "
                                         "networks, three Adam steps, Polyak), numpy minibatch in, losses out, host-synchronous, "
                                         "one CUDA-graph launch per update",
                      "cfg5_sweep8_updates_per_sec": world * 8 * 1e3 / sweep_ms, "cfg5_sweep8_ms_per_round": sweep_ms,
                      "cfg5_definition": "8 independent cfg1 agents per GPU (own handles, streams and graph each), one full update "
                                         "each per round, launched back to back then awaited; replicas only",
                      "cfg1_run_env_steps_per_sec": run_steps / run1_s,
                      "cfg5_runs8_env_steps_per_sec": world * 8 * run_steps / run8_s,
                      "device_run_definition": "whole runs of the README command on the device (Pendulum-v0 + ReverseKL, %d steps "
                                               "each): per step env.step + replay add + minibatch of 32 from the reference's index "
                                               "stream + full update_network + Polyak + sample_action, evaluation sessions (10 x "
                                               "200 greedy steps every 500 steps) inside the timed region; 1 run / 8 interleaved "
                                               "runs per GPU (rlcontrol_b200.device_loop)" % run_steps,
                      "critic_update_rows_per_rank": B,
                      "critic_update_allreduce": "nccl sum of theta_Q grads" if world > 1 else "none (1 rank)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_arm(params, s_np, a_np, w_np, (mean_np, lstd_np), alpha)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_LINE_OUT = [None]     # the process's original stdout (set in main)


def emit(line: dict):
    out = _LINE_OUT[0] or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    # stdout carries exactly ONE line (the JSON): libraries that write to file descriptor 1 themselves (NCCL prints its
    # version banner there when NCCL_DEBUG is set) are sent to stderr; the JSON line goes to the saved descriptor
    sys.stdout.flush()
    _LINE_OUT[0] = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
