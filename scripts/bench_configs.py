"""Secondary measurements on the other BASELINE.json configs (cfg2 Actor-Expert, cfg3 QT-Opt CEM), beside
the oracle's CPU restatement of the same arithmetic (TF 1.15 is not installable: 'reference-equivalent CPU
restatement, not TF').  One JSON line per config.  bench.py stays the headline (cfg4)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rlcontrol_b200 as rb
from oracle import oracle_np as onp


def gpu_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def tmid_params(rng, S, A, H1, H2):
    k1, k2 = np.sqrt(3 / S), np.sqrt(3 / (H1 + A))
    u = lambda k, *sh: rng.uniform(-k, k, sh).astype(np.float32)
    return [u(k1, S, H1), u(k1, H1), u(k2, H1 + A, H2), u(k2, H2), u(0.3, H2, 1), u(0.3, 1)]


def main():
    eng = rb.Engine(0)
    rng = np.random.RandomState(0)
    # ---- cfg3: QT-Opt CEM, S=17, A=6, B=256, N=1024, 3 iterations, 400-300 T-mid, top_m=6, num_modal=2
    S, A, H1, H2, B, N, iters, top_m, M = 17, 6, 400, 300, 256, 1024, 3, 6, 2
    p = tmid_params(rng, S, A, H1, H2)
    smin, smax = -10 * np.ones(S), 10 * np.ones(S)
    cr = rb.Critic(eng, rb.TMID, S, A, H1, H2, smin, smax).load(*p, rb.LAYOUT_IN_OUT)
    s = rng.randn(B, S).astype(np.float32)
    u0 = rng.uniform(size=(B, N, A)).astype(np.float32)
    noise = rng.randn(iters - 1, B, N, A).astype(np.float32)
    cu = rng.uniform(size=(iters - 1, B, N)).astype(np.float32)
    t = lambda x: torch.as_tensor(x, device="cuda")
    sd, u0d, nd, cud = t(s), t(u0), t(noise), t(cu)
    dt = gpu_time(lambda: cr.cem(sd, u0d, nd, cud, top_m, M, -np.ones(A), np.ones(A)))
    # CPU: oracle restatement on a bounded sample of the states (un-hoisted critic as TF runs it + numpy EM)
    bs = 16
    qf = lambda st, ac: onp.tmid_eval(st, ac, p, smin, smax)
    t0 = time.perf_counter()
    onp.cem_iterate(qf, s[:bs], u0[:bs], noise[:, :bs], cu[:, :bs], top_m, M, -np.ones(A), np.ones(A))
    cpu = (time.perf_counter() - t0) * B / bs
    print(json.dumps({"config": "cfg3 QT-Opt CEM B=256 N=1024 x3 iters 400-300 T-mid (one predict_action)",
                      "gpu_ms": dt * 1e3, "gpu_q_evals_per_sec": B * N * iters / dt,
                      "cpu_oracle_ms_extrapolated": cpu * 1e3, "cpu_sample": f"{bs} of {B} states, numpy oracle (1 thread BLAS as configured)",
                      "note": "survey probe of the real reference (sklearn GMM fit+sample 20 ms each): 15.6 s + 1.6 s of critic evals per call"}))
    # ---- cfg2: Actor-Expert, S=1, A=1, B=32, N=120, k=6, 200-200 T-mid: predict_q on the stack + per-state top-k + gather
    S, A, H1, H2, B, N, k = 1, 1, 200, 200, 32, 120, 6
    p = tmid_params(rng, S, A, H1, H2)
    cr = rb.Critic(eng, rb.TMID, S, A, H1, H2).load(*p, rb.LAYOUT_IN_OUT)
    s = rng.randn(B, S).astype(np.float32)
    acts = rng.uniform(-1, 1, (B, N, A)).astype(np.float32)
    sd, ad = t(s), t(acts)

    def ae():
        q = cr.eval(sd, ad, "fp32")
        return eng.topk(q, k, ad)
    dt = gpu_time(ae, reps=200)
    t0 = time.perf_counter()
    for _ in range(20):
        q = onp.tmid_eval(s, acts, p)
        idx = onp.topk_desc(q, k)
        onp.gather_elites(acts, idx)
    cpu = (time.perf_counter() - t0) / 20
    print(json.dumps({"config": "cfg2 Actor-Expert B=32 N=120 k=6 200-200 T-mid (predict_q + top-k + elite gather)",
                      "gpu_ms": dt * 1e3, "gpu_q_evals_per_sec": B * N / dt, "cpu_oracle_ms": cpu * 1e3,
                      "note": "launch-latency-bound on the GPU (2 kernels + host overhead per call)"}))
    # the same step with the sampling on the device too: rlc_ae_expert_step (sample -> Q -> top-k -> gather, one launch
    # after the state-term launch), preallocated draws; CPU side = numpy choice/normal + the above
    M = 1
    alpha = np.ones((B, M), np.float32)
    mean = np.tanh(rng.randn(B, M, A)).astype(np.float32)
    sigma = np.exp(rng.uniform(-2, 0, (B, M, A))).astype(np.float32)
    cu, nz = rng.random_sample((B, N)).astype(np.float32), rng.standard_normal((B, N, A)).astype(np.float32)
    al, me, si, cud, nzd = t(alpha), t(mean), t(sigma), t(cu), t(nz)
    lo, hi = -np.ones(A), np.ones(A)
    dt = gpu_time(lambda: cr.ae_expert_step(sd, k, al, me, si, cud, nzd, lo, hi), reps=200)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        cr.ae_expert_step(sd, k, al, me, si, cud, nzd, lo, hi)
    e1.record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        a2, _ = onp.mixture_sample(alpha, mean, sigma, cu, nz, lo, hi)
        q = onp.tmid_eval(s, a2.astype(np.float32), p)
        onp.gather_elites(a2, onp.topk_desc(q, k))
    cpu = (time.perf_counter() - t0) / 20
    print(json.dumps({"config": "cfg2 Actor-Expert fused expert step (mixture sampling + predict_q + top-k + elite gather), B=32 N=120 k=6",
                      "gpu_ms_host_loop": dt * 1e3, "gpu_ms_device": e0.elapsed_time(e1) / 200, "cpu_oracle_ms": cpu * 1e3}))


if __name__ == "__main__":
    main()
