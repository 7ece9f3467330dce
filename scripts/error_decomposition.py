"""Error decomposition of the tensor-core arithmetic of K1-grid on the cfg4 bench workload (CPU only, fp64).

Each operand rounding of the stated arithmetic (oracle_np.tin_eval_rounded(head="grid")) is toggled on its
own and in the combinations the strict modes use; every variant is compared with the exact fp64 oracle on
the metric of tests/conftest.py (|dq| / max(|q|, rms_state q)).  Output: one JSON line per variant
(profiles/r02_error_decomposition.jsonl).

    python scripts/error_decomposition.py [n_states]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402
from conftest import rel_err, golden  # noqa: E402

r16 = lambda z: onp.round_operand(z, "fp16")
rbf = lambda z: onp.round_operand(z, "bf16")


def split(x, r=r16, terms=2):
    """x ~ sum of `terms` operand-type values (hi, lo, ...), each the rounding of the remainder."""
    out, rem = [], np.asarray(x, np.float64)
    for _ in range(terms):
        h = r(rem)
        out.append(h)
        rem = rem - h
    return out


def variants(s, a, params):
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, np.float64) for p in params]
    S = s.shape[1]
    B, N = s.shape[0], a.shape[0]
    f32 = np.float32
    ps32 = (np.asarray(s, np.float64) @ W1[:, :S].T + b1).astype(f32).astype(np.float64)     # fp32 tables
    pa32 = (np.asarray(a, np.float64) @ W1[:, S:].T).astype(f32).astype(np.float64)
    w3 = W3.reshape(-1)
    mx = np.abs(w3).max()
    scale = np.ldexp(1.0, 1 - int(np.frexp(mx)[1]))
    sw = scale * np.abs(w3)
    W2s = (sw[:, None] * W2).astype(f32).astype(np.float64)      # folded head, fp32 product
    bs = (sw * b2).astype(f32).astype(np.float64)
    sign = np.where(w3 < 0, -1.0, 1.0)

    def finish(h1_terms, W_terms, pairs, b_term):
        """z = sum over the (i, j) operand-term pairs of h1_i . W_j + bias; q = folded head."""
        z = np.zeros((B * N, W2.shape[0]))
        for (i, j) in pairs:
            z += h1_terms[i] @ W_terms[j].T
        z = np.maximum(z + b_term, 0)
        return ((z @ sign) / scale + b3.reshape(())).reshape(B, N)

    pre_exact = np.maximum(ps32[:, None, :] + pa32[None, :, :], 0).reshape(B * N, -1)
    pre_3r = r16(np.maximum(r16(ps32)[:, None, :] + r16(pa32)[None, :, :], 0)).reshape(B * N, -1)
    out = {}
    # the shipped fast mode and its single toggles
    out["fast: r(r(PS)+r(PA)), r(W2') [shipped fp16 grid mode]"] = finish([pre_3r], [r16(W2s)], [(0, 0)], r16(bs))
    out["tables fp16 x3 roundings only (W2' exact)"] = finish([pre_3r], [W2s], [(0, 0)], bs)
    out["h1 one rounding only (fp32 tables; W2' exact)"] = finish([r16(pre_exact)], [W2s], [(0, 0)], bs)
    out["W2' rounding only (h1 exact)"] = finish([pre_exact], [r16(W2s)], [(0, 0)], r16(bs))
    out["fp32 tables + r(h1), r(W2') [1 MMA]"] = finish([r16(pre_exact)], [r16(W2s)], [(0, 0)], r16(bs))
    # split modes
    hh = split(pre_exact)
    ww = split(W2s)
    bb = split(bs)
    out["h1 hi+lo, r(W2') [2 MMA]"] = finish(hh, [r16(W2s)], [(0, 0), (1, 0)], r16(bs))
    out["r(h1), W2' hi+lo [2 MMA]"] = finish([r16(pre_exact)], ww, [(0, 0), (0, 1)], bb[0] + bb[1])
    out["h1 hi+lo, W2' hi+lo, no lo.lo [3 MMA]"] = finish(hh, ww, [(0, 0), (1, 0), (0, 1)], bb[0] + bb[1])
    # bf16 three-term for reference
    hb = split(pre_exact, rbf)
    wb = split(W2s, rbf)
    out["bf16 h1 hi+lo, W2' hi+lo, no lo.lo [3 MMA]"] = finish(hb, wb, [(0, 0), (1, 0), (0, 1)], wb[0][:, 0] * 0 + bs)
    return out


def main():
    nB = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    W = bench.WORKLOAD
    rng = np.random.RandomState(0)
    params = bench.make_params(rng, W["S"], W["A"], W["H1"], W["H2"])
    s, a, w, _ = bench.make_inputs(np.random.RandomState(1000), W["B_per_gpu"], W["N"], W["S"], W["A"])
    a = bench.make_inputs(np.random.RandomState(1), 1, W["N"], W["S"], W["A"])[1]
    rows = np.arange(0, W["B_per_gpu"], W["B_per_gpu"] // nB)[:nB]
    cases = [("cfg4 bench workload", s[rows], a, params)]
    g = golden("tin_400_300.npz")
    cases.append(("golden tin_400_300", g["s"], g["a"], [g[k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")]))
    lines = []
    for name, s_, a_, p_ in cases:
        ref = onp.tin_eval(s_, a_, p_, dtype=np.float64)
        for k, q in variants(s_, a_, p_).items():
            e = rel_err(q, ref)
            rec = {"case": name, "rows": int(ref.size), "variant": k, "rel_err_rms": float(np.sqrt((e ** 2).mean())),
                   "rel_err_max": float(e.max())}
            lines.append(rec)
            print(json.dumps(rec))
    with open(os.path.join(ROOT, "profiles", "r02_error_decomposition.jsonl"), "w") as f:
        for rec in lines:
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
