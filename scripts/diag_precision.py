"""Diagnostic: how exactly does the tcgen05 path follow 'fp16 operands, fp32 accumulate'?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import rlcontrol_b200 as rb
from oracle import oracle_np as onp
from conftest import golden, rel_err

eng = rb.Engine(0)

def tin(p, S, A, H1, H2):
    return rb.Critic(eng, rb.TIN, S, A, H1, H2).load(*p, rb.LAYOUT_OUT_IN)

# 1. integer exactness: every product and partial sum is exactly representable
rng = np.random.RandomState(0)
for (S, A, H1, H2) in ((3, 1, 200, 200), (17, 6, 400, 300)):
    ints = lambda lo, hi, *sh: rng.randint(lo, hi, sh).astype(np.float32)
    p = [ints(-2, 3, H1, S + A), ints(-2, 3, H1), ints(-1, 2, H2, H1), ints(-2, 3, H2), ints(-2, 3, 1, H2), ints(-2, 3, 1)]
    s, a = ints(-3, 4, 16, S), ints(-3, 4, 128, A)
    ref = onp.tin_eval(s, a, p, dtype=np.float64)
    for prec in ("fp16", "bf16", "fp32"):
        q = tin(p, S, A, H1, H2).eval(s, a, prec).cpu().numpy()
        print(f"[ints {H1}-{H2}] {prec}: max |dq| = {np.abs(q - ref).max():.3e}  (|q| max {np.abs(ref).max():.0f})  umma_err={eng.umma_error()}")

# 2. accumulate dynamic range: q = sum_j h1_j * W2[0,j] with one big and many small terms
S, A, H1, H2 = 1, 1, 400, 32
for big in (1.0, 64.0, 2048.0, 32768.0):
    for small in (2.0 ** -10, 2.0 ** -14):
        W1 = np.zeros((H1, 2), np.float32); b1 = np.ones(H1, np.float32); b1[0] = big
        W2 = np.zeros((H2, H1), np.float32); W2[0, :] = small; W2[0, 0] = 1.0
        b2 = np.zeros(H2, np.float32); W3 = np.zeros((1, H2), np.float32); W3[0, 0] = 1.0; b3 = np.zeros(1, np.float32)
        p = [W1, b1, W2, b2, W3, b3]
        s = np.zeros((2, 1), np.float32); a = np.zeros((128, 1), np.float32)
        exact = big + (H1 - 1) * small
        q = tin(p, S, A, H1, H2).eval(s, a, "fp16").cpu().numpy()
        print(f"[range] big={big:8.0f} small=2^{int(np.log2(small))}: exact {exact:.8f} gpu {q[0,0]:.8f} rel err {(q[0,0]-exact)/exact:+.3e}")

# 3. golden: kernel vs emulations
for name, dims in (("tin_cfg1.npz", (3, 1, 200, 200)), ("tin_400_300.npz", (17, 6, 400, 300))):
    g = golden(name)
    p = [g[k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")]
    q = tin(p, *dims).eval(g["s"], g["a"], "fp16").cpu().numpy()
    rnd = onp.tin_eval_rounded(g["s"], g["a"], p, "fp16")
    e = rel_err(q, rnd)
    print(f"[{name}] vs rounded-oracle: max {e.max():.3e} rms {np.sqrt((e**2).mean()):.3e}; vs exact max {rel_err(q, g['q']).max():.3e}")
    # emulate truncating accumulation: after each K=16 block, keep `bits` mantissa bits (round toward zero)
    W1, b1, W2, b2, W3, b3 = [np.asarray(x, np.float64) for x in p]
    s, a = g["s"], g["a"]; B = s.shape[0]; N = a.shape[0]
    x = np.concatenate([onp.stack_state_major(s, N), onp.stack_actions(a, B)], 1)
    r = lambda z: onp.round_operand(z, "fp16")
    h1 = r(np.maximum(r(x) @ r(W1).T + r(b1), 0))
    W2r = r(W2)
    def trunc(v, bits):
        m, ex = np.frexp(v)
        return np.ldexp(np.trunc(m * 2.0 ** bits) / 2.0 ** bits, ex)
    for bits in (24, 16, 13, 12, 11):
        acc = np.zeros((h1.shape[0], W2r.shape[0]))
        for k0 in range(0, h1.shape[1], 16):
            acc = trunc(acc + h1[:, k0:k0 + 16] @ W2r[:, k0:k0 + 16].T, bits)
        qe = (np.maximum(acc + b2, 0) @ W3.reshape(-1) + b3.reshape(())).reshape(B, N)
        print(f"    acc truncated to {bits} bits per K16 block: kernel-vs-emu max {rel_err(q, qe).max():.3e}")
