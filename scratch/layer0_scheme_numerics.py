"""
Numerics of a cheaper layer-0 operand scheme for ffae_infer_tc.cu (CPU emulation, no GPU needed).

Today (224 TMEM columns per tile slot):   D = tf32(A_lo)*W_hi + A_hi*W_hi            (kind::tf32, 64 + 64 columns of A)
                                            + bf16(A)*bf16(W - W_hi)                 (kind::f16,  32 columns)
Candidate (128 columns with A_hi read from the x box by an SS-form MMA, scratch/ffae_infer_tc_v16_ahi_from_xbox.cu):
                                          D = A_hi*W_hi                              (kind::tf32, A from shared memory)
                                            + bf16(A_lo)*bf16(W_hi)                  (kind::f16,  32 columns)
                                            + bf16(A)*bf16(W - W_hi)                 (kind::f16,  32 columns)
A_hi = A with the low 13 mantissa bits cleared (what the tensor core does to fp32 data), A_lo = A - A_hi (exact),
W_hi = W rounded to TF32.  16 MMA instructions for K=64 instead of 20, and three (even four) tile slots fit in TMEM.

Prints the error of both schemes against float64 for layer 0 of the BASELINE net (64 -> 53) on data of several magnitudes.
"""
import numpy as np


def trunc_tf32(a):
    return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def round_tf32(a):
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def round_bf16(a):
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def dot32(a, w):
    """fp32 accumulation of exact products (products of <=11-bit x <=11-bit operands are exact in fp32)."""
    acc = np.zeros((a.shape[0], w.shape[1]), np.float32)
    for k in range(a.shape[1]):
        acc += (a[:, k:k + 1].astype(np.float64) * w[k:k + 1, :].astype(np.float64)).astype(np.float32)
    return acc


def main():
    rng = np.random.default_rng(0)
    K, N, rows = 64, 53, 4096
    lim = np.sqrt(6.0 / (K + N))
    W = rng.uniform(-lim, lim, (K, N)).astype(np.float32)
    W_hi = round_tf32(W)
    W_lo = (W - W_hi).astype(np.float32)
    cases = {
        "unit range [0,1]": rng.random((rows, K)),
        "min-max scaled sensor (0.5 +- 0.4)": 0.5 + 0.4 * np.sin(rng.random((rows, K)) * 6),
        "offset dominated (1000 +- 1)": 1000 + rng.normal(0, 1, (rows, K)),
        "large (1e6 scale)": rng.normal(0, 1e6, (rows, K)),
        "tiny (1e-4 scale)": rng.normal(0, 1e-4, (rows, K)),
    }
    print(f"{'data':38s} {'|z| rms':>10s} {'today: max err':>15s} {'rms':>10s} {'candidate: max':>15s} {'rms':>10s}   (relative to |z| rms)")
    for name, x in cases.items():
        A = x.astype(np.float32)
        want = A.astype(np.float64) @ W.astype(np.float64)
        A_hi = trunc_tf32(A)
        A_lo = (A - A_hi).astype(np.float32)
        corr = dot32(round_bf16(A), round_bf16(W_lo))
        today = dot32(trunc_tf32(A_lo), W_hi) + dot32(A_hi, W_hi) + corr
        cand = dot32(A_hi, W_hi) + dot32(round_bf16(A_lo), round_bf16(W_hi)) + corr
        mag = np.sqrt((want ** 2).mean())
        e0, e1 = np.abs(today - want) / mag, np.abs(cand - want) / mag
        print(f"{name:38s} {mag:10.3e} {e0.max():15.2e} {np.sqrt((e0**2).mean()):10.2e} {e1.max():15.2e} {np.sqrt((e1**2).mean()):10.2e}")


if __name__ == "__main__":
    main()
