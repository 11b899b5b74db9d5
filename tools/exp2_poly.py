"""Design aid for the queued attention work (performance queue item 2): coefficients of the polynomial that
would evaluate a share of the softmax exponentials on the FMA pipe instead of MUFU.EX2.

2^x = 2^floor(x) * 2^f, f in [0, 1): the integer part goes straight into the exponent field, 2^f is a low-degree
polynomial.  P is stored as fp16 (relative precision 2^-11 = 4.9e-4), so a max relative error of ~1e-4 is invisible.
Prints near-minimax coefficients (least squares on Chebyshev nodes, then a few Remez-style exchange steps) and the
max relative error in float32 Horner arithmetic, for degrees 2..4.
"""
import numpy as np


def fit(deg, iters=20):
    k = np.arange(4096)
    x = 0.5 - 0.5 * np.cos(np.pi * (k + 0.5) / k.size)  # Chebyshev nodes on [0, 1]
    w = np.ones_like(x)
    for _ in range(iters):  # iteratively re-weighted least squares on the relative error -> near-minimax
        A = np.vander(x, deg + 1, increasing=True) / (2.0 ** x)[:, None]
        c, *_ = np.linalg.lstsq(A * w[:, None], w, rcond=None)
        err = np.abs(A @ c - 1.0)
        w = w * (1.0 + 4.0 * err / err.max())
    return c


def max_rel_err_f32(c):
    x = np.linspace(0.0, 1.0, 1 << 20, endpoint=False, dtype=np.float32)
    acc = np.full_like(x, np.float32(c[-1]))
    for a in c[-2::-1]:
        acc = acc * x + np.float32(a)  # one FFMA per coefficient
    return float(np.max(np.abs(acc.astype(np.float64) / 2.0 ** x.astype(np.float64) - 1.0)))


if __name__ == "__main__":
    for deg in (2, 3, 4):
        c = fit(deg)
        print(f"degree {deg}: max rel err (fp32 Horner) {max_rel_err_f32(c):.3e}  coeffs " +
              " ".join(f"{v:.9f}" for v in c))
