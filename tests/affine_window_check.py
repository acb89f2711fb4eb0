"""Empirical check of the 4-bit trace window of the packed affine kernels (test infrastructure; numpy only).
For every scoring the host admits to 4 trace bits (packed_affine_trace_bits in seqa_cuda.cu) fill Gotoh matrices of
random / adversarial pairs and record the largest value of every difference the walk tests on low nibbles:
  D1 = H(i,j) - sim - H(i-1,j-1)   (diag test, must be in [0,15])      D2 = H(i,j) - Ix(i,j)  ([0,15]; same for Iy)
  D3 = Ix(i,j) - ge - Ix(i-1,j)    (extend test, [0,15]; same for Iy)  D4 = H(i,j) - H(i,j-1) (row scan, [-8,7])"""
import itertools, sys
import numpy as np

NEG = -10000

def fill(a, b, go, ge, m, x, allow, local):
    M, N = len(a), len(b)
    H = np.zeros((M + 1, N + 1), np.int64); Ix = np.full((M + 1, N + 1), NEG, np.int64); Iy = Ix.copy()
    if not local:
        for i in range(1, M + 1): H[i, 0] = go + i * ge
        for j in range(1, N + 1): H[0, j] = go + j * ge
    for i in range(1, M + 1):
        for j in range(1, N + 1):
            Ix[i, j] = max(H[i - 1, j] + go + ge, Ix[i - 1, j] + ge)
            Iy[i, j] = max(H[i, j - 1] + go + ge, Iy[i, j - 1] + ge)
            eq = a[i - 1] == b[j - 1]
            d = H[i - 1, j - 1] + (m if eq else -x) if (allow or eq) else -(1 << 40)
            v = max(d, Ix[i, j], Iy[i, j])
            H[i, j] = max(v, 0) if local else v
    return H, Ix, Iy

def main():
    rng = np.random.default_rng(1)
    worst = {}
    combos = [(go, ge, m, x, allow) for go in (0, -1, -3, -5) for ge in (-1, -2, -3) for m in (1, 2, 3, 4) for x in (1, 2, 4)
              for allow in (True, False)]
    n = 0
    for go, ge, m, x, allow in combos:
        g = -(go + ge)
        xx = x if allow else 0
        if not (m + xx + 2 * g <= 12 and m + g <= 7):
            continue
        n += 1
        for local in (False, True):
            for trial in range(6):
                L1, L2 = rng.integers(5, 40, 2)
                alpha = "ACGT" if trial % 3 else "AC"
                a = "".join(alpha[k] for k in rng.integers(0, len(alpha), L1))
                b = ("".join(alpha[k] for k in rng.integers(0, len(alpha), L2)) if trial != 4 else a[: L1 // 2] + a[L1 // 2 + 3:])
                if trial == 5:
                    a, b = "A" * L1, "A" * L2
                if not b: b = "A"
                H, Ix, Iy = fill(a, b, go, ge, m, x, allow, local)
                M, N = len(a), len(b)
                for i in range(1, M + 1):
                    for j in range(1, N + 1):
                        eq = a[i - 1] == b[j - 1]
                        if (allow or eq) and i > 1 and j > 1:
                            d1 = H[i, j] - (m if eq else -x) - H[i - 1, j - 1]
                            worst["D1"] = max(worst.get("D1", 0), d1); assert d1 >= 0
                        for P in (Ix, Iy):
                            d2 = H[i, j] - P[i, j]
                            worst["D2"] = max(worst.get("D2", 0), d2); assert d2 >= 0
                        if i > 1:
                            d3 = Ix[i, j] - ge - Ix[i - 1, j]
                            worst["D3"] = max(worst.get("D3", 0), d3); assert d3 >= 0
                        if j > 1:
                            d3 = Iy[i, j] - ge - Iy[i, j - 1]
                            worst["D3"] = max(worst.get("D3", 0), d3)
                            d4 = H[i, j] - H[i, j - 1]
                            worst["D4max"] = max(worst.get("D4max", -99), d4); worst["D4min"] = min(worst.get("D4min", 99), d4)
    print("scorings admitted:", n, "worst:", {k: int(v) for k, v in worst.items()})
    assert worst["D1"] <= 15 and worst["D2"] <= 15 and worst["D3"] <= 15 and -8 <= worst["D4min"] and worst["D4max"] <= 7
    print("ok: every difference fits the 4-bit window")

if __name__ == "__main__":
    main()
