"""Independent Python big-int restatement of the arithmetic underneath the hot path,
used only to cross-check the C++ oracle (and, through it, the CUDA kernels).
Nothing here is derived from oracle/ code: plain integers mod l / mod q."""
L = 2**252 + 27742317777372353535851937790883648493
Q = 2**255 - 19
D = (-121665 * pow(121666, -1, Q)) % Q
BX = 15112221349535400772501151409588531511454012693041857206046113283949847762202
BY = (4 * pow(5, -1, Q)) % Q


def eq_evals(r):
    """naive bitwise product, r[0] <-> MSB (dense_mlpoly.rs:499-516 compute_chis_at_r)"""
    ell = len(r)
    out = []
    for i in range(1 << ell):
        acc = 1
        for j in range(ell):
            bit = (i >> (ell - 1 - j)) & 1
            acc = acc * (r[j] if bit else (1 - r[j])) % L
        out.append(acc)
    return out


def bind_top(Z, r):
    n = len(Z) // 2
    return [(Z[i] + r * (Z[i + n] - Z[i])) % L for i in range(n)]


def bind_bot(Z, r):
    n = len(Z) // 2
    return [(Z[2 * i] + r * (Z[2 * i + 1] - Z[2 * i])) % L for i in range(n)]


def interpolate(evals):
    """coefficients of the unique poly of degree < n through (0, e0), (1, e1), ... (Lagrange, mod L)"""
    n = len(evals)
    coeffs = [0] * n
    for i in range(n):
        # basis polynomial l_i(x) = prod_{j != i} (x - j) / (i - j)
        num = [1]
        den = 1
        for j in range(n):
            if j == i:
                continue
            num = [(a - j * b) % L for a, b in zip([0] + num, num + [0])]
            den = den * (i - j) % L
        scale = evals[i] * pow(den, -1, L) % L
        for k in range(n):
            coeffs[k] = (coeffs[k] + num[k] * scale) % L
    return coeffs


# ---- twisted Edwards -x^2 + y^2 = 1 + d x^2 y^2, affine formulas ----
def te_add(P1, P2):
    x1, y1 = P1
    x2, y2 = P2
    t = D * x1 * x2 * y1 * y2 % Q
    x3 = (x1 * y2 + y1 * x2) * pow(1 + t, -1, Q) % Q
    y3 = (y1 * y2 + x1 * x2) * pow(1 - t, -1, Q) % Q
    return (x3, y3)


def te_mul(P, k):
    acc = (0, 1)
    while k:
        if k & 1:
            acc = te_add(acc, P)
        P = te_add(P, P)
        k >>= 1
    return acc


def te_msm(points, scalars):
    acc = (0, 1)
    for P, k in zip(points, scalars):
        acc = te_add(acc, te_mul(P, k % L))
    return acc


def rfc8032_encode(P):
    x, y = P
    return (y | ((x & 1) << 255)).to_bytes(32, "little")


def ark_encode(P):
    """ark-serialize TE compression: flag set iff x > -x as canonical integers"""
    x, y = P
    neg = x > (Q - x) % Q
    return (y | (int(neg) << 255)).to_bytes(32, "little")
