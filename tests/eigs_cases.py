"""Input matrices shared by the eigs parity tests (GPU, host emulation, golden fixtures). Seeded; the shapes cover
ArrayKernel.eigs' branches: real roots, complex pairs, already-Hessenberg / triangular input (zero Householder
columns), the all-zero matrix (norm == 0: no back-substitution, LinearAlgebraOpsEigs.cpp:457-459), repeated roots,
plane rotations (purely imaginary pairs), a companion matrix, badly scaled rows and the ad hoc shifts."""
import numpy as np


def cases(sizes=(1, 2, 3, 5, 8, 16, 33, 64)):
    out = []
    for n in sizes:  # seeded per size: a subset of sizes gives the same matrices
        out.append(("uniform%d" % n, np.random.default_rng(20261020 + n).uniform(0, 1, (n, n))))  # MatrixTest.java:208-221
    rng = np.random.default_rng(20261021)  # the special matrices do not depend on `sizes`
    out.append(("zeros1", np.zeros((1, 1))))  # the one all-zero matrix the reference's hqr2 returns from
    out.append(("eye5", np.eye(5)))
    out.append(("upper6", np.triu(rng.uniform(-1, 1, (6, 6)))))
    out.append(("lower6", np.tril(rng.uniform(-1, 1, (6, 6)))))
    s = rng.uniform(-1, 1, (12, 12))
    out.append(("symmetric12", s + s.T))
    out.append(("skew10", s[:10, :10] - s[:10, :10].T))
    c, si = np.cos(0.3), np.sin(0.3)
    rot = np.eye(7)
    rot[2:4, 2:4] = [[c, -si], [si, c]]
    rot[5:7, 5:7] = [[0.0, -2.0], [2.0, 0.0]]
    out.append(("rotations7", rot))
    comp = np.zeros((9, 9))
    comp[0, :] = -rng.uniform(-1, 1, 9)
    comp[np.arange(1, 9), np.arange(0, 8)] = 1.0
    out.append(("companion9", comp))
    out.append(("scaled11", rng.uniform(-1, 1, (11, 11)) * (10.0 ** rng.integers(-6, 7, (11, 1)))))
    out.append(("repeated6", np.diag([2.0, 2.0, 2.0, -1.0, -1.0, 0.0]) + np.diag([1.0, 0, 0, 1.0, 0], 1)))
    # slow convergence: a cyclic shift has all its eigenvalues on the unit circle (exercises the iter == 10 shift)
    cyc = np.zeros((8, 8))
    cyc[np.arange(8), (np.arange(8) + 1) % 8] = 1.0
    out.append(("cyclic8", cyc))
    out.append(("hilbert10", 1.0 / (np.arange(10)[:, None] + np.arange(10)[None, :] + 1.0)))
    out.append(("normal24", rng.standard_normal((24, 24))))
    return [(name, np.ascontiguousarray(a, dtype=np.float64)) for name, a in out]


def non_converging():
    """Inputs on which the reference's uncapped hqr2 loop never returns (do NOT hand these to oracle.ref())."""
    nan = np.random.default_rng(5).uniform(0, 1, (5, 5))
    nan[3, 2] = np.nan  # spreads over the whole matrix in the Hessenberg reduction; NaN < eps * s never holds
    return [("zeros4", np.zeros((4, 4))), ("nan5", nan)]


def residual(a, vec, val):
    """MatrixTest.java:216-219: sum |A V - V D| with D block diagonal as RealArray.mEigs builds it
    (RealArray.java:239-262)."""
    n = a.shape[0]
    d = np.zeros((n, n))
    i = 0
    while i < n:
        if i < n - 1 and val[2 * i] == val[2 * i + 2] and val[2 * i + 1] > 0 and val[2 * i + 3] < 0:
            d[i, i], d[i, i + 1], d[i + 1, i + 1], d[i + 1, i] = val[2 * i], val[2 * i + 1], val[2 * i + 2], val[2 * i + 3]
            i += 2
        else:
            d[i, i] = val[2 * i]
            i += 1
    v = vec.reshape(n, n)
    return np.abs(a @ v - v @ d).sum()
