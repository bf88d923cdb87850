"""shared helpers of the parity tests"""
import numpy as np


def rel_inf(a, b):
    """difference relative to the vector's infinity norm (elements pass through 0)"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


class V2Row:
    def __init__(self, rows, N, M):
        self.it = rows[:, 0]; self.mu = rows[:, 1]; self.beta = rows[:, 2:2 + M]
        self.sigmaE = rows[:, 2 + M]; self.sigmaG = rows[:, 3 + M]
        self.comp = rows[:, 4 + M:4 + 2 * M]; self.eps = rows[:, 4 + 2 * M:4 + 2 * M + N]


class GroupsRow:
    def __init__(self, rows, N, M, G, F, restart=False):
        self.it = rows[:, 0]; self.mu = rows[:, 1]; self.beta = rows[:, 2:2 + M]
        self.sigmaE = rows[:, 2 + M]; self.comp = rows[:, 3 + M:3 + 2 * M]
        self.sigmaG = rows[:, 3 + 2 * M:3 + 2 * M + G]
        self.eps = rows[:, 3 + 2 * M + G:3 + 2 * M + G + N]
        if not restart:
            self.alpha = rows[:, 3 + 2 * M + G + N:3 + 2 * M + G + N + F]
            self.sigmaF = rows[:, 3 + 2 * M + G + N + F]


class HsRow:
    def __init__(self, rows, N, M):
        self.it = rows[:, 0]; self.mu = rows[:, 1]; self.beta = rows[:, 2:2 + M]
        self.sigmaE = rows[:, 2 + M]; self.tau = rows[:, 3 + M]
        self.lam = rows[:, 4 + M:4 + 2 * M]; self.eps = rows[:, 4 + 2 * M:4 + 2 * M + N]


def assert_trace_close(name, got, want, tol=1e-9):
    """per-iteration comparison, each iteration relative to that iteration's infinity norm"""
    got = np.atleast_2d(got.T).T if got.ndim == 1 else got
    want = np.atleast_2d(want.T).T if want.ndim == 1 else want
    assert got.shape == want.shape, (name, got.shape, want.shape)
    worst = 0.0
    for t in range(got.shape[0]):
        worst = max(worst, rel_inf(got[t], want[t]))
    assert worst <= tol, "%s: worst per-iteration relative error %.3e > %.1e" % (name, worst, tol)
    return worst
