"""ctypes binding of the CPU oracle (oracle/bayesrr_oracle.c).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the shipped package bayesrrcpp_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

S_INIT_U, S_MU, S_MARK_U, S_MARK_Z, S_GAMMA, S_PERM, S_FIX_Z, S_FIXPERM, S_HS_NU, S_HS_LAM, S_INIT_G = range(11)
SRC_PHILOX, SRC_SEQ, SRC_REPLAY = 0, 1, 2
KIND_V2, KIND_GROUPS, KIND_GRSTART, KIND_HORSESHOE = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(target="all"):
    subprocess.run(["make", "-s", "-C", _DIR, target], check=True)


def lib(native=False):
    name = "liboracle_native.so" if native else "liboracle.so"
    if name not in _LIBS:
        path = os.path.join(_DIR, name)
        if not os.path.exists(path):
            build("native" if native else "all")
        L = C.CDLL(path)
        L.orc_api_px_uniform.restype = C.c_double
        L.orc_api_px_uniform.argtypes = [C.c_uint64, C.c_int, C.c_int64, C.c_int64]
        L.orc_api_px_normal.restype = C.c_double
        L.orc_api_px_normal.argtypes = [C.c_uint64, C.c_int, C.c_int64, C.c_int64]
        L.orc_api_px_gamma.restype = C.c_double
        L.orc_api_px_gamma.argtypes = [C.c_uint64, C.c_int, C.c_int64, C.c_int64, C.c_double]
        L.orc_api_px_shuffle.restype = None
        L.orc_api_px_shuffle.argtypes = [C.c_uint64, C.c_int, C.c_int64, _ip, C.c_int64]
        L.orc_philox_raw.restype = None
        L.orc_philox_raw.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_format_header.restype = C.c_size_t
        L.orc_format_header.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_char_p, C.c_size_t]
        L.orc_format_row.restype = C.c_size_t
        L.orc_format_row.argtypes = [_dp, C.c_int64, C.c_char_p, C.c_size_t]
        _LIBS[name] = L
    return _LIBS[name]


class Tables(C.Structure):
    _fields_ = [("n_iter", C.c_int64), ("M", C.c_int64), ("F", C.c_int64), ("n_gam", C.c_int64),
                ("n_init_u", C.c_int64), ("n_init_g", C.c_int64),
                ("mark_u", _dp), ("mark_z", _dp), ("mu_z", _dp), ("gam", _dp), ("fix_z", _dp),
                ("hs_nu", _dp), ("hs_lam", _dp), ("init_u", _dp), ("init_g", _dp),
                ("perm", _ip), ("fixperm", _ip)]


class DrawTables:
    """numpy-backed draw tables (record target / replay source); layout in bayesrr_oracle.h."""

    def __init__(self, n_iter, M, n_gam, F=0, n_init_u=0, n_init_g=0, horseshoe=False):
        self.n_iter, self.M, self.F, self.n_gam = n_iter, M, F, n_gam
        nan = np.nan
        self.mark_u = np.full((n_iter, M), nan)
        self.mark_z = np.full((n_iter, M), nan)
        self.mu_z = np.full(n_iter, nan)
        self.gam = np.full((n_iter, max(n_gam, 1)), nan)
        self.fix_z = np.full((n_iter, max(F, 1)), nan)
        self.hs_nu = np.full((n_iter, M), nan) if horseshoe else None
        self.hs_lam = np.full((n_iter, M), nan) if horseshoe else None
        self.init_u = np.full(max(n_init_u, 1), nan)
        self.init_g = np.full(max(n_init_g, 1), nan)
        self.perm = np.zeros((n_iter, M), dtype=np.int32)
        self.fixperm = np.zeros((n_iter, max(F, 1)), dtype=np.int32)
        self.n_init_u, self.n_init_g = n_init_u, n_init_g

    def cstruct(self):
        def p(a):
            return a.ctypes.data_as(_dp) if a is not None else _dp()
        t = Tables(self.n_iter, self.M, self.F, self.n_gam, self.n_init_u, self.n_init_g,
                   p(self.mark_u), p(self.mark_z), p(self.mu_z), p(self.gam), p(self.fix_z),
                   p(self.hs_nu), p(self.hs_lam), p(self.init_u), p(self.init_g),
                   self.perm.ctypes.data_as(_ip), self.fixperm.ctypes.data_as(_ip))
        return t


class _V2Args(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("burn_in", C.c_int), ("thinning", C.c_int),
                ("N", C.c_int64), ("M", C.c_int64), ("X", _dp), ("Y", _dp),
                ("sigma0", C.c_double), ("v0E", C.c_double), ("s02E", C.c_double), ("v0G", C.c_double), ("s02G", C.c_double),
                ("cva", _dp), ("ncva", C.c_int), ("pi_init", _dp), ("emit_all", C.c_int), ("pi_trace", _dp)]


class _GroupsArgs(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("burn_in", C.c_int), ("thinning", C.c_int),
                ("N", C.c_int64), ("M", C.c_int64), ("X", _dp), ("Y", _dp),
                ("sigma0", C.c_double), ("v0E", C.c_double), ("s02E", C.c_double), ("v0G", C.c_double), ("s02G", C.c_double),
                ("cva", _dp), ("ncva", C.c_int), ("groups", C.c_int), ("gAssign", _ip),
                ("fixed", _dp), ("F", C.c_int64), ("emit_all", C.c_int), ("pi_trace", _dp)]


class _GrstartArgs(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("burn_in", C.c_int), ("thinning", C.c_int),
                ("mu", C.c_double), ("beta", _dp), ("sigmaE", C.c_double), ("sigmaGG", _dp),
                ("N", C.c_int64), ("M", C.c_int64), ("X", _dp), ("epsilon", _dp), ("components", _dp),
                ("sigma0", C.c_double), ("v0E", C.c_double), ("s02E", C.c_double), ("v0G", C.c_double), ("s02G", C.c_double),
                ("cva", _dp), ("ncva", C.c_int), ("groups", C.c_int), ("gAssign", _ip),
                ("emit_all", C.c_int), ("pi_trace", _dp)]


class _HsArgs(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("burn_in", C.c_int), ("thinning", C.c_int),
                ("N", C.c_int64), ("M", C.c_int64), ("X", _dp), ("Y", _dp),
                ("A", C.c_double), ("v0E", C.c_double), ("s02E", C.c_double), ("vL", C.c_double), ("vT", C.c_double),
                ("c2", C.c_double), ("vC", C.c_double), ("sC", C.c_double),
                ("emit_all", C.c_int), ("hyper_trace", _dp)]


def _f64(a, order="C"):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)) if order == "C" else np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp)


def _n_rows(max_it, burn_in, thinning, emit_all):
    if emit_all:
        return max_it
    return sum(1 for it in range(max_it) if it >= burn_in and it % thinning == 0)


NATIVE = False     # bench.py's CPU arm switches to the -O3 -march=native build for its second measurement


def _call(fn, args, source, seed, tables, record, n_rows, rowlen, want_rows=True):
    L = lib(NATIVE)
    rows = np.zeros((n_rows, rowlen)) if want_rows else None
    got = C.c_int64(0)
    secs = C.c_double(0)
    ct = tables.cstruct() if tables is not None else None
    f = getattr(L, fn)
    f.restype = C.c_int
    rc = f(C.byref(args), C.c_int(source), C.c_uint64(seed), C.byref(ct) if ct is not None else None,
           C.c_int(1 if record else 0), _p(rows) if rows is not None else _dp(), C.c_int64(n_rows if want_rows else 0),
           C.byref(got), C.byref(secs))
    return rc, rows, got.value, secs.value


def default_pi_init(cva):
    """Evident intent of src/BayesRv2.cpp:148-150: [0.5, 0.5*cva/sum(cva)]."""
    cva = np.asarray(cva, dtype=np.float64)
    return np.concatenate([[0.5], 0.5 * cva / cva.sum()])


def run_v2(X, Y, cva, max_iterations, burn_in=1, thinning=1, sigma0=0.01, v0E=1e-4, s02E=1e-3, v0G=1e-4, s02G=1e-3,
           pi_init=None, source=SRC_PHILOX, seed=1, tables=None, record=False, emit_all=True, want_rows=True):
    X = _f64(X, "F"); Y = _f64(Y); cva = _f64(cva)
    N, M = X.shape; K = len(cva) + 1
    pi0 = _f64(default_pi_init(cva) if pi_init is None else pi_init)
    pit = np.zeros((max_iterations, K))
    a = _V2Args(max_iterations, burn_in, thinning, N, M, _p(X), _p(Y), sigma0, v0E, s02E, v0G, s02G,
                _p(cva), len(cva), _p(pi0), 1 if emit_all else 0, _p(pit))
    nr = _n_rows(max_iterations, burn_in, thinning, emit_all)
    rc, rows, got, secs = _call("orc_api_v2", a, source, seed, tables, record, nr, 2 * M + 4 + N, want_rows)
    return dict(rc=rc, rows=rows, n_rows=got, pi=pit, seconds=secs, N=N, M=M, K=K)


def run_groups(X, Y, cva, groups, gAssign, fixed, max_iterations, burn_in=1, thinning=1, sigma0=0.01, v0E=1e-4,
               s02E=1e-3, v0G=1e-4, s02G=1e-3, source=SRC_PHILOX, seed=1, tables=None, record=False,
               emit_all=True, want_rows=True):
    X = _f64(X, "F"); Y = _f64(Y); cva = _f64(np.atleast_2d(cva), "F")
    N, M = X.shape; K = cva.shape[1] + 1
    fixed = _f64(fixed, "F") if fixed is not None else np.zeros((N, 0), order="F")
    F = fixed.shape[1]
    gA = np.ascontiguousarray(gAssign, dtype=np.int32)
    pit = np.zeros((max_iterations, groups, K))
    a = _GroupsArgs(max_iterations, burn_in, thinning, N, M, _p(X), _p(Y), sigma0, v0E, s02E, v0G, s02G,
                    _p(cva), K - 1, groups, gA.ctypes.data_as(_ip), _p(fixed) if F else _dp(), F,
                    1 if emit_all else 0, _p(pit))
    nr = _n_rows(max_iterations, burn_in, thinning, emit_all)
    rc, rows, got, secs = _call("orc_api_groups", a, source, seed, tables, record, nr, 2 * M + 3 + groups + N + F + 1, want_rows)
    return dict(rc=rc, rows=rows, n_rows=got, pi=pit, seconds=secs, N=N, M=M, K=K, G=groups, F=F)


def run_grstart(mu, beta, sigmaE, sigmaGG, X, epsilon, components, cva, groups, gAssign, max_iterations, burn_in=1,
                thinning=1, sigma0=0.01, v0E=1e-4, s02E=1e-3, v0G=1e-4, s02G=1e-3, source=SRC_PHILOX, seed=1,
                tables=None, record=False, emit_all=True, want_rows=True):
    X = _f64(X, "F"); cva = _f64(np.atleast_2d(cva), "F")
    N, M = X.shape; K = cva.shape[1] + 1
    beta = _f64(beta).ravel(); eps = _f64(epsilon); comp = _f64(components); sg = _f64(sigmaGG)
    gA = np.ascontiguousarray(gAssign, dtype=np.int32)
    pit = np.zeros((max_iterations, groups, K))
    a = _GrstartArgs(max_iterations, burn_in, thinning, mu, _p(beta), sigmaE, _p(sg), N, M, _p(X), _p(eps), _p(comp),
                     sigma0, v0E, s02E, v0G, s02G, _p(cva), K - 1, groups, gA.ctypes.data_as(_ip),
                     1 if emit_all else 0, _p(pit))
    nr = _n_rows(max_iterations, burn_in, thinning, emit_all)
    rc, rows, got, secs = _call("orc_api_grstart", a, source, seed, tables, record, nr, 2 * M + 3 + groups + N, want_rows)
    return dict(rc=rc, rows=rows, n_rows=got, pi=pit, seconds=secs, N=N, M=M, K=K, G=groups)


def run_horseshoe(X, Y, A, max_iterations, burn_in=1, thinning=1, v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0,
                  vC=10.0, sC=10.0, source=SRC_PHILOX, seed=1, tables=None, record=False, emit_all=True,
                  want_rows=True):
    X = _f64(X, "F"); Y = _f64(Y)
    N, M = X.shape
    ht = np.zeros((max_iterations, 3))
    a = _HsArgs(max_iterations, burn_in, thinning, N, M, _p(X), _p(Y), A, v0E, s02E, vL, vT, c2, vC, sC,
                1 if emit_all else 0, _p(ht))
    nr = _n_rows(max_iterations, burn_in, thinning, emit_all)
    rc, rows, got, secs = _call("orc_api_horseshoe", a, source, seed, tables, record, nr, 2 * M + 4 + N, want_rows)
    return dict(rc=rc, rows=rows, n_rows=got, hyper=ht, seconds=secs, N=N, M=M)


def format_header(kind, N, M, G=0, F=0):
    L = lib()
    n = L.orc_format_header(kind, N, M, G, F, None, 0)
    buf = C.create_string_buffer(n + 1)
    L.orc_format_header(kind, N, M, G, F, buf, n)
    return buf.raw[:n].decode()


def format_row(row):
    L = lib()
    row = _f64(row)
    n = L.orc_format_row(_p(row), len(row), None, 0)
    buf = C.create_string_buffer(n + 1)
    L.orc_format_row(_p(row), len(row), buf, n)
    return buf.raw[:n].decode()


def philox_raw(key, ctr):
    L = lib()
    k = (C.c_uint32 * 2)(*key); c = (C.c_uint32 * 4)(*ctr); o = (C.c_uint32 * 4)()
    L.orc_philox_raw(k, c, o)
    return list(o)


# ---------------------------------------------------------------------------------------------
# synthetic data: g ~ Binomial(2, p_j), p_j ~ U(0.05, 0.5); columns standardised with
# the N-1 sd (R scale()); 10 % causal, b ~ N(0, h2/M_causal), y = Xb + e, then y centred and scaled.
def synth(N, M, seed, h2=0.5, causal_frac=0.1):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.05, 0.5, size=M)
    G = rng.binomial(2, p, size=(N, M)).astype(np.int8)
    for j in range(M):
        while G[:, j].min() == G[:, j].max():
            G[:, j] = rng.binomial(2, p[j], size=N)
    mean = G.mean(axis=0)
    sd = G.std(axis=0, ddof=1)
    X = np.asfortranarray((G - mean) / sd)
    mc = max(1, int(round(causal_frac * M)))
    idx = rng.choice(M, mc, replace=False)
    b = np.zeros(M)
    b[idx] = rng.normal(0, np.sqrt(h2 / mc), size=mc)
    g = X @ b
    y = g + rng.normal(0, np.sqrt(max(1e-12, 1 - h2)), size=N)
    y = (y - y.mean()) / y.std(ddof=1)
    return dict(G=G, X=X, y=y, b=b, mean=mean, sd=sd)


# ---------------------------------------------------------------------------------------------
# oracle/_ref: the reference's OWN sources compiled against oracle/shim (only where /root/reference exists)
REF_LIB = os.path.join(_DIR, "_ref", "libbayesrr_ref.so")


def ref_available():
    if not os.path.exists(REF_LIB) and os.path.isdir("/root/reference/src"):
        try:
            build("ref")
        except Exception:
            return False
    return os.path.exists(REF_LIB)


def _ref():
    if "ref" not in _LIBS:
        lib()                                          # liboracle.so first: _ref links against it
        L = C.CDLL(REF_LIB)
        for f in ("ref_v2", "ref_groups", "ref_grstart", "ref_horseshoe"):
            getattr(L, f).restype = C.c_long
        _LIBS["ref"] = L
    return _LIBS["ref"]


def ref_v2(out, seed, max_it, burn_in, thinning, X, Y, cva, sigma0=0.01, v0E=1e-4, s02E=1e-3, v0G=1e-4, s02G=1e-3):
    X = _f64(X, "F"); Y = _f64(Y); cva = _f64(cva); N, M = X.shape
    nr = _n_rows(max_it, burn_in, thinning, False)
    rows = np.zeros((nr, 2 * M + 4 + N))
    n = _ref().ref_v2(os.fsencode(out), C.c_uint64(seed), max_it, burn_in, thinning, _p(X), C.c_long(N), C.c_long(M), _p(Y),
                      C.c_double(sigma0), C.c_double(v0E), C.c_double(s02E), C.c_double(v0G), C.c_double(s02G), _p(cva),
                      len(cva), _p(rows), C.c_long(nr))
    return rows, n


def ref_groups(out, seed, max_it, burn_in, thinning, X, Y, cva, groups, gAssign, fixed, sigma0=0.01, v0E=1e-4, s02E=1e-3,
               v0G=1e-4, s02G=1e-3):
    X = _f64(X, "F"); Y = _f64(Y); cva = _f64(np.atleast_2d(cva), "F"); N, M = X.shape
    fixed = _f64(fixed, "F") if fixed is not None else np.zeros((N, 0), order="F")
    F = fixed.shape[1]
    gA = np.ascontiguousarray(gAssign, dtype=np.int32)
    nr = _n_rows(max_it, burn_in, thinning, False)
    rows = np.zeros((nr, 2 * M + 3 + groups + N + F + 1))
    n = _ref().ref_groups(os.fsencode(out), C.c_uint64(seed), max_it, burn_in, thinning, _p(X), C.c_long(N), C.c_long(M), _p(Y),
                          C.c_double(sigma0), C.c_double(v0E), C.c_double(s02E), C.c_double(v0G), C.c_double(s02G), _p(cva),
                          cva.shape[1], groups, gA.ctypes.data_as(_ip), _p(fixed) if F else _dp(), C.c_long(F), _p(rows), C.c_long(nr))
    return rows, n


def ref_grstart(out, seed, max_it, burn_in, thinning, mu, beta, sigmaE, sigmaGG, X, epsilon, components, cva, groups, gAssign,
                sigma0=0.01, v0E=1e-4, s02E=1e-3, v0G=1e-4, s02G=1e-3):
    X = _f64(X, "F"); cva = _f64(np.atleast_2d(cva), "F"); N, M = X.shape
    beta = _f64(beta).ravel(); eps = _f64(epsilon); comp = _f64(components); sg = _f64(sigmaGG)
    gA = np.ascontiguousarray(gAssign, dtype=np.int32)
    nr = _n_rows(max_it, burn_in, thinning, False)
    rows = np.zeros((nr, 2 * M + 3 + groups + N))
    n = _ref().ref_grstart(os.fsencode(out), C.c_uint64(seed), max_it, burn_in, thinning, C.c_double(mu), _p(beta), C.c_double(sigmaE),
                           _p(sg), _p(X), C.c_long(N), C.c_long(M), _p(eps), _p(comp), C.c_double(sigma0), C.c_double(v0E),
                           C.c_double(s02E), C.c_double(v0G), C.c_double(s02G), _p(cva), cva.shape[1], groups,
                           gA.ctypes.data_as(_ip), _p(rows), C.c_long(nr))
    return rows, n


def ref_horseshoe(out, seed, max_it, burn_in, thinning, X, Y, A, v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0):
    X = _f64(X, "F"); Y = _f64(Y); N, M = X.shape
    nr = _n_rows(max_it, burn_in, thinning, False)
    rows = np.zeros((nr, 2 * M + 4 + N))
    n = _ref().ref_horseshoe(os.fsencode(out), C.c_uint64(seed), max_it, burn_in, thinning, _p(X), C.c_long(N), C.c_long(M), _p(Y),
                             C.c_double(A), C.c_double(v0E), C.c_double(s02E), C.c_double(vL), C.c_double(vT), C.c_double(c2),
                             C.c_double(vC), C.c_double(sC), _p(rows), C.c_long(nr))
    return rows, n
