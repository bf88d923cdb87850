"""bayesrrcpp_b200 -- B200-native BayesR / BayesRR / Horseshoe Gibbs samplers.

Host-side mirror of the reference's R interface (R/RcppExports.R:25,49,70,74): the four entry points keep the
reference's names, argument order and meaning; numpy arrays stand in for R matrices (column-major like Eigen).
Everything computes in libbayesrr_b200.so (hand-written sm_100a kernels behind the C ABI of include/bayesrr_b200.h),
loaded with ctypes.  There is no CPU fallback: without the built library, or without a B200, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRR_LIB") or os.path.join(_HERE, "libbayesrr_b200.so")   # BRR_LIB: a variant build of the library (bayesrrcpp_b200/build.py)
_lib = None

OK, E_ITER, E_ARG, E_GENO, E_CUDA, E_IO, E_SIZE = range(7)
V2, GROUPS, GRSTART, HORSESHOE = range(4)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


class BayesRRError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bayesrr_b200 error %d: %s" % (code, msg))
        self.code = code


class _Config(C.Structure):
    _fields_ = [("kind", C.c_int), ("seed", C.c_uint64),
                ("max_iterations", C.c_int), ("burn_in", C.c_int), ("thinning", C.c_int),
                ("Y", _dp),
                ("sigma0", C.c_double), ("v0E", C.c_double), ("s02E", C.c_double), ("v0G", C.c_double), ("s02G", C.c_double),
                ("cva", _dp), ("ncva", C.c_int), ("groups", C.c_int), ("gAssign", _ip),
                ("fixed", _dp), ("F", C.c_int64), ("pi_init", _dp),
                ("mu0", C.c_double), ("beta0", _dp), ("sigmaE0", C.c_double), ("sigmaGG0", _dp),
                ("epsilon0", _dp), ("components0", _dp),
                ("A", C.c_double), ("vL", C.c_double), ("vT", C.c_double), ("c2", C.c_double), ("vC", C.c_double), ("sC", C.c_double),
                ("block", C.c_int), ("gram_impl", C.c_int), ("workers", C.c_int)]


class _Replay(C.Structure):
    _fields_ = [("n_iter", C.c_int64), ("M", C.c_int64), ("F", C.c_int64), ("n_gam", C.c_int64),
                ("n_init_u", C.c_int64), ("n_init_g", C.c_int64),
                ("mark_u", _dp), ("mark_z", _dp), ("mu_z", _dp), ("gam", _dp), ("fix_z", _dp),
                ("hs_nu", _dp), ("hs_lam", _dp), ("init_u", _dp), ("init_g", _dp),
                ("perm", _ip), ("fixperm", _ip)]


def lib():
    """Load the C-ABI library; fails loudly when it has not been built (python -m bayesrrcpp_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m bayesrrcpp_b200.build` "
                              "(there is no Python/CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.brr_last_error.restype = C.c_char_p
        L.brr_chain_row_len.restype = C.c_int64
        L.brr_chain_row_len.argtypes = [C.c_void_p]
        L.brr_geno_free.restype = None
        L.brr_geno_free.argtypes = [C.c_void_p]
        L.brr_chain_destroy.restype = None
        L.brr_chain_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != OK:
        raise BayesRRError(rc, lib().brr_last_error().decode(errors="replace"))


def _f64(a, fortran=False):
    a = np.asarray(a, dtype=np.float64)
    return np.asfortranarray(a) if fortran else np.ascontiguousarray(a)


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else _dp()


class Genotypes:
    """2-bit packed, column-major genotype store in HBM (the storage/packing layer)."""

    def __init__(self, handle):
        self._h = handle
        n, m, s = C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().brr_geno_dims(self._h, C.byref(n), C.byref(m), C.byref(s)))
        self.N, self.M, self.stride = n.value, m.value, s.value

    @classmethod
    def from_dense(cls, X, device=0):
        X = _f64(X, fortran=True)
        h = C.c_void_p()
        _check(lib().brr_geno_from_dense(_p(X), C.c_int64(X.shape[0]), C.c_int64(X.shape[1]), C.c_int(device), C.byref(h)))
        return cls(h)

    @classmethod
    def from_packed(cls, packed, N, mean=None, sd=None, device=0):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)   # (M, col_stride_bytes)
        M, cs = packed.shape
        mean = _f64(mean) if mean is not None else None
        sd = _f64(sd) if sd is not None else None
        h = C.c_void_p()
        _check(lib().brr_geno_from_packed(packed.ctypes.data_as(_bp), C.c_int64(cs), C.c_int64(N), C.c_int64(M),
                                          _p(mean), _p(sd), C.c_int(device), C.byref(h)))
        return cls(h)

    @classmethod
    def from_bed(cls, path, N=None, M=None, rows=None, impute_missing=False, device=0):
        """PLINK 1 binary genotypes.  `path`: the .bed file or the common prefix of .bed/.bim/.fam; N, M default to the line counts
        of the .fam / .bim files; rows = (row0, n_rows) reads one row shard."""
        prefix = path[:-4] if path.endswith(".bed") else path
        bed = prefix + ".bed"

        def lines(p):
            with open(p, "rb") as f:
                return sum(1 for ln in f if ln.strip())
        N = lines(prefix + ".fam") if N is None else N
        M = lines(prefix + ".bim") if M is None else M
        row0, n = rows if rows is not None else (0, 0)
        h, miss = C.c_void_p(), C.c_int64()
        _check(lib().brr_geno_from_bed(os.fsencode(bed), C.c_int64(N), C.c_int64(M), C.c_int64(row0), C.c_int64(n),
                                       C.c_int(1 if impute_missing else 0), C.c_int(device), C.byref(h), C.byref(miss)))
        g = cls(h)
        g.n_missing = miss.value
        return g

    @classmethod
    def synthetic(cls, N, M, seed, row0=0, device=0):
        h = C.c_void_p()
        _check(lib().brr_geno_synthetic(C.c_int64(N), C.c_int64(M), C.c_uint64(seed), C.c_int64(row0), C.c_int(device), C.byref(h)))
        return cls(h)

    def set_dense_columns(self, cols, values):
        """make the markers `cols` (ascending) dense fp64 columns holding `values` (N x len(cols), taken as given): continuous
        covariates beside packed genotypes"""
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        values = _f64(values, fortran=True)
        assert values.shape == (self.N, len(cols))
        _check(lib().brr_geno_set_dense_columns(self._h, cols.ctypes.data_as(_ip), C.c_int64(len(cols)), _p(values)))
        return self

    def dense_columns(self):
        """per-marker index of its dense column, or -1 (packed genotype column)"""
        n = C.c_int64()
        idx = np.zeros(self.M, dtype=np.int32)
        _check(lib().brr_geno_dense_columns(self._h, C.byref(n), idx.ctypes.data_as(_ip)))
        return idx

    def shard_stats(self, comm):
        """make the per-SNP statistics those of the whole matrix (collective over `comm`, a sharded.Comm)"""
        _check(lib().brr_geno_shard_stats(self._h, comm.byref()))
        return self

    def stats(self):
        out = {k: np.zeros(self.M) for k in ("mean", "sd", "a", "d", "xsq")}
        _check(lib().brr_geno_stats(self._h, _p(out["mean"]), _p(out["sd"]), _p(out["a"]), _p(out["d"]), _p(out["xsq"])))
        return out

    def codes(self):
        buf = np.zeros((self.M, self.stride), dtype=np.uint8)
        _check(lib().brr_geno_codes(self._h, buf.ctypes.data_as(_bp)))
        return buf

    def unpack(self):
        """N x M int8 matrix of codes (tests only)."""
        b = self.codes()
        c = np.stack([(b >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(self.M, -1)[:, :self.N]
        return np.ascontiguousarray(c.T.astype(np.int8))

    def matvec(self, b):
        b = _f64(b)
        y = np.zeros(self.N)
        _check(lib().brr_geno_matvec(self._h, _p(b), _p(y)))
        return y

    def xt_eps(self, eps):
        eps = _f64(eps)
        r = np.zeros(self.M)
        ms = C.c_double()
        _check(lib().brr_xt_eps(self._h, _p(eps), _p(r), C.byref(ms)))
        return r, ms.value

    def gram_blocks(self, order, block=128, impl=0):
        order = np.ascontiguousarray(order, dtype=np.int32)
        nb = (len(order) + block - 1) // block
        G = np.zeros((nb, block, block), dtype=np.int32)
        ms = C.c_double()
        _check(lib().brr_gram_blocks(self._h, order.ctypes.data_as(_ip), C.c_int64(len(order)), C.c_int(block), C.c_int(impl),
                                     G.ctypes.data_as(_ip), C.byref(ms)))
        return G, ms.value

    def gram_cross_blocks(self, order, block=128, impl=0):
        order = np.ascontiguousarray(order, dtype=np.int32)
        nb = (len(order) + block - 1) // block
        G = np.zeros((nb, block, block), dtype=np.int32)
        X = np.zeros((nb, lookahead(block), block), dtype=np.int32)
        ms = C.c_double()
        _check(lib().brr_gram_cross_blocks(self._h, order.ctypes.data_as(_ip), C.c_int64(len(order)), C.c_int(block), C.c_int(impl),
                                           G.ctypes.data_as(_ip), X.ctypes.data_as(_ip), C.byref(ms)))
        return G, X, ms.value

    def close(self):
        if self._h:
            lib().brr_geno_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Chain:
    """One Markov chain on one device (what the four entry points are built from)."""

    def __init__(self, geno, kind, max_iterations, burn_in=1, thinning=1, seed=1, Y=None,
                 sigma0=0.01, v0E=1e-4, s02E=1e-3, v0G=1e-4, s02G=1e-3, cva=None, groups=1, gAssign=None, fixed=None,
                 pi_init=None, mu=0.0, beta=None, sigmaE=0.0, sigmaGG=None, epsilon=None, components=None,
                 A=0.0, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0, block=0, gram_impl=0, workers=0, comm=None):
        """comm: a sharded.Comm -> row-sharded chain (Y / fixed / epsilon hold this rank's rows; collective call)"""
        self.geno, self.kind, self.comm = geno, kind, comm
        keep = []

        def arr(a, fortran=False):
            if a is None:
                return None
            a = _f64(a, fortran)
            keep.append(a)
            return a
        cfg = _Config()
        cfg.kind, cfg.seed = kind, seed
        cfg.max_iterations, cfg.burn_in, cfg.thinning = max_iterations, burn_in, thinning
        cfg.Y = _p(arr(Y))
        cfg.sigma0, cfg.v0E, cfg.s02E, cfg.v0G, cfg.s02G = sigma0, v0E, s02E, v0G, s02G
        self.K = 0
        if kind != HORSESHOE:
            cv = arr(np.atleast_2d(cva) if kind != V2 else np.asarray(cva).ravel(), fortran=kind != V2)
            cfg.cva = _p(cv)
            cfg.ncva = cv.shape[1] if kind != V2 else cv.shape[0]
            self.K = cfg.ncva + 1
        cfg.groups = groups
        if gAssign is not None:
            ga = np.ascontiguousarray(gAssign, dtype=np.int32)
            keep.append(ga)
            cfg.gAssign = ga.ctypes.data_as(_ip)
        fx = arr(fixed, fortran=True)
        cfg.fixed = _p(fx)
        cfg.F = 0 if fx is None else fx.shape[1]
        cfg.pi_init = _p(arr(pi_init))
        cfg.mu0, cfg.sigmaE0 = mu, sigmaE
        cfg.beta0, cfg.sigmaGG0 = _p(arr(None if beta is None else np.ravel(beta))), _p(arr(sigmaGG))
        cfg.epsilon0, cfg.components0 = _p(arr(epsilon)), _p(arr(components))
        cfg.A, cfg.vL, cfg.vT, cfg.c2, cfg.vC, cfg.sC = A, vL, vT, c2, vC, sC
        cfg.block, cfg.gram_impl, cfg.workers = block, gram_impl, workers
        self.G = groups if kind in (GROUPS, GRSTART) else 1
        self.F = cfg.F
        h = C.c_void_p()
        if comm is None:
            _check(lib().brr_chain_create(C.byref(cfg), geno._h, C.byref(h)))
        else:
            _check(lib().brr_chain_create_sharded(C.byref(cfg), geno._h, comm.byref(), C.byref(h)))
        self._h = h
        self._keep = keep
        self.row_len = lib().brr_chain_row_len(self._h)

    def set_replay(self, t):
        """t: object carrying the brr_replay tables as numpy arrays (attributes n_iter, M, F, n_gam, mark_u, ...)."""
        keep = []

        def p(a, ptr=_dp, dt=np.float64):
            if a is None:
                return ptr()
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(ptr)
        r = _Replay(t.n_iter, t.M, t.F, t.n_gam, t.n_init_u, t.n_init_g,
                    p(t.mark_u), p(t.mark_z), p(t.mu_z), p(t.gam), p(t.fix_z), p(t.hs_nu), p(t.hs_lam),
                    p(t.init_u), p(t.init_g), p(t.perm, _ip, np.int32), p(t.fixperm, _ip, np.int32))
        _check(lib().brr_chain_set_replay(self._h, C.byref(r)))

    def open_output(self, path):
        _check(lib().brr_chain_open_output(self._h, os.fsencode(path)))

    def open_binary_output(self, path):
        _check(lib().brr_chain_open_binary_output(self._h, os.fsencode(path)))

    def close_output(self):
        _check(lib().brr_chain_close_output(self._h))

    def save(self, path):
        """lossless checkpoint of the chain after its last completed iteration"""
        _check(lib().brr_chain_save(self._h, os.fsencode(path)))

    def load(self, path):
        """continue the chain a checkpoint was taken from (call on a freshly created chain of the same configuration)"""
        _check(lib().brr_chain_load(self._h, os.fsencode(path)))

    def run(self, n_iter, emit_all=False, max_rows=None):
        max_rows = n_iter if max_rows is None else max_rows
        rows = np.zeros((max_rows, self.row_len))
        got = C.c_int64()
        _check(lib().brr_chain_run(self._h, C.c_int(n_iter), C.c_int(1 if emit_all else 0), _p(rows), C.c_int64(max_rows), C.byref(got)))
        return rows[:min(got.value, max_rows)]

    def run_discard(self, n_iter):
        got = C.c_int64()
        _check(lib().brr_chain_run(self._h, C.c_int(n_iter), C.c_int(0), _dp(), C.c_int64(0), C.byref(got)))
        return got.value

    def pi(self):
        out = np.zeros((self.G, self.K))
        _check(lib().brr_chain_get_pi(self._h, _p(out)))
        return out

    def hyper(self):
        out = np.zeros(3)
        _check(lib().brr_chain_get_hyper(self._h, _p(out)))
        return out

    def last_timing(self):
        ms, n = C.c_double(), C.c_int64()
        _check(lib().brr_chain_last_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def kernel_ms(self):
        out = np.zeros(3)
        _check(lib().brr_chain_kernel_ms(self._h, _p(out)))
        return dict(gram=out[0], sweep=out[1], hyper=out[2])

    def sweep_profile(self):
        o = np.zeros(16)
        _check(lib().brr_chain_sweep_profile(self._h, _p(o)))
        return dict(gather=o[0], gather_first_chunk=o[1], serial_pass=o[2], publish=o[3], windows=o[4], full_steps=o[5], blocks=o[6],
                    gather_last_chunk=o[7], worker_wait=o[8], eval_cycles=o[9], fp64_draws=o[9], resolve_cycles=o[14], prologue_cycles=o[15], worker_dots=o[10], bookkeeping=o[11], chunks_received=o[12], worker_reduce=o[13])

    def geometry(self):
        b, w, r, s = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().brr_chain_geometry(self._h, C.byref(b), C.byref(w), C.byref(r), C.byref(s)))
        return dict(block=b.value, workers=w.value, rows_per_worker_max=r.value, smem_bytes=s.value)

    def close(self):
        if self._h:
            lib().brr_chain_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------
# The four entry points, named and ordered as in the reference (R/RcppExports.R:25,49,70,74).
def BayesRSamplerV2(outputFile, seed, max_iterations, burn_in, thinning, X, Y, sigma0, v0E, s02E, v0G, s02G, cva):
    X = _f64(X, fortran=True); Y = _f64(Y); cva = _f64(cva).ravel()
    _check(lib().brr_BayesRSamplerV2(os.fsencode(outputFile), C.c_int(seed), C.c_int(max_iterations), C.c_int(burn_in),
                                     C.c_int(thinning), _p(X), C.c_int64(X.shape[0]), C.c_int64(X.shape[1]), _p(Y),
                                     C.c_double(sigma0), C.c_double(v0E), C.c_double(s02E), C.c_double(v0G), C.c_double(s02G),
                                     _p(cva), C.c_int(len(cva))))


def BayesRSamplerV2Groups(outputFile, seed, max_iterations, burn_in, thinning, X, Y, sigma0, v0E, s02E, v0G, s02G, cva,
                          groups, gAssign, fixed):
    X = _f64(X, fortran=True); Y = _f64(Y); cva = _f64(np.atleast_2d(cva), fortran=True)
    gA = np.ascontiguousarray(gAssign, dtype=np.int32)
    fixed = _f64(fixed, fortran=True) if fixed is not None else np.zeros((X.shape[0], 0), order="F")
    _check(lib().brr_BayesRSamplerV2Groups(os.fsencode(outputFile), C.c_int(seed), C.c_int(max_iterations), C.c_int(burn_in),
                                           C.c_int(thinning), _p(X), C.c_int64(X.shape[0]), C.c_int64(X.shape[1]), _p(Y),
                                           C.c_double(sigma0), C.c_double(v0E), C.c_double(s02E), C.c_double(v0G), C.c_double(s02G),
                                           _p(cva), C.c_int(cva.shape[1]), C.c_int(groups), gA.ctypes.data_as(_ip),
                                           _p(fixed) if fixed.shape[1] else _dp(), C.c_int64(fixed.shape[1])))


def BRV2Grstart(outputFile, seed, max_iterations, burn_in, thinning, mu, beta, sigmaE, sigmaGG, X, epsilon, components,
                sigma0, v0E, s02E, v0G, s02G, cva, groups, gAssign):
    X = _f64(X, fortran=True); cva = _f64(np.atleast_2d(cva), fortran=True)
    beta = _f64(beta).ravel(); sg = _f64(sigmaGG); eps = _f64(epsilon); comp = _f64(components)
    gA = np.ascontiguousarray(gAssign, dtype=np.int32)
    _check(lib().brr_BRV2Grstart(os.fsencode(outputFile), C.c_int(seed), C.c_int(max_iterations), C.c_int(burn_in),
                                 C.c_int(thinning), C.c_double(mu), _p(beta), C.c_double(sigmaE), _p(sg),
                                 _p(X), C.c_int64(X.shape[0]), C.c_int64(X.shape[1]), _p(eps), _p(comp),
                                 C.c_double(sigma0), C.c_double(v0E), C.c_double(s02E), C.c_double(v0G), C.c_double(s02G),
                                 _p(cva), C.c_int(cva.shape[1]), C.c_int(groups), gA.ctypes.data_as(_ip)))


def HorseshoeR(outputFile, seed, max_iterations, burn_in, thinning, X, Y, A, v0E, s02E, vL, vT, c2, vC, sC):
    X = _f64(X, fortran=True); Y = _f64(Y)
    _check(lib().brr_HorseshoeR(os.fsencode(outputFile), C.c_int(seed), C.c_int(max_iterations), C.c_int(burn_in),
                                C.c_int(thinning), _p(X), C.c_int64(X.shape[0]), C.c_int64(X.shape[1]), _p(Y),
                                C.c_double(A), C.c_double(v0E), C.c_double(s02E), C.c_double(vL), C.c_double(vT),
                                C.c_double(c2), C.c_double(vC), C.c_double(sC)))


_MSG_T = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)
_msg_keep = None


def set_message_handler(fn):
    """fn(text) receives the reference's console messages ("iteration: <n>", "duration: <s>s"; src/BayesRv2.cpp:173-175,276-278)
    from the four entry points; None silences them (the default)."""
    global _msg_keep
    L = lib()
    L.brr_set_message_handler.restype = None
    L.brr_set_message_handler.argtypes = [_MSG_T, C.c_void_p]
    cb = _MSG_T(lambda ctx, text: fn(text.decode())) if fn is not None else C.cast(None, _MSG_T)
    L.brr_set_message_handler(cb, None)
    _msg_keep = cb


def read_binary_samples(path):
    """(meta, rows) of a file written through Chain.open_binary_output"""
    with open(path, "rb") as f:
        h = f.read(64)
        assert h[:7] == b"BRRSMP1", "not a binary sample file"
        kind, G = np.frombuffer(h, dtype=np.int32, count=2, offset=8)
        N, M, F, L = np.frombuffer(h, dtype=np.int64, count=4, offset=16)
        rows = np.fromfile(f, dtype=np.float64)
    return dict(kind=int(kind), groups=int(G), N=int(N), M=int(M), F=int(F), row_len=int(L)), rows.reshape(-1, int(L))


def lookahead(block):
    """markers of a Gibbs block whose deltas reach the next block through the cross-Gram correction (csrc/common.cuh; a
    build-time constant of the library: the whole block by default)"""
    la = int(lib().brr_lookahead(C.c_int(block)))
    if la <= 0:
        raise ValueError("block must be 32, 64 or 128")
    return la


def peak_fp64(device=0):
    """measured fp64 FMA throughput (thread-level DFMAs per second)"""
    v = C.c_double()
    _check(lib().brr_peak_fp64(C.c_int(device), C.byref(v)))
    return v.value


def draws_sample(seed, stream, it, idx0, n, kind, shape=1.0):
    out = np.zeros(n)
    _check(lib().brr_draws_sample(C.c_uint64(seed), C.c_int(stream), C.c_int64(it), C.c_int64(idx0), C.c_int64(n),
                                  C.c_int(kind), C.c_double(shape), _p(out)))
    return out


def shuffle_host(seed, stream, it, order):
    order = np.ascontiguousarray(order, dtype=np.int32).copy()
    _check(lib().brr_shuffle_host(C.c_uint64(seed), C.c_int(stream), C.c_int64(it), order.ctypes.data_as(_ip), C.c_int64(len(order))))
    return order
