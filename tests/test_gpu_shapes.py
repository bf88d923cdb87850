"""GPU parity at the worker geometries and row counts of the BASELINE shapes.

The sweep kernel is instantiated per (block, TW): TW = 1, 2, 4 packed 16-row words per lane, chosen by the rows a worker CTA
holds (<= 512 / <= 1024 / <= 2048, csrc/chain.cu::choose_geometry).  BASELINE configs[2] and configs[4] run TW = 2 (896 rows
per worker) -- so every sampler is compared with the oracle at TW = 2 and TW = 4 here, both by forcing few workers at
N = 2,000 and at the full row counts of the BASELINE shapes on a column sample the oracle finishes in seconds
(reference src/BayesRv2.cpp:186-245, src/BayesRv2Groups.cpp:232-298, src/HorseshoeR.cpp:219-240).
Bar: assignments exact, beta / residuals / variances within 1e-9 (inf-norm relative) per iteration."""
import numpy as np
import pytest

from conftest import CVA, HYP
from helpers import GroupsRow, HsRow, V2Row, assert_trace_close, rel_inf
from test_gpu_parity import _compare_groups, _compare_v2, _groups_case

pytestmark = pytest.mark.gpu
TOL = 1e-9
HS = dict(v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0)


def _compare_hs(o, rows, N, M, hyper=None):
    a, b = HsRow(rows, N, M), HsRow(o["rows"], N, M)
    for name in ("beta", "eps", "lam"):
        assert_trace_close(name, getattr(a, name), getattr(b, name), TOL)
    assert np.all(np.abs(a.tau / b.tau - 1) <= TOL) and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL) and rel_inf(a.mu, b.mu) <= TOL
    if hyper is not None:
        assert rel_inf(hyper, o["hyper"][-1]) <= TOL


# ------------------------------------------------------------------------------------------------ forced geometries, N = 2,000
# workers = 2 -> 1,024 rows per worker (TW = 2); workers = 1 -> 2,048 rows (TW = 4)
GEOM = [(2, 128, 1024), (2, 64, 1024), (1, 128, 2048), (1, 64, 2048), (1, 32, 2048)]


@pytest.mark.parametrize("workers,block,rows_pw", GEOM)
def test_v2_wide_worker_geometries(po, brr, workers, block, rows_pw):
    N, M, T = 2000, 420, 12
    d = po.synth(N, M, seed=401)
    o = po.run_v2(d["X"], d["y"], CVA, T, seed=402, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.V2, T, seed=402, Y=d["y"], cva=CVA, block=block, workers=workers, **HYP)
    geo = c.geometry()
    assert geo["workers"] == workers and geo["rows_per_worker_max"] == rows_pw and geo["block"] == block
    _compare_v2(o, c.run(T, emit_all=True), N, M, c.pi()[0])


@pytest.mark.parametrize("workers,block,rows_pw", GEOM[:4])
def test_groups_with_fixed_effects_wide_worker_geometries(po, brr, workers, block, rows_pw):
    N, M, G, F, T = 2000, 300, 3, 3, 10
    d, gA, cva, fixed = _groups_case(po, N, M, G, F, seed=410)
    o = po.run_groups(d["X"], d["y"], cva, G, gA, fixed, T, seed=411, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.GROUPS, T, seed=411, Y=d["y"], cva=cva, groups=G, gAssign=gA, fixed=fixed, block=block, workers=workers, **HYP)
    assert c.geometry()["rows_per_worker_max"] == rows_pw
    _compare_groups(o, c.run(T, emit_all=True), N, M, G, F)
    assert rel_inf(c.pi(), o["pi"][-1]) <= TOL


@pytest.mark.parametrize("workers,block,rows_pw", GEOM[:4])
def test_horseshoe_wide_worker_geometries(po, brr, workers, block, rows_pw):
    N, M, T = 2000, 260, 10
    d = po.synth(N, M, seed=420)
    A = (1 / np.sqrt(N)) * (0.1 * M) / (M - 0.1 * M)
    o = po.run_horseshoe(d["X"], d["y"], A, T, seed=421, **HS)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.HORSESHOE, T, seed=421, Y=d["y"], A=A, block=block, workers=workers, **HS)
    assert c.geometry()["rows_per_worker_max"] == rows_pw
    _compare_hs(o, c.run(T, emit_all=True), N, M, c.hyper())


# ------------------------------------------------------------------------------------------------ full row counts of the BASELINE shapes
def _synth_rows(N, M, seed, h2=0.5):
    """like pyoracle.synth, built column by column (the N x M int8 / fp64 temporaries of the vectorised form are large here)"""
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.05, 0.5, size=M)
    X = np.empty((N, M), order="F")
    for j in range(M):
        gcol = rng.binomial(2, p[j], size=N).astype(np.float64)
        X[:, j] = (gcol - gcol.mean()) / gcol.std(ddof=1)
    mc = max(1, M // 10)
    b = np.zeros(M); b[rng.choice(M, mc, replace=False)] = rng.normal(0, np.sqrt(h2 / mc), size=mc)
    y = X @ b + rng.normal(0, np.sqrt(1 - h2), size=N)
    return X, (y - y.mean()) / y.std(ddof=1)


def test_v2_config2_rows_column_sample(po, brr):
    """BASELINE configs[1] rows: N = 50,000, default geometry (98 workers x 512 rows on a B200), M = 1,024 columns"""
    N, M, T = 50000, 1024, 4
    X, y = _synth_rows(N, M, seed=430)
    o = po.run_v2(X, y, CVA, T, seed=431, **HYP)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.V2, T, seed=431, Y=y, cva=CVA, **HYP)
    _compare_v2(o, c.run(T, emit_all=True), N, M, c.pi()[0])


def test_groups_config3_rows_column_sample(po, brr):
    """BASELINE configs[2] rows: N = 100,000, 22 groups, the vignette's N x 1 zero fixed matrix, default geometry (TW = 2), M = 512"""
    N, M, G, T = 100000, 512, 22, 3
    X, y = _synth_rows(N, M, seed=440)
    gA = (np.arange(M) * G // M).astype(np.int32)
    cva = np.tile(np.array(CVA), (G, 1))
    fixed = np.zeros((N, 1))
    o = po.run_groups(X, y, cva, G, gA, fixed, T, seed=441, **HYP)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.GROUPS, T, seed=441, Y=y, cva=cva, groups=G, gAssign=gA, fixed=fixed, **HYP)
    assert c.geometry()["rows_per_worker_max"] > 512
    _compare_groups(o, c.run(T, emit_all=True), N, M, G, 1)
    assert rel_inf(c.pi(), o["pi"][-1]) <= TOL


def test_horseshoe_config5_rows_column_sample(po, brr):
    """BASELINE configs[4] rows: N = 100,000, default geometry (TW = 2), M = 512"""
    N, M, T = 100000, 512, 3
    X, y = _synth_rows(N, M, seed=450)
    A = (1 / np.sqrt(N)) * (0.1 * M) / (M - 0.1 * M)
    o = po.run_horseshoe(X, y, A, T, seed=451, **HS)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.HORSESHOE, T, seed=451, Y=y, A=A, **HS)
    assert c.geometry()["rows_per_worker_max"] > 512
    _compare_hs(o, c.run(T, emit_all=True), N, M, c.hyper())


# ------------------------------------------------------------------------------------------------ the categorical draw at its boundaries
def _first_iteration_tables(po, N, M, seed):
    K = 4
    rng = np.random.default_rng(seed)
    t = po.DrawTables(1, M, n_gam=2 + K, n_init_u=1)
    t.mark_u[:] = rng.uniform(size=(1, M)); t.mark_z[:] = rng.normal(size=(1, M)); t.mu_z[:] = rng.normal(size=1)
    t.gam[:, 0] = rng.gamma(50.0, size=1); t.gam[:, 1] = rng.gamma((1e-4 + N) / 2, size=1); t.gam[:, 2:] = rng.gamma(20.0, size=(1, K))
    t.init_u[0] = 0.41
    t.perm[0] = rng.permutation(M)
    return t


def test_categorical_draw_one_ulp_either_side_of_the_oracle_boundary(po, brr):
    """The sampler's walk decides with u * sum_l e_l <= prefix_k (and, for a marker outside the model, with one comparison of
    |num| against a bisected threshold) where the reference compares u with sum_k 1 / sum_l exp(logL_l - logL_k)
    (src/BayesRv2.cpp:216-242).  The two agree except for rounding, i.e. they may name different components only for a uniform
    within a few ulps of a boundary of the reference's cumulative probabilities.  This test finds the oracle's boundaries of one
    marker by bisection over the bit patterns of u, then replays u on the GPU
      * 1e-12 (relative) either side: the assignments MUST agree (4,500 ulps: far outside any rounding difference),
      * 1 ulp either side: the pick must be one of the two components that meet at the boundary, and a disagreement with the
        oracle is reported (DESIGN.md section 4 states the limit)."""
    N, M = 400, 48
    d = po.synth(N, M, seed=460, h2=0.6, causal_frac=0.2)
    t = _first_iteration_tables(po, N, M, seed=461)
    score = np.abs(d["X"].T @ d["y"])
    markers_done = 0
    for marker in [int(m) for m in np.argsort(-score)[:12]]:                   # markers with visible effects: their boundaries lie inside (0, 1)
        pos = int(np.nonzero(t.perm[0] == marker)[0][0])

        def oracle_pick(u):
            t.mark_u[0, pos] = u
            o = po.run_v2(d["X"], d["y"], CVA, 1, source=po.SRC_REPLAY, tables=t, **HYP)
            return int(V2Row(o["rows"], N, M).comp[0, marker]), o

        def gpu_pick(u):
            t.mark_u[0, pos] = u
            g = brr.Genotypes.from_dense(d["X"])
            c = brr.Chain(g, brr.V2, 1, Y=d["y"], cva=CVA, block=32, **HYP)
            c.set_replay(t)
            rows = c.run(1, emit_all=True)
            c.close(); g.close()
            return int(V2Row(rows, N, M).comp[0, marker]), rows

        u_keep = float(t.mark_u[0, pos])
        lo_pick, hi_pick = oracle_pick(1e-300)[0], oracle_pick(1.0 - 2.0 ** -53)[0]
        checked = 0
        for k in range(lo_pick, hi_pick):                    # boundary between "pick <= k" and "pick > k"
            lo_b, hi_b = int(np.float64(1e-300).view(np.int64)), int(np.float64(1.0 - 2.0 ** -53).view(np.int64))
            while hi_b - lo_b > 1:                            # positive doubles are ordered like their bit patterns
                mid = lo_b + (hi_b - lo_b) // 2
                if oracle_pick(np.int64(mid).view(np.float64))[0] <= k:
                    lo_b = mid
                else:
                    hi_b = mid
            b_in, b_out = float(np.int64(lo_b).view(np.float64)), float(np.int64(hi_b).view(np.float64))   # last u with pick <= k, first with pick > k
            if not (1e-6 < b_in < 1 - 1e-6):
                continue                                      # a boundary squeezed against 0 or 1 has no room for the margins
            k_in, k_out = oracle_pick(b_in)[0], oracle_pick(b_out)[0]
            assert k_in <= k < k_out
            for u in (b_in * (1 - 1e-12), b_out * (1 + 1e-12)):
                want, o = oracle_pick(u)
                got, rows = gpu_pick(u)
                assert got == want, "marker %d, u = %.17g: GPU picks %d, oracle %d" % (marker, u, got, want)
                _compare_v2(o, rows, N, M)
            for u, want in ((b_in, k_in), (b_out, k_out)):
                got, _ = gpu_pick(u)
                assert got in (k_in, k_out), (marker, u, got, k_in, k_out)
                if got != want:
                    print("boundary %d|%d of marker %d: at u = %.17g (1 ulp from the oracle's boundary) the GPU picks %d, the oracle %d"
                          % (k_in, k_out, marker, u, got, want))
            checked += 1
        t.mark_u[0, pos] = u_keep
        markers_done += checked > 0
        if markers_done >= 2:
            break
    assert markers_done >= 2, "fewer than two markers with a boundary inside (1e-6, 1 - 1e-6)"
