"""Dense fp64 covariate columns through the same sweep: the reference's X is any Eigen::MatrixXd
(src/BayesRv2.cpp:60, src/BayesRv2Groups.cpp:75); its vignette binds scale()d Gaussian "methylation" probes to the scale()d genotypes
and runs BayesRSamplerV2Groups with one group each (vignettes/BayesRR.Rmd:47-57,150-167), its Rd examples pass X = rnorm.  Genotype-like
columns are packed to 2 bits, every other column stays fp64; both take the same Gibbs blocks, the Gram tiles are then fp64.
Bar as everywhere: assignments exact, traces within 1e-9 (inf-norm relative) of the oracle, which sweeps the dense matrix itself."""
import numpy as np
import pytest

from conftest import CVA, HYP
from helpers import GroupsRow, HsRow, V2Row, assert_trace_close, rel_inf
from test_gpu_parity import _compare_groups, _compare_v2

pytestmark = pytest.mark.gpu
TOL = 1e-9
HS = dict(v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0)


def _mixed(po, N, Mg, Mc, seed, order="blocks"):
    """Mg scale()d genotype columns + Mc scale()d Gaussian columns, phenotype from both (the vignette's y2)"""
    d = po.synth(N, Mg, seed=seed, h2=0.4)
    rng = np.random.default_rng(seed + 7)
    X2 = rng.normal(size=(N, Mc))
    X2 = (X2 - X2.mean(0)) / X2.std(0, ddof=1)
    b2 = np.zeros(Mc); k = max(1, Mc // 5); b2[rng.choice(Mc, k, replace=False)] = rng.normal(0, np.sqrt(0.3 / k), size=k)
    y = d["X"] @ d["b"] + X2 @ b2 + rng.normal(0, np.sqrt(0.3), size=N)
    y = (y - y.mean()) / y.std(ddof=1)
    X = np.asfortranarray(np.hstack([d["X"], X2]))
    is_dense = np.r_[np.zeros(Mg, bool), np.ones(Mc, bool)]
    if order == "interleaved":
        perm = rng.permutation(Mg + Mc)
        X, is_dense = np.asfortranarray(X[:, perm]), is_dense[perm]
    return X, y, is_dense


def test_store_keeps_continuous_columns_dense(po, brr):
    N, Mg, Mc = 1003, 90, 40
    X, y, is_dense = _mixed(po, N, Mg, Mc, seed=501, order="interleaved")
    g = brr.Genotypes.from_dense(X)
    di = g.dense_columns()
    assert np.array_equal(di >= 0, is_dense) and np.array_equal(np.sort(di[di >= 0]), np.arange(Mc))
    st = g.stats()
    assert rel_inf(st["xsq"], (X ** 2).sum(axis=0)) < 1e-12
    assert np.all(st["a"][is_dense] == 0.0) and np.all(st["d"][is_dense] == 1.0)
    codes = g.unpack()
    assert not codes[:, is_dense].any()
    eps = np.random.default_rng(3).normal(size=N)
    assert rel_inf(g.xt_eps(eps)[0], X.T @ eps) < 1e-12
    b = np.zeros(Mg + Mc); b[[1, 17, 60, 101, 129]] = [0.5, -1.0, 0.25, 2.0, -0.75]
    assert rel_inf(g.matvec(b), X @ b) < 1e-12


@pytest.mark.parametrize("N,Mg,Mc,order", [(1500, 200, 180, "blocks"), (1100, 150, 150, "interleaved"), (800, 0, 256, "blocks")])
def test_v2_with_dense_columns_matches_oracle(po, brr, N, Mg, Mc, order):
    """mixed and all-dense X (man/BayesRSamplerV2.Rd runs the sampler on X = rnorm)"""
    T = 12
    if Mg:
        X, y, _ = _mixed(po, N, Mg, Mc, seed=510, order=order)
    else:
        rng = np.random.default_rng(511)
        X = np.asfortranarray(rng.normal(size=(N, Mc)))
        y = X[:, :8] @ rng.normal(size=8) * 0.3 + rng.normal(size=N)
        y = (y - y.mean()) / y.std(ddof=1)
    M = Mg + Mc
    o = po.run_v2(X, y, CVA, T, seed=512, **HYP)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.V2, T, seed=512, Y=y, cva=CVA, **HYP)
    assert c.geometry()["block"] == 64
    _compare_v2(o, c.run(T, emit_all=True), N, M, c.pi()[0])


def test_vignette_two_group_example(po, brr, tmp_path):
    """vignettes/BayesRR.Rmd:150-167 at a reduced size: group 0 genotypes, group 1 methylation probes, K = 5 (four variances per group),
    an N x 1 zero fixed matrix -- through the chain API and through the entry point brr_BayesRSamplerV2Groups (CSV file)"""
    N, Mg, Mc, G, T = 1200, 160, 400, 2, 14
    X, y, is_dense = _mixed(po, N, Mg, Mc, seed=520)
    M = Mg + Mc
    cva = np.array([[1e-5, 1e-4, 1e-3, 1e-2], [1e-5, 1e-4, 1e-3, 1e-2]])
    gA = np.r_[np.zeros(Mg, np.int32), np.ones(Mc, np.int32)]
    fixed = np.zeros((N, 1))
    o = po.run_groups(X, y, cva, G, gA, fixed, T, seed=521, **HYP)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.GROUPS, T, seed=521, Y=y, cva=cva, groups=G, gAssign=gA, fixed=fixed, **HYP)
    _compare_groups(o, c.run(T, emit_all=True), N, M, G, 1)
    assert rel_inf(c.pi(), o["pi"][-1]) <= TOL
    out = tmp_path / "simGroups.csv"
    brr.BayesRSamplerV2Groups(str(out), 2, 20, 10, 5, X, y, 0.01, 1e-4, 1e-3, 1e-4, 1e-3, cva, G, gA, fixed)
    o2 = po.run_groups(X, y, cva, G, gA, fixed, 20, burn_in=10, thinning=5, seed=2, emit_all=False, **HYP)
    lines = open(out).read().split("\n")
    assert lines[0] + "\n" == po.format_header(po.KIND_GROUPS, N, M, G, 1)
    rows = [np.array([float(x) for x in ln.split(", ")]) for ln in lines[1:-1]]
    assert len(rows) == o2["n_rows"] == 2 and all(np.allclose(r, o2["rows"][i], rtol=2e-5, atol=1e-12) for i, r in enumerate(rows))


def test_groups_dense_with_real_fixed_effects_and_wide_workers(po, brr):
    """dense columns at the TW = 2 worker geometry, F = 3 fixed effects (the vignette's y3)"""
    N, Mg, Mc, G, F, T = 2000, 120, 136, 2, 3, 10
    X, y, _ = _mixed(po, N, Mg, Mc, seed=530, order="interleaved")
    rng = np.random.default_rng(531)
    fixed = rng.normal(size=(N, F)); fixed = (fixed - fixed.mean(0)) / fixed.std(0, ddof=1)
    y = y + fixed @ rng.normal(scale=0.2, size=F)
    gA = rng.integers(0, G, size=Mg + Mc).astype(np.int32)
    cva = np.tile(np.array(CVA), (G, 1))
    o = po.run_groups(X, y, cva, G, gA, fixed, T, seed=532, **HYP)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.GROUPS, T, seed=532, Y=y, cva=cva, groups=G, gAssign=gA, fixed=fixed, workers=2, **HYP)
    assert c.geometry()["rows_per_worker_max"] == 1024
    _compare_groups(o, c.run(T, emit_all=True), N, Mg + Mc, G, F)


def test_horseshoe_with_dense_columns(po, brr):
    N, Mg, Mc, T = 900, 100, 92, 10
    X, y, _ = _mixed(po, N, Mg, Mc, seed=540, order="interleaved")
    M = Mg + Mc
    A = (1 / np.sqrt(N)) * (0.1 * M) / (M - 0.1 * M)
    o = po.run_horseshoe(X, y, A, T, seed=541, **HS)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.HORSESHOE, T, seed=541, Y=y, A=A, **HS)
    a, b = HsRow(c.run(T, emit_all=True), N, M), HsRow(o["rows"], N, M)
    for name in ("beta", "eps", "lam"):
        assert_trace_close(name, getattr(a, name), getattr(b, name), TOL)
    assert np.all(np.abs(a.tau / b.tau - 1) <= TOL) and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL)


def test_dense_columns_beside_packed_codes_row_sharded(po, brr):
    """two thread ranks: each packs its genotype rows, adds its rows of the continuous columns (brr_geno_set_dense_columns), takes the
    statistics of all rows; the sharded chain equals the oracle on the whole matrix and the ranks are bit-identical (fp64 Gram partials
    summed in rank order)"""
    from bayesrrcpp_b200 import sharded
    from test_sharded import pack_codes
    N, Mg, Mc, T = 1400, 130, 126, 8
    d = po.synth(N, Mg, seed=550, h2=0.4)
    rng = np.random.default_rng(551)
    X2 = rng.normal(size=(N, Mc)); X2 = (X2 - X2.mean(0)) / X2.std(0, ddof=1)
    y = d["X"] @ d["b"] + X2[:, :10] @ rng.normal(scale=0.2, size=10) + rng.normal(0, 0.6, size=N)
    y = (y - y.mean()) / y.std(ddof=1)
    M = Mg + Mc
    X = np.asfortranarray(np.hstack([d["X"], X2]))
    o = po.run_v2(X, y, CVA, T, seed=552, **HYP)
    bounds = sharded.shard_bounds(N, 2)
    codes = np.hstack([d["G"], np.zeros((N, Mc), dtype=np.int8)])

    def fn(r, comm):
        lo, hi = bounds[r]
        g = brr.Genotypes.from_packed(pack_codes(codes[lo:hi]), hi - lo)
        g.set_dense_columns(np.arange(Mg, M), X2[lo:hi]).shard_stats(comm)
        c = brr.Chain(g, brr.V2, T, seed=552, Y=y[lo:hi], cva=CVA, workers=10, comm=comm, **HYP)
        rows = c.run(T, emit_all=True)
        c.close(); g.close()
        return rows
    try:
        res = sharded.ThreadGroup(2).run(fn)
    except brr.BayesRRError as e:          # thread ranks share one device: see tests/test_sharded.py::_run_sharded
        if "watchdog" not in str(e):
            raise
        res = sharded.ThreadGroup(2).run(fn)
    assert np.array_equal(res[0], res[1]), "ranks diverged"
    _compare_v2(o, res[0], N, M)
