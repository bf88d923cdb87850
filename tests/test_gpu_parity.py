"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs.  Bar (BASELINE.json north_star): mixture-component assignments EXACT; beta, residuals and the variance /
intercept / pi traces within 1e-9 per iteration, relative to each vector's infinity norm; integer work bit-exact."""
import numpy as np
import pytest

from conftest import CVA, HYP
from helpers import GroupsRow, HsRow, V2Row, assert_trace_close, rel_inf

pytestmark = pytest.mark.gpu
TOL = 1e-9


# ------------------------------------------------------------------------------------------------ storage layer
def test_pack_dense_roundtrip_and_stats(po, brr):
    d = po.synth(1003, 301, seed=11)
    g = brr.Genotypes.from_dense(d["X"])
    assert (g.N, g.M) == (1003, 301) and g.stride % 128 == 0
    codes = g.unpack()
    st = g.stats()
    # codes are an affine relabelling of the genotypes: x = a + d * code reproduces X
    Xr = st["a"][None, :] + st["d"][None, :] * codes
    assert rel_inf(Xr, d["X"]) < 1e-13
    assert rel_inf(st["xsq"], (d["X"] ** 2).sum(axis=0)) < 1e-12
    # three-valued columns keep the genotype itself
    three = np.array([len(np.unique(d["G"][:, j])) == 3 for j in range(g.M)])
    assert np.array_equal(codes[:, three], d["G"][:, three])


def test_non_genotype_columns_stay_dense(po, brr):
    """a column that is not a + d * code is kept as a dense fp64 column (tests/test_gpu_dense.py), not rejected"""
    X = np.random.default_rng(0).normal(size=(64, 4))
    X[:, 2] = np.random.default_rng(1).integers(0, 3, size=64)          # one genotype column among them
    g = brr.Genotypes.from_dense(X)
    assert np.array_equal(g.dense_columns() >= 0, [True, True, False, True])
    assert rel_inf(g.stats()["xsq"], (np.asfortranarray(X) ** 2).sum(axis=0)) < 1e-13


def test_from_packed_matches_dense_and_rejects_missing(po, brr):
    d = po.synth(777, 130, seed=12)
    N, M = d["G"].shape
    cs = (N + 3) // 4 + 5                                    # odd stride on purpose
    packed = np.zeros((M, cs), dtype=np.uint8)
    for q in range(4):
        rows = np.arange(q, N, 4)
        packed[:, :len(rows)] |= (d["G"][rows, :].T.astype(np.uint8) << (2 * q))
    g = brr.Genotypes.from_packed(packed, N)
    assert np.array_equal(g.unpack(), d["G"])
    st = g.stats()
    assert rel_inf(st["mean"], d["mean"]) < 1e-14 and rel_inf(st["sd"], d["sd"]) < 1e-13
    assert rel_inf(st["xsq"], np.full(M, N - 1.0)) < 1e-12   # scale()d columns
    packed[3, 0] |= 3
    with pytest.raises(brr.BayesRRError) as e:
        brr.Genotypes.from_packed(packed, N)
    assert e.value.code == brr.E_GENO


def test_from_packed_page_locked_source(po, brr):
    """A page-locked host matrix is read in place by the copy engine (no staging): pitched and store-pitch sources."""
    import torch
    d = po.synth(1030, 70, seed=13)
    N, M = d["G"].shape
    ref = brr.Genotypes.from_dense(d["X"])
    for cs in ((N + 3) // 4 + 3, ref.stride):
        pinned = torch.zeros((M, cs), dtype=torch.uint8).pin_memory()
        packed = pinned.numpy()
        for q in range(4):
            rows = np.arange(q, N, 4)
            packed[:, :len(rows)] |= (d["G"][rows, :].T.astype(np.uint8) << (2 * q))
        g = brr.Genotypes.from_packed(packed, N)
        assert np.array_equal(g.codes(), ref.codes())
        assert np.array_equal(g.unpack(), d["G"])
        g.close()
    ref.close()


def _write_plink(prefix, codes, missing=None):
    """codes: N x M counts of the A1 allele (0/1/2); missing: boolean N x M.  Writes prefix.bed/.bim/.fam (SNP-major)."""
    N, M = codes.shape
    bed = np.full((N, M), 0, dtype=np.uint8)
    bed[codes == 2] = 0b00; bed[codes == 1] = 0b10; bed[codes == 0] = 0b11
    if missing is not None:
        bed[missing] = 0b01
    width = (N + 3) // 4
    out = np.zeros((M, width), dtype=np.uint8)
    for q in range(4):
        rows = np.arange(q, N, 4)
        out[:, :len(rows)] |= (bed[rows, :].T << (2 * q)).astype(np.uint8)
    with open(prefix + ".bed", "wb") as f:
        f.write(bytes([0x6c, 0x1b, 0x01])); f.write(out.tobytes())
    with open(prefix + ".bim", "w") as f:
        for j in range(M):
            f.write("1\trs%d\t0\t%d\tA\tG\n" % (j, j + 1))
    with open(prefix + ".fam", "w") as f:
        for i in range(N):
            f.write("F%d I%d 0 0 0 -9\n" % (i, i))


def test_plink_bed_ingest(po, brr, tmp_path):
    """.bed -> packed store without a dense detour; same store as from_dense on scale()d columns; row shards;
    missing genotypes rejected or imputed to the rounded column mean"""
    d = po.synth(1003, 140, seed=16)
    prefix = str(tmp_path / "toy")
    _write_plink(prefix, d["G"])
    g = brr.Genotypes.from_bed(prefix)
    assert (g.N, g.M, g.n_missing) == (1003, 140, 0)
    assert np.array_equal(g.unpack(), d["G"])
    st, ref = g.stats(), brr.Genotypes.from_dense(d["X"]).stats()
    assert rel_inf(st["mean"], d["mean"]) < 1e-14 and rel_inf(st["sd"], d["sd"]) < 1e-13 and rel_inf(st["xsq"], ref["xsq"]) < 1e-12
    eps = np.random.default_rng(2).normal(size=g.N)
    assert rel_inf(g.xt_eps(eps)[0], d["X"].T @ eps) < 1e-12
    # a row shard reads only its rows
    sh = brr.Genotypes.from_bed(prefix + ".bed", rows=(512, 300))
    assert np.array_equal(sh.unpack(), d["G"][512:812])
    # missing genotypes
    miss = np.zeros(d["G"].shape, dtype=bool)
    miss[[3, 77, 500, 1002], [0, 0, 5, 139]] = True
    _write_plink(prefix + "_m", d["G"], miss)
    with pytest.raises(brr.BayesRRError) as e:
        brr.Genotypes.from_bed(prefix + "_m")
    assert e.value.code == brr.E_GENO and "4 missing" in str(e.value)
    gi = brr.Genotypes.from_bed(prefix + "_m", impute_missing=True)
    assert gi.n_missing == 4
    got = gi.unpack()
    want = d["G"].copy()
    for i, j in zip(*np.nonzero(miss)):
        obs = d["G"][~miss[:, j], j]
        want[i, j] = int(obs.mean() + 0.5)
    assert np.array_equal(got, want)
    # not a .bed file / wrong dimensions
    bad = tmp_path / "bad.bed"
    bad.write_bytes(b"\x00\x01\x01" + bytes(5000))
    with pytest.raises(brr.BayesRRError) as e:
        brr.Genotypes.from_bed(str(bad), N=100, M=100)
    assert e.value.code == brr.E_IO
    with pytest.raises(brr.BayesRRError) as e:
        brr.Genotypes.from_bed(prefix + ".bed", N=5000, M=140)
    assert e.value.code == brr.E_IO


def test_synthetic_store_is_row_shard_consistent(brr):
    whole = brr.Genotypes.synthetic(1024, 64, seed=5).unpack()
    lo = brr.Genotypes.synthetic(512, 64, seed=5, row0=0).unpack()
    hi = brr.Genotypes.synthetic(512, 64, seed=5, row0=512).unpack()
    assert np.array_equal(np.vstack([lo, hi]), whole)
    assert set(np.unique(whole)) <= {0, 1, 2}
    maf = whole.mean(axis=0) / 2
    assert 0.03 < maf.min() and maf.max() < 0.55


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("block", [128, 64, 32])
def test_gram_tensor_core_is_bit_exact(po, brr, block):
    d = po.synth(1500, 333, seed=13)
    g = brr.Genotypes.from_dense(d["X"])
    codes = g.unpack().astype(np.int64)
    order = np.random.default_rng(3).permutation(g.M).astype(np.int32)
    G_tc, _ = g.gram_blocks(order, block=block, impl=0)
    G_cc, _ = g.gram_blocks(order, block=block, impl=1)
    nb = (g.M + block - 1) // block
    for b in range(nb):
        idx = order[b * block:(b + 1) * block]
        want = np.zeros((block, block), dtype=np.int64)
        want[:len(idx), :len(idx)] = codes[:, idx].T @ codes[:, idx]
        assert np.array_equal(G_cc[b], want), "dp4a gram, block %d" % b
        assert np.array_equal(G_tc[b], want), "tcgen05 gram, block %d" % b


@pytest.mark.parametrize("block", [128, 64, 32])
def test_gram_cross_products_with_previous_block_tail(po, brr, block):
    """look-ahead tiles: X[b][jl][k] = codes[:, order[b*B-LA+jl]] . codes[:, order[b*B+k]], tensor cores == bit-sliced CUDA cores == numpy"""
    LA = brr.lookahead(block)
    d = po.synth(1100, 300, seed=15)
    g = brr.Genotypes.from_dense(d["X"])
    codes = g.unpack().astype(np.int64)
    order = np.random.default_rng(4).permutation(g.M).astype(np.int32)
    G_tc, X_tc, _ = g.gram_cross_blocks(order, block=block, impl=0)
    G_cc, X_cc, _ = g.gram_cross_blocks(order, block=block, impl=1)
    assert np.array_equal(G_tc, G_cc) and np.array_equal(X_tc, X_cc)
    nb = (g.M + block - 1) // block
    assert not X_tc[0].any()
    for b in range(1, nb):
        prev = order[b * block - LA:b * block]
        cur = order[b * block:(b + 1) * block]
        want = np.zeros((LA, block), dtype=np.int64)
        want[:, :len(cur)] = codes[:, prev].T @ codes[:, cur]
        assert np.array_equal(X_tc[b], want), "block %d" % b


def test_gram_large_rows_property(brr):
    """full-size property: diagonal of the exact Gram == sum of squared codes, symmetry, at N = 50,000"""
    g = brr.Genotypes.synthetic(50000, 256, seed=9)
    order = np.arange(256, dtype=np.int32)
    G, _ = g.gram_blocks(order, block=128, impl=0)
    st = g.stats()
    q = st["mean"] * 0  # placeholder to keep names short
    codes = g.unpack().astype(np.int64)
    for b in range(2):
        assert np.array_equal(G[b], G[b].T)
        assert np.array_equal(np.diag(G[b]), (codes[:, b * 128:(b + 1) * 128] ** 2).sum(axis=0))
    del q


def test_xt_eps_kernel(po, brr):
    d = po.synth(2050, 200, seed=14)
    g = brr.Genotypes.from_dense(d["X"])
    eps = np.random.default_rng(1).normal(size=g.N)
    r, _ = g.xt_eps(eps)
    assert rel_inf(r, d["X"].T @ eps) < 1e-12
    b = np.zeros(g.M); b[[3, 50, 199]] = [0.5, -1.25, 2.0]
    assert rel_inf(g.matvec(b), d["X"] @ b) < 1e-13


def test_device_draws_match_oracle_generator(po, brr):
    L = po.lib()
    seed = 0x1234567890ABCDEF
    u = brr.draws_sample(seed, po.S_MARK_U, 7, 100, 257, 0)
    z = brr.draws_sample(seed, po.S_MARK_Z, 7, 100, 257, 1)
    ou = np.array([L.orc_api_px_uniform(seed, po.S_MARK_U, 7, 100 + i) for i in range(257)])
    oz = np.array([L.orc_api_px_normal(seed, po.S_MARK_Z, 7, 100 + i) for i in range(257)])
    assert np.array_equal(u, ou)                       # integer -> double map: bit-exact
    assert np.max(np.abs(z - oz)) < 1e-13              # log / cos / sqrt differ by ulps between libm and the device
    for shape in (0.3, 1.0, 2.5, 500.25, 25000.5):
        gdev = brr.draws_sample(seed, po.S_GAMMA, 3, 0, 64, 2, shape)
        gor = np.array([L.orc_api_px_gamma(seed, po.S_GAMMA, 3, i, shape) for i in range(64)])
        assert rel_inf(gdev, gor) < 1e-12, shape
    order = np.arange(1000, dtype=np.int32)
    mine = brr.shuffle_host(seed, po.S_PERM, 2, order)
    ref = order.copy()
    L.orc_api_px_shuffle(seed, po.S_PERM, 2, ref.ctypes.data_as(po._ip), len(ref))
    assert np.array_equal(mine, ref) and sorted(mine) == list(range(1000))


# ------------------------------------------------------------------------------------------------ BayesRSamplerV2
def _compare_v2(o, rows, N, M, pi_gpu=None):
    a, b = V2Row(rows, N, M), V2Row(o["rows"], N, M)
    assert np.array_equal(a.it, b.it)
    assert np.array_equal(a.comp, b.comp), "mixture-component assignments differ in %d places" % int((a.comp != b.comp).sum())
    assert_trace_close("beta", a.beta, b.beta, TOL)
    assert_trace_close("epsilon", a.eps, b.eps, TOL)
    assert rel_inf(a.mu, b.mu) <= TOL and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL) and np.all(np.abs(a.sigmaG / b.sigmaG - 1) <= TOL)
    if pi_gpu is not None:
        assert rel_inf(pi_gpu, o["pi"][-1]) <= TOL


@pytest.mark.parametrize("N,M,block,iters", [(2000, 1000, 128, 40), (1003, 517, 64, 25), (640, 150, 32, 25)])
def test_v2_philox_chain_matches_oracle(po, brr, N, M, block, iters):
    d = po.synth(N, M, seed=1001)
    o = po.run_v2(d["X"], d["y"], CVA, iters, seed=2001, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.V2, iters, seed=2001, Y=d["y"], cva=CVA, block=block, **HYP)
    rows = c.run(iters, emit_all=True)
    _compare_v2(o, rows, N, M, c.pi()[0])


def test_v2_config1_full_chain(po, brr):
    """BASELINE config 1: N=2,000 x M=1,000, K=4, 500 iterations, burn-in 250, thinning 5 (kept rows only)."""
    d = po.synth(2000, 1000, seed=1001)
    o = po.run_v2(d["X"], d["y"], CVA, 500, burn_in=250, thinning=5, seed=2001, emit_all=False, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.V2, 500, burn_in=250, thinning=5, seed=2001, Y=d["y"], cva=CVA, **HYP)
    rows = c.run(500, emit_all=False)
    assert rows.shape == o["rows"].shape == (50, 2 * 1000 + 4 + 2000)
    _compare_v2(o, rows, 2000, 1000, c.pi()[0])
    r = V2Row(rows, 2000, 1000)
    h2 = np.mean(r.sigmaG / (r.sigmaG + r.sigmaE))
    assert 0.3 < h2 < 0.7                                # vignette acceptance: PVE near the simulated h2 = 0.5


def test_v2_replay_of_foreign_draws(po, brr):
    """draw-replay mode with tables that do NOT come from the Philox scheme (numpy generator)."""
    N, M, T, K = 900, 260, 12, 4
    d = po.synth(N, M, seed=21)
    rng = np.random.default_rng(99)
    t = po.DrawTables(T, M, n_gam=2 + K, n_init_u=1)
    t.mark_u[:] = rng.uniform(size=(T, M)); t.mark_z[:] = rng.normal(size=(T, M)); t.mu_z[:] = rng.normal(size=T)
    t.gam[:, 0] = rng.gamma(50.0, size=T); t.gam[:, 1] = rng.gamma((1e-4 + N) / 2, size=T); t.gam[:, 2:] = rng.gamma(20.0, size=(T, K))
    t.init_u[0] = 0.37
    for i in range(T):
        t.perm[i] = rng.permutation(M)
    o = po.run_v2(d["X"], d["y"], CVA, T, source=po.SRC_REPLAY, tables=t, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.V2, T, Y=d["y"], cva=CVA, **HYP)
    c.set_replay(t)
    rows = c.run(T, emit_all=True)
    _compare_v2(o, rows, N, M, c.pi()[0])


def test_v2_uninitialised_pi_quirk_and_fall_through(po, brr):
    """with NaN mixture proportions (zeroed heap in the reference) no marker is assigned in
    iteration 0 -- beta, components stay as they were -- and the chain recovers from iteration 1 on."""
    N, M, T = 512, 140, 6
    d = po.synth(N, M, seed=22)
    pi0 = [0.5, np.nan, np.nan, np.nan]
    o = po.run_v2(d["X"], d["y"], CVA, T, seed=5, pi_init=pi0, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.V2, T, seed=5, Y=d["y"], cva=CVA, pi_init=pi0, **HYP)
    rows = c.run(T, emit_all=True)
    r = V2Row(rows, N, M)
    assert np.all(r.beta[0] == 0) and np.all(r.comp[0] == 0)
    assert np.any(r.comp[1:] != 0) or np.any(r.beta[1:] != 0)
    _compare_v2(o, rows, N, M)


@pytest.mark.parametrize("cva", [[1e-3], [1e-3, 1e-2], [1e-5, 1e-4, 1e-3, 1e-2, 1e-1], [1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 0.5]])
def test_v2_other_component_counts(po, brr, cva):
    N, M, T = 700, 200, 15
    d = po.synth(N, M, seed=23)
    o = po.run_v2(d["X"], d["y"], cva, T, seed=31, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.V2, T, seed=31, Y=d["y"], cva=cva, **HYP)
    _compare_v2(o, c.run(T, emit_all=True), N, M, c.pi()[0])


def test_v2_overflow_guard(po, brr):
    """huge effects drive |logL_k - logL_0| past 700: the guard of src/BayesRv2.cpp:216,235 zeroes probability mass"""
    N, M, T = 600, 64, 8
    d = po.synth(N, M, seed=24, h2=0.99, causal_frac=0.05)
    y = d["y"] * 60.0
    o = po.run_v2(d["X"], y, CVA, T, seed=41, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.V2, T, seed=41, Y=y, cva=CVA, **HYP)
    _compare_v2(o, c.run(T, emit_all=True), N, M)


def test_v2_chain_is_run_to_run_deterministic_and_resumable(po, brr):
    N, M = 800, 300
    d = po.synth(N, M, seed=25)
    g = brr.Genotypes.from_dense(d["X"])
    a = brr.Chain(g, brr.V2, 20, seed=7, Y=d["y"], cva=CVA, **HYP).run(20, emit_all=True)
    c = brr.Chain(g, brr.V2, 20, seed=7, Y=d["y"], cva=CVA, **HYP)
    b = np.vstack([c.run(7, emit_all=True), c.run(13, emit_all=True)])
    assert np.array_equal(a, b)
    w = brr.Chain(g, brr.V2, 20, seed=7, Y=d["y"], cva=CVA, workers=5, **HYP).run(20, emit_all=True)
    assert np.array_equal(V2Row(a, N, M).comp, V2Row(w, N, M).comp)
    assert_trace_close("beta across geometries", V2Row(w, N, M).beta, V2Row(a, N, M).beta, TOL)


def test_v2_full_size_residual_identity(brr):
    """BASELINE config 2 size (N = M = 50,000; too large for the CPU oracle): the defining identity of the sampler's state,
    eps = y - mu - X beta (reference src/BayesRv2.cpp:168,191,243), must hold after every iteration although eps is only ever
    updated incrementally (block dots, look-ahead corrections, streamed deltas); components and beta must be consistent."""
    N = M = 50000
    g = brr.Genotypes.synthetic(N, M, seed=77)
    rng = np.random.default_rng(3)
    b = np.zeros(M); idx = rng.choice(M, 5000, replace=False); b[idx] = rng.normal(0, np.sqrt(0.5 / 5000), size=5000)
    y = g.matvec(b) + rng.normal(0, np.sqrt(0.5), size=N)
    y = (y - y.mean()) / y.std(ddof=1)
    T = 4
    c = brr.Chain(g, brr.V2, T, seed=5, Y=y, cva=CVA, **HYP)
    rows = c.run(T, emit_all=True)
    r = V2Row(rows, N, M)
    for t in range(T):
        resid = y - r.mu[t] - g.matvec(r.beta[t])
        assert rel_inf(r.eps[t], resid) < 1e-9, t
        assert np.array_equal(r.comp[t] != 0, r.beta[t] != 0)          # component 0 <=> beta == 0 (:226-231)
        assert set(np.unique(r.comp[t])) <= {0.0, 1.0, 2.0, 3.0}
    assert (r.beta[-1] != 0).sum() > 100 and np.all(r.sigmaE > 0) and np.all(r.sigmaG > 0)
    # the same chain with other geometry (64-marker blocks, fewer workers): assignments identical, traces within tolerance
    c2 = brr.Chain(g, brr.V2, T, seed=5, Y=y, cva=CVA, block=64, workers=100, **HYP)
    r2 = V2Row(c2.run(T, emit_all=True), N, M)
    assert np.array_equal(r.comp, r2.comp)
    assert_trace_close("beta across geometries", r2.beta, r.beta, TOL)
    assert_trace_close("epsilon across geometries", r2.eps, r.eps, TOL)


def test_groups_full_size_residual_identity(brr):
    """BASELINE config 3 size: BayesRSamplerV2Groups, N = 100,000 x M = 200,000, 22 groups, the vignette's N x 1 zero fixed matrix
    (two 64-row words per lane in the workers): eps = y - mu - X beta - fixed alpha after every iteration"""
    N, M, G, T = 100000, 200000, 22, 2
    g = brr.Genotypes.synthetic(N, M, seed=78)
    rng = np.random.default_rng(4)
    b = np.zeros(M); idx = rng.choice(M, 2000, replace=False); b[idx] = rng.normal(0, np.sqrt(0.5 / 2000), size=2000)
    y = g.matvec(b) + rng.normal(0, np.sqrt(0.5), size=N)
    y = (y - y.mean()) / y.std(ddof=1)
    gA = (np.arange(M) * G // M).astype(np.int32)
    cva = np.tile(np.array(CVA), (G, 1))
    c = brr.Chain(g, brr.GROUPS, T, seed=6, Y=y, cva=cva, groups=G, gAssign=gA, fixed=np.zeros((N, 1)), **HYP)
    assert c.geometry()["rows_per_worker_max"] > 512
    r = GroupsRow(c.run(T, emit_all=True), N, M, G, 1)
    for t in range(T):
        keep = np.nonzero(r.beta[t])[0]
        bt = np.zeros(M); bt[keep] = r.beta[t][keep]
        assert rel_inf(r.eps[t], y - r.mu[t] - g.matvec(bt)) < 1e-9, t
        assert np.array_equal(r.comp[t] != 0, r.beta[t] != 0)
    assert np.all(r.sigmaG > 0) and np.all(r.sigmaE > 0) and r.sigmaG.shape == (T, G)


def test_horseshoe_full_size_residual_identity(brr):
    """BASELINE config 5 size: HorseshoeR, N = 100,000 x M = 100,000 (every marker moves every sweep)"""
    N, M, T = 100000, 100000, 2
    g = brr.Genotypes.synthetic(N, M, seed=79)
    rng = np.random.default_rng(5)
    b = np.zeros(M); idx = rng.choice(M, 1000, replace=False); b[idx] = rng.normal(0, np.sqrt(0.5 / 1000), size=1000)
    y = g.matvec(b) + rng.normal(0, np.sqrt(0.5), size=N)
    y = (y - y.mean()) / y.std(ddof=1)
    p0 = 0.1 * M
    c = brr.Chain(g, brr.HORSESHOE, T, seed=7, Y=y, A=(1 / np.sqrt(N)) * p0 / (M - p0), v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0)
    r = HsRow(c.run(T, emit_all=True), N, M)
    for t in range(T):
        assert rel_inf(r.eps[t], y - r.mu[t] - g.matvec(r.beta[t])) < 1e-9, t
    assert np.all(r.lam > 0) and np.all(r.tau > 0) and np.all(r.sigmaE > 0)


# ------------------------------------------------------------------------------------------------ Groups / restart
def _groups_case(po, N, M, G, F, seed):
    d = po.synth(N, M, seed=seed)
    rng = np.random.default_rng(seed + 1)
    gA = np.sort(rng.integers(0, G, size=M)).astype(np.int32)
    cva = np.tile(np.array(CVA), (G, 1)) * (1.0 + 0.5 * np.arange(G))[:, None]
    fixed = None
    if F:
        fixed = rng.normal(size=(N, F))
        fixed = (fixed - fixed.mean(0)) / fixed.std(0, ddof=1)
        d["y"] = d["y"] + fixed @ rng.normal(scale=0.3, size=F)
    return d, gA, cva, fixed


def _compare_groups(o, rows, N, M, G, F, restart=False):
    a, b = GroupsRow(rows, N, M, G, F, restart), GroupsRow(o["rows"], N, M, G, F, restart)
    assert np.array_equal(a.comp, b.comp), "assignments differ in %d places" % int((a.comp != b.comp).sum())
    assert_trace_close("beta", a.beta, b.beta, TOL)
    assert_trace_close("epsilon", a.eps, b.eps, TOL)
    assert_trace_close("sigmaG", a.sigmaG, b.sigmaG, TOL)
    assert rel_inf(a.mu, b.mu) <= TOL and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL)
    if not restart:
        if F:   # with F == 0 sigmaF is a draw from an inverse gamma of shape v0E/2 that nothing reads (often inf)
            assert_trace_close("alpha", a.alpha, b.alpha, TOL)
            assert np.all(np.abs(a.sigmaF / b.sigmaF - 1) <= TOL)


@pytest.mark.parametrize("N,M,G,F", [(1200, 400, 3, 4), (900, 330, 22, 0), (700, 200, 2, 1)])
def test_groups_chain_matches_oracle(po, brr, N, M, G, F):
    T = 20
    d, gA, cva, fixed = _groups_case(po, N, M, G, F, seed=50 + G)
    o = po.run_groups(d["X"], d["y"], cva, G, gA, fixed, T, seed=77, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.GROUPS, T, seed=77, Y=d["y"], cva=cva, groups=G, gAssign=gA, fixed=fixed, **HYP)
    rows = c.run(T, emit_all=True)
    _compare_groups(o, rows, N, M, G, F)
    assert rel_inf(c.pi(), o["pi"][-1]) <= TOL


def test_groups_with_all_zero_fixed_matrix(po, brr):
    """vignettes/BayesRR.Rmd:166 passes an N x 1 zero matrix when there are no fixed effects"""
    N, M, G, T = 600, 150, 2, 10
    d, gA, cva, _ = _groups_case(po, N, M, G, 0, seed=60)
    fixed = np.zeros((N, 1))
    o = po.run_groups(d["X"], d["y"], cva, G, gA, fixed, T, seed=78, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.GROUPS, T, seed=78, Y=d["y"], cva=cva, groups=G, gAssign=gA, fixed=fixed, **HYP)
    _compare_groups(o, c.run(T, emit_all=True), N, M, G, 1)


def test_grstart_continues_a_groups_chain(po, brr):
    N, M, G, T = 1000, 350, 4, 15
    d, gA, cva, _ = _groups_case(po, N, M, G, 0, seed=70)
    first = po.run_groups(d["X"], d["y"], cva, G, gA, None, 12, seed=5, **HYP)
    last = GroupsRow(first["rows"][-1:], N, M, G, 0)
    st = dict(mu=float(last.mu[0]), beta=last.beta[0], sigmaE=float(last.sigmaE[0]), sigmaGG=last.sigmaG[0],
              epsilon=last.eps[0], components=last.comp[0])
    o = po.run_grstart(st["mu"], st["beta"], st["sigmaE"], st["sigmaGG"], d["X"], st["epsilon"], st["components"],
                       cva, G, gA, T, seed=6, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.GRSTART, T, seed=6, cva=cva, groups=G, gAssign=gA, **st, **HYP)
    rows = c.run(T, emit_all=True)
    _compare_groups(o, rows, N, M, G, 0, restart=True)
    assert rel_inf(c.pi(), o["pi"][-1]) <= TOL


def test_grstart_residuals_outgrow_the_digit_scale(po, brr):
    """the workers cut their residual slice into fixed-point digits against a scale taken at the start of the sweep (2^10 of
    headroom); a restart from residuals seven orders of magnitude smaller than what the first sweep makes of them (BRV2Grstart takes
    epsilon and beta independently, src/BRv2Grstart.cpp:61-67) must take a new scale inside the sweep -- never a wrong dot"""
    N, M, G, T = 1000, 350, 4, 6
    d, gA, cva, _ = _groups_case(po, N, M, G, 0, seed=71)
    first = po.run_groups(d["X"], d["y"], cva, G, gA, None, 12, seed=5, **HYP)
    last = GroupsRow(first["rows"][-1:], N, M, G, 0)
    st = dict(mu=float(last.mu[0]), beta=last.beta[0], sigmaE=float(last.sigmaE[0]), sigmaGG=last.sigmaG[0],
              epsilon=last.eps[0] * 1e-7, components=last.comp[0])
    o = po.run_grstart(st["mu"], st["beta"], st["sigmaE"], st["sigmaGG"], d["X"], st["epsilon"], st["components"],
                       cva, G, gA, T, seed=6, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.GRSTART, T, seed=6, cva=cva, groups=G, gAssign=gA, **st, **HYP)
    rows = c.run(T, emit_all=True)
    _compare_groups(o, rows, N, M, G, 0, restart=True)


# ------------------------------------------------------------------------------------------------ Horseshoe
@pytest.mark.parametrize("N,M,block", [(1000, 300, 128), (650, 129, 64)])
def test_horseshoe_chain_matches_oracle(po, brr, N, M, block):
    T = 20
    d = po.synth(N, M, seed=80)
    p0 = 0.1 * M
    A = (1 / np.sqrt(N)) * p0 / (M - p0)
    kw = dict(v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0)
    o = po.run_horseshoe(d["X"], d["y"], A, T, seed=91, **kw)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.HORSESHOE, T, seed=91, Y=d["y"], A=A, block=block, **kw)
    rows = c.run(T, emit_all=True)
    a, b = HsRow(rows, N, M), HsRow(o["rows"], N, M)
    assert_trace_close("beta", a.beta, b.beta, TOL)
    assert_trace_close("epsilon", a.eps, b.eps, TOL)
    assert_trace_close("lambda", a.lam, b.lam, TOL)
    assert np.all(np.abs(a.tau / b.tau - 1) <= TOL) and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL) and rel_inf(a.mu, b.mu) <= TOL
    assert rel_inf(c.hyper(), o["hyper"][-1]) <= TOL


def test_horseshoe_replay(po, brr):
    N, M, T = 500, 120, 8
    d = po.synth(N, M, seed=81)
    A = 0.02
    t = po.DrawTables(T, M, n_gam=4, n_init_u=1, n_init_g=2 * M + 2, horseshoe=True)
    o = po.run_horseshoe(d["X"], d["y"], A, T, seed=3, tables=t, record=True)
    o2 = po.run_horseshoe(d["X"], d["y"], A, T, source=po.SRC_REPLAY, tables=t)
    assert np.array_equal(o["rows"], o2["rows"])
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.HORSESHOE, T, Y=d["y"], A=A, v0E=1e-3, s02E=1e-3)
    c.set_replay(t)
    rows = c.run(T, emit_all=True)
    a, b = HsRow(rows, N, M), HsRow(o["rows"], N, M)
    assert_trace_close("beta", a.beta, b.beta, TOL)
    assert_trace_close("lambda", a.lam, b.lam, TOL)


# ------------------------------------------------------------------------------------------------ entry points + writer
def _parse(path, header=True):
    lines = open(path).read().split("\n")
    assert lines[-1] == ""
    body = lines[1:-1] if header else lines[:-1]
    return (lines[0] if header else None), [np.array([float(x) for x in ln.split(", ")]) for ln in body]


def test_entry_point_v2_writes_the_reference_file(po, brr, tmp_path):
    N, M = 400, 90
    d = po.synth(N, M, seed=30)
    out = tmp_path / "chain.csv"
    brr.BayesRSamplerV2(str(out), 17, 40, 20, 5, d["X"], d["y"], 0.01, 1e-4, 1e-3, 1e-4, 1e-3, CVA)
    o = po.run_v2(d["X"], d["y"], CVA, 40, burn_in=20, thinning=5, seed=17, emit_all=False, **HYP)
    header, rows = _parse(out)
    assert header + "\n" == po.format_header(po.KIND_V2, N, M)
    assert len(rows) == o["n_rows"] == 4
    text = open(out).read().split("\n")[1:-1]
    for i, r in enumerate(rows):
        assert np.allclose(r, o["rows"][i], rtol=2e-5, atol=1e-12)
        want = po.format_row(o["rows"][i]).rstrip("\n").split(", ")
        got = text[i].split(", ")
        same = sum(a == b for a, b in zip(got, want))
        assert same >= 0.999 * len(want)            # "%g" text equal except for values straddling a rounding boundary


def test_entry_points_emit_the_reference_messages(po, brr, tmp_path):
    """src/BayesRv2.cpp:173-175,276-278: "iteration: <n>" before every iteration n > 0 with n % (max_iterations / 10) == 0 and
    "duration: <s>s" at the end; HorseshoeR adds "initial eta" / "initial tau" and tau / eta / sigmaE lines (HorseshoeR.cpp:191-206)"""
    N, M = 300, 60
    d = po.synth(N, M, seed=35)
    got = []
    brr.set_message_handler(got.append)
    try:
        brr.BayesRSamplerV2(str(tmp_path / "m.csv"), 1, 45, 10, 5, d["X"], d["y"], 0.01, 1e-4, 1e-3, 1e-4, 1e-3, CVA)
        assert got[:-1] == ["iteration: %d\n" % i for i in range(4, 45, 4)]
        assert got[-1].startswith("duration: ") and got[-1].endswith("s\n")
        del got[:]
        brr.HorseshoeR(str(tmp_path / "mh.csv"), 1, 20, 10, 5, d["X"], d["y"], 0.05, 1e-3, 1e-3, 1.0, 1.0, 1.0, 10.0, 10.0)
        assert got[0].startswith("initial eta ") and got[1].startswith("initial tau ")
        assert got[2:6] == ["iteration: 2\n"] + got[3:6] and got[3].startswith(" tau ") and got[4].startswith(" eta ") and got[5].startswith("sigmaE")
        assert sum(1 for g in got if g.startswith("iteration: ")) == 9 and got[-1].startswith("duration: ")
        o = po.run_horseshoe(d["X"], d["y"], 0.05, 20, burn_in=10, thinning=5, seed=1, emit_all=False)
        rows = [np.array([float(x) for x in ln.split(", ")]) for ln in open(tmp_path / "mh.csv").read().split("\n")[1:-1]]
        assert len(rows) == 2 and all(np.allclose(r, o["rows"][i], rtol=2e-5, atol=1e-12) for i, r in enumerate(rows))   # chunked run == one run
    finally:
        brr.set_message_handler(None)
    del got[:]
    brr.BayesRSamplerV2(str(tmp_path / "m2.csv"), 1, 12, 2, 5, d["X"], d["y"], 0.01, 1e-4, 1e-3, 1e-4, 1e-3, CVA)
    assert got == []


def test_entry_points_groups_restart_horseshoe_files(po, brr, tmp_path):
    N, M, G, F = 300, 70, 2, 2
    d, gA, cva, fixed = _groups_case(po, N, M, G, F, seed=90)
    out = tmp_path / "g.csv"
    brr.BayesRSamplerV2Groups(str(out), 3, 30, 10, 10, d["X"], d["y"], 0.01, 1e-4, 1e-3, 1e-4, 1e-3, cva, G, gA, fixed)
    o = po.run_groups(d["X"], d["y"], cva, G, gA, fixed, 30, burn_in=10, thinning=10, seed=3, emit_all=False, **HYP)
    header, rows = _parse(out)
    assert header + "\n" == po.format_header(po.KIND_GROUPS, N, M, G, F)
    assert len(rows) == 2 and all(np.allclose(r, o["rows"][i], rtol=2e-5, atol=1e-12) for i, r in enumerate(rows))
    # restart from the last kept row: no header line in the reference (initialize_file is never called, src/BRv2Grstart.cpp)
    last = GroupsRow(o["rows"][-1:], N, M, G, F)
    out2 = tmp_path / "r.csv"
    brr.BRV2Grstart(str(out2), 4, 12, 2, 2, float(last.mu[0]), last.beta[0], float(last.sigmaE[0]), last.sigmaG[0], d["X"],
                    last.eps[0], last.comp[0], 0.01, 1e-4, 1e-3, 1e-4, 1e-3, cva, G, gA)
    o2 = po.run_grstart(float(last.mu[0]), last.beta[0], float(last.sigmaE[0]), last.sigmaG[0], d["X"], last.eps[0], last.comp[0],
                        cva, G, gA, 12, burn_in=2, thinning=2, seed=4, emit_all=False, **HYP)
    _, rows2 = _parse(out2, header=False)
    assert len(rows2) == o2["n_rows"] == 5 and all(np.allclose(r, o2["rows"][i], rtol=2e-5, atol=1e-12) for i, r in enumerate(rows2))
    out3 = tmp_path / "h.csv"
    brr.HorseshoeR(str(out3), 9, 20, 10, 5, d["X"], d["y"], 0.05, 1e-3, 1e-3, 1.0, 1.0, 1.0, 10.0, 10.0)
    o3 = po.run_horseshoe(d["X"], d["y"], 0.05, 20, burn_in=10, thinning=5, seed=9, emit_all=False)
    lines = open(out3).read().split("\n")
    assert lines[0] + "\n" == po.format_header(po.KIND_HORSESHOE, N, M) and lines[0].endswith(",")
    rows3 = [np.array([float(x) for x in ln.split(", ")]) for ln in lines[1:-1]]
    assert len(rows3) == o3["n_rows"] == 2 and all(np.allclose(r, o3["rows"][i], rtol=2e-5, atol=1e-12) for i, r in enumerate(rows3))


# ------------------------------------------------------------------------------------------------ binary sink, checkpoint / resume
def test_binary_sink_is_lossless(po, brr, tmp_path):
    """the binary file holds the sample rows bit for bit (the CSV keeps 6 significant digits)"""
    N, M, T = 500, 130, 12
    d = po.synth(N, M, seed=33)
    g = brr.Genotypes.from_dense(d["X"])
    c = brr.Chain(g, brr.V2, T, burn_in=2, thinning=3, seed=8, Y=d["y"], cva=CVA, **HYP)
    c.open_output(str(tmp_path / "s.csv")); c.open_binary_output(str(tmp_path / "s.bin"))
    rows = c.run(T)
    c.close_output()
    meta, got = brr.read_binary_samples(str(tmp_path / "s.bin"))
    assert meta == dict(kind=brr.V2, groups=1, N=N, M=M, F=0, row_len=2 * M + 4 + N)
    assert got.shape == rows.shape == (3, 2 * M + 4 + N) and np.array_equal(got, rows)        # iterations 3, 6, 9
    _, text_rows = _parse(tmp_path / "s.csv")
    assert len(text_rows) == 3 and np.allclose(text_rows[-1], rows[-1], rtol=2e-5, atol=1e-12)


@pytest.mark.parametrize("kind", ["v2", "groups", "horseshoe"])
def test_checkpoint_resume_continues_the_chain_bit_for_bit(po, brr, tmp_path, kind):
    """run 5 iterations, save, destroy; a fresh chain that loads the file produces exactly the rows the
    uninterrupted chain produces (Philox draws are keyed by the iteration, the marker order is part of the state)"""
    N, M, T, cut = 700, 300, 13, 5
    if kind == "groups":
        d, gA, cva, fixed = _groups_case(po, N, M, 3, 2, seed=95)
        mk = lambda g: brr.Chain(g, brr.GROUPS, T, seed=21, Y=d["y"], cva=cva, groups=3, gAssign=gA, fixed=fixed, **HYP)
    elif kind == "horseshoe":
        d = po.synth(N, M, seed=96)
        mk = lambda g: brr.Chain(g, brr.HORSESHOE, T, seed=22, Y=d["y"], A=0.03, v0E=1e-3, s02E=1e-3)
    else:
        d = po.synth(N, M, seed=97)
        mk = lambda g: brr.Chain(g, brr.V2, T, seed=23, Y=d["y"], cva=CVA, **HYP)
    g = brr.Genotypes.from_dense(d["X"])
    whole = mk(g).run(T, emit_all=True)
    a = mk(g)
    first = a.run(cut, emit_all=True)
    a.save(str(tmp_path / "chain.ckpt"))
    a.close()
    b = mk(g)
    b.load(str(tmp_path / "chain.ckpt"))
    rest = np.vstack([b.run(3, emit_all=True), b.run(T - cut - 3, emit_all=True)])
    assert np.array_equal(first, whole[:cut]) and np.array_equal(rest, whole[cut:])
    # a checkpoint of another chain is refused
    other = brr.Chain(g, brr.V2, T, seed=99, Y=d["y"], cva=CVA, **HYP)
    with pytest.raises(brr.BayesRRError) as e:
        other.load(str(tmp_path / "chain.ckpt"))
    assert e.value.code == brr.E_ARG


# ------------------------------------------------------------------------------------------------ golden vectors of the reference's own sources
def _gold(name):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"), allow_pickle=False)
    return z, np.asfortranarray((z["G"] - z["mean"]) / z["sd"])


def test_gpu_replays_reference_golden_v2(po, brr):
    """tests/golden/v2.npz holds rows packed by the reference's own BayesRv2.cpp; the draws it consumed are re-derived by
    running the oracle on the same sequential stream with recording on, then replayed on the GPU."""
    z, X = _gold("v2")
    N, M = X.shape
    T, burn, thin = int(z["max_iterations"]), int(z["burn_in"]), int(z["thinning"])
    t = po.DrawTables(T, M, n_gam=6, n_init_u=1)
    pi0 = [0.5, np.nan, np.nan, np.nan]
    po.run_v2(X, z["y"], z["cva"], T, source=po.SRC_SEQ, seed=int(z["seed"]), tables=t, record=True, pi_init=pi0, **HYP)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.V2, T, burn_in=burn, thinning=thin, Y=z["y"], cva=z["cva"], pi_init=pi0, **HYP)
    c.set_replay(t)
    rows = c.run(T, emit_all=False)
    a, b = V2Row(rows, N, M), V2Row(z["rows"], N, M)
    assert rows.shape == z["rows"].shape and np.array_equal(a.comp, b.comp)
    assert_trace_close("beta", a.beta, b.beta, TOL); assert_trace_close("epsilon", a.eps, b.eps, TOL)
    assert rel_inf(a.sigmaG, b.sigmaG) <= TOL and rel_inf(a.sigmaE, b.sigmaE) <= TOL and rel_inf(a.mu, b.mu) <= TOL


def test_gpu_replays_reference_golden_groups_and_restart(po, brr):
    z, X = _gold("groups")
    N, M = X.shape; G, F, K = 3, 2, 4
    T, burn, thin = int(z["max_iterations"]), int(z["burn_in"]), int(z["thinning"])
    t = po.DrawTables(T, M, n_gam=2 + G * (K + 1), F=F, n_init_u=G + 1)
    po.run_groups(X, z["y"], z["cva"], G, z["gAssign"], z["fixed"], T, source=po.SRC_SEQ, seed=int(z["seed"]), tables=t, record=True, **HYP)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.GROUPS, T, burn_in=burn, thinning=thin, Y=z["y"], cva=z["cva"], groups=G, gAssign=z["gAssign"], fixed=z["fixed"], **HYP)
    c.set_replay(t)
    _compare_groups(dict(rows=z["rows"]), c.run(T, emit_all=False), N, M, G, F)
    z, X = _gold("grstart")
    T, burn, thin = int(z["max_iterations"]), int(z["burn_in"]), int(z["thinning"])
    t = po.DrawTables(T, M, n_gam=2 + G * (K + 1), n_init_g=G * (K + 1))
    st = dict(mu=float(z["mu"]), beta=z["beta"], sigmaE=float(z["sigmaE"]), sigmaGG=z["sigmaGG"], epsilon=z["epsilon"], components=z["components"])
    po.run_grstart(st["mu"], st["beta"], st["sigmaE"], st["sigmaGG"], X, st["epsilon"], st["components"], z["cva"], G, z["gAssign"], T,
                   source=po.SRC_SEQ, seed=int(z["seed"]), tables=t, record=True, **HYP)
    c = brr.Chain(g, brr.GRSTART, T, burn_in=burn, thinning=thin, cva=z["cva"], groups=G, gAssign=z["gAssign"], **st, **HYP)
    c.set_replay(t)
    _compare_groups(dict(rows=z["rows"]), c.run(T, emit_all=False), N, M, G, 0, restart=True)


def test_gpu_replays_reference_golden_horseshoe(po, brr):
    z, X = _gold("horseshoe")
    N, M = X.shape
    T, burn, thin = int(z["max_iterations"]), int(z["burn_in"]), int(z["thinning"])
    t = po.DrawTables(T, M, n_gam=4, n_init_u=1, n_init_g=2 * M + 2, horseshoe=True)
    po.run_horseshoe(X, z["y"], float(z["A"]), T, source=po.SRC_SEQ, seed=int(z["seed"]), tables=t, record=True)
    g = brr.Genotypes.from_dense(X)
    c = brr.Chain(g, brr.HORSESHOE, T, burn_in=burn, thinning=thin, Y=z["y"], A=float(z["A"]), v0E=1e-3, s02E=1e-3)
    c.set_replay(t)
    rows = c.run(T, emit_all=False)
    a, b = HsRow(rows, N, M), HsRow(z["rows"], N, M)
    for name in ("beta", "eps", "lam"):
        assert_trace_close(name, getattr(a, name), getattr(b, name), TOL)
    assert rel_inf(a.tau, b.tau) <= TOL and rel_inf(a.sigmaE, b.sigmaE) <= TOL and rel_inf(a.mu, b.mu) <= TOL
