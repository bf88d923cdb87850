"""Row-sharded chains: host plumbing on CPU (thread ranks, gloo world_size 2), and on the GPU the invariants
  * after every iteration all ranks hold bit-identical rows (beta, components, mu, sigma, residuals of ALL ranks),
  * the sharded chain agrees with the unsharded CPU oracle: assignments exact, traces within 1e-9 (inf-norm relative).
Two ranks run as threads on one device (peer exchange through same-process addresses) and, where two devices exist, as two
processes over CUDA IPC / NVLink."""
import os
import sys

import numpy as np
import pytest

from conftest import CVA, HYP
from helpers import GroupsRow, HsRow, V2Row, assert_trace_close, rel_inf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-9


def pack_codes(G):
    """N x M codes in {0,1,2} -> (M, ceil(N/4)) bytes, individual i in bits 2*(i%4) of byte i/4"""
    N, M = G.shape
    out = np.zeros((M, (N + 3) // 4), dtype=np.uint8)
    for q in range(4):
        rows = np.arange(q, N, 4)
        out[:, :len(rows)] |= (G[rows, :].T.astype(np.uint8) << (2 * q))
    return out


# ------------------------------------------------------------------------------------------------ CPU: host plumbing
def test_shard_bounds_cover_all_rows():
    from bayesrrcpp_b200.sharded import shard_bounds
    for n, w in [(1000, 3), (100000, 2), (64, 2), (400000, 8), (130, 4)]:
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(lo % 64 == 0 for lo, _ in b)


def test_thread_ranks_callbacks_through_the_c_abi(brr):
    from bayesrrcpp_b200 import sharded
    tg = sharded.ThreadGroup(3)
    res = tg.run(lambda r, comm: comm.selftest(np.arange(5.0) * (r + 1) + 0.1 * r, 100 + r))
    for buf, got in res:
        assert np.array_equal(buf, res[0][0]) and np.array_equal(got, [100, 101, 102])
    assert np.allclose(res[0][0], np.arange(5.0) * 6 + 0.3)


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bayesrrcpp_b200 import sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = sharded.torch_comm()
        buf, got = comm.selftest(np.array([1.0 + rank, 0.25 * rank, 3.0]), 7 + rank)
        q.put((rank, buf.tolist(), got.tolist()))
    finally:
        dist.destroy_process_group()


def test_torch_distributed_gloo_world2_callbacks(brr):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for rank, buf, got in res:
        assert buf == [3.0, 0.25, 6.0] and got == [7, 8]


# ------------------------------------------------------------------------------------------------ GPU: thread ranks on one device
def _shard_geno(brr, d, lo, hi, comm, device=0):
    g = brr.Genotypes.from_packed(pack_codes(d["G"][lo:hi]), hi - lo, device=device)
    return g.shard_stats(comm)


def _run_sharded(brr, world, d, make_chain, T, device_of=lambda r: 0):
    from bayesrrcpp_b200 import sharded
    bounds = sharded.shard_bounds(d["G"].shape[0], world)
    tg = sharded.ThreadGroup(world)

    def fn(r, comm):
        lo, hi = bounds[r]
        g = _shard_geno(brr, d, lo, hi, comm, device_of(r))
        c = make_chain(g, comm, lo, hi)
        rows = c.run(T, emit_all=True)
        extra = c.pi() if c.kind != brr.HORSESHOE else c.hyper()
        st = g.stats()
        c.close(); g.close()
        return rows, extra, st
    # The ranks' persistent kernels wait for each other, and CUDA does not promise that kernels of several streams of ONE device
    # run side by side (three ranks with 128-marker blocks reproducibly do not, while four real GPUs do: tools/gpu_multi.sh).
    # A watchdog time-out here is therefore retried once before it counts as a failure.
    # A second time-out skips the test: it says nothing about the protocol (one process per GPU, test_sharded_one_process_per_gpu, has no
    # such dependence and carries the parity weight), and a chain that failed keeps its exchange window, so nothing later is affected.
    try:
        return tg.run(fn)
    except brr.BayesRRError as e:
        if "watchdog" not in str(e):
            raise
        print("first attempt:", e, file=sys.stderr)
    try:
        return sharded.ThreadGroup(world).run(fn)
    except brr.BayesRRError as e:
        if "watchdog" not in str(e):
            raise
        pytest.skip("the device did not run the %d ranks' persistent kernels side by side (twice): co-scheduling of kernels that share one device is not promised by CUDA" % world)


@pytest.mark.gpu
@pytest.mark.parametrize("world,N,M,block", [(2, 1500, 420, 64), (2, 2000, 300, 128), (3, 1500, 420, 64)])
def test_sharded_v2_thread_ranks_match_oracle_and_each_other(po, brr, world, N, M, block):
    T = 12
    d = po.synth(N, M, seed=301)
    o = po.run_v2(d["X"], d["y"], CVA, T, seed=302, **HYP)
    res = _run_sharded(brr, world, d, lambda g, comm, lo, hi: brr.Chain(
        g, brr.V2, T, seed=302, Y=d["y"][lo:hi], cva=CVA, block=block, workers=12, comm=comm, **HYP), T)
    rows0, pi0, st0 = res[0]
    assert rel_inf(st0["mean"], d["mean"]) < 1e-14 and rel_inf(st0["sd"], d["sd"]) < 1e-13      # statistics of ALL rows
    for rows, pi, _ in res[1:]:
        assert np.array_equal(rows, rows0) and np.array_equal(pi, pi0), "ranks diverged"
    a, b = V2Row(rows0, N, M), V2Row(o["rows"], N, M)
    assert np.array_equal(a.comp, b.comp)
    assert_trace_close("beta", a.beta, b.beta, TOL)
    assert_trace_close("epsilon", a.eps, b.eps, TOL)
    assert rel_inf(a.mu, b.mu) <= TOL and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL) and np.all(np.abs(a.sigmaG / b.sigmaG - 1) <= TOL)
    assert rel_inf(pi0, o["pi"][-1]) <= TOL


@pytest.mark.gpu
def test_sharded_groups_with_fixed_effects_thread_ranks(po, brr):
    from test_gpu_parity import _groups_case
    N, M, G, F, T = 1300, 260, 3, 3, 10
    d, gA, cva, fixed = _groups_case(po, N, M, G, F, seed=310)
    o = po.run_groups(d["X"], d["y"], cva, G, gA, fixed, T, seed=311, **HYP)
    res = _run_sharded(brr, 2, d, lambda g, comm, lo, hi: brr.Chain(
        g, brr.GROUPS, T, seed=311, Y=d["y"][lo:hi], cva=cva, groups=G, gAssign=gA, fixed=fixed[lo:hi], workers=12, comm=comm, **HYP), T)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    a, b = GroupsRow(res[0][0], N, M, G, F), GroupsRow(o["rows"], N, M, G, F)
    assert np.array_equal(a.comp, b.comp)
    for name in ("beta", "eps", "sigmaG", "alpha"):
        assert_trace_close(name, getattr(a, name), getattr(b, name), TOL)
    assert rel_inf(a.mu, b.mu) <= TOL and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL) and np.all(np.abs(a.sigmaF / b.sigmaF - 1) <= TOL)


@pytest.mark.gpu
def test_sharded_horseshoe_thread_ranks(po, brr):
    N, M, T = 1100, 200, 10
    d = po.synth(N, M, seed=320)
    A = (1 / np.sqrt(N)) * (0.1 * M) / (M - 0.1 * M)
    kw = dict(v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0)
    o = po.run_horseshoe(d["X"], d["y"], A, T, seed=321, **kw)
    res = _run_sharded(brr, 2, d, lambda g, comm, lo, hi: brr.Chain(
        g, brr.HORSESHOE, T, seed=321, Y=d["y"][lo:hi], A=A, workers=12, comm=comm, **kw), T)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    a, b = HsRow(res[0][0], N, M), HsRow(o["rows"], N, M)
    for name in ("beta", "eps", "lam"):
        assert_trace_close(name, getattr(a, name), getattr(b, name), TOL)
    assert np.all(np.abs(a.tau / b.tau - 1) <= TOL) and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL)


@pytest.mark.gpu
def test_sharded_chain_from_bed_shards_with_checkpoint(po, brr, tmp_path):
    """two ranks read their own rows of one PLINK .bed file, run a sharded chain, checkpoint it (one file per rank), and fresh
    chains resumed from the files finish it: the whole thing equals the unsharded oracle run (assignments exact, 1e-9)"""
    from test_gpu_parity import _write_plink
    from bayesrrcpp_b200 import sharded
    N, M, T, cut = 1400, 330, 10, 4
    d = po.synth(N, M, seed=340)
    prefix = str(tmp_path / "cohort")
    _write_plink(prefix, 2 - d["G"])                 # the file counts the OTHER allele: codes = 2 - G, i.e. x -> -x after scaling
    o = po.run_v2(-d["X"], d["y"], CVA, T, seed=341, **HYP)
    bounds = sharded.shard_bounds(N, 2)

    def run(first):
        def fn(r, comm):
            lo, hi = bounds[r]
            g = brr.Genotypes.from_bed(prefix, rows=(lo, hi - lo)).shard_stats(comm)
            c = brr.Chain(g, brr.V2, T, seed=341, Y=d["y"][lo:hi], cva=CVA, workers=12, comm=comm, **HYP)
            ck = str(tmp_path / ("rank%d.ckpt" % r))
            if first:
                rows = c.run(cut, emit_all=True)
                c.save(ck)
            else:
                c.load(ck)
                rows = c.run(T - cut, emit_all=True)
            c.close(); g.close()
            return rows
        try:
            return sharded.ThreadGroup(2).run(fn)
        except brr.BayesRRError as e:      # see _run_sharded: co-scheduling of the ranks' kernels on one device is not guaranteed
            if "watchdog" not in str(e):
                raise
        try:
            return sharded.ThreadGroup(2).run(fn)
        except brr.BayesRRError as e:
            if "watchdog" not in str(e):
                raise
            pytest.skip("the device did not run the two ranks' persistent kernels side by side (twice)")
    a, b = run(True), run(False)
    assert np.array_equal(a[0], a[1]) and np.array_equal(b[0], b[1]), "ranks diverged"
    got, want = V2Row(np.vstack([a[0], b[0]]), N, M), V2Row(o["rows"], N, M)
    assert np.array_equal(got.comp, want.comp)
    assert_trace_close("beta", got.beta, want.beta, TOL)
    assert_trace_close("epsilon", got.eps, want.eps, TOL)
    assert np.all(np.abs(got.sigmaE / want.sigmaE - 1) <= TOL) and np.all(np.abs(got.sigmaG / want.sigmaG - 1) <= TOL)


@pytest.mark.gpu
def test_bed_row_shards_impute_missing_genotypes_like_the_whole_file(po, brr, tmp_path):
    """a missing genotype is filled with the rounded mean over the observed genotypes of the WHOLE column, whatever the sharding
    (per-shard means made the data depend on the world size)"""
    from test_gpu_parity import _write_plink
    from bayesrrcpp_b200 import sharded
    N, M = 1200, 90
    rng = np.random.default_rng(7)
    G = np.zeros((N, M), dtype=np.int8)
    # columns whose upper and lower halves have different allele frequencies: shard means round differently from the global mean
    G[:N // 2] = rng.binomial(2, 0.12, size=(N // 2, M)); G[N // 2:] = rng.binomial(2, 0.62, size=(N - N // 2, M))
    miss = rng.uniform(size=(N, M)) < 0.01
    miss[:, 5] = False; miss[3, 5] = True                       # a column with a single missing genotype, in the first shard only
    prefix = str(tmp_path / "gaps")
    _write_plink(prefix, G, miss)
    whole = brr.Genotypes.from_bed(prefix, impute_missing=True)
    want, st_want = whole.unpack(), whole.stats()
    for world in (2, 3):
        bounds = sharded.shard_bounds(N, world)

        def fn(r, comm):
            lo, hi = bounds[r]
            g = brr.Genotypes.from_bed(prefix, rows=(lo, hi - lo), impute_missing=True)
            with pytest.raises(brr.BayesRRError):                # not usable before the collective call
                g.stats()
            g.shard_stats(comm)
            out = g.unpack(), g.stats(), g.n_missing
            g.close()
            return out
        res = sharded.ThreadGroup(world).run(fn)
        got = np.vstack([r[0] for r in res])
        assert np.array_equal(got, want), "world %d: %d genotypes differ from the unsharded imputation" % (world, int((got != want).sum()))
        assert sum(r[2] for r in res) == whole.n_missing == int(miss.sum())
        for r in res:
            assert rel_inf(r[1]["mean"], st_want["mean"]) < 1e-14 and rel_inf(r[1]["sd"], st_want["sd"]) < 1e-13
    # the per-shard rule would have given other data: the test is not vacuous
    lo, hi = sharded.shard_bounds(N, 2)[0]
    local = G[lo:hi].astype(float); local[miss[lo:hi]] = np.nan
    shard_fill = np.floor(np.nanmean(local, axis=0) + 0.5)
    glob = G.astype(float); glob[miss] = np.nan
    assert np.any(shard_fill != np.floor(np.nanmean(glob, axis=0) + 0.5))


# ------------------------------------------------------------------------------------------------ GPU: one process per device
PROC_N, PROC_M, PROC_T = 6000, 700, 12
HS_KW = dict(v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0)


def _proc_case(po, case):
    """inputs of one process-per-GPU case (every rank and the checking process build the same ones)"""
    from test_gpu_parity import _groups_case
    N, M = PROC_N, PROC_M
    if case == "groups":
        d, gA, cva, fixed = _groups_case(po, N, M, 3, 2, seed=333)
        return dict(d=d, gA=gA, cva=cva, fixed=fixed, G=3, F=2)
    return dict(d=po.synth(N, M, seed=331))


def _proc_worker(rank, world, port, q, case):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import bayesrrcpp_b200 as brr
    from bayesrrcpp_b200 import sharded
    from oracle import pyoracle as po
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        comm = sharded.torch_comm()
        N, M, T = PROC_N, PROC_M, PROC_T
        cs = _proc_case(po, case)
        d = cs["d"]
        lo, hi = sharded.shard_bounds(N, world)[rank]
        g = brr.Genotypes.from_packed(pack_codes(d["G"][lo:hi]), hi - lo, device=rank).shard_stats(comm)
        if case == "groups":
            c = brr.Chain(g, brr.GROUPS, T, seed=332, Y=d["y"][lo:hi], cva=cs["cva"], groups=cs["G"], gAssign=cs["gA"],
                          fixed=cs["fixed"][lo:hi], comm=comm, **HYP)
        elif case == "horseshoe":
            A = (1 / np.sqrt(N)) * (0.1 * M) / (M - 0.1 * M)
            c = brr.Chain(g, brr.HORSESHOE, T, seed=332, Y=d["y"][lo:hi], A=A, comm=comm, **HS_KW)
        else:
            c = brr.Chain(g, brr.V2, T, seed=332, Y=d["y"][lo:hi], cva=CVA, comm=comm, **HYP)
        rows = c.run(T, emit_all=True)
        # device-side check over NCCL that the ranks are bit-identical: max == min of every entry
        t = torch.from_numpy(rows).cuda()
        hi_t, lo_t = t.clone(), t.clone()
        dist.all_reduce(hi_t, op=dist.ReduceOp.MAX); dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(hi_t, lo_t))
        q.put((rank, same, rows if rank == 0 else None))
        c.close(); g.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["v2", "groups", "horseshoe"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_one_process_per_gpu(po, brr, world, case):
    """`world` processes, one per GPU, CUDA-IPC exchange windows over NVLink: all ranks bit-identical (checked on the devices
    over NCCL) and equal to the unsharded oracle (assignments exact, 1e-9); reference src/BayesRv2.cpp:186-245,
    src/BayesRv2Groups.cpp:216-298, src/HorseshoeR.cpp:219-240"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run under gpurun --gpus %d)" % (world, world))
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + world * 3 + len(case)) % 2000
    ps = [ctx.Process(target=_proc_worker, args=(r, world, port, q, case)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=900) for _ in range(world)), key=lambda x: x[0])
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert all(same for _, same, _ in res), "ranks diverged"
    N, M, T = PROC_N, PROC_M, PROC_T
    cs = _proc_case(po, case)
    d = cs["d"]
    if case == "groups":
        G, F = cs["G"], cs["F"]
        o = po.run_groups(d["X"], d["y"], cs["cva"], G, cs["gA"], cs["fixed"], T, seed=332, **HYP)
        a, b = GroupsRow(res[0][2], N, M, G, F), GroupsRow(o["rows"], N, M, G, F)
        assert np.array_equal(a.comp, b.comp)
        for name in ("beta", "eps", "sigmaG", "alpha"):
            assert_trace_close(name, getattr(a, name), getattr(b, name), TOL)
        assert rel_inf(a.mu, b.mu) <= TOL and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL) and np.all(np.abs(a.sigmaF / b.sigmaF - 1) <= TOL)
    elif case == "horseshoe":
        A = (1 / np.sqrt(N)) * (0.1 * M) / (M - 0.1 * M)
        o = po.run_horseshoe(d["X"], d["y"], A, T, seed=332, **HS_KW)
        a, b = HsRow(res[0][2], N, M), HsRow(o["rows"], N, M)
        for name in ("beta", "eps", "lam"):
            assert_trace_close(name, getattr(a, name), getattr(b, name), TOL)
        assert np.all(np.abs(a.tau / b.tau - 1) <= TOL) and np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL)
    else:
        o = po.run_v2(d["X"], d["y"], CVA, T, seed=332, **HYP)
        a, b = V2Row(res[0][2], N, M), V2Row(o["rows"], N, M)
        assert np.array_equal(a.comp, b.comp)
        assert_trace_close("beta", a.beta, b.beta, TOL)
        assert_trace_close("epsilon", a.eps, b.eps, TOL)
        assert np.all(np.abs(a.sigmaE / b.sigmaE - 1) <= TOL) and np.all(np.abs(a.sigmaG / b.sigmaG - 1) <= TOL)
