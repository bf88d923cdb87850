#!/usr/bin/env python
"""bench.py -- SNP-updates/s of the per-SNP Gibbs sweep (BASELINE.json metric) on N B200s of one node.

A "step" is one Gibbs iteration: one pass of the hot path (block Gram -> persistent sweep -> hyper draws) over all M
markers.  N=1 workload = BASELINE.json configs[1]: BayesRSamplerV2, N=50,000 x M=50,000 synthetic genotypes, simulated
phenotype h2=0.5, K=4.  A multi-GPU run is ONE chain, row-sharded (rank r holds a contiguous slice of the individuals of the
same virtual matrix), every rank replicates the chain, the per-block partial dots are exchanged over NVLink peer memory inside
the sweep kernel:
  default          weak scaling in individuals: --rows (50,000) per GPU, N_total = rows x world
  --total-rows T   strong scaling: T individuals split over the ranks (BASELINE configs[2] / configs[4] at 1/2/4/8 GPUs)

  value     : whole-job SNP-updates/s, genotypes resident in HBM, device-timed (CUDA events on the chain's stream)
  e2e       : same metric through the C ABI with HOST buffers (page-locked): packed genotypes H2D, chain creation, per-iteration
              permutation upload, sample rows D2H + CSV writer (thinning 5) all inside the timed region (wall clock), >= 100 steps
  roofline  : the persistent sweep kernel against the measured HBM copy bandwidth (it is bound by the serial chain, not by HBM --
              DESIGN.md section 3.2); `rooflines` adds the block-Gram kernel against the int8 tensor peak measured here and the
              workers' dot stage against the fp64 pipe peak measured here
  cpu_baseline / --impl reference : the CPU oracle (C restatement of the reference's Eigen sampler, pinned against the
              reference's own sources built over a minimal Eigen/Rcpp shim -- that shim is not a fair timing of Eigen, so the
              port is what is timed) on a bounded column sample of the same workload, 1 core like the reference's sampler thread.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(N=50000, M=50000, cva=[1e-4, 1e-3, 1e-2], h2=0.5, causal_frac=0.1,
           hyp=dict(sigma0=0.01, v0E=1e-4, s02E=1e-3, v0G=1e-4, s02G=1e-3), data_seed=1002, chain_seed=2002,
           chain_burn=20)
HS = dict(v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0)
WORKLOAD = "BayesRSamplerV2 N=50000 x M=50000 K=4 synthetic 2-bit genotypes, simulated phenotype h2=0.5 (BASELINE configs[1])"
CPU_SAMPLE_BYTES = 1.6e9     # dense fp64 column sample the CPU arm sweeps (full N): 4,000 columns at N = 50,000


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measure_int8_peak(dev):
    """dense int8 tensor throughput of this device: cuBLASLt int8 GEMM 8192^3 through torch._int_mm, best of 10 (the same
    recipe MEASURED_PEAKS.json uses for bf16; a library call measures the peak, it is not on the product path)"""
    import torch
    try:
        n = 8192
        a = torch.randint(-3, 4, (n, n), dtype=torch.int8, device="cuda:%d" % dev)
        b = torch.randint(-3, 4, (n, n), dtype=torch.int8, device="cuda:%d" % dev)
        torch._int_mm(a, b); torch._int_mm(a, b)
        best = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(a, b); e1.record(); e1.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3))
        del a, b
        return best / 1e12, "measured here: torch._int_mm (cuBLASLt int8 -> int32) 8192^3, best of 10"
    except Exception as e:          # keep the line: fall back to twice the measured bf16 burst
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                return 2.0 * float(json.load(f)["bf16_tflops"]), "2 x measured bf16 burst (int8 GEMM unavailable: %s)" % type(e).__name__
        except Exception:
            return 4500.0, "nominal 4.5 POP/s (no measurement available)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.all_rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        """launch nvidia-smi early (its start-up takes a while and would otherwise fall into the timed region)"""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.all_rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def __enter__(self):
        if self.proc is None:
            self.start()
        self.t0 = time.time()
        return self

    def __exit__(self, *a):
        self.t1 = time.time()
        if self.proc:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)      # gone before anything else is timed
            except Exception:
                pass

    @property
    def rows(self):
        """samples taken inside the timed region (plus the one right after it when the region is shorter than the sampling period)"""
        inside = [r for t, r in self.all_rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.05]
        return inside if inside else [r for _, r in self.all_rows[-1:]]

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def simulate_phenotype(geno, seed, h2, causal_frac, rank=0, allreduce=None):
    """y = X b + e over this rank's rows (same effects b on every rank), centred and scaled over ALL rows"""
    rng = np.random.default_rng(seed)
    M, N = geno.M, geno.N
    mc = max(1, int(round(causal_frac * M)))
    b = np.zeros(M)
    b[rng.choice(M, mc, replace=False)] = rng.normal(0, np.sqrt(h2 / mc), size=mc)
    y = geno.matvec(b) + np.random.default_rng(seed + 104729 * (rank + 1)).normal(0, np.sqrt(1 - h2), size=N)
    mom = np.array([float(N), y.sum(), (y * y).sum()])
    if allreduce is not None:
        allreduce(mom)
    mean = mom[1] / mom[0]
    sd = np.sqrt((mom[2] - mom[0] * mean * mean) / (mom[0] - 1))
    return (y - mean) / sd


def cpu_sample_data(N, M_s, seed):
    """dense fp64 column sample of the workload for the CPU arm (numpy only: runs without a GPU)"""
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.05, 0.5, size=M_s)
    X = np.empty((N, M_s), order="F")
    for j in range(M_s):
        g = rng.binomial(2, p[j], size=N).astype(np.float64)
        sd = g.std(ddof=1)
        X[:, j] = (g - g.mean()) / (sd if sd > 0 else 1.0)
    b = np.zeros(M_s)
    idx = rng.choice(M_s, max(1, M_s // 10), replace=False)
    b[idx] = rng.normal(0, np.sqrt(0.5 / len(idx)), size=len(idx))
    y = X @ b + rng.normal(0, np.sqrt(0.5), size=N)
    return X, (y - y.mean()) / y.std(ddof=1)


def run_cpu_arm(steps, warmup, sampler="v2", N=None):
    """the reference's CPU algorithm (oracle port), single thread like the reference's sampler thread (its second OpenMP thread
    only writes the CSV, src/BayesRv2.cpp:105-107).  Returns (SNP-updates/s, timed seconds, sample columns, description)."""
    from oracle import pyoracle as po
    po.build()
    N = CFG["N"] if N is None else N
    M_s = int(max(256, min(4000, CPU_SAMPLE_BYTES // (8 * N))))
    X, y = cpu_sample_data(N, M_s, CFG["data_seed"])
    kw = dict(CFG["hyp"])

    def once(n_it, seed):
        if sampler == "groups":
            G = 22
            return po.run_groups(X, y, np.tile(np.array(CFG["cva"]), (G, 1)), G, (np.arange(M_s) * G // M_s).astype(np.int32),
                                 np.zeros((N, 1)), n_it, seed=seed, want_rows=False, **kw)
        if sampler == "horseshoe":
            p0 = 0.1 * M_s
            return po.run_horseshoe(X, y, (1 / np.sqrt(N)) * p0 / (M_s - p0), n_it, seed=seed, want_rows=False, **HS)
        return po.run_v2(X, y, CFG["cva"], n_it, seed=seed, want_rows=False, **kw)
    if warmup > 0:
        once(warmup, 1)                                                  # warm-up (page-in, caches)
    r = once(steps, CFG["chain_seed"])
    rate = M_s * steps / r["seconds"]
    # second flavour: -O3 -march=native (the first is -O2, R's default for packages); the faster one is the baseline
    flavour, secs = "-O2", r["seconds"]
    try:
        # always rebuilt on the machine that runs it: a -march=native library from another host could use missing instructions
        subprocess.run(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), "native"], check=True, capture_output=True)
        po.NATIVE = True
        once(1, 1)
        rn = once(steps, CFG["chain_seed"])
        rate_n = M_s * steps / rn["seconds"]
        if rate_n > rate:
            rate, secs, flavour = rate_n, rn["seconds"], "-O3 -march=native"
    except Exception:
        pass
    finally:
        po.NATIVE = False
    sample = ("%s: full N=%d rows x %d-column dense fp64 sample, %d timed iterations after %d warm-up, oracle built %s (faster of -O2 and "
              "-O3 -march=native; the per-marker cost is independent of M and beta: SNP-updates/s extrapolates to the full M)"
              ) % (sampler, N, M_s, steps, warmup, flavour)
    return rate, secs, M_s, sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=100, help="iterations of the end-to-end leg (at least --steps)")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--workers", type=int, default=0, help="worker CTAs of the sweep kernel (default: the library's split of the SMs)")
    ap.add_argument("--markers", type=int, default=CFG["M"], help="markers M (default: BASELINE configs[1]; 500000 with --gpus 8 = configs[3])")
    ap.add_argument("--rows", type=int, default=CFG["N"], help="individuals per GPU (default: BASELINE configs[1]); weak scaling")
    ap.add_argument("--total-rows", type=int, default=0, help="individuals of the whole chain, split over the GPUs: strong scaling "
                    "(configs[2]: --sampler groups --total-rows 100000 --markers 200000; configs[4]: --sampler horseshoe --total-rows 100000 --markers 100000)")
    ap.add_argument("--sampler", default="v2", choices=["v2", "groups", "horseshoe"])
    ap.add_argument("--burn", type=int, default=CFG["chain_burn"], help="untimed chain burn-in iterations before the warm-up")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        rate, secs, M_s, sample = run_cpu_arm(args.steps, W, args.sampler, args.total_rows or args.rows)
        M = args.markers
        print(json.dumps({"impl": "reference", "metric": "SNP-updates/sec", "value": rate, "unit": "SNP-updates/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": W,
                          "ms_per_step": 1e3 * secs / args.steps,            # TIMED: one sweep over the sample's columns
                          "extrapolated_ms_per_iteration": 1e3 * secs / args.steps * M / M_s, "sample_markers": M_s, "same_config": False,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": WORKLOAD if args.sampler == "v2" and M == CFG["M"] else "%s N=%d x M=%d" % (args.sampler, args.total_rows or args.rows, M),
                                     "sample": sample},
                          "cpu_baseline": {"value": rate, "unit": "SNP-updates/s", "cores": 1, "kind": "port", "sample": sample},
                          "e2e": {"value": rate, "unit": "SNP-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import bayesrrcpp_b200 as brr
    from bayesrrcpp_b200 import sharded
    dist, comm, host_allreduce = None, None, None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = sharded.torch_comm()                    # gloo group for the library's host call-backs (set-up time only)

        def host_allreduce(a):
            t = torch.from_numpy(a).cuda()             # NCCL over NVLink
            dist.all_reduce(t)
            a[:] = t.cpu().numpy()
    dev = local
    M = args.markers
    if args.total_rows:                                # strong scaling: one fixed matrix split over the ranks
        row0, row1 = sharded.shard_bounds(args.total_rows, world)[rank]
        N, N_total = row1 - row0, args.total_rows
    else:                                              # weak scaling: fixed rows per GPU
        N, N_total = args.rows, args.rows * world
        row0 = rank * N
    geno = brr.Genotypes.synthetic(N, M, CFG["data_seed"], row0=row0, device=dev)
    if comm is not None:
        geno.shard_stats(comm)
    y = simulate_phenotype(geno, CFG["data_seed"], CFG["h2"], CFG["causal_frac"], rank, host_allreduce)

    def make_chain(g, n_iter, **kw):
        common = dict(seed=CFG["chain_seed"], Y=y, block=args.block, workers=args.workers, comm=comm, **kw)
        if args.sampler == "groups":   # 22 chromosome-like groups, identical ladders, the vignette's N x 1 zero fixed matrix
            G = 22
            return brr.Chain(g, brr.GROUPS, n_iter, cva=np.tile(np.array(CFG["cva"]), (G, 1)), groups=G,
                             gAssign=(np.arange(M) * G // M).astype(np.int32), fixed=np.zeros((N, 1)), **common, **CFG["hyp"])
        if args.sampler == "horseshoe":
            p0 = 0.1 * M
            return brr.Chain(g, brr.HORSESHOE, n_iter, A=(1 / np.sqrt(N_total)) * p0 / (M - p0), **common, **HS)
        return brr.Chain(g, brr.V2, n_iter, cva=CFG["cva"], **common, **CFG["hyp"])

    total_iters = args.burn + W + args.steps + 1
    chain = make_chain(geno, total_iters)
    clk = ClockSampler(dev).start()        # nvidia-smi is up and sampling before the timed region begins
    young = None
    if args.burn > 0:
        chain.run_discard(args.burn)       # untimed chain burn-in: the timed steps see a settled sparsity pattern
        yms, _ = chain.last_timing()
        yprof = chain.sweep_profile()
        young = {"ms_per_step": yms / args.burn, "iterations": "0..%d" % (args.burn - 1),
                 "state_changing_marker_fraction": yprof["full_steps"] / (M * args.burn),
                 "note": "the chain's first iterations, before the sparsity pattern settles (device time of the burn-in run, this rank)"}
    chain.run_discard(W)                           # warm-up steps

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    barrier()
    with clk:
        chain.run_discard(args.steps)              # timed: device time between CUDA events on the chain's stream
        barrier()
    ms, launches = chain.last_timing()
    kms = chain.kernel_ms()
    prof = chain.sweep_profile()
    geom = chain.geometry()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda:%d" % dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    chain_rate = M * args.steps / (ms_max * 1e-3)
    value = chain_rate if args.total_rows else world * chain_rate

    # ---------------- correctness carried by the line itself: one more iteration, its whole sample row (beta, components, sigma,
    # the residuals of ALL ranks) compared across the ranks on the devices over NCCL
    row = chain.run(1, emit_all=True)
    ranks_identical = None
    if dist is not None:
        tr = torch.from_numpy(row).cuda()
        hi_t, lo_t = tr.clone(), tr.clone()
        dist.all_reduce(hi_t, op=dist.ReduceOp.MAX); dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
        ranks_identical = bool(torch.equal(hi_t, lo_t))
    finite_state = bool(np.isfinite(row).all())

    # ---------------- stand-alone measurements for the extra roofline entries (rank 0 uses them)
    gram_alone_ms = None
    if rank == 0:
        try:
            order = np.random.default_rng(1).permutation(M).astype(np.int32)[:min(M, 50000)]
            _, _, gram_alone_ms = geno.gram_cross_blocks(order, block=geom["block"], impl=0)
            gram_alone_markers = len(order)
        except Exception:
            gram_alone_ms = None

    # ---------------- end to end through the C ABI with host buffers (every rank; max time over ranks)
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(args.e2e_steps, args.steps)
        codes_pinned = torch.from_numpy(geno.codes()).pin_memory()     # host packed genotypes, page-locked (outside the timed region)
        codes = codes_pinned.numpy()
        st = geno.stats()
        thin = 5
        tmp = tempfile.NamedTemporaryFile(suffix=".csv", delete=False); tmp.close()
        barrier()
        t0 = time.perf_counter()
        g2 = brr.Genotypes.from_packed(codes, N, mean=st["mean"], sd=st["sd"], device=dev)    # H2D of the packed matrix
        if comm is not None:
            g2.shard_stats(comm)
        t1 = time.perf_counter()
        c2 = make_chain(g2, e2e_steps, burn_in=1, thinning=thin)
        if rank == 0:
            c2.open_output(tmp.name)
        t2 = time.perf_counter()
        kept = c2.run_discard(e2e_steps)                               # perm H2D per step, kept rows D2H + CSV writer
        t3 = time.perf_counter()
        if rank == 0:
            c2.close_output()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        e2e_split = {"genotypes_h2d_ms": 1e3 * (t1 - t0), "chain_create_ms": 1e3 * (t2 - t1), "iterations_ms": 1e3 * (t3 - t2),
                     "writer_drain_ms": 1e3 * (time.perf_counter() - t3)}
        te = torch.tensor([dt], dtype=torch.float64, device="cuda:%d" % dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        row_bytes = 8 * c2.row_len
        e2e_chain = M * e2e_steps / dt
        e2e = {"value": e2e_chain if args.total_rows else world * e2e_chain, "unit": "SNP-updates/s", "steps": e2e_steps,
               "h2d_bytes_per_step": int(codes.nbytes / e2e_steps + 4 * M + 8 * N / e2e_steps),
               "d2h_bytes_per_step": int(row_bytes * kept / e2e_steps),
               "split_ms_rank0": e2e_split,
               "note": "brr_geno_from_packed(host codes) + brr_chain_create + %d iterations with CSV rows every %d; wall clock, max over ranks" % (e2e_steps, thin)}
        c2.close(); g2.close()
        os.unlink(tmp.name)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    shape_key = "%s:%d:%d:%d" % (args.sampler, N, M, world)            # rows per GPU : markers : ranks
    traffic = None                                  # DRAM bytes of one sweep launch from a committed ncu --set full capture of THIS shape
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(shape_key, {}).get("sweep_kernel_dram_bytes_per_launch")
    except Exception:
        pass
    nbytes_col = (N + 3) // 4
    algo_bytes = M * nbytes_col + 16 * N + 24 * M                      # per sweep launch and GPU
    sweep_ms = kms["sweep"] / args.steps
    achieved = algo_bytes / (sweep_ms * 1e-3) / 1e9
    B, LA = geom["block"], brr.lookahead(geom["block"])
    nblocks = max(prof["blocks"], 1)
    sm_hz = (clk.summary()["sm_mhz"] or 1965.0) * 1e6
    int8_peak, int8_src = measure_int8_peak(dev)
    fp64_peak = brr.peak_fp64(dev)                                     # thread-level DFMA / s, whole device
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    gram_ops = 2.0 * (B + LA) * M * N                                  # int8 multiply-adds x 2 per iteration and GPU (self + look-ahead columns)
    gram_ms = kms["gram"] / args.steps
    dots_cycles_per_block = prof["worker_dots"] / nblocks
    rows_w0 = geom["rows_per_worker_max"]
    rooflines = [
        {"kernel": "sweep_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": traffic, "algorithmic_bytes": algo_bytes, "ms_per_launch": sweep_ms},
        {"kernel": "gram_tc_kernel", "bound": "tensor", "unit": "TOP/s (int8)", "peak": int8_peak, "peak_source": int8_src,
         "achieved": gram_ops / (gram_ms * 1e-3) / 1e12 if gram_ms > 0 else None,
         "frac": gram_ops / (gram_ms * 1e-3) / 1e12 / int8_peak if gram_ms > 0 else None,
         "ms_per_launch": gram_ms, "algorithmic_ops": gram_ops,
         "note": "in situ: a persistent grid on the %d SMs the sweep leaves free, beside the previous iteration's sweep (off the critical path while "
                 "gram < sweep); the fraction is against the WHOLE device's peak" % max(1, sms - geom["workers"] - 5),
         "stand_alone": None if not gram_alone_ms else {
             "ms": gram_alone_ms, "markers": gram_alone_markers,
             "achieved": 2.0 * (B + LA) * gram_alone_markers * N / (gram_alone_ms * 1e-3) / 1e12,
             "frac": 2.0 * (B + LA) * gram_alone_markers * N / (gram_alone_ms * 1e-3) / 1e12 / int8_peak,
             "note": "the same kernel alone on all SMs (brr_gram_cross_blocks, CUDA events)"}},
        {"kernel": "sweep_kernel / workers' dot stage (code_b^T eps, exact int8 contraction on tcgen05)", "bound": "shared memory",
         "unit": "B/clk per SM", "peak": 128.0, "peak_source": "B200 shared-memory bandwidth per SM (128 B/clk)",
         "achieved": 2.25 * B * rows_w0 / dots_cycles_per_block if dots_cycles_per_block > 0 else None,
         "frac": 2.25 * B * rows_w0 / dots_cycles_per_block / 128.0 if dots_cycles_per_block > 0 else None,
         "fp64_pipe_equivalent": {
             "unit": "G genotype-MACs/s per SM", "achieved": B * rows_w0 / (dots_cycles_per_block / sm_hz) / 1e9 if dots_cycles_per_block > 0 else None,
             "fp64_peak": fp64_peak / sms / 1e9, "frac": B * rows_w0 / (dots_cycles_per_block / sm_hz) / (fp64_peak / sms) if dots_cycles_per_block > 0 else None,
             "note": "the same stage as fp64 FMAs per genotype (round 1's formulation) against the measured DFMA rate of one SM (brr_peak_fp64)"},
         "note": "first worker CTA: %d markers x %d rows per block in %.0f SM cycles (residual digits, 2-bit -> int8 unpack, MMAs, TMEM read-out, "
                 "partial sends), from the kernel's own cycle counters; bytes through shared memory per block = 2.25 x markers x rows (operand tile "
                 "written once and read once by the tensor core, packed columns read once)" % (B, rows_w0, dots_cycles_per_block)}]
    if args.total_rows:
        wl = "%s N=%d x M=%d synthetic 2-bit genotypes (%s), %d rows per GPU" % (
            {"v2": "BayesRSamplerV2", "groups": "BayesRSamplerV2Groups (22 groups)", "horseshoe": "HorseshoeR"}[args.sampler], N_total, M,
            "BASELINE configs[2]" if (args.sampler, N_total, M) == ("groups", 100000, 200000) else
            "BASELINE configs[4]" if (args.sampler, N_total, M) == ("horseshoe", 100000, 100000) else "fixed total size", N)
    elif args.sampler != "v2" or N != CFG["N"]:
        wl = "%s N=%d x M=%d synthetic 2-bit genotypes (non-default sampler / shape)" % (args.sampler, N_total, M)
    elif M == CFG["M"]:
        wl = WORKLOAD
    else:
        wl = WORKLOAD.replace("M=50000", "M=%d" % M).replace("BASELINE configs[1]", "BASELINE configs[3] shape" if M == 500000 else "non-default M")
    out = {"metric": "SNP-updates/sec", "value": value, "unit": "SNP-updates/s", "n_gpus": world, "steps": args.steps, "warmup": W,
           "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if args.total_rows else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "chain_snp_updates_per_s": chain_rate,
           "genotype_cells_per_s": float(N_total) * chain_rate,
           "gibbs_iterations_per_s": 1e3 * args.steps / ms_max,
           "ranks_bit_identical": ranks_identical, "state_finite": finite_state,
           "config": {"workload": wl, "block": geom["block"], "workers": geom["workers"],
                      "rows_per_worker_max": geom["rows_per_worker_max"], "chain_burn_in_iterations": args.burn,
                      "l2": "inputs larger than L2: %d MB of packed genotypes per GPU are re-read every step" % (M * nbytes_col // 1000000),
                      "parallelism": "1 GPU" if world == 1 else
                      "ONE chain over N_total=%d individuals, row-sharded over %d GPUs (%d rows on rank 0), chain replicated, per-block "
                      "exchange of partial dots over NVLink peer memory inside the sweep kernel" % (N_total, world, N),
                      "n_total": N_total,
                      "value_definition": ("SNP-updates/s of the chain (fixed total size: strong scaling)" if args.total_rows else
                                           "SNP-updates/s in units of one marker update over one GPU's %d-row shard: world x M x steps / time (weak scaling in "
                                           "individuals; the chain itself advances chain_snp_updates_per_s markers per second)" % N)},
           "gpu_launches": int(launches),
           "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items()},
           "cycles_per_block": {k: prof[k] / nblocks for k in ("gather", "gather_first_chunk", "gather_last_chunk", "serial_pass", "publish", "eval_cycles", "resolve_cycles", "prologue_cycles", "bookkeeping", "chunks_received", "worker_wait", "worker_dots", "worker_reduce")},
           "markers_per_speculative_window": (M * args.steps) / max(prof["windows"], 1),
           "state_changing_marker_fraction": prof["full_steps"] / (M * args.steps),
           "fp64_draw_fraction": prof["fp64_draws"] / max(prof["full_steps"], 1),
           "young_chain": young,
           "roofline": dict(rooflines[0], peak_source=peak_src,
                            note="serial Gibbs chain: the kernel is latency-bound (M dependent marker steps), see DESIGN.md 3.2"),
           "rooflines": rooflines,
           "clocks": clk.summary()}
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu and world == 1:
        rate, secs, M_s, sample = run_cpu_arm(10, 1, args.sampler, N_total)
        out["cpu_baseline"] = {"value": rate, "unit": "SNP-updates/s", "cores": 1, "kind": "port", "sample": sample}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
