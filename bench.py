#!/usr/bin/env python
"""bench.py -- SNP-updates/s of the per-SNP Gibbs sweep (BASELINE.json metric) on N B200s of one node.

A "step" is one Gibbs iteration: one pass of the hot path (block Gram -> persistent sweep -> hyper draws) over all M
markers.  N=1 workload = BASELINE.json configs[1]: BayesRSamplerV2, N=50,000 x M=50,000 synthetic genotypes, simulated
phenotype h2=0.5, K=4.  A multi-GPU run is ONE chain over N_total = 50,000 x world individuals, row-sharded: rank r holds
rows [50,000 r, 50,000 (r+1)) of the same virtual matrix (weak scaling: fixed rows per GPU), every rank replicates the chain,
the per-block partial dots are exchanged over NVLink peer memory inside the sweep kernel.

  value     : whole-job SNP-updates/s, genotypes resident in HBM, device-timed (CUDA events on the chain's stream)
  e2e       : same metric through the C ABI with HOST buffers (page-locked): packed genotypes H2D, chain creation, per-iteration
              permutation upload, sample rows D2H + CSV writer (thinning 5) all inside the timed region (wall clock)
  roofline  : the persistent sweep kernel against the measured HBM copy bandwidth (it is bound by the serial chain,
              not by HBM -- DESIGN.md section 4)
  cpu_baseline / --impl reference : the CPU oracle (C restatement of the reference's Eigen sampler, pinned against the
              reference's own sources built over a minimal Eigen/Rcpp shim -- that shim is not a fair timing of Eigen, so the
              port is what is timed) on a bounded column sample of the same workload, 1 core like the reference.
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
WORKLOAD = "BayesRSamplerV2 N=50000 x M=50000 K=4 synthetic 2-bit genotypes, simulated phenotype h2=0.5 (BASELINE configs[1])"
CPU_SAMPLE_M = 4000          # columns of the dense fp64 sample the CPU arm sweeps (full N)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def run_cpu_arm(steps, warmup):
    """the reference's CPU algorithm (oracle port), single thread like the reference's sampler thread"""
    from oracle import pyoracle as po
    po.build()
    X, y = cpu_sample_data(CFG["N"], CPU_SAMPLE_M, CFG["data_seed"])
    kw = dict(CFG["hyp"])
    po.run_v2(X, y, CFG["cva"], max(1, warmup), seed=1, want_rows=False, **kw)          # warm-up (page-in, caches)
    r = po.run_v2(X, y, CFG["cva"], steps, seed=CFG["chain_seed"], want_rows=False, **kw)
    rate = CPU_SAMPLE_M * steps / r["seconds"]
    # second flavour: -O3 -march=native (the first is -O2, R's default for packages); the faster one is the baseline
    flavour, secs = "-O2", r["seconds"]
    try:
        # always rebuilt on the machine that runs it: a -march=native library from another host could use missing instructions
        subprocess.run(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), "native"], check=True, capture_output=True)
        po.NATIVE = True
        po.run_v2(X, y, CFG["cva"], 1, seed=1, want_rows=False, **kw)
        rn = po.run_v2(X, y, CFG["cva"], steps, seed=CFG["chain_seed"], want_rows=False, **kw)
        rate_n = CPU_SAMPLE_M * steps / rn["seconds"]
        if rate_n > rate:
            rate, secs, flavour = rate_n, rn["seconds"], "-O3 -march=native"
    except Exception:
        pass
    finally:
        po.NATIVE = False
    sample = ("full N=%d rows x %d-column dense fp64 sample, %d iterations, oracle built %s (faster of -O2 and -O3 -march=native; "
              "per-marker cost is independent of M: extrapolates)") % (CFG["N"], CPU_SAMPLE_M, steps, flavour)
    return rate, secs, sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--workers", type=int, default=0, help="worker CTAs of the sweep kernel (default: the library's split of the SMs)")
    ap.add_argument("--markers", type=int, default=CFG["M"], help="markers M (default: BASELINE configs[1]; 500000 with --gpus 8 = configs[3])")
    ap.add_argument("--rows", type=int, default=CFG["N"], help="individuals per GPU (default: BASELINE configs[1])")
    ap.add_argument("--sampler", default="v2", choices=["v2", "groups", "horseshoe"],
                    help="non-default samplers are for the other BASELINE shapes (configs[2]: groups 100000 x 200000, configs[4]: horseshoe 100000 x 100000)")
    ap.add_argument("--burn", type=int, default=CFG["chain_burn"], help="untimed chain burn-in iterations before the warm-up")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 20))
        rate, secs, sample = run_cpu_arm(steps, 1)
        print(json.dumps({"impl": "reference", "metric": "SNP-updates/sec", "value": rate, "unit": "SNP-updates/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": 1e3 * secs / steps * CFG["M"] / CPU_SAMPLE_M,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": WORKLOAD, "sample": sample},
                          "cpu_baseline": {"value": rate, "unit": "SNP-updates/s", "cores": 1, "kind": "port", "sample": sample},
                          "e2e": {"value": rate, "unit": "SNP-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import bayesrrcpp_b200 as brr
    dist, comm, host_allreduce = None, None, None
    if world > 1:
        import torch.distributed as dist
        from bayesrrcpp_b200 import sharded
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = sharded.torch_comm()                    # gloo group for the library's host call-backs (set-up time only)

        def host_allreduce(a):
            t = torch.from_numpy(a).cuda()             # NCCL over NVLink
            dist.all_reduce(t)
            a[:] = t.cpu().numpy()
    dev = local
    N, M = args.rows, args.markers                     # N: rows per GPU
    geno = brr.Genotypes.synthetic(N, M, CFG["data_seed"], row0=rank * N, device=dev)
    if comm is not None:
        geno.shard_stats(comm)
    y = simulate_phenotype(geno, CFG["data_seed"], CFG["h2"], CFG["causal_frac"], rank, host_allreduce)
    total_iters = args.burn + W + args.steps
    if args.sampler == "groups":       # 22 chromosome-like groups, identical ladders, the vignette's N x 1 zero fixed matrix
        G = 22
        chain = brr.Chain(geno, brr.GROUPS, total_iters, seed=CFG["chain_seed"], Y=y, cva=np.tile(np.array(CFG["cva"]), (G, 1)), groups=G,
                          gAssign=(np.arange(M) * G // M).astype(np.int32), fixed=np.zeros((N, 1)), block=args.block, workers=args.workers, comm=comm, **CFG["hyp"])
    elif args.sampler == "horseshoe":
        p0 = 0.1 * M
        chain = brr.Chain(geno, brr.HORSESHOE, total_iters, seed=CFG["chain_seed"], Y=y, A=(1 / np.sqrt(N * world)) * p0 / (M - p0),
                          v0E=1e-3, s02E=1e-3, vL=1.0, vT=1.0, c2=1.0, vC=10.0, sC=10.0, block=args.block, workers=args.workers, comm=comm)
    else:
        chain = brr.Chain(geno, brr.V2, total_iters, seed=CFG["chain_seed"], Y=y, cva=CFG["cva"], block=args.block, workers=args.workers, comm=comm, **CFG["hyp"])
    clk = ClockSampler(dev).start()        # nvidia-smi is up and sampling before the timed region begins
    chain.run_discard(args.burn)           # untimed chain burn-in: the timed steps see a settled sparsity pattern
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
    value = world * M * args.steps / (ms_max * 1e-3)

    # ---------------- end to end through the C ABI with host buffers (every rank; max time over ranks)
    e2e = None
    if not args.no_e2e and args.sampler == "v2":
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
        c2 = brr.Chain(g2, brr.V2, args.steps, burn_in=1, thinning=thin, seed=CFG["chain_seed"], Y=y, cva=CFG["cva"],
                       block=args.block, workers=args.workers, comm=comm, **CFG["hyp"])
        if rank == 0:
            c2.open_output(tmp.name)
        t2 = time.perf_counter()
        kept = c2.run_discard(args.steps)                              # perm H2D per step, kept rows D2H + CSV writer
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
        row_bytes = 8 * (2 * M + 4 + N * world)
        e2e = {"value": world * M * args.steps / dt, "unit": "SNP-updates/s",
               "h2d_bytes_per_step": int(codes.nbytes / args.steps + 4 * M + 8 * N / args.steps),
               "d2h_bytes_per_step": int(row_bytes * kept / args.steps),
               "split_ms_rank0": e2e_split,
               "note": "brr_geno_from_packed(host codes) + brr_chain_create + %d iterations with CSV rows every %d; wall clock" % (args.steps, thin)}
        c2.close(); g2.close()
        os.unlink(tmp.name)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    traffic = None                                  # DRAM bytes of one sweep launch from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_full_v8.json")) as f:
            k = [v for n, v in json.load(f).items() if "sweep_kernel" in n][0]
            traffic = int(1e6 * (float(k["dram__bytes_read.sum"]["value"]) + float(k["dram__bytes_write.sum"]["value"])))
    except Exception:
        pass
    algo_bytes = M * ((N + 3) // 4) + 16 * N + 24 * M                  # per sweep launch and GPU
    sweep_ms = kms["sweep"] / args.steps
    achieved = algo_bytes / (sweep_ms * 1e-3) / 1e9
    out = {"metric": "SNP-updates/sec", "value": value, "unit": "SNP-updates/s", "n_gpus": world, "steps": args.steps, "warmup": W,
           "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": ("%s N=%d x M=%d synthetic 2-bit genotypes (non-default sampler / shape)" % (args.sampler, N * world, M)) if (args.sampler != "v2" or N != CFG["N"]) else WORKLOAD if M == CFG["M"] else WORKLOAD.replace("M=50000", "M=%d" % M).replace("BASELINE configs[1]", "BASELINE configs[3] shape" if M == 500000 else "non-default M"), "block": geom["block"], "workers": geom["workers"],
                      "rows_per_worker_max": geom["rows_per_worker_max"], "chain_burn_in_iterations": args.burn,
                      "l2": "inputs larger than L2: %d MB of packed genotypes per GPU are re-read every step" % (M * ((N + 3) // 4) // 1000000),
                      "parallelism": "1 GPU" if world == 1 else
                      "ONE chain over N_total=%d individuals, row-sharded over %d GPUs (%d rows each), chain replicated, per-block "
                      "exchange of partial dots over NVLink peer memory inside the sweep kernel" % (N * world, world, N),
                      "n_total": N * world,
                      "value_definition": "SNP-updates/s in units of one marker update over one GPU's 50,000-row shard: world x M x steps / time "
                                          "(weak scaling in individuals; the chain itself advances chain_snp_updates_per_s markers per second)",
                      "chain_snp_updates_per_s": M * args.steps / (ms_max * 1e-3),
                      "genotype_cells_per_s": float(N) * world * M * args.steps / (ms_max * 1e-3),
                      "gibbs_iterations_per_s": 1e3 * args.steps / ms_max},
           "gpu_launches": int(launches),
           "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items()},
           "cycles_per_block": {k: prof[k] / max(prof["blocks"], 1) for k in ("gather", "gather_first_chunk", "gather_last_chunk", "serial_pass", "publish", "eval_cycles", "resolve_cycles", "prologue_cycles", "bookkeeping", "chunks_received", "worker_wait", "worker_dots", "worker_reduce")},
           "markers_per_speculative_window": (M * args.steps) / max(prof["windows"], 1),
           "state_changing_marker_fraction": prof["full_steps"] / (M * args.steps),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "algorithmic_bytes": algo_bytes, "kernel": "sweep_kernel", "peak_source": peak_src,
                        "note": "serial Gibbs chain: the kernel is latency-bound (M dependent marker steps), see DESIGN.md 3.2"},
           "clocks": clk.summary()}
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu and world == 1:
        rate, secs, sample = run_cpu_arm(10, 1)
        out["cpu_baseline"] = {"value": rate, "unit": "SNP-updates/s", "cores": 1, "kind": "port", "sample": sample}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
