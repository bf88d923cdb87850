import sys, time, tempfile, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bayesrrcpp_b200 as brr
if os.environ.get("PROBE_TORCH"):
    import torch; torch.cuda.set_device(0); torch.zeros(1, device="cuda")
N = M = 50000
g = brr.Genotypes.synthetic(N, M, 7)
y = np.random.default_rng(0).normal(size=N)
codes = g.codes(); st = g.stats()
if os.environ.get("PROBE_PINNED"):
    import torch; keep = torch.from_numpy(codes).pin_memory(); codes = keep.numpy()
hyp = dict(sigma0=0.01, v0E=1e-4, s02E=1e-3, v0G=1e-4, s02G=1e-3)
for rep in range(2):
    t = [time.perf_counter()]
    g2 = brr.Genotypes.from_packed(codes, N, mean=st["mean"], sd=st["sd"]); t.append(time.perf_counter())
    c = brr.Chain(g2, brr.V2, 20, burn_in=1, thinning=5, seed=3, Y=y, cva=[1e-4, 1e-3, 1e-2], **hyp); t.append(time.perf_counter())
    tmp = tempfile.NamedTemporaryFile(suffix=".csv", delete=False); tmp.close()
    c.open_output(tmp.name); t.append(time.perf_counter())
    c.run_discard(1); t.append(time.perf_counter())
    c.run_discard(19); t.append(time.perf_counter())
    c.close_output(); t.append(time.perf_counter())
    names = ["from_packed", "Chain()", "open_output", "first run(1)", "run(19)", "close_output"]
    print("rep", rep, {n: round(1e3 * (b - a), 1) for n, a, b in zip(names, t[:-1], t[1:])})
    c.close(); g2.close(); os.unlink(tmp.name)
