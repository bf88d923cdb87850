"""Generate the golden vectors of tests/golden/ from the reference's OWN source files.

Runs only where /root/reference exists: `make -C oracle ref` compiles src/BayesRv2.cpp, src/BayesRv2Groups.cpp,
src/BRv2Grstart.cpp, src/HorseshoeR.cpp and src/distributions.cpp unmodified against the minimal Eigen/Rcpp/R shim of
oracle/shim/ (those libraries are not installed here), with R's nmath draws re-routed to a sequential Philox stream and
std::random_shuffle's rand() seeded with the chain seed.  Each fixture stores the inputs, the sample rows the reference
packed (captured at its `sample << ...` statement) and the CSV text its ofstream wrote.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

HYP = dict(sigma0=0.01, v0E=1e-4, s02E=1e-3, v0G=1e-4, s02G=1e-3)
CVA = [1e-4, 1e-3, 1e-2]


def main():
    assert po.ref_available(), "needs /root/reference (oracle/_ref)"
    tmp = tempfile.mkdtemp()

    def csv(name):
        return os.path.join(tmp, name)

    # --- BayesRSamplerV2
    d = po.synth(320, 90, seed=101)
    rows, n = po.ref_v2(csv("v2.csv"), 11, 40, 10, 5, d["X"], d["y"], CVA, **HYP)
    np.savez_compressed(os.path.join(HERE, "v2.npz"), G=d["G"], mean=d["mean"], sd=d["sd"], y=d["y"], cva=CVA, seed=11,
                        max_iterations=40, burn_in=10, thinning=5, rows=rows[:n], csv=open(csv("v2.csv")).read())
    # --- BayesRSamplerV2Groups (3 groups, 2 fixed effects)
    rng = np.random.default_rng(7)
    d = po.synth(280, 75, seed=102)
    gA = np.sort(rng.integers(0, 3, size=75)).astype(np.int32)
    cva = np.tile(np.array(CVA), (3, 1)) * np.array([1.0, 2.0, 0.5])[:, None]
    fixed = rng.normal(size=(280, 2)); fixed = (fixed - fixed.mean(0)) / fixed.std(0, ddof=1)
    y = d["y"] + fixed @ np.array([0.4, -0.2])
    rows, n = po.ref_groups(csv("g.csv"), 12, 30, 10, 4, d["X"], y, cva, 3, gA, fixed, **HYP)
    np.savez_compressed(os.path.join(HERE, "groups.npz"), G=d["G"], mean=d["mean"], sd=d["sd"], y=y, cva=cva, gAssign=gA,
                        fixed=fixed, seed=12, max_iterations=30, burn_in=10, thinning=4, rows=rows[:n],
                        csv=open(csv("g.csv")).read())
    # --- BRV2Grstart from the last kept row above
    M, N, G, F = 75, 280, 3, 2
    last = rows[n - 1]
    st = dict(mu=last[1], beta=last[2:2 + M], sigmaE=last[2 + M], components=last[3 + M:3 + 2 * M],
              sigmaGG=last[3 + 2 * M:3 + 2 * M + G], epsilon=last[3 + 2 * M + G:3 + 2 * M + G + N])
    rows2, n2 = po.ref_grstart(csv("r.csv"), 13, 24, 4, 4, st["mu"], st["beta"], st["sigmaE"], st["sigmaGG"], d["X"], st["epsilon"],
                               st["components"], cva, 3, gA, **HYP)
    np.savez_compressed(os.path.join(HERE, "grstart.npz"), G=d["G"], mean=d["mean"], sd=d["sd"], cva=cva, gAssign=gA, seed=13,
                        max_iterations=24, burn_in=4, thinning=4, rows=rows2[:n2], csv=open(csv("r.csv")).read(), **st)
    # --- HorseshoeR (the reference writes only the header without OpenMP, SURVEY.md Q10: rows come from the hook)
    d = po.synth(260, 70, seed=103)
    A = (1 / np.sqrt(260)) * 7 / (70 - 7)
    rows, n = po.ref_horseshoe(csv("h.csv"), 14, 30, 10, 5, d["X"], d["y"], A)
    np.savez_compressed(os.path.join(HERE, "horseshoe.npz"), G=d["G"], mean=d["mean"], sd=d["sd"], y=d["y"], A=A, seed=14,
                        max_iterations=30, burn_in=10, thinning=5, rows=rows[:n], csv=open(csv("h.csv")).read())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
