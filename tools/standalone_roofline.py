"""HBM-bound / tensor-bound stand-alone kernels against the measured peaks: X^T eps over the packed matrix, block Gram."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bayesrrcpp_b200 as brr
N = M = 50000
g = brr.Genotypes.synthetic(N, M, 7)
eps = np.random.default_rng(0).normal(size=N)
best = min(g.xt_eps(eps)[1] for _ in range(5))
bytes_x = M * ((N + 3) // 4) + 8 * N + 8 * M
peak = 6547.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
order = np.random.default_rng(1).permutation(M).astype(np.int32)
gram = min(g.gram_blocks(order, block=128)[1] for _ in range(3))
gramx = min(g.gram_cross_blocks(order, block=128)[2] for _ in range(3))
out = {"xt_eps": {"ms": best, "algorithmic_bytes": bytes_x, "GBps": bytes_x / best / 1e6, "frac_of_measured_hbm_peak": bytes_x / best / 1e6 / peak, "peak_GBps": peak},
       "gram_128": {"ms": gram, "int8_TOPs": 2.0 * 128 * M * N / gram / 1e9, "GBps": M * ((N + 3) // 4) / gram / 1e6},
       "gram_128_with_lookahead_cross": {"ms": gramx, "int8_TOPs": 2.0 * (128 + brr.lookahead(128)) * M * N / gramx / 1e9}}
print(json.dumps(out))
