"""stand-alone block-Gram kernel at the BASELINE configs[1] shape (for ncu captures and CUDA-event timing)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bayesrrcpp_b200 as brr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 12800
B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
g = brr.Genotypes.synthetic(N, M, seed=5)
order = np.random.default_rng(1).permutation(M).astype(np.int32)
best = 1e9
for _ in range(3):
    _, _, ms = g.gram_cross_blocks(order, block=B, impl=0)
    best = min(best, ms)
ops = 2.0 * (B + brr.lookahead(B)) * M * N
print("gram stand-alone N=%d M=%d B=%d: %.3f ms, %.1f int8 TOP/s" % (N, M, B, best, ops / (best * 1e-3) / 1e12))
