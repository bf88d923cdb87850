import sys; sys.path.insert(0,'/root/repo')
import numpy as np, bayesrrcpp_b200 as brr
g = brr.Genotypes.synthetic(50000, 50000, 7)
order = np.random.default_rng(1).permutation(50000).astype(np.int32)
for blk in (128, 64):
    _, ms0 = g.gram_blocks(order, block=blk)
    _, _, ms1 = g.gram_cross_blocks(order, block=blk)
    print("block", blk, "self only %.3f ms, with cross %.3f ms" % (ms0, ms1))
