"""SASS evidence of libbayesrr_b200.so without a GPU: cuobjdump -sass, per kernel the counts of the tensor-core / TMEM / TMA / mbarrier
instructions (python tools/sass_summary.py > profiles/<name>.txt)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bayesrrcpp_b200", "libbayesrr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCIMMA", "LDTM", "UTCBAR", "UBLKCP", "UTCATOMSWS", "SYNCS", "MEMBAR", "MUFU.EX2", "DFMA", "ELECT"]
print("# SASS evidence of libbayesrr_b200.so (cuobjdump -sass, sm_100a): tensor-core, TMEM and TMA instructions per kernel")
print("# UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops,")
print("# UTCATOMSWS = tcgen05.alloc/dealloc, MEMBAR = fences\n")
name, counts, n = None, collections.Counter(), 0


def flush():
    if name is not None:
        print(name)
        print("    instructions %d  " % n + "  ".join("%s %d" % (k, counts[k]) for k in KEYS if counts[k]))


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, counts, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        n += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[k] += 1
flush()
