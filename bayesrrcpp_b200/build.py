"""In-tree build of libbayesrr_b200.so (nvcc, sm_100a only).  No torch, no JIT cache: the .so lies next to this file
so that it travels with the repository snapshot."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# BRR_LIB_SUFFIX=_la64 (with e.g. BRR_LOOKAHEAD128=64): a variant build beside the default one (own object directory, own .so;
# BRR_LIB=<path> makes the package load it)
SUFFIX = os.environ.get("BRR_LIB_SUFFIX", "")
OBJ = os.path.join(HERE, "_build" + SUFFIX)
LIB = os.path.join(HERE, "libbayesrr_b200%s.so" % SUFFIX)
SOURCES = ["geno.cu", "gram.cu", "sweep.cu", "hyper.cu", "shard.cu", "chain.cu", "writer.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = (["-DBRR_ROUND_PROFILE=1"] if os.environ.get("BRR_ROUND_PROFILE") else []) + \
        (["-DBRR_TENSOR_DOTS=%s" % os.environ["BRR_TENSOR_DOTS"]] if os.environ.get("BRR_TENSOR_DOTS") else []) + \
        (["-DBRR_LOOKAHEAD128=%s" % os.environ["BRR_LOOKAHEAD128"]] if os.environ.get("BRR_LOOKAHEAD128") else []) + \
        (["-DBRR_PHASE_PROFILE=1"] if os.environ.get("BRR_PHASE_PROFILE") else []) + \
        (["-DBRR_COLUMN_STAGES=%s" % os.environ["BRR_COLUMN_STAGES"]] if os.environ.get("BRR_COLUMN_STAGES") else []) + \
        (["-DBRR_MBAR_HANDOVER=%s" % os.environ["BRR_MBAR_HANDOVER"]] if os.environ.get("BRR_MBAR_HANDOVER") else []) + (["-DBRR_DOT_PROFILE=1"] if os.environ.get("BRR_DOT_PROFILE") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas", "-Xptxas", "-v"]


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdr.append(os.path.join(HERE, "..", "include", "bayesrr_b200.h"))
    return max(os.path.getmtime(h) for h in hdr)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "flags.txt")             # a changed flag set (BRR_TENSOR_DOTS, BRR_ROUND_PROFILE) rebuilds everything
    flags = " ".join(FLAGS)
    if not os.path.exists(stamp) or open(stamp).read() != flags:
        force = True
    hdr_time = _deps()
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC] + FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        log = os.path.join(OBJ, os.path.basename(src) + ".ptxas.log")
        with open(log, "w") as f:
            f.write(r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=6) as ex:
        outs = list(ex.map(compile_one, jobs))
    if verbose:
        for o in outs:
            sys.stderr.write(o)
    with open(stamp, "w") as f:
        f.write(flags)
    objs = [os.path.join(OBJ, s + ".o") for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
