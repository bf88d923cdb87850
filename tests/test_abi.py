"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bayesrr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(brr_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_declares_the_four_reference_entry_points():
    syms = declared_symbols()
    for name in ("brr_BayesRSamplerV2", "brr_BayesRSamplerV2Groups", "brr_BRV2Grstart", "brr_HorseshoeR"):
        assert name in syms


def test_library_exports_every_declared_symbol(brr):
    lib = ctypes.CDLL(brr.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.brr_abi_version() == 1


def test_no_cpu_fallback_without_device(brr):
    """Without a GPU every computing call must fail loudly with BRR_E_CUDA (never a silent CPU path)."""
    import numpy as np
    import pytest
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(brr.BayesRRError) as e:
        brr.Genotypes.from_dense(np.zeros((8, 2)))
    assert e.value.code == brr.E_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_references_the_oracle():
    """The shipped package must not import, include or link anything under oracle/."""
    pkg = os.path.join(ROOT, "bayesrrcpp_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in text and "oracle/" not in text, os.path.join(dirpath, f)
                assert "liboracle" not in text and "orc_" not in text, os.path.join(dirpath, f)


def test_entry_point_iteration_validation_matches_reference(brr, tmp_path):
    """src/BayesRv2.cpp:69-80: V2 truncates the file and writes the header before rejecting the iteration arguments;
    HorseshoeR (src/HorseshoeR.cpp:119-123) returns before touching the file.  Validation precedes any device work."""
    import numpy as np
    X = np.zeros((8, 3)); Y = np.zeros(8)
    out = tmp_path / "v2.csv"
    try:
        brr.BayesRSamplerV2(str(out), 1, 5, 10, 1, X, Y, 0.01, 1e-4, 1e-3, 1e-4, 1e-3, [1e-4, 1e-3, 1e-2])
        raise AssertionError("expected an error")
    except brr.BayesRRError as e:
        assert e.code == brr.E_ITER
    text = out.read_text()
    assert text.startswith("iteration,mu,beta[1],beta[2],beta[3],sigmaE,sigmaG,comp[1],") and text.endswith("epsilon[8]\n")
    out2 = tmp_path / "hs.csv"
    try:
        brr.HorseshoeR(str(out2), 1, 5, 0, 1, X, Y, 0.1, 1e-3, 1e-3, 1, 1, 1, 10, 10)
        raise AssertionError("expected an error")
    except brr.BayesRRError as e:
        assert e.code == brr.E_ITER
    assert not out2.exists()


def test_row_text_is_printf_g(brr):
    """the writer's number formatting (fast paths + std::to_chars) against C's "%g" (what Eigen's IOFormat / ofstream precision 6 emit)"""
    import numpy as np
    lib = ctypes.CDLL(brr.LIB_PATH)
    lib.brr_format_row.restype = ctypes.c_int64
    rng = np.random.default_rng(5)
    vals = np.concatenate([
        [0.0, -0.0, 1.0, 2.0, 3.0, 17.0, 99999.0, 100000.0, 999999.0, 1e6, 1234567.0, -5.0, 0.5, 1e-5, 1.5e-5, 123456.5, 0.1, 1 / 3,
         2.5e-310, 1e300, -1e-300, np.inf, -np.inf, np.nan, 9.999995, 99999.95, 0.000123456789, 4503599627370496.0],
        rng.normal(size=2000), rng.normal(size=500) * 1e-7, rng.normal(size=500) * 1e9, np.exp(rng.uniform(-700, 700, size=1000)),
        np.round(rng.uniform(0, 2e6, size=300))])
    buf = ctypes.create_string_buffer(64 * len(vals))
    a = np.ascontiguousarray(vals, dtype=np.float64)
    n = lib.brr_format_row(a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_int64(len(a)), buf, ctypes.c_int64(len(buf)))
    got = buf.value.decode().split(", ")
    assert n == len(buf.value) and len(got) == len(vals)
    want = ["%g" % v for v in vals]
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, bad[:10]


def test_writer_long_rows_keep_text_and_order(brr, tmp_path):
    """rows of more than 65,536 values are formatted by several threads, span by span: the file must hold exactly the text the
    single-threaded formatter gives, every row, in order (the reference's consumer writes rows in queue order, src/BayesRv2.cpp:282-289);
    short rows and the binary sink go through the same queue"""
    import numpy as np
    lib = ctypes.CDLL(brr.LIB_PATH)
    lib.brr_format_row.restype = ctypes.c_int64
    rng = np.random.default_rng(11)
    dp = ctypes.POINTER(ctypes.c_double)
    for length, nrows in [(200_003, 5), (70_000, 3), (37, 40)]:
        row = np.where(rng.uniform(size=length) < 0.6, 0.0, rng.normal(size=length) * np.exp(rng.uniform(-20, 20, size=length)))
        row[1:5] = [1.0, 2.0, 123456.0, -0.0]
        path = str(tmp_path / ("rows_%d.csv" % length))
        rc = lib.brr_writer_selftest(path.encode(), row.ctypes.data_as(dp), ctypes.c_int64(length), ctypes.c_int64(nrows), ctypes.c_int(0))
        assert rc == 0
        lines = open(path).read().split("\n")
        assert lines[-1] == "" and len(lines) == nrows + 1
        buf = ctypes.create_string_buffer(32 * length)
        for i in range(nrows):
            r = row.copy(); r[0] = float(i)
            n = lib.brr_format_row(r.ctypes.data_as(dp), ctypes.c_int64(length), buf, ctypes.c_int64(len(buf)))
            assert n == len(buf.value) and lines[i] == buf.value.decode(), "row %d of length %d differs" % (i, length)
    row = rng.normal(size=70_001)
    path = str(tmp_path / "rows.bin")
    assert lib.brr_writer_selftest(path.encode(), row.ctypes.data_as(dp), ctypes.c_int64(len(row)), ctypes.c_int64(4), ctypes.c_int(1)) == 0
    got = np.fromfile(path, dtype=np.float64).reshape(4, -1)
    assert np.array_equal(got[:, 1:], np.tile(row[1:], (4, 1))) and np.array_equal(got[:, 0], np.arange(4.0))


def test_writer_is_clean_under_thread_sanitizer():
    """SURVEY.md section 5 (race detection): the queue-backed writer, compiled as plain C++ with -fsanitize=thread, takes 1,000-row
    enqueue storms through its capacity-8 ring and long rows through the multi-threaded formatter without a report (tools/tsan_writer.sh)"""
    import shutil, subprocess
    import pytest
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["bash", os.path.join(root, "tools", "tsan_writer.sh")], capture_output=True, text=True, timeout=600)
    if r.returncode != 0 and "ThreadSanitizer" not in r.stderr + r.stdout and ("tsan" in r.stderr or "sanitize" in r.stderr):
        pytest.skip("this g++ cannot link -fsanitize=thread: " + r.stderr[-200:])
    assert r.returncode == 0 and "WARNING: ThreadSanitizer" not in r.stderr + r.stdout, (r.stdout + r.stderr)[-2000:]


def test_bed_reader_reports_io_errors_before_touching_a_device(brr, tmp_path):
    """missing / truncated / non-.bed files are BRR_E_IO (the file is checked first; no device is needed to find that out)"""
    import pytest
    with pytest.raises(brr.BayesRRError) as e:
        brr.Genotypes.from_bed(str(tmp_path / "nope.bed"), N=10, M=10)
    assert e.value.code == brr.E_IO
    short = tmp_path / "short.bed"
    short.write_bytes(bytes([0x6c, 0x1b, 0x01, 0, 0]))
    with pytest.raises(brr.BayesRRError) as e:
        brr.Genotypes.from_bed(str(short), N=100, M=100)
    assert e.value.code == brr.E_IO and "shorter" in str(e.value)
    notbed = tmp_path / "x.bed"
    notbed.write_bytes(bytes(3 + 25 * 100))
    with pytest.raises(brr.BayesRRError) as e:
        brr.Genotypes.from_bed(str(notbed), N=100, M=100)
    assert e.value.code == brr.E_IO and "magic" in str(e.value)


def test_rcpp_shim_compiles_with_the_reference_signatures_and_links(brr, tmp_path):
    """shim/rcpp_shim.cpp (what a maintainer drops into the R package's src/) is compiled against the minimal Rcpp / Eigen
    stand-ins the oracle's reference build uses (R, Rcpp and Eigen are absent here), its four functions are type-checked against
    the reference's exact C++ signatures (src/RcppExports.cpp:10,39,61,86 declare them the same way) and the object is linked
    against libbayesrr_b200.so, so every brr_* call it makes resolves with matching argument types."""
    import subprocess
    check = tmp_path / "check.cpp"
    check.write_text(r'''
#include <RcppEigen.h>
#include <iostream>
#include <string>
namespace Eigen { void (*shim_row_hook)(const double *, long) = nullptr; }
namespace Rcpp { std::ostream &Rcout = std::cout; std::ostream &Rcerr = std::cerr; }
// declarations exactly as the reference's generated glue has them (src/RcppExports.cpp:10,39,61,86)
void BRV2Grstart(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning, double mu, Eigen::MatrixXd beta, double sigmaE, Eigen::VectorXd sigmaGG, Eigen::MatrixXd X, Eigen::VectorXd epsilon, Eigen::VectorXd components, double sigma0, double v0E, double s02E, double v0G, double s02G, Eigen::MatrixXd cva, int groups, Eigen::VectorXi gAssign);
void BayesRSamplerV2(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning, Eigen::MatrixXd X, Eigen::VectorXd Y, double sigma0, double v0E, double s02E, double v0G, double s02G, Eigen::VectorXd cva);
void BayesRSamplerV2Groups(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning, Eigen::MatrixXd X, Eigen::VectorXd Y, double sigma0, double v0E, double s02E, double v0G, double s02G, Eigen::MatrixXd cva, int groups, Eigen::VectorXi gAssign, Eigen::MatrixXd fixed);
void HorseshoeR(std::string outputFile, int seed, int max_iterations, int burn_in, int thinning, Eigen::MatrixXd X, Eigen::VectorXd Y, double A, double v0E, double s02E, double vL, double vT, double c2, double vC, double sC);
int main(int argc, char **)
{
    if (argc > 100) {   // never executed: the calls only have to compile and link
        Eigen::MatrixXd X(4, 2); Eigen::VectorXd v(4); Eigen::VectorXi g(2);
        BayesRSamplerV2("o", 1, 2, 1, 1, X, v, 0.01, 1e-4, 1e-3, 1e-4, 1e-3, v);
        BayesRSamplerV2Groups("o", 1, 2, 1, 1, X, v, 0.01, 1e-4, 1e-3, 1e-4, 1e-3, X, 1, g, X);
        BRV2Grstart("o", 1, 2, 1, 1, 0.0, X, 1.0, v, X, v, v, 0.01, 1e-4, 1e-3, 1e-4, 1e-3, X, 1, g);
        HorseshoeR("o", 1, 2, 1, 1, X, v, 0.1, 1e-3, 1e-3, 1.0, 1.0, 1.0, 10.0, 10.0);
    }
    std::cout << "linked\\n";
    return 0;
}
''')
    exe = tmp_path / "shim_check"
    libdir = os.path.dirname(brr.LIB_PATH)
    cmd = ["g++", "-std=c++11", "-Wall", "-Werror=return-type", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "shim", "rcpp_shim.cpp"), str(check), "-L" + libdir, "-l:libbayesrr_b200.so", "-Wl,-rpath," + libdir,
           "-Wl,--no-undefined", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nm = subprocess.run(["nm", "-C", "--defined-only", str(exe)], capture_output=True, text=True).stdout
    for f in ("BayesRSamplerV2(", "BayesRSamplerV2Groups(", "BRV2Grstart(", "HorseshoeR("):
        assert f in nm, f
    und = subprocess.run(["nm", "-C", "--undefined-only", str(exe)], capture_output=True, text=True).stdout
    for f in ("brr_BayesRSamplerV2", "brr_BayesRSamplerV2Groups", "brr_BRV2Grstart", "brr_HorseshoeR", "brr_set_message_handler", "brr_last_error"):
        assert f in und, f
    out = subprocess.run([str(exe)], capture_output=True, text=True)        # loads libbayesrr_b200.so (and the CUDA runtime in it)
    assert out.returncode == 0 and "linked" in out.stdout, out.stderr


def test_message_handler_is_exported_and_silent_by_default(brr):
    import ctypes as C
    L = brr.lib()
    FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)
    got = []
    cb = FN(lambda ctx, text: got.append(text))
    L.brr_set_message_handler.restype = None
    L.brr_set_message_handler.argtypes = [FN, C.c_void_p]
    L.brr_set_message_handler(cb, None)
    L.brr_set_message_handler(C.cast(None, FN), None)
    assert got == []


def test_gram_tile_layout_is_a_bijection_and_lookahead_is_exported(brr, tmp_path):
    """csrc/common.cuh: a block's self Gram tile is stored as its block-upper trapezoid.  Every (i, j) whose column sub-window is not
    before its row's maps to a distinct entry of [0, gram_tile_entries), rows are contiguous from the first column of their own
    sub-window on (what the sampler's register walk and the kernels' 128-bit stores assume), everything else maps to -1 while its mirror
    image is stored; the look-ahead depths the library reports are legal (multiples of 32, at most the block)."""
    import subprocess
    src = tmp_path / "layout.cu"
    src.write_text(r'''
#include "%s/bayesrrcpp_b200/csrc/common.cuh"
#include <cstdio>
#include <vector>
int main()
{
    for (int B : {32, 64, 128}) {
        const int TE = brr::gram_tile_entries(B);
        std::vector<int> seen(TE, 0);
        for (int i = 0; i < B; ++i)
            for (int j = 0; j < B; ++j) {
                const int idx = brr::gram_tile_index(B, i, j);
                if (j / 32 < i / 32) { if (idx != -1 || brr::gram_tile_index(B, j, i) < 0) { printf("mirror %%d %%d %%d\n", B, i, j); return 1; } continue; }
                if (idx < 0 || idx >= TE || seen[idx]++) { printf("index %%d %%d %%d -> %%d\n", B, i, j, idx); return 1; }
                if (j > 32 * (i / 32) && idx != brr::gram_tile_index(B, i, j - 1) + 1) { printf("row not contiguous %%d %%d %%d\n", B, i, j); return 1; }
                if (j == 32 * (i / 32) && i %% 32 == 0 && idx != brr::gram_subwindow_offset(B, i / 32)) { printf("offset %%d %%d\n", B, i); return 1; }
            }
        for (int v : seen) if (v != 1) { printf("hole %%d\n", B); return 1; }
        printf("%%d %%d %%d\n", B, TE, brr::lookahead(B));
    }
    return 0;
}
''' % ROOT)
    exe = tmp_path / "layout"
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    assert out[0].split()[:2] == ["32", "1024"] and out[1].split()[:2] == ["64", "3072"] and out[2].split()[:2] == ["128", "10240"]
    for block in (32, 64, 128):
        la = brr.lookahead(block)
        assert la % 32 == 0 and 32 <= la <= block
    with pytest.raises(ValueError):
        brr.lookahead(48)


def test_bench_reference_arm_line_keeps_the_driver_contract():
    """`bench.py --impl reference` (the oracle port on the host cores; runs without a GPU): one JSON line with the arm's keys -- the
    TIMED ms_per_step of the column sample, the extrapolated full-M figure under its own name, an e2e object that repeats the line's
    value with zero copy bytes, and a cpu_baseline describing the run."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "SNP-updates/sec" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["warmup"] == 3 and line["n_gpus"] == 1 and line["dtype"] == "f64"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["same_config"] is False
    # the timed step is the sample's, the extrapolation scales it by M / sample_markers
    assert abs(line["extrapolated_ms_per_iteration"] / line["ms_per_step"] - 50000 / line["sample_markers"]) < 1e-6
    assert abs(line["value"] - line["sample_markers"] / (line["ms_per_step"] * 1e-3)) / line["value"] < 1e-6
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == line["value"] and "sample" in cb
    assert "workload" in line["config"]
