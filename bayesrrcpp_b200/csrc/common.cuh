// Internal declarations shared by the translation units of libbayesrr_b200.so.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <stdexcept>
#include "../../include/bayesrr_b200.h"
#include "philox.cuh"

namespace brr {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);

#define BRR_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            throw brr::Error(BRR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) +       \
                                             " (" __FILE__ ":" + std::to_string(__LINE__) + ")");   \
    } while (0)

#define BRR_REQUIRE(cond, code, msg)                                                                \
    do { if (!(cond)) throw brr::Error((code), (msg)); } while (0)

// run `body`, translate exceptions into a return code + last-error string
template <class F> int guarded(F &&body)
{
    // errors of earlier calls were reported by those calls: they must not surface again in this one's cudaGetLastError() checks
    (void)cudaGetLastError();
    try { body(); return BRR_OK; }
    catch (const Error &e) { set_last_error(e.what()); (void)cudaGetLastError(); return e.code; }
    catch (const std::exception &e) { set_last_error(e.what()); (void)cudaGetLastError(); return BRR_E_ARG; }
}

// throws BRR_E_CUDA unless `device` exists and is compute capability 10.x; makes it current
void require_device(int device);

// Force the (lazily loaded) kernel into the context now.  Loading a kernel on its first launch can wait for the device to go
// idle; a chain whose persistent sweep kernel is already spinning on a peer rank of the same device would then never see
// that peer's launch.  Every kernel of the iteration loop is therefore loaded when the chain is created.
template <class K> void preload_kernel(K *kernel)
{
    cudaFuncAttributes a;
    BRR_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void *>(kernel)));
}

// Raise (never lower) a kernel's dynamic shared-memory limit on the current device; cached per (function, device).  A call per launch
// is not free: the driver may order it against running instances of the function -- and ranks that are threads of one process launch
// persistent kernels that WAIT for each other, so a rank whose launch sits behind a peer's running kernel times out -- and a smaller
// value set for one chain must never undercut the larger one another chain of the same geometry needs.
void ensure_dynamic_smem(const void *fn, size_t bytes);

// BRR_TRACE_SETUP=1: wall-clock milliseconds of the set-up stages on stderr (where the end-to-end time outside the iterations goes)
struct SetupTrace {
    bool on; const char *what; std::chrono::steady_clock::time_point t;
    explicit SetupTrace(const char *w) : on(getenv("BRR_TRACE_SETUP") != nullptr), what(w), t(std::chrono::steady_clock::now()) {}
    void mark(const char *stage)
    {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[brr setup] %s / %s: %.2f ms\n", what, stage, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// Look-ahead depth of the sweep for Gibbs blocks of B markers: the deltas of the last lookahead(B) markers of a block reach the
// next block through the cross-Gram correction instead of through the workers' dots (sweep.cu, gram.cu).
// 128-marker blocks: the whole block (the workers form the dots of block b + 1 while the sampler walks block b, so the per-block
// latency loop deltas -> residual update -> dots -> reducer -> (NVLink) -> sampler has a block's time to turn over; with 64 the
// sampler waited 1.6k cycles per block on one GPU and 5.7k on eight).  -DBRR_LOOKAHEAD128=64|96|128 (env BRR_LOOKAHEAD128 of
// bayesrrcpp_b200.build).
#ifndef BRR_LOOKAHEAD128
#define BRR_LOOKAHEAD128 128
#endif
static_assert(BRR_LOOKAHEAD128 == 32 || BRR_LOOKAHEAD128 == 64 || BRR_LOOKAHEAD128 == 96 || BRR_LOOKAHEAD128 == 128, "look-ahead depth of 128-marker blocks");
__host__ __device__ constexpr int lookahead(int B) { return B >= 128 ? BRR_LOOKAHEAD128 : B >= 64 ? 64 : 32; }

// Layout of a block's self Gram tile (gram.cu -> sweep.cu).  The walk reads row j of the tile only at the columns of j's own
// 32-marker sub-window and the later ones (the markers not yet visited), so a tile is stored as its block-upper trapezoid: the 32
// rows of sub-window q, each cut to the B - 32 q columns from 32 q on, one sub-window after the other -- 10,240 of 16,384 entries
// at B = 128.  One flat bulk copy per block as before, and the sampler's shared memory holds the deeper look-ahead's cross tile.
__host__ __device__ constexpr int gram_tile_entries(int B) { return 32 * (B * (B / 32) - 16 * (B / 32) * (B / 32 - 1)); }
__host__ __device__ constexpr int gram_subwindow_offset(int B, int q) { return 32 * q * B - 512 * q * (q - 1); }   // first entry of sub-window q's rows
// entry (i, j) of a tile, or -1 when it is not stored (j's sub-window precedes i's); the tile is symmetric: (j, i) is stored then
__host__ __device__ constexpr int gram_tile_index(int B, int i, int j)
{
    return j / 32 < i / 32 ? -1 : gram_subwindow_offset(B, i / 32) + (i % 32) * (B - 32 * (i / 32)) + (j - 32 * (i / 32));
}

constexpr int ROW_PAD = 512;   // rows per column are padded to a multiple of this (codes 0): 128-byte column stride

}  // namespace brr

// Genotype store (one device).  x[i, j] = a[j] + d[j] * code[i, j], code in {0,1,2}, 2 bits each, column-major.
// Columns that are not genotype-like (continuous covariates: the "methylation" group of vignettes/BayesRR.Rmd:47-57,150-167) are
// kept as dense fp64 columns beside the packed matrix: such a column j has dense_idx[j] >= 0, all-zero codes,
// a = 0, d = 1, and "code" stands for the value itself in every formula (S = sum x, Q = sum x^2).
struct brr_geno {
    int device = 0;
    int64_t N = 0, M = 0;          // local rows, markers
    int64_t Npad = 0, stride = 0;  // padded rows (multiple of ROW_PAD), bytes per column (= Npad / 4)
    uint8_t *d_packed = nullptr;
    // per-marker constants (device, fp64): a, d, S = sum code, Q = sum code^2, xsq = ||x||^2, csum = sum x
    double *d_a = nullptr, *d_d = nullptr, *d_S = nullptr, *d_Q = nullptr, *d_xsq = nullptr, *d_csum = nullptr;
    std::vector<double> h_a, h_d, h_S, h_Q, h_xsq;
    double n_total = 0;            // rows the statistics refer to (== N unless sharded)
    int64_t Md = 0;                // dense fp64 columns
    double *d_dense = nullptr;     // Npad x Md, column-major, padding rows 0
    int32_t *d_dense_idx = nullptr;   // M entries: index of the marker's dense column, or -1 (null when Md == 0)
    std::vector<int32_t> h_dense_idx;
    // A row shard of a .bed file with missing genotypes to impute: the fill value of a column is the rounded mean over the observed
    // genotypes of ALL ranks, so the codes 3 stay in place (and the statistics undefined) until brr_geno_shard_stats has summed
    // the per-column counts n0, n1, n2, n_missing (4 per marker) over the ranks.
    bool pending_impute = false;
    std::vector<double> pending_cnt;
};

namespace brr {
// recompute xsq / csum (device + host mirrors) from a, d, S, Q
void geno_finalize_stats(brr_geno *g);
// recompute a, d from S, Q and n_total (sd with the n_total - 1 denominator), then xsq / csum
void geno_affine_from_stats(brr_geno *g);
// resolve a pending imputation with the counts of all ranks (4 per marker), then the local code statistics
void geno_impute_pending(brr_geno *g, const std::vector<double> &cnt_all);
}
