// Genotype storage / packing layer.
// The reference keeps X as a dense fp64 Eigen::MatrixXd passed by value (reference src/BayesRv2.cpp:60;
// two deep copies, src/RcppExports.cpp:48-49).  Here X lives in HBM as 2-bit codes, column-major, with a
// per-marker affine map x = a + d*code, so a standardised column costs N/4 bytes instead of 8N.
#include "common.cuh"
#include <map>
#include <mutex>
#include <utility>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <thread>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace brr {

static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }
const char *last_error_cstr() { return g_last_error.c_str(); }

void ensure_dynamic_smem(const void *fn, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> have;
    int dev = 0;
    BRR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = have[std::make_pair(fn, dev)];
    if (bytes <= cur) return;
    BRR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
}

void require_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Error(BRR_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                    "): bayesrr_b200 has no CPU fallback");
    BRR_REQUIRE(device >= 0 && device < n, BRR_E_CUDA, "device index out of range");
    int major = 0, minor = 0;                     // two attribute reads: cudaGetDeviceProperties costs milliseconds per call
    BRR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    BRR_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10) {
        cudaDeviceProp p;
        BRR_CUDA(cudaGetDeviceProperties(&p, device));
        throw Error(BRR_E_CUDA, std::string("device ") + p.name + " is sm_" + std::to_string(major) + std::to_string(minor) +
                                    "; this library is built for sm_100a (B200) only");
    }
    BRR_CUDA(cudaSetDevice(device));
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread16(uint32_t x)   // bit i -> bit 2i
{
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

__device__ __forceinline__ double warp_min(double v) { for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ double warp_max(double v) { for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

// One CTA per dense column: find the (at most three, equally spaced) values, emit 2-bit codes with
// coalesced reads + ballot packing, and the integer statistics S = sum code, Q = sum code^2.
__global__ void __launch_bounds__(256) pack_dense_kernel(const double *__restrict__ X, int64_t N, int64_t ld,
                                                         uint8_t *__restrict__ packed, int64_t stride,
                                                         double *__restrict__ a_out, double *__restrict__ d_out,
                                                         double *__restrict__ S_out, double *__restrict__ Q_out,
                                                         int *__restrict__ not_geno /* per column: 1 = not of the form a + d * code */)
{
    __shared__ double s_lo[8], s_hi[8];
    __shared__ unsigned long long s_n1, s_n2;
    __shared__ int s_mid, s_bad;
    const int64_t col = blockIdx.x;
    const double *x = X + col * ld;
    uint32_t *out = reinterpret_cast<uint32_t *>(packed + col * stride);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_n1 = 0; s_n2 = 0; s_mid = 0; s_bad = 0; }
    double lo = INFINITY, hi = -INFINITY;
    int nanseen = 0;
    for (int64_t i = tid; i < N; i += 256) { const double v = x[i]; nanseen |= (v != v); lo = fmin(lo, v); hi = fmax(hi, v); }
    lo = warp_min(lo); hi = warp_max(hi);
    if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
    __syncthreads();
    if (nanseen) atomicOr(&s_bad, 1);
    lo = s_lo[0]; hi = s_hi[0];
    for (int w = 1; w < 8; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
    const double mid = lo + 0.5 * (hi - lo);
    const double tol = 1e-9 * fmax(fmax(fabs(lo), fabs(hi)), 1e-300);
    unsigned long long n1 = 0, n2 = 0;
    int anymid = 0, anybad = 0;
    const int64_t nwarp_rows = (N + 31) / 32;
    for (int64_t wr = warp; wr < nwarp_rows; wr += 8) {
        const int64_t i = wr * 32 + lane;
        int c = 0;
        if (i < N) {
            const double v = x[i];
            if (fabs(v - lo) <= tol) c = 0;
            else if (fabs(v - hi) <= tol) c = 2;
            else if (fabs(v - mid) <= tol) { c = 1; anymid = 1; }
            else anybad = 1;
        }
        const uint32_t b0 = __ballot_sync(0xffffffffu, c & 1), b1 = __ballot_sync(0xffffffffu, c & 2);
        if (lane == 0) {
            out[wr * 2] = spread16(b0) | (spread16(b1) << 1);
            out[wr * 2 + 1] = spread16(b0 >> 16) | (spread16(b1 >> 16) << 1);
            n1 += __popc(b0); n2 += __popc(b1);
        }
    }
    if (anymid) atomicOr(&s_mid, 1);
    if (anybad) atomicOr(&s_bad, 1);
    if (lane == 0) { atomicAdd(&s_n1, n1); atomicAdd(&s_n2, n2); }
    __syncthreads();
    const bool three = s_mid != 0;
    if (!three && hi > lo) {
        // two-valued column: only codes 0 and 2 (binary 10) were written; recode 2 -> 1
        for (int64_t w = tid; w < nwarp_rows * 2; w += 256) out[w] = (out[w] >> 1) & 0x55555555u;
    }
    if (tid == 0) {
        double c1 = (double)s_n1, c2 = (double)s_n2;
        if (!three) { c1 = c2; c2 = 0; }
        a_out[col] = lo;
        d_out[col] = hi > lo ? (three ? 0.5 * (hi - lo) : (hi - lo)) : 0.0;
        S_out[col] = c1 + 2.0 * c2;
        Q_out[col] = c1 + 4.0 * c2;
        not_geno[col] = s_bad;
    }
}

// Dense (continuous) columns, SURVEY.md 8f-n4: S = sum x, Q = sum x^2 in a fixed order (thread-strided sums, then a fixed tree);
// a = 0, d = 1 so that x = a + d * "code" holds with the value itself as the code.  One CTA per dense column.
__global__ void __launch_bounds__(256) dense_stats_kernel(const double *__restrict__ dense, int64_t Npad, int64_t N,
                                                          const int32_t *__restrict__ cols /* marker of every dense column */,
                                                          double *__restrict__ a, double *__restrict__ d,
                                                          double *__restrict__ S, double *__restrict__ Q)
{
    __shared__ double s1[8], s2[8];
    const double *x = dense + (int64_t)blockIdx.x * Npad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double u = 0.0, v = 0.0;
    for (int64_t i = tid; i < N; i += 256) { const double t = x[i]; u += t; v = fma(t, t, v); }
    for (int o = 16; o; o >>= 1) { u += __shfl_xor_sync(0xffffffffu, u, o); v += __shfl_xor_sync(0xffffffffu, v, o); }
    if (lane == 0) { s1[warp] = u; s2[warp] = v; }
    __syncthreads();
    if (tid == 0) {
        double A = 0.0, B = 0.0;
        for (int w = 0; w < 8; ++w) { A += s1[w]; B += s2[w]; }
        const int m = cols[blockIdx.x];
        S[m] = A; Q[m] = B; a[m] = 0.0; d[m] = 1.0;
    }
}

// One CTA per packed column: clear bits beyond row N, count codes; code 3 -> bad.
__global__ void __launch_bounds__(256) stats_packed_kernel(uint8_t *__restrict__ packed, int64_t stride, int64_t N,
                                                           double *__restrict__ S_out, double *__restrict__ Q_out,
                                                           int *__restrict__ bad)
{
    __shared__ unsigned long long s_n[3];
    const int64_t col = blockIdx.x;
    uint32_t *w = reinterpret_cast<uint32_t *>(packed + col * stride);
    const int tid = threadIdx.x;
    if (tid < 3) s_n[tid] = 0;
    __syncthreads();
    const int64_t nwords = stride / 4, full = N / 16;
    unsigned long long n1 = 0, n2 = 0, n3 = 0;
    for (int64_t i = tid; i < nwords; i += 256) {
        uint32_t v = w[i];
        if (i >= full) {
            const uint32_t keep = (i == full && (N & 15)) ? ((1u << (2 * (N & 15))) - 1u) : 0u;
            if ((v & ~keep) != 0) { v &= keep; w[i] = v; }
        }
        const uint32_t lo = v & 0x55555555u, hi = (v >> 1) & 0x55555555u;
        n1 += __popc(lo & ~hi); n2 += __popc(hi & ~lo); n3 += __popc(lo & hi);
    }
    atomicAdd(&s_n[0], n1); atomicAdd(&s_n[1], n2); atomicAdd(&s_n[2], n3);
    __syncthreads();
    if (tid == 0) {
        S_out[col] = (double)s_n[0] + 2.0 * (double)s_n[1];
        Q_out[col] = (double)s_n[0] + 4.0 * (double)s_n[1];
        if (s_n[2]) atomicAdd(bad, 1);
    }
}

// a = -mean/sd, d = 1/sd from the code statistics (sd with the N-1 denominator, R scale()); optional overrides.
__global__ void affine_from_stats_kernel(int64_t M, double n, const double *__restrict__ S, const double *__restrict__ Q,
                                         const double *__restrict__ mean_in, const double *__restrict__ sd_in,
                                         double *__restrict__ a, double *__restrict__ d, const int32_t *__restrict__ dense_idx)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= M) return;
    if (dense_idx && dense_idx[j] >= 0) { a[j] = 0.0; d[j] = 1.0; return; }   // a dense column is taken as given (the caller scaled it)
    const double mean = mean_in ? mean_in[j] : S[j] / n;
    double sd = sd_in ? sd_in[j] : sqrt(fmax(0.0, (Q[j] - S[j] * S[j] / n) / (n - 1.0)));
    if (!(sd > 0.0)) { a[j] = 0.0; d[j] = 0.0; return; }   // monomorphic column: x = 0
    a[j] = -mean / sd; d[j] = 1.0 / sd;
}

__global__ void finalize_stats_kernel(int64_t M, double n, const double *__restrict__ a, const double *__restrict__ d,
                                      const double *__restrict__ S, const double *__restrict__ Q,
                                      double *__restrict__ xsq, double *__restrict__ csum)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= M) return;
    // ||x||^2 = sum (a + d c)^2 = n a^2 + 2 a d S + d^2 Q  (the reference computes X.colwise().squaredNorm(), src/BayesRv2.cpp:170)
    xsq[j] = n * a[j] * a[j] + 2.0 * a[j] * d[j] * S[j] + d[j] * d[j] * Q[j];
    csum[j] = n * a[j] + d[j] * S[j];
}

// Synthetic genotypes, generated where they will live.  Thread = (marker, 16-row word).
__global__ void synth_kernel(uint8_t *__restrict__ packed, int64_t stride, int64_t N, int64_t M, int64_t row0, PhiloxKey key)
{
    const int64_t wpc = stride / 4;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= wpc * M) return;
    const int64_t col = t / wpc, wi = t % wpc;
    uint32_t word = 0;
    if (wi * 16 < N) {
        const double p = 0.05 + 0.45 * draw_uniform(key, 20, -1, col);
        const uint32_t thr = (uint32_t)(p * 4294967296.0);
        for (int q = 0; q < 8; ++q) {
            const int64_t r = wi * 16 + 2 * q;   // local rows r, r+1
            uint32_t w[4];
            philox4x32_10(key, (uint32_t)(((uint64_t)(row0 + r) >> 1) & 0xffffffffu), (uint32_t)col, 21u,
                          (uint32_t)((uint64_t)col >> 32) ^ (uint32_t)(((uint64_t)(row0 + r) >> 33) << 8), w);
            const uint32_t c0 = (w[0] < thr) + (w[1] < thr), c1 = (w[2] < thr) + (w[3] < thr);
            if (r < N) word |= c0 << (4 * q);
            if (r + 1 < N) word |= c1 << (4 * q + 2);
        }
    }
    reinterpret_cast<uint32_t *>(packed + col * stride)[wi] = word;
}

// y = X b over the non-zero entries of b: thread = 16-row word, coalesced across threads for each column.
__global__ void matvec_kernel(const uint8_t *__restrict__ packed, int64_t stride, int64_t N, int nnz,
                              const int32_t *__restrict__ cols, const double *__restrict__ vals,
                              const double *__restrict__ a, const double *__restrict__ d, double *__restrict__ y,
                              const double *__restrict__ dense, const int32_t *__restrict__ dense_idx, int64_t Npad)
{
    const int64_t wi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (wi * 16 >= N) return;
    double acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.0;
    for (int e = 0; e < nnz; ++e) {
        const int64_t col = cols[e];
        if (dense_idx && dense_idx[col] >= 0) {
            const double *x = dense + (int64_t)dense_idx[col] * Npad + wi * 16;
            const double b = vals[e];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] += b * x[q];
            continue;
        }
        const uint32_t w = reinterpret_cast<const uint32_t *>(packed + col * stride)[wi];
        const double b = vals[e], aa = a[col] * b, dd = d[col] * b;
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] += aa + dd * (double)((w >> (2 * q)) & 3u);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) if (wi * 16 + q < N) y[wi * 16 + q] = acc[q];
}

// r[j] = x_j^T eps: one warp per marker, 128-bit loads of the packed column, fp64 accumulation,
// warp-shuffle reduction.  (Stand-alone form of the dot the sweep kernel fuses; used for tests.)
__global__ void __launch_bounds__(256) xt_eps_kernel(const uint8_t *__restrict__ packed, int64_t stride, int64_t N, int64_t M,
                                                      const double *__restrict__ eps, double eps_sum,
                                                      const double *__restrict__ a, const double *__restrict__ d,
                                                      double *__restrict__ r,
                                                      const double *__restrict__ dense, const int32_t *__restrict__ dense_idx, int64_t Npad)
{
    const int64_t col = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (col >= M) return;
    const int lane = threadIdx.x & 31;
    if (dense_idx && dense_idx[col] >= 0) {
        const double *x = dense + (int64_t)dense_idx[col] * Npad;
        double s = 0.0;
        for (int64_t i = lane; i < N; i += 32) s = fma(x[i], eps[i], s);
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) r[col] = s;
        return;
    }
    const uint4 *p = reinterpret_cast<const uint4 *>(packed + col * stride);
    const int64_t nv = (N + 63) / 64;
    double s = 0.0;
    for (int64_t v = lane; v < nv; v += 32) {
        const uint4 q = p[v];
        const uint32_t w[4] = { q.x, q.y, q.z, q.w };
        const int64_t base = v * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int64_t i = base + k * 16 + t;
                const uint32_t c = (w[k] >> (2 * t)) & 3u;
                if (c && i < N) s += (double)c * eps[i];
            }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) r[col] = a[col] * eps_sum + d[col] * s;
}

void geno_finalize_stats(brr_geno *g)
{
    const int64_t M = g->M;
    const int tb = 256; const unsigned nb = (unsigned)((M + tb - 1) / tb);
    finalize_stats_kernel<<<nb, tb>>>(M, g->n_total, g->d_a, g->d_d, g->d_S, g->d_Q, g->d_xsq, g->d_csum);
    BRR_CUDA(cudaGetLastError());
    g->h_a.resize(M); g->h_d.resize(M); g->h_S.resize(M); g->h_Q.resize(M); g->h_xsq.resize(M);
    BRR_CUDA(cudaMemcpy(g->h_a.data(), g->d_a, M * 8, cudaMemcpyDeviceToHost));
    BRR_CUDA(cudaMemcpy(g->h_d.data(), g->d_d, M * 8, cudaMemcpyDeviceToHost));
    BRR_CUDA(cudaMemcpy(g->h_S.data(), g->d_S, M * 8, cudaMemcpyDeviceToHost));
    BRR_CUDA(cudaMemcpy(g->h_Q.data(), g->d_Q, M * 8, cudaMemcpyDeviceToHost));
    BRR_CUDA(cudaMemcpy(g->h_xsq.data(), g->d_xsq, M * 8, cudaMemcpyDeviceToHost));
}

void geno_affine_from_stats(brr_geno *g)
{
    const int64_t M = g->M;
    affine_from_stats_kernel<<<(unsigned)((M + 255) / 256), 256>>>(M, g->n_total, g->d_S, g->d_Q, nullptr, nullptr, g->d_a, g->d_d, g->d_dense_idx);
    BRR_CUDA(cudaGetLastError());
    geno_finalize_stats(g);
}

static brr_geno *geno_alloc(int64_t N, int64_t M, int device)
{
    BRR_REQUIRE(N >= 2 && M >= 1, BRR_E_ARG, "genotype matrix needs N >= 2 rows and M >= 1 markers");
    BRR_REQUIRE(N < (int64_t)1 << 31 && M < (int64_t)1 << 31, BRR_E_SIZE, "N and M must fit 31 bits");
    require_device(device);
    brr_geno *g = new brr_geno();
    g->device = device; g->N = N; g->M = M; g->n_total = (double)N;
    g->Npad = (N + ROW_PAD - 1) / ROW_PAD * ROW_PAD;
    g->stride = g->Npad / 4;
    try {
        BRR_CUDA(cudaMalloc(&g->d_packed, (size_t)g->stride * M));
        BRR_CUDA(cudaMemset(g->d_packed, 0, (size_t)g->stride * M));
        for (double **p : { &g->d_a, &g->d_d, &g->d_S, &g->d_Q, &g->d_xsq, &g->d_csum }) BRR_CUDA(cudaMalloc(p, M * 8));
    } catch (...) { brr_geno_free(g); throw; }
    return g;
}

}  // namespace brr

using namespace brr;

extern "C" const char *brr_last_error(void) { return last_error_cstr(); }
extern "C" int brr_abi_version(void) { return 1; }

extern "C" void brr_geno_free(brr_geno *g)
{
    if (!g) return;
    cudaSetDevice(g->device);
    cudaFree(g->d_packed); cudaFree(g->d_dense); cudaFree(g->d_dense_idx);
    for (double *p : { g->d_a, g->d_d, g->d_S, g->d_Q, g->d_xsq, g->d_csum }) cudaFree(p);
    delete g;
}

// Make the markers `cols` (ascending, distinct) of the store dense fp64 columns holding `values` (host, N x n column-major with
// leading dimension ld): their packed codes are cleared, a = 0, d = 1, S / Q from the values.  Replaces earlier dense columns.
static void geno_set_dense(brr_geno *g, const std::vector<int32_t> &cols, const double *values, int64_t ld)
{
    const int64_t M = g->M, N = g->N, n = (int64_t)cols.size();
    cudaFree(g->d_dense); cudaFree(g->d_dense_idx); g->d_dense = nullptr; g->d_dense_idx = nullptr; g->Md = 0;
    g->h_dense_idx.assign((size_t)M, -1);
    if (n == 0) return;
    int32_t *d_cols = nullptr;
    try {
        for (int64_t k = 0; k < n; ++k) {
            BRR_REQUIRE(cols[k] >= 0 && cols[k] < M && (k == 0 || cols[k] > cols[k - 1]), BRR_E_ARG, "dense column list must be ascending, distinct and inside [0, M)");
            g->h_dense_idx[cols[k]] = (int32_t)k;
        }
        BRR_CUDA(cudaMalloc(&g->d_dense, (size_t)g->Npad * n * 8));
        BRR_CUDA(cudaMemset(g->d_dense, 0, (size_t)g->Npad * n * 8));
        BRR_CUDA(cudaMemcpy2D(g->d_dense, (size_t)g->Npad * 8, values, (size_t)ld * 8, (size_t)N * 8, (size_t)n, cudaMemcpyHostToDevice));
        BRR_CUDA(cudaMalloc(&g->d_dense_idx, (size_t)M * 4));
        BRR_CUDA(cudaMemcpy(g->d_dense_idx, g->h_dense_idx.data(), (size_t)M * 4, cudaMemcpyHostToDevice));
        BRR_CUDA(cudaMalloc(&d_cols, (size_t)n * 4));
        BRR_CUDA(cudaMemcpy(d_cols, cols.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
        for (int64_t k = 0; k < n; ++k) BRR_CUDA(cudaMemsetAsync(g->d_packed + (size_t)cols[k] * g->stride, 0, (size_t)g->stride, 0));
        dense_stats_kernel<<<(unsigned)n, 256>>>(g->d_dense, g->Npad, N, d_cols, g->d_a, g->d_d, g->d_S, g->d_Q);
        BRR_CUDA(cudaGetLastError());
        BRR_CUDA(cudaDeviceSynchronize());
        g->Md = n;
    } catch (...) { cudaFree(d_cols); throw; }
    cudaFree(d_cols);
}

extern "C" int brr_geno_from_dense(const double *X, int64_t N, int64_t M, int device, brr_geno **out)
{
    return guarded([&] {
        BRR_REQUIRE(X && out, BRR_E_ARG, "brr_geno_from_dense: null pointer");
        brr_geno *g = geno_alloc(N, M, device);
        double *d_chunk = nullptr; int *d_flag = nullptr;
        try {
            const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(M, ((int64_t)256 << 20) / (8 * N)));
            BRR_CUDA(cudaMalloc(&d_chunk, (size_t)chunk * N * 8));
            BRR_CUDA(cudaMalloc(&d_flag, (size_t)M * sizeof(int)));
            for (int64_t c0 = 0; c0 < M; c0 += chunk) {
                const int64_t nc = std::min(chunk, M - c0);
                BRR_CUDA(cudaMemcpy(d_chunk, X + c0 * N, (size_t)nc * N * 8, cudaMemcpyHostToDevice));
                pack_dense_kernel<<<(unsigned)nc, 256>>>(d_chunk, N, N, g->d_packed + c0 * g->stride, g->stride,
                                                          g->d_a + c0, g->d_d + c0, g->d_S + c0, g->d_Q + c0, d_flag + c0);
                BRR_CUDA(cudaGetLastError());
            }
            // Columns that are not a + d * code with code in {0,1,2} -- continuous covariates, as the reference's Eigen::MatrixXd X
            // admits (vignettes/BayesRR.Rmd:150-167 binds scale()d methylation probes to the genotypes) -- stay dense fp64.
            std::vector<int> flag((size_t)M);
            BRR_CUDA(cudaMemcpy(flag.data(), d_flag, (size_t)M * sizeof(int), cudaMemcpyDeviceToHost));
            std::vector<int32_t> dense_cols;
            for (int64_t j = 0; j < M; ++j) if (flag[j]) dense_cols.push_back((int32_t)j);
            if (!dense_cols.empty()) {
                // gather the dense columns on the host side by side (they are scattered in X), then one strided upload
                std::vector<double> vals((size_t)N * dense_cols.size());
                for (size_t k = 0; k < dense_cols.size(); ++k) memcpy(&vals[k * (size_t)N], X + (size_t)dense_cols[k] * N, (size_t)N * 8);
                geno_set_dense(g, dense_cols, vals.data(), N);
            }
            geno_finalize_stats(g);
        } catch (...) { cudaFree(d_chunk); cudaFree(d_flag); brr_geno_free(g); throw; }
        cudaFree(d_chunk); cudaFree(d_flag);
        *out = g;
    });
}

extern "C" int brr_geno_set_dense_columns(brr_geno *g, const int32_t *cols, int64_t n_cols, const double *values)
{
    return guarded([&] {
        BRR_REQUIRE(g && (n_cols == 0 || (cols && values)) && n_cols >= 0, BRR_E_ARG, "brr_geno_set_dense_columns: bad arguments");
        BRR_REQUIRE(!g->pending_impute, BRR_E_ARG, "a .bed row shard with missing genotypes needs brr_geno_shard_stats first");
        BRR_CUDA(cudaSetDevice(g->device));
        geno_set_dense(g, std::vector<int32_t>(cols, cols + n_cols), values, g->N);
        geno_finalize_stats(g);
    });
}

extern "C" int brr_geno_dense_columns(const brr_geno *g, int64_t *n_dense, int32_t *dense_idx)
{
    return guarded([&] {
        BRR_REQUIRE(g, BRR_E_ARG, "null genotype store");
        if (n_dense) *n_dense = g->Md;
        if (dense_idx) for (int64_t j = 0; j < g->M; ++j) dense_idx[j] = g->Md ? g->h_dense_idx[j] : -1;
    });
}

static void stats_from_codes(brr_geno *g, const double *mean, const double *sd)
{
    int *d_bad = nullptr; double *d_mean = nullptr, *d_sd = nullptr;
    const int64_t M = g->M;
    try {
        BRR_CUDA(cudaMalloc(&d_bad, sizeof(int)));
        BRR_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
        stats_packed_kernel<<<(unsigned)M, 256>>>(g->d_packed, g->stride, g->N, g->d_S, g->d_Q, d_bad);
        BRR_CUDA(cudaGetLastError());
        int bad = 0;
        BRR_CUDA(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
        BRR_REQUIRE(bad == 0, BRR_E_GENO, std::to_string(bad) + " column(s) contain code 3 (missing genotypes are not supported)");
        if (mean) { BRR_CUDA(cudaMalloc(&d_mean, M * 8)); BRR_CUDA(cudaMemcpy(d_mean, mean, M * 8, cudaMemcpyHostToDevice)); }
        if (sd) { BRR_CUDA(cudaMalloc(&d_sd, M * 8)); BRR_CUDA(cudaMemcpy(d_sd, sd, M * 8, cudaMemcpyHostToDevice)); }
        affine_from_stats_kernel<<<(unsigned)((M + 255) / 256), 256>>>(M, g->n_total, g->d_S, g->d_Q, d_mean, d_sd, g->d_a, g->d_d, g->d_dense_idx);
        BRR_CUDA(cudaGetLastError());
        geno_finalize_stats(g);
    } catch (...) { cudaFree(d_bad); cudaFree(d_mean); cudaFree(d_sd); throw; }
    cudaFree(d_bad); cudaFree(d_mean); cudaFree(d_sd);
}

// Host (pageable) columns -> device columns with the padded stride.  cudaMemcpy2D from pageable memory moves ~1.4 GB/s here;
// instead the columns are re-strided by a few host threads into two pinned staging buffers whose 1-D copies overlap the next
// chunk's staging.
static void upload_columns(uint8_t *d_dst, size_t dpitch, const uint8_t *src, size_t spitch, size_t width, size_t ncols)
{
    // page-locked source (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch tensor): the copy engine reads it in place
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
        if (spitch == dpitch) BRR_CUDA(cudaMemcpy(d_dst, src, ncols * dpitch - (dpitch - width), cudaMemcpyHostToDevice));
        else BRR_CUDA(cudaMemcpy2D(d_dst, dpitch, src, spitch, width, ncols, cudaMemcpyHostToDevice));
        return;
    }
    (void)cudaGetLastError();
    const size_t chunk_bytes = (size_t)16 << 20;
    const size_t cols_per = std::max<size_t>(1, chunk_bytes / dpitch);
    // the two pinned staging buffers are kept for the life of the process (pinning memory costs milliseconds per call and varies)
    static std::mutex stage_mu;
    static uint8_t *stage_buf[2] = { nullptr, nullptr };
    static size_t stage_cap = 0;
    std::lock_guard<std::mutex> stage_lock(stage_mu);
    const size_t want = std::min(cols_per, ncols) * dpitch;
    if (want > stage_cap) {
        for (int i = 0; i < 2; ++i) { if (stage_buf[i]) cudaFreeHost(stage_buf[i]); stage_buf[i] = nullptr; }
        stage_cap = 0;
        for (int i = 0; i < 2; ++i) BRR_CUDA(cudaMallocHost(&stage_buf[i], std::max(want, chunk_bytes + dpitch)));
        stage_cap = std::max(want, chunk_bytes + dpitch);
    }
    uint8_t *buf[2] = { stage_buf[0], stage_buf[1] }; cudaEvent_t ev[2] = { nullptr, nullptr }; cudaStream_t st = nullptr;
    try {
        for (int i = 0; i < 2; ++i) BRR_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        BRR_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        int i = 0; bool used[2] = { false, false };
        for (size_t c0 = 0; c0 < ncols; c0 += cols_per, i ^= 1) {
            const size_t n = std::min(cols_per, ncols - c0);
            if (used[i]) BRR_CUDA(cudaEventSynchronize(ev[i]));
            const int nthreads = 4;
            std::vector<std::thread> th;
            for (int t = 0; t < nthreads; ++t)
                th.emplace_back([=] {
                    const size_t a = n * t / nthreads, b = n * (t + 1) / nthreads;
                    if (spitch == dpitch) { if (b > a) memcpy(buf[i] + a * dpitch, src + (c0 + a) * spitch, (b - a) * dpitch - (dpitch - width)); }
                    else for (size_t c = a; c < b; ++c) memcpy(buf[i] + c * dpitch, src + (c0 + c) * spitch, width);
                });
            for (auto &t : th) t.join();
            BRR_CUDA(cudaMemcpyAsync(d_dst + c0 * dpitch, buf[i], n * dpitch - (dpitch - width), cudaMemcpyHostToDevice, st));
            BRR_CUDA(cudaEventRecord(ev[i], st));
            used[i] = true;
        }
        BRR_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        for (int i = 0; i < 2; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
        if (st) cudaStreamDestroy(st);
        throw;
    }
    for (int i = 0; i < 2; ++i) cudaEventDestroy(ev[i]);
    cudaStreamDestroy(st);
}

extern "C" int brr_geno_from_packed(const uint8_t *packed, int64_t col_stride_bytes, int64_t N, int64_t M,
                                    const double *mean, const double *sd, int device, brr_geno **out)
{
    return guarded([&] {
        BRR_REQUIRE(packed && out, BRR_E_ARG, "brr_geno_from_packed: null pointer");
        const int64_t width = (N + 3) / 4;
        BRR_REQUIRE(col_stride_bytes >= width, BRR_E_ARG, "col_stride_bytes smaller than ceil(N/4)");
        SetupTrace tr("geno_from_packed");
        brr_geno *g = geno_alloc(N, M, device);
        tr.mark("device store (allocation, zero fill)");
        try {
            upload_columns(g->d_packed, (size_t)g->stride, packed, (size_t)col_stride_bytes, (size_t)width, (size_t)M);
            tr.mark("host -> device copy of the packed columns");
            stats_from_codes(g, mean, sd);
            tr.mark("per-marker statistics");
        } catch (...) { brr_geno_free(g); throw; }
        *out = g;
    });
}

// ---- PLINK .bed ingest.  The file is already 2 bits per genotype, SNP-major, four individuals per byte, low
// bits first -- the layout of this store -- with another code book: 00 homozygous A1, 10 heterozygous, 11 homozygous A2,
// 01 missing.  Codes here count A1 alleles (PLINK's additive coding): 00 -> 2, 10 -> 1, 11 -> 0, 01 -> 3 (missing, resolved below).
__global__ void __launch_bounds__(256) bed_remap_kernel(uint8_t *__restrict__ packed, int64_t stride, int64_t N, int64_t M,
                                                        unsigned long long *__restrict__ cnt /* [M][4]: n0, n1, n2, n_missing */)
{
    __shared__ unsigned long long s_n[4];
    const int64_t col = blockIdx.x;
    uint32_t *w = reinterpret_cast<uint32_t *>(packed + col * stride);
    const int tid = threadIdx.x;
    if (tid < 4) s_n[tid] = 0;
    __syncthreads();
    const int64_t nwords = stride / 4, full = N / 16;
    unsigned long long n1 = 0, n2 = 0, n3 = 0;
    for (int64_t i = tid; i < nwords; i += 256) {
        const uint32_t v = w[i];
        const uint32_t lo = v & 0x55555555u, hi = (v >> 1) & 0x55555555u;
        uint32_t c_hi = ~hi & ~lo & 0x55555555u, c_lo = hi & ~lo, miss = ~hi & lo & 0x55555555u;     // per 2-bit field, in the low bit
        uint32_t keep = 0xffffffffu;                                                                 // fields of real individuals
        if (i >= full) keep = (i == full && (N & 15)) ? ((1u << (2 * (N & 15))) - 1u) : 0u;
        const uint32_t k1 = keep & 0x55555555u;
        c_hi &= k1; c_lo &= k1; miss &= k1;
        w[i] = (c_hi << 1) | c_lo | miss | (miss << 1);                                              // missing -> code 3 for now
        n1 += __popc(c_lo); n2 += __popc(c_hi); n3 += __popc(miss);
    }
    atomicAdd(&s_n[1], n1); atomicAdd(&s_n[2], n2); atomicAdd(&s_n[3], n3);
    __syncthreads();
    if (tid < 4) cnt[col * 4 + tid] = tid == 0 ? (unsigned long long)N - s_n[1] - s_n[2] - s_n[3] : s_n[tid];
}
// missing genotypes -> the integer code nearest to the column's mean over the observed genotypes
__global__ void __launch_bounds__(256) bed_impute_kernel(uint8_t *__restrict__ packed, int64_t stride, int64_t M,
                                                         const unsigned long long *__restrict__ cnt)
{
    const int64_t col = blockIdx.x;
    const unsigned long long n1 = cnt[col * 4 + 1], n2 = cnt[col * 4 + 2], nm = cnt[col * 4 + 3], n0 = cnt[col * 4];
    if (nm == 0) return;
    const unsigned long long obs = n0 + n1 + n2;
    const uint32_t fill = obs ? (uint32_t)((double)(n1 + 2 * n2) / (double)obs + 0.5) : 0u;             // 0, 1 or 2
    const uint32_t pat = fill * 0x55555555u;                                                             // the code in every field
    uint32_t *w = reinterpret_cast<uint32_t *>(packed + col * stride);
    for (int64_t i = threadIdx.x; i < stride / 4; i += 256) {
        const uint32_t v = w[i];
        const uint32_t m = v & (v >> 1) & 0x55555555u;          // fields holding 3
        if (m) { const uint32_t mask = m | (m << 1); w[i] = (v & ~mask) | (pat & mask); }
    }
}

extern "C" int brr_geno_from_bed(const char *bed_path, int64_t N_total, int64_t M, int64_t row0, int64_t N, int impute_missing,
                                 int device, brr_geno **out, int64_t *n_missing)
{
    return guarded([&] {
        BRR_REQUIRE(bed_path && out && N_total > 0 && M > 0, BRR_E_ARG, "brr_geno_from_bed: bad arguments");
        if (N <= 0) { row0 = 0; N = N_total; }
        BRR_REQUIRE(row0 >= 0 && row0 % 4 == 0 && row0 + N <= N_total, BRR_E_ARG, "row shard must start at a multiple of 4 and lie inside the file's individuals");
        const int64_t width_total = (N_total + 3) / 4, width = (N + 3) / 4;
        const int fd = open(bed_path, O_RDONLY);
        BRR_REQUIRE(fd >= 0, BRR_E_IO, std::string("cannot open '") + bed_path + "'");
        struct stat stt;
        const bool ok_stat = fstat(fd, &stt) == 0;
        const int64_t need = 3 + width_total * M;
        if (!ok_stat || (int64_t)stt.st_size < need) { close(fd); throw Error(BRR_E_IO, std::string("'") + bed_path + "' is shorter than 3 + ceil(N/4) * M bytes: wrong N or M?"); }
        void *map = mmap(nullptr, (size_t)need, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        BRR_REQUIRE(map != MAP_FAILED, BRR_E_IO, std::string("cannot map '") + bed_path + "'");
        const uint8_t *file = static_cast<const uint8_t *>(map);
        brr_geno *g = nullptr; unsigned long long *d_cnt = nullptr;
        try {
            BRR_REQUIRE(file[0] == 0x6c && file[1] == 0x1b, BRR_E_IO, "not a PLINK .bed file (magic number)");
            BRR_REQUIRE(file[2] == 0x01, BRR_E_IO, "individual-major .bed files are not supported (re-export SNP-major: the PLINK 1.9 default)");
            g = geno_alloc(N, M, device);
            upload_columns(g->d_packed, (size_t)g->stride, file + 3 + row0 / 4, (size_t)width_total, (size_t)width, (size_t)M);
            BRR_CUDA(cudaMalloc(&d_cnt, (size_t)M * 4 * sizeof(unsigned long long)));
            bed_remap_kernel<<<(unsigned)M, 256>>>(g->d_packed, g->stride, N, M, d_cnt);
            BRR_CUDA(cudaGetLastError());
            std::vector<unsigned long long> cnt((size_t)M * 4);
            BRR_CUDA(cudaMemcpy(cnt.data(), d_cnt, cnt.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            int64_t missing = 0;
            for (int64_t j = 0; j < M; ++j) missing += (int64_t)cnt[(size_t)j * 4 + 3];
            if (n_missing) *n_missing = missing;
            const bool shard = row0 != 0 || N != N_total;
            if (missing > 0) BRR_REQUIRE(impute_missing != 0, BRR_E_GENO, std::to_string(missing) + " missing genotypes in '" + bed_path +
                            "' (the reference has no notion of missing data; pass impute_missing to fill them with the rounded column mean)");
            if (impute_missing != 0 && shard) {
                // the fill value is the rounded mean over the observed genotypes of the WHOLE column: decided in brr_geno_shard_stats,
                // once the counts of all row shards are known (a shard without missing genotypes still contributes its counts)
                g->pending_impute = true;
                g->pending_cnt.assign(cnt.begin(), cnt.end());
            } else {
                if (missing > 0) {
                    bed_impute_kernel<<<(unsigned)M, 256>>>(g->d_packed, g->stride, M, d_cnt);
                    BRR_CUDA(cudaGetLastError());
                }
                stats_from_codes(g, nullptr, nullptr);
            }
        } catch (...) { munmap(map, (size_t)need); cudaFree(d_cnt); brr_geno_free(g); throw; }
        munmap(map, (size_t)need); cudaFree(d_cnt);
        *out = g;
    });
}

namespace brr {
void geno_impute_pending(brr_geno *g, const std::vector<double> &cnt_all)
{
    const int64_t M = g->M;
    BRR_REQUIRE((int64_t)cnt_all.size() == 4 * M, BRR_E_ARG, "imputation counts do not match the store");
    std::vector<unsigned long long> cnt((size_t)4 * M);
    for (size_t i = 0; i < cnt.size(); ++i) cnt[i] = (unsigned long long)cnt_all[i];
    unsigned long long *d_cnt = nullptr;
    try {
        BRR_CUDA(cudaMalloc(&d_cnt, cnt.size() * sizeof(unsigned long long)));
        BRR_CUDA(cudaMemcpy(d_cnt, cnt.data(), cnt.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        // the kernel skips columns whose (global) missing count is zero; a column this shard has no missing genotype in is left as it is
        bed_impute_kernel<<<(unsigned)M, 256>>>(g->d_packed, g->stride, M, d_cnt);
        BRR_CUDA(cudaGetLastError());
        stats_from_codes(g, nullptr, nullptr);
    } catch (...) { cudaFree(d_cnt); throw; }
    cudaFree(d_cnt);
    g->pending_impute = false; g->pending_cnt.clear();
}
}  // namespace brr

extern "C" int brr_geno_synthetic(int64_t N, int64_t M, uint64_t seed, int64_t row0, int device, brr_geno **out)
{
    return guarded([&] {
        BRR_REQUIRE(out, BRR_E_ARG, "brr_geno_synthetic: null pointer");
        brr_geno *g = geno_alloc(N, M, device);
        try {
            const PhiloxKey key{ (uint32_t)seed, (uint32_t)(seed >> 32) };
            const int64_t total = g->stride / 4 * M;
            synth_kernel<<<(unsigned)((total + 255) / 256), 256>>>(g->d_packed, g->stride, N, M, row0, key);
            BRR_CUDA(cudaGetLastError());
            stats_from_codes(g, nullptr, nullptr);
        } catch (...) { brr_geno_free(g); throw; }
        *out = g;
    });
}

extern "C" int brr_geno_dims(const brr_geno *g, int64_t *N, int64_t *M, int64_t *col_stride_bytes)
{
    return guarded([&] {
        BRR_REQUIRE(g, BRR_E_ARG, "null genotype store");
        if (N) *N = g->N;
        if (M) *M = g->M;
        if (col_stride_bytes) *col_stride_bytes = g->stride;
    });
}

extern "C" int brr_geno_stats(const brr_geno *g, double *mean, double *sd, double *a, double *d, double *xsq)
{
    return guarded([&] {
        BRR_REQUIRE(g, BRR_E_ARG, "null genotype store");
        BRR_REQUIRE(!g->pending_impute, BRR_E_ARG, "statistics of a .bed row shard with missing genotypes exist only after brr_geno_shard_stats");
        const double n = g->n_total;
        for (int64_t j = 0; j < g->M; ++j) {
            if (mean) mean[j] = g->h_S[j] / n;
            if (sd) sd[j] = std::sqrt(std::max(0.0, (g->h_Q[j] - g->h_S[j] * g->h_S[j] / n) / (n - 1.0)));
            if (a) a[j] = g->h_a[j];
            if (d) d[j] = g->h_d[j];
            if (xsq) xsq[j] = g->h_xsq[j];
        }
    });
}

extern "C" int brr_geno_codes(const brr_geno *g, uint8_t *packed_out)
{
    return guarded([&] {
        BRR_REQUIRE(g && packed_out, BRR_E_ARG, "null pointer");
        BRR_CUDA(cudaSetDevice(g->device));
        BRR_CUDA(cudaMemcpy(packed_out, g->d_packed, (size_t)g->stride * g->M, cudaMemcpyDeviceToHost));
    });
}

extern "C" int brr_geno_matvec(const brr_geno *g, const double *b, double *y)
{
    return guarded([&] {
        BRR_REQUIRE(g && b && y, BRR_E_ARG, "null pointer");
        BRR_REQUIRE(!g->pending_impute, BRR_E_ARG, "a .bed row shard with missing genotypes needs brr_geno_shard_stats first");
        BRR_CUDA(cudaSetDevice(g->device));
        std::vector<int32_t> cols; std::vector<double> vals;
        for (int64_t j = 0; j < g->M; ++j) if (b[j] != 0.0) { cols.push_back((int32_t)j); vals.push_back(b[j]); }
        int32_t *d_cols = nullptr; double *d_vals = nullptr, *d_y = nullptr;
        try {
            const size_t nnz = cols.size();
            BRR_CUDA(cudaMalloc(&d_cols, std::max<size_t>(1, nnz) * 4));
            BRR_CUDA(cudaMalloc(&d_vals, std::max<size_t>(1, nnz) * 8));
            BRR_CUDA(cudaMalloc(&d_y, g->N * 8));
            if (nnz) {
                BRR_CUDA(cudaMemcpy(d_cols, cols.data(), nnz * 4, cudaMemcpyHostToDevice));
                BRR_CUDA(cudaMemcpy(d_vals, vals.data(), nnz * 8, cudaMemcpyHostToDevice));
            }
            const int64_t nw = (g->N + 15) / 16;
            matvec_kernel<<<(unsigned)((nw + 127) / 128), 128>>>(g->d_packed, g->stride, g->N, (int)nnz, d_cols, d_vals, g->d_a, g->d_d, d_y, g->d_dense, g->d_dense_idx, g->Npad);
            BRR_CUDA(cudaGetLastError());
            BRR_CUDA(cudaMemcpy(y, d_y, g->N * 8, cudaMemcpyDeviceToHost));
        } catch (...) { cudaFree(d_cols); cudaFree(d_vals); cudaFree(d_y); throw; }
        cudaFree(d_cols); cudaFree(d_vals); cudaFree(d_y);
    });
}

extern "C" int brr_xt_eps(const brr_geno *g, const double *eps, double *r, double *ms)
{
    return guarded([&] {
        BRR_REQUIRE(g && eps && r, BRR_E_ARG, "null pointer");
        BRR_REQUIRE(!g->pending_impute, BRR_E_ARG, "a .bed row shard with missing genotypes needs brr_geno_shard_stats first");
        BRR_CUDA(cudaSetDevice(g->device));
        double *d_eps = nullptr, *d_r = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
        try {
            BRR_CUDA(cudaMalloc(&d_eps, g->N * 8)); BRR_CUDA(cudaMalloc(&d_r, g->M * 8));
            BRR_CUDA(cudaMemcpy(d_eps, eps, g->N * 8, cudaMemcpyHostToDevice));
            double es = 0.0; for (int64_t i = 0; i < g->N; ++i) es += eps[i];
            BRR_CUDA(cudaEventCreate(&e0)); BRR_CUDA(cudaEventCreate(&e1));
            BRR_CUDA(cudaEventRecord(e0));
            xt_eps_kernel<<<(unsigned)((g->M + 7) / 8), 256>>>(g->d_packed, g->stride, g->N, g->M, d_eps, es, g->d_a, g->d_d, d_r, g->d_dense, g->d_dense_idx, g->Npad);
            BRR_CUDA(cudaEventRecord(e1));
            BRR_CUDA(cudaGetLastError());
            BRR_CUDA(cudaEventSynchronize(e1));
            float t = 0; BRR_CUDA(cudaEventElapsedTime(&t, e0, e1)); if (ms) *ms = t;
            BRR_CUDA(cudaMemcpy(r, d_r, g->M * 8, cudaMemcpyDeviceToHost));
        } catch (...) { cudaFree(d_eps); cudaFree(d_r); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); throw; }
        cudaFree(d_eps); cudaFree(d_r); cudaEventDestroy(e0); cudaEventDestroy(e1);
    });
}

// ---- measured fp64 pipe peak: independent DFMA chains on every SM (the roofline the workers' dot stage is reported against)
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int brr_peak_fp64(int device, double *dfma_per_s)
{
    return guarded([&] {
        BRR_REQUIRE(dfma_per_s, BRR_E_ARG, "null pointer");
        require_device(device);
        int sms = 0;
        BRR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        const int blocks = sms * 8, iters = 1 << 14;
        double *d = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
        try {
            BRR_CUDA(cudaMalloc(&d, (size_t)blocks * 256 * 8));
            BRR_CUDA(cudaEventCreate(&e0)); BRR_CUDA(cudaEventCreate(&e1));
            double best = 0.0;
            for (int rep = 0; rep < 4; ++rep) {
                BRR_CUDA(cudaEventRecord(e0));
                fp64_peak_kernel<<<blocks, 256>>>(d, iters, 1.0 + rep);
                BRR_CUDA(cudaEventRecord(e1));
                BRR_CUDA(cudaGetLastError());
                BRR_CUDA(cudaEventSynchronize(e1));
                float ms = 0; BRR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0) best = std::max(best, (double)blocks * 256 * 8 * iters / (ms * 1e-3));
            }
            *dfma_per_s = best;
        } catch (...) { cudaFree(d); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); throw; }
        cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
    });
}

extern "C" int brr_draws_sample(uint64_t seed, int stream, int64_t it, int64_t idx0, int64_t n, int kind, double shape, double *out);
extern "C" int brr_shuffle_host(uint64_t seed, int stream, int64_t it, int32_t *order, int64_t n)
{
    return guarded([&] {
        BRR_REQUIRE(order, BRR_E_ARG, "null pointer");
        shuffle_host(PhiloxKey{ (uint32_t)seed, (uint32_t)(seed >> 32) }, stream, it, order, n);
    });
}

__global__ void draws_sample_kernel(PhiloxKey key, int stream, int64_t it, int64_t idx0, int64_t n, int kind, double shape, double *out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = kind == 0 ? draw_uniform(key, stream, it, idx0 + i)
           : kind == 1 ? draw_normal(key, stream, it, idx0 + i)
                       : draw_gamma(key, stream, it, idx0 + i, shape);
}

extern "C" int brr_draws_sample(uint64_t seed, int stream, int64_t it, int64_t idx0, int64_t n, int kind, double shape, double *out)
{
    return guarded([&] {
        BRR_REQUIRE(out && n > 0, BRR_E_ARG, "bad arguments");
        require_device(0);
        double *d = nullptr;
        try {
            BRR_CUDA(cudaMalloc(&d, n * 8));
            draws_sample_kernel<<<(unsigned)((n + 127) / 128), 128>>>(PhiloxKey{ (uint32_t)seed, (uint32_t)(seed >> 32) }, stream, it, idx0, n, kind, shape, d);
            BRR_CUDA(cudaGetLastError());
            BRR_CUDA(cudaMemcpy(out, d, n * 8, cudaMemcpyDeviceToHost));
        } catch (...) { cudaFree(d); throw; }
        cudaFree(d);
    });
}
