// Counter-based draws (Philox4x32-10) shared by host and device code of the sampler.
// Replaces the reference's R::rgamma / R::rnorm / R::rbeta / R::runif wrappers
// (reference src/distributions.cpp:12-39,60-65) with draws that are a pure function of
// (seed, stream, iteration, index): every rank of a row-sharded chain, and the CPU oracle in its
// Philox mode, obtain the same value without communicating.
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define BRR_HD __host__ __device__ __forceinline__
#else
#define BRR_HD inline
#endif

namespace brr {

// stream ids (counter word c2, low byte).  Numbering is part of the replay contract.
enum : int {
    S_INIT_U = 0, S_MU = 1, S_MARK_U = 2, S_MARK_Z = 3, S_GAMMA = 4, S_PERM = 5,
    S_FIX_Z = 6, S_FIXPERM = 7, S_HS_NU = 8, S_HS_LAM = 9, S_INIT_G = 10
};

struct PhiloxKey { uint32_t k0, k1; };

BRR_HD void philox4x32_10(PhiloxKey key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__)
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// counter layout: c0 = idx low, c1 = it + 1 (0 = before the first iteration), c2 = stream | sub << 8, c3 = idx high
BRR_HD void draw_words(PhiloxKey key, int stream, int sub, int64_t it, int64_t idx, uint32_t w[4])
{
    philox4x32_10(key, (uint32_t)((uint64_t)idx & 0xffffffffu), (uint32_t)(it + 1),
                  (uint32_t)stream | ((uint32_t)sub << 8), (uint32_t)((uint64_t)idx >> 32), w);
}
// 52-bit uniform strictly inside (0, 1)
BRR_HD double u52(uint32_t a, uint32_t b)
{
    const uint64_t v = ((uint64_t)(a >> 6) << 26) | (uint64_t)(b >> 6);
    return ((double)v + 0.5) * (1.0 / 4503599627370496.0);
}
BRR_HD double box_muller(const uint32_t w[4])
{
    const double u1 = u52(w[0], w[1]), u2 = u52(w[2], w[3]);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}
BRR_HD double draw_uniform(PhiloxKey key, int stream, int64_t it, int64_t idx)
{
    uint32_t w[4]; draw_words(key, stream, 0, it, idx, w); return u52(w[0], w[1]);
}
BRR_HD double draw_normal(PhiloxKey key, int stream, int64_t it, int64_t idx)
{
    uint32_t w[4]; draw_words(key, stream, 0, it, idx, w); return box_muller(w);
}
// Unit-scale Gamma(shape): Marsaglia & Tsang (2000), attempts indexed by the counter (sub = 1+2t normal,
// 2+2t acceptance uniform), bounded at 64; shape < 1 boosted with the sub-0 uniform.  Fixed map from
// (key, stream, it, idx, shape) to the variate: replicas stay in lock-step.
BRR_HD double draw_gamma(PhiloxKey key, int stream, int64_t it, int64_t idx, double shape)
{
    const double a = shape < 1.0 ? shape + 1.0 : shape;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double res = d;
    uint32_t w[4];
    for (int t = 0; t < 64; ++t) {
        draw_words(key, stream, 1 + 2 * t, it, idx, w);
        const double z = box_muller(w);
        double v = 1.0 + c * z;
        if (v <= 0.0) continue;
        v = v * v * v;
        draw_words(key, stream, 2 + 2 * t, it, idx, w);
        const double u = u52(w[0], w[1]);
        if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) { res = d * v; break; }
    }
    if (shape < 1.0) {
        draw_words(key, stream, 0, it, idx, w);
        res *= pow(u52(w[0], w[1]), 1.0 / shape);
    }
    return res;
}

// In-place shuffle in std::random_shuffle's form (reference src/BayesRv2.cpp:182; libstdc++ swaps a[i] with
// a[r % (i+1)] for i = 1..n-1) with r taken from the Philox stream.  Host only: O(M), overlapped with the GPU sweep.
inline void shuffle_host(PhiloxKey key, int stream, int64_t it, int32_t *order, int64_t n)
{
    for (int64_t i = 1; i < n; ++i) {
        uint32_t w[4]; draw_words(key, stream, 0, it, i, w);
        const uint64_t r = ((uint64_t)w[1] << 32) | w[0];
        const int64_t j = (int64_t)(r % (uint64_t)(i + 1));
        const int32_t t = order[i]; order[i] = order[j]; order[j] = t;
    }
}

}  // namespace brr
