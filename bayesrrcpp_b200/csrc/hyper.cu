// Variance-component, mixture-proportion and shrinkage draws that follow each sweep, on the device with
// counter-based Philox draws (or replayed variates).  Reference sites:
//   BayesRSamplerV2        src/BayesRv2.cpp:247-255          (m0, sigmaG, sigmaE, pi)
//   BayesRSamplerV2Groups  src/BayesRv2Groups.cpp:301-312    (sigmaF, sigmaE, per-group sigmaG + pi, interleaved)
//   BRV2Grstart            src/BRv2Grstart.cpp:254-262
//   HorseshoeR             src/HorseshoeR.cpp:217-218,242-253 (eta, nu, lambda, tau, c2, sigmaE)
//   intercept              src/BayesRv2.cpp:177-179          (drawn here for the NEXT iteration; applied as a shift)
//   distributions          src/distributions.cpp:12-39       (parameterisations restated inline)
#include "hyper.cuh"

namespace brr {

namespace {

__device__ double block_sum(double v, double *scratch)   // deterministic: fixed tree
{
    const int tid = threadIdx.x;
    scratch[tid] = v;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (tid < s) scratch[tid] += scratch[tid + s];
        __syncthreads();
    }
    const double r = scratch[0];
    __syncthreads();
    return r;
}
__device__ double strided_sum_sq(const double *x, int64_t n, double *scratch)
{
    // eight loads in flight per thread: with one, the 400 KB of beta at M = 50,000 cost this single CTA 200 dependent L2 round trips
    // per thread (~55 of the kernel's 58 us, 1.7 % of a config-2 step)
    double s[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    const int64_t st = blockDim.x;
    int64_t i = threadIdx.x;
    for (; i + 7 * st < n; i += 8 * st) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = x[i + u * st];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = fma(v[u], v[u], s[u]);
    }
    for (; i < n; i += st) s[0] = fma(x[i], x[i], s[0]);
    return block_sum(((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])), scratch);
}
// 1 / R::rgamma(shape, 1/rate)  (distributions.cpp:27-32; inv_gamma_rng(shape, scale) has the same form, :21-23)
__device__ __forceinline__ double inv_gamma(double g_unit, double rate) { return 1.0 / ((1.0 / rate) * g_unit); }
// inv_scaled_chisq_rng(dof, scale) = inv_gamma_rng(dof/2, dof*scale/2)  (distributions.cpp:34-36)
__device__ __forceinline__ double inv_scaled_chisq(double g_unit, double dof, double scale) { return inv_gamma(g_unit, 0.5 * dof * scale); }

__device__ __forceinline__ double gam_draw(const HyperParams &h, int64_t slot, double shape)
{
    return h.tbl_gam ? h.tbl_gam[slot] : draw_gamma(h.key, S_GAMMA, h.it, slot, shape);
}
// sum in the order of the oracle's / Eigen's 4-lane reduction so that pi matches to the last bits where possible
__device__ double sum4(const double *x, int n)
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0; int i = 0;
    for (; i + 4 <= n; i += 4) { s0 += x[i]; s1 += x[i + 1]; s2 += x[i + 2]; s3 += x[i + 3]; }
    double s = (s0 + s2) + (s1 + s3);
    for (; i < n; ++i) s += x[i];
    return s;
}

__device__ void next_intercept(const HyperParams &h, double eps_sum, double sigmaE)
{
    IterScalars *sc = h.sc;
    const double n = h.n_total, mu = sc->mu;
    const double z = h.tbl_mu_z_next ? *h.tbl_mu_z_next : draw_normal(h.key, S_MU, h.it + 1, 0);
    // eps += mu; mu' = N(sum(eps)/N, sigmaE/N); eps -= mu'     (reference src/BayesRv2.cpp:177-179)
    const double mu_next = (eps_sum + n * mu) / n + sqrt(sigmaE / n) * z;
    sc->mu_next = mu_next;
    sc->shift = mu - mu_next;
    sc->eps_sum = eps_sum + n * (mu - mu_next);
}

__global__ void __launch_bounds__(256) hyper_mixture_kernel(const HyperParams h)
{
    __shared__ double scratch[256];
    __shared__ double s_g[KMAX + 1];
    const int tid = threadIdx.x;
    IterScalars *sc = h.sc;
    double a = 0.0, b = 0.0;
    for (int w = tid; w < h.nW; w += blockDim.x) { a += h.fin[2 * w]; b += h.fin[2 * w + 1]; }
    const double eps_sum = block_sum(a, scratch), eps_sq = block_sum(b, scratch);
    const double beta_sq = h.kind == BRR_V2 ? strided_sum_sq(h.beta, h.M, scratch) : 0.0;     // (the Groups samplers use the per-group sums of the sweep)
    const double alpha_sq = h.F > 0 ? strided_sum_sq(h.alpha, h.F, scratch) : 0.0;
    const int K = h.K, G = h.G;
    const double N = h.n_total;
    if (h.kind == BRR_V2) {
        if (tid == 0) {
            const int m0 = (int)((double)h.M - h.vcount[0]);                                         // :247
            const double dofG = h.v0G + m0;
            sc->beta_sq = beta_sq;
            h.sigmaG[0] = inv_scaled_chisq(gam_draw(h, 0, 0.5 * dofG), dofG, (beta_sq * m0 + h.v0G * h.s02G) / dofG);   // :248 (Q6)
        }
        if (tid < K) s_g[tid] = 1.0 * gam_draw(h, 2 + tid, h.vcount[tid] + 1.0);                     // :255, distributions.cpp:12-20
        __syncthreads();
        if (tid == 0) { const double s = sum4(s_g, K); for (int k = 0; k < K; ++k) h.pi[k] = s_g[k] / s; }
    } else {
        // per-group draws are independent given the counts: one thread per group
        for (int g = tid; g < G; g += blockDim.x) {
            const double *v = h.vcount + (size_t)g * K;
            double vs[KMAX];
            for (int k = 0; k < K; ++k) vs[k] = v[k];
            const int m0 = (int)(sum4(vs, K) - v[0]);                                                // Groups:308
            const double dofG = h.v0G + m0;
            const int64_t slot0 = 2 + (int64_t)g * (K + 1);
            h.sigmaG[g] = inv_scaled_chisq(gam_draw(h, slot0, 0.5 * dofG), dofG, (h.betaAcum[g] * m0 + h.v0G * h.s02G) / dofG);   // :309
            double gg[KMAX];
            for (int k = 0; k < K; ++k) gg[k] = 1.0 * gam_draw(h, slot0 + 1 + k, v[k] + 1.0);        // :310
            const double s = sum4(gg, K);
            for (int k = 0; k < K; ++k) h.pi[(size_t)g * K + k] = gg[k] / s;
        }
        if (tid == 0 && h.kind == BRR_GROUPS) {
            const double dofF = h.v0E + (double)h.F;
            sc->sigmaF = inv_scaled_chisq(gam_draw(h, 0, 0.5 * dofF), dofF, (alpha_sq + h.v0E * h.s02E) / dofF);   // :301 (Q8)
        }
    }
    if (tid == 0) {
        const double dofE = h.v0E + N;
        const double sigmaE = inv_scaled_chisq(gam_draw(h, 1, 0.5 * dofE), dofE, (eps_sq + h.v0E * h.s02E) / dofE);    // :251
        sc->sigmaE = sigmaE; sc->eps_sq = eps_sq; sc->it_done = h.it + 1;
        next_intercept(h, eps_sum, sigmaE);
    }
}

// Horseshoe, element-wise part: lambda_j of this iteration, then nu_j of the next one (it depends on lambda_j only)
__global__ void __launch_bounds__(256) hs_local_kernel(const HyperParams h)
{
    __shared__ double scratch[256];
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double t1 = 0.0, t2 = 0.0;
    if (j < h.M) {
        const double tau = h.sc->tau, b = h.beta[j];
        const double shape = 0.5 + 0.5 * h.vL;
        const double gl = h.tbl_hs_lam ? h.tbl_hs_lam[j] : draw_gamma(h.key, S_HS_LAM, h.it, j, shape);
        const double lam = inv_gamma(gl, h.vL * (1.0 / h.nu[j]) + (0.5 * (b * b)) * (1.0 / tau));    // HorseshoeR.cpp:242
        h.lambda[j] = lam;
        const double gn = h.tbl_hs_nu_next ? h.tbl_hs_nu_next[j] : draw_gamma(h.key, S_HS_NU, h.it + 1, j, shape);
        h.nu[j] = inv_gamma(gn, h.vL / lam + 1.0);                                                   // :218 (next iteration)
        t1 = (b * b) / lam; t2 = b * b;
    }
    const double s1 = block_sum(t1, scratch), s2 = block_sum(t2, scratch);
    if (threadIdx.x == 0) { h.hs_part[2 * blockIdx.x] = s1; h.hs_part[2 * blockIdx.x + 1] = s2; }
}

__global__ void __launch_bounds__(256) hs_global_kernel(const HyperParams h, int nparts)
{
    __shared__ double scratch[256];
    const int tid = threadIdx.x;
    IterScalars *sc = h.sc;
    double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
    for (int w = tid; w < h.nW; w += blockDim.x) { a += h.fin[2 * w]; b += h.fin[2 * w + 1]; }
    for (int i = tid; i < nparts; i += blockDim.x) { c += h.hs_part[2 * i]; d += h.hs_part[2 * i + 1]; }
    const double eps_sum = block_sum(a, scratch), eps_sq = block_sum(b, scratch);
    const double bl = block_sum(c, scratch), bsq = block_sum(d, scratch);
    if (tid == 0) {
        const double M = (double)h.M, N = h.n_total;
        const double eta = sc->eta_next;
        sc->eta = eta;
        const double tau = inv_gamma(gam_draw(h, 1, 0.5 * (M + h.vT)), h.vT / eta + ((0.5) * bl));               // :245
        const double c2 = inv_gamma(gam_draw(h, 2, 0.5 * h.vC + 0.5 * M), h.vC * h.sC * 0.5 + 0.5 * bsq);        // :248
        const double dofE = h.v0E + N;
        const double sigmaE = inv_scaled_chisq(gam_draw(h, 3, 0.5 * dofE), dofE, (eps_sq + h.v0E * h.s02E) / dofE);   // :253
        sc->tau = tau; sc->c2 = c2; sc->sigmaE = sigmaE; sc->eps_sq = eps_sq; sc->beta_sq = bsq; sc->it_done = h.it + 1;
        next_intercept(h, eps_sum, sigmaE);
        // eta of the next iteration (:217) uses the new sigmaE and tau
        const double ge = h.tbl_gam_next ? h.tbl_gam_next[0] : draw_gamma(h.key, S_GAMMA, h.it + 1, 0, 0.5 + 0.5 * h.vT);
        sc->eta_next = inv_gamma(ge, (1.0 / (sigmaE * h.A * h.A)) + h.vT / tau);
    }
}

}  // namespace

void launch_hyper(const HyperParams &h, cudaStream_t stream)
{
    if (h.kind == BRR_HORSESHOE) {
        const int nb = (int)((h.M + 255) / 256);
        hs_local_kernel<<<nb, 256, 0, stream>>>(h);
        BRR_CUDA(cudaGetLastError());
        hs_global_kernel<<<1, 256, 0, stream>>>(h, nb);
    } else {
        hyper_mixture_kernel<<<1, 256, 0, stream>>>(h);
    }
    BRR_CUDA(cudaGetLastError());
}
void preload_hyper(int kind)
{
    if (kind == BRR_HORSESHOE) { preload_kernel(hs_local_kernel); preload_kernel(hs_global_kernel); }
    else preload_kernel(hyper_mixture_kernel);
}
int hyper_launch_count(int kind) { return kind == BRR_HORSESHOE ? 2 : 1; }

}  // namespace brr
