/*
 * bayesrr_oracle.c -- TEST INFRASTRUCTURE (see bayesrr_oracle.h).  CPU restatement of the
 * reference's four samplers, one literal pass structure per reference line; file:line
 * citations are relative to /root/reference.
 */
#include "bayesrr_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* =====================================================================================
 * Philox4x32-10 (Salmon et al., SC'11 -- Random123).  Counter layout used everywhere:
 *   c0 = low 32 bits of idx, c1 = it+1 (0 for pre-iteration draws),
 *   c2 = stream | (sub << 8)   (sub: 0 = primary, 1.. = rejection-sampler attempts),
 *   c3 = high 32 bits of idx;   key = 64-bit seed.
 * ===================================================================================== */
void orc_philox_raw(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox_init(orc_philox *p, uint64_t seed)
{
    p->key[0] = (uint32_t)seed; p->key[1] = (uint32_t)(seed >> 32);
}

static void px_words(const orc_philox *p, int stream, int sub, int64_t it, int64_t idx, uint32_t w[4])
{
    uint32_t c[4] = { (uint32_t)((uint64_t)idx & 0xffffffffu), (uint32_t)(it + 1),
                      (uint32_t)stream | ((uint32_t)sub << 8), (uint32_t)((uint64_t)idx >> 32) };
    orc_philox_raw(p->key, c, w);
}
/* 52-bit uniform strictly inside (0,1) */
static double u52(uint32_t a, uint32_t b)
{
    uint64_t v = ((uint64_t)(a >> 6) << 26) | (uint64_t)(b >> 6);
    return ((double)v + 0.5) * (1.0 / 4503599627370496.0);
}
static double bm_normal(const uint32_t w[4])
{
    double u1 = u52(w[0], w[1]), u2 = u52(w[2], w[3]);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}
static double px_uniform(void *ctx, int stream, int64_t it, int64_t idx)
{
    uint32_t w[4]; px_words((orc_philox *)ctx, stream, 0, it, idx, w); return u52(w[0], w[1]);
}
static double px_normal(void *ctx, int stream, int64_t it, int64_t idx)
{
    uint32_t w[4]; px_words((orc_philox *)ctx, stream, 0, it, idx, w); return bm_normal(w);
}
/* Marsaglia & Tsang (2000) with counter-indexed attempts: attempt t uses sub = 1+2t (normal) and
 * 2+2t (acceptance uniform); shape < 1 is boosted with the sub-0 uniform.  Bounded at 64 attempts. */
static double px_gamma(void *ctx, int stream, int64_t it, int64_t idx, double shape)
{
    const orc_philox *p = (const orc_philox *)ctx;
    double a = shape < 1.0 ? shape + 1.0 : shape;
    double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d), res = d;
    for (int t = 0; t < 64; ++t) {
        uint32_t w[4];
        px_words(p, stream, 1 + 2 * t, it, idx, w);
        double z = bm_normal(w);
        double v = 1.0 + c * z;
        if (v <= 0.0) continue;
        v = v * v * v;
        px_words(p, stream, 2 + 2 * t, it, idx, w);
        double u = u52(w[0], w[1]);
        if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) { res = d * v; break; }
    }
    if (shape < 1.0) {
        uint32_t w[4]; px_words(p, stream, 0, it, idx, w);
        res *= pow(u52(w[0], w[1]), 1.0 / shape);
    }
    return res;
}
/* keyed Fisher-Yates in std::random_shuffle's form (libstdc++ stl_algo.h: for i=1..n-1 swap(a[i], a[r%(i+1)])) */
static void px_shuffle(void *ctx, int stream, int64_t it, int32_t *order, int64_t n)
{
    const orc_philox *p = (const orc_philox *)ctx;
    for (int64_t i = 1; i < n; ++i) {
        uint32_t w[4]; px_words(p, stream, 0, it, i, w);
        uint64_t r = ((uint64_t)w[1] << 32) | w[0];
        int64_t j = (int64_t)(r % (uint64_t)(i + 1));
        int32_t t = order[i]; order[i] = order[j]; order[j] = t;
    }
}
orc_draws orc_philox_source(orc_philox *p)
{
    orc_draws d = { p, px_uniform, px_normal, px_gamma, px_shuffle };
    return d;
}

/* ---- sequential source */
void orc_seq_init(orc_seq *s, uint64_t seed)
{
    orc_philox_init(&s->px, seed); s->n_u = s->n_z = s->n_g = 0;
    srand((unsigned)seed);
}
double orc_seq_next_uniform(orc_seq *s) { return px_uniform(&s->px, 100, -1, (int64_t)s->n_u++); }
double orc_seq_next_normal(orc_seq *s)  { return px_normal(&s->px, 101, -1, (int64_t)s->n_z++); }
double orc_seq_next_gamma(orc_seq *s, double shape) { return px_gamma(&s->px, 102, -1, (int64_t)s->n_g++, shape); }
static double sq_uniform(void *c, int st, int64_t it, int64_t idx) { (void)st; (void)it; (void)idx; return orc_seq_next_uniform((orc_seq *)c); }
static double sq_normal(void *c, int st, int64_t it, int64_t idx)  { (void)st; (void)it; (void)idx; return orc_seq_next_normal((orc_seq *)c); }
static double sq_gamma(void *c, int st, int64_t it, int64_t idx, double a) { (void)st; (void)it; (void)idx; return orc_seq_next_gamma((orc_seq *)c, a); }
static void sq_shuffle(void *c, int st, int64_t it, int32_t *order, int64_t n)
{
    (void)c; (void)st; (void)it;
    for (int64_t i = 1; i < n; ++i) {           /* libstdc++ std::random_shuffle, stl_algo.h */
        int64_t j = rand() % (i + 1);
        if (i != j) { int32_t t = order[i]; order[i] = order[j]; order[j] = t; }
    }
}
orc_draws orc_seq_source(orc_seq *s)
{
    orc_draws d = { s, sq_uniform, sq_normal, sq_gamma, sq_shuffle };
    return d;
}

/* ---- recorder / replay */
static double *tbl_slot(orc_tables *t, int stream, int64_t it, int64_t idx)
{
    switch (stream) {
    case ORC_S_INIT_U:  return (t->init_u && idx < t->n_init_u) ? &t->init_u[idx] : NULL;
    case ORC_S_INIT_G:  return (t->init_g && idx < t->n_init_g) ? &t->init_g[idx] : NULL;
    case ORC_S_MU:      return t->mu_z ? &t->mu_z[it] : NULL;
    case ORC_S_MARK_U:  return t->mark_u ? &t->mark_u[it * t->M + idx] : NULL;
    case ORC_S_MARK_Z:  return t->mark_z ? &t->mark_z[it * t->M + idx] : NULL;
    case ORC_S_GAMMA:   return (t->gam && idx < t->n_gam) ? &t->gam[it * t->n_gam + idx] : NULL;
    case ORC_S_FIX_Z:   return t->fix_z ? &t->fix_z[it * t->F + idx] : NULL;
    case ORC_S_HS_NU:   return t->hs_nu ? &t->hs_nu[it * t->M + idx] : NULL;
    case ORC_S_HS_LAM:  return t->hs_lam ? &t->hs_lam[it * t->M + idx] : NULL;
    default: return NULL;
    }
}
static double rec_uniform(void *c, int st, int64_t it, int64_t idx)
{
    orc_recorder *r = (orc_recorder *)c; double v = r->inner.uniform(r->inner.ctx, st, it, idx);
    double *p = tbl_slot(r->t, st, it, idx); if (p) *p = v; return v;
}
static double rec_normal(void *c, int st, int64_t it, int64_t idx)
{
    orc_recorder *r = (orc_recorder *)c; double v = r->inner.normal(r->inner.ctx, st, it, idx);
    double *p = tbl_slot(r->t, st, it, idx); if (p) *p = v; return v;
}
static double rec_gamma(void *c, int st, int64_t it, int64_t idx, double a)
{
    orc_recorder *r = (orc_recorder *)c; double v = r->inner.gamma(r->inner.ctx, st, it, idx, a);
    double *p = tbl_slot(r->t, st, it, idx); if (p) *p = v; return v;
}
static void rec_shuffle(void *c, int st, int64_t it, int32_t *order, int64_t n)
{
    orc_recorder *r = (orc_recorder *)c; r->inner.shuffle(r->inner.ctx, st, it, order, n);
    if (st == ORC_S_PERM && r->t->perm) memcpy(&r->t->perm[it * r->t->M], order, (size_t)n * sizeof(int32_t));
    if (st == ORC_S_FIXPERM && r->t->fixperm) memcpy(&r->t->fixperm[it * r->t->F], order, (size_t)n * sizeof(int32_t));
}
orc_draws orc_record_source(orc_recorder *r)
{
    orc_draws d = { r, rec_uniform, rec_normal, rec_gamma, rec_shuffle };
    return d;
}
static double rep_get(void *c, int st, int64_t it, int64_t idx)
{
    double *p = tbl_slot((orc_tables *)c, st, it, idx); return p ? *p : NAN;
}
static double rep_gamma(void *c, int st, int64_t it, int64_t idx, double a) { (void)a; return rep_get(c, st, it, idx); }
static void rep_shuffle(void *c, int st, int64_t it, int32_t *order, int64_t n)
{
    orc_tables *t = (orc_tables *)c;
    if (st == ORC_S_PERM) memcpy(order, &t->perm[it * t->M], (size_t)n * sizeof(int32_t));
    else if (st == ORC_S_FIXPERM) memcpy(order, &t->fixperm[it * t->F], (size_t)n * sizeof(int32_t));
}
orc_draws orc_replay_source(orc_tables *t)
{
    orc_draws d = { t, rep_get, rep_get, rep_gamma, rep_shuffle };
    return d;
}

/* =====================================================================================
 * Vector helpers.  Eigen reduces with two SSE2 packets of two doubles (redux, linear vectorised
 * traversal): four running partial sums combined as (s0+s2)+(s1+s3).  Mimicked here so that the
 * CPU-baseline timing has Eigen's instruction-level parallelism; not claimed bit-identical.
 * ===================================================================================== */
static double vsum(const double *x, int64_t n)
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0; int64_t i = 0;
    for (; i + 4 <= n; i += 4) { s0 += x[i]; s1 += x[i + 1]; s2 += x[i + 2]; s3 += x[i + 3]; }
    double s = (s0 + s2) + (s1 + s3);
    for (; i < n; ++i) s += x[i];
    return s;
}
static double vdot(const double *x, const double *y, int64_t n)
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0; int64_t i = 0;
    for (; i + 4 <= n; i += 4) { s0 += x[i] * y[i]; s1 += x[i + 1] * y[i + 1]; s2 += x[i + 2] * y[i + 2]; s3 += x[i + 3] * y[i + 3]; }
    double s = (s0 + s2) + (s1 + s3);
    for (; i < n; ++i) s += x[i] * y[i];
    return s;
}
static double vsqnorm(const double *x, int64_t n) { return vdot(x, x, n); }

/* ---- src/distributions.cpp */
/* :21-23  inv_gamma_rng(shape, scale) = 1 / R::rgamma(shape, 1/scale);  R::rgamma(a, s) = s * g(a) */
static double inv_gamma_rng(const orc_draws *d, int st, int64_t it, int64_t idx, double shape, double scale)
{
    return 1.0 / ((1.0 / scale) * d->gamma(d->ctx, st, it, idx, shape));
}
/* :27-32  inv_gamma_rate_rng(shape, rate) = 1 / R::rgamma(shape, 1/rate) */
static double inv_gamma_rate_rng(const orc_draws *d, int st, int64_t it, int64_t idx, double shape, double rate)
{
    return 1.0 / ((1.0 / rate) * d->gamma(d->ctx, st, it, idx, shape));
}
/* :34-36 */
static double inv_scaled_chisq_rng(const orc_draws *d, int st, int64_t it, int64_t idx, double dof, double scale)
{
    return inv_gamma_rng(d, st, it, idx, 0.5 * dof, 0.5 * dof * scale);
}
/* :37-39  second argument is a VARIANCE; R::rnorm(m, s) = m + s * z */
static double norm_rng(const orc_draws *d, int st, int64_t it, int64_t idx, double mean, double sigma2)
{
    return mean + sqrt(sigma2) * d->normal(d->ctx, st, it, idx);
}
/* :12-20  dirichilet_rng: g_i = R::rgamma(alpha_i, 1) in index order, normalised by their sum */
static void dirichlet_rng(const orc_draws *d, int64_t it, int64_t slot0, const double *alpha, int K, double *out)
{
    for (int i = 0; i < K; ++i) out[i] = 1.0 * d->gamma(d->ctx, ORC_S_GAMMA, it, slot0 + i, alpha[i]);
    double s = vsum(out, K);
    for (int i = 0; i < K; ++i) out[i] /= s;
}
static void dirichlet_rng_st(const orc_draws *d, int st, int64_t it, int64_t slot0, const double *alpha, int K, double *out)
{
    for (int i = 0; i < K; ++i) out[i] = 1.0 * d->gamma(d->ctx, st, it, slot0 + i, alpha[i]);
    double s = vsum(out, K);
    for (int i = 0; i < K; ++i) out[i] /= s;
}

static int check_iters(int max_iterations, int burn_in, int thinning)
{
    /* src/BayesRv2.cpp:76-80 (same in all four) -- the only validation that returns */
    if (max_iterations < burn_in || max_iterations < 1 || burn_in < 1) return ORC_ERR_ITER;
    if (thinning < 1) return ORC_ERR_ARG;   /* the reference would divide by zero (:259) */
    return ORC_OK;
}
static int keep(int it, int burn_in, int thinning, int emit_all)
{
    return emit_all || (it >= burn_in && it % thinning == 0);   /* :257-259 */
}

/* The mixture step shared by V2 / Groups / Grstart: src/BayesRv2.cpp:195-242 (Groups:247-294,
 * Grstart:198-246).  Returns the chosen component or -1 when the CDF walk falls through (Q5).
 * muk/denom are outputs used by the caller for the beta draw.                                 */
static int mixture_pick(int K, const double *logpi_src, const double *cVa, const double *cVaI,
                        double xsq, double num, double sigmaE, double sigmaG, double p,
                        double *logL, double *muk, double *denom, double *tmp)
{
    muk[0] = 0.0;                                                           /* :195 */
    for (int k = 1; k < K; ++k) denom[k - 1] = xsq + (sigmaE / sigmaG) * cVaI[k];   /* :199 */
    for (int k = 1; k < K; ++k) muk[k] = num / denom[k - 1];                /* :203 */
    for (int k = 0; k < K; ++k) logL[k] = log(logpi_src[k]);                /* :207 */
    for (int k = 1; k < K; ++k)                                             /* :211 */
        logL[k] = logL[k] - 0.5 * log(((sigmaG / sigmaE) * xsq) * cVa[k] + 1.0) + 0.5 * (muk[k] * num) / sigmaE;
    double acum;
    int big = 0;
    for (int k = 1; k < K; ++k) if (fabs(logL[k] - logL[0]) > 700) big = 1; /* :216 */
    if (big) acum = 0;
    else { for (int l = 0; l < K; ++l) tmp[l] = exp(logL[l] - logL[0]); acum = 1.0 / vsum(tmp, K); }  /* :219 */
    for (int k = 0; k < K; ++k) {                                           /* :222 */
        if (p <= acum) return k;
        if (k + 1 < K) {      /* the reference reads logL[K] here when k = K-1; the value is never used (Q4) */
            big = 0;
            for (int l = 1; l < K; ++l) if (fabs(logL[l] - logL[k + 1]) > 700) big = 1;     /* :235 */
            if (big) acum += 0;
            else { for (int l = 0; l < K; ++l) tmp[l] = exp(logL[l] - logL[k + 1]); acum += 1.0 / vsum(tmp, K); }  /* :239 */
        }
    }
    return -1;
}

/* =====================================================================================
 * BayesRSamplerV2 -- src/BayesRv2.cpp:60-294
 * ===================================================================================== */
int orc_v2_run(const orc_v2_args *a, const orc_draws *d, orc_row_fn sink, void *sctx)
{
    int rc = check_iters(a->max_iterations, a->burn_in, a->thinning);
    if (rc) return rc;
    const int64_t N = a->N, M = a->M; const int K = a->ncva + 1;                    /* :64-65,73 */
    const int64_t L = orc_v2_rowlen(N, M);
    double *pi = calloc(K, 8), *cVa = calloc(K, 8), *cVaI = calloc(K, 8), *logL = calloc(K + 1, 8);
    double *muk = calloc(K, 8), *denom = calloc(K, 8), *v = calloc(K, 8), *tmp = calloc(K, 8), *alpha = calloc(K, 8);
    double *beta = calloc(M, 8), *yt = calloc(N, 8), *eps = calloc(N, 8), *xsq = calloc(M, 8);
    double *comp = calloc(M, 8), *row = calloc(L, 8);
    int32_t *markerI = malloc(M * sizeof(int32_t));
    for (int64_t i = 0; i < M; ++i) markerI[i] = (int32_t)i;                        /* :137-140 */

    cVa[0] = 0; for (int k = 1; k < K; ++k) cVa[k] = a->cva[k - 1];                 /* :152-153 */
    cVaI[0] = 0; for (int k = 1; k < K; ++k) cVaI[k] = 1.0 / cVa[k];                /* :155-156 */
    double mu = 0;                                                                  /* :158,160 (beta already 0) */
    double sigmaG = d->uniform(d->ctx, ORC_S_INIT_U, -1, 0);                        /* :162 beta_rng(1,1) */
    for (int k = 0; k < K; ++k) pi[k] = a->pi_init[k];                              /* :150,164 (Q1: explicit) */
    /* :168  epsilon = Y - mu - X*beta  (beta == 0) */
    for (int64_t i = 0; i < N; ++i) eps[i] = a->Y[i] - mu - 0.0;
    double sigmaE = vsqnorm(eps, N) / N * 0.5;                                      /* :169 */
    for (int64_t j = 0; j < M; ++j) xsq[j] = vsqnorm(a->X + j * N, N);              /* :170 */

    for (int it = 0; it < a->max_iterations; ++it) {                                /* :171 */
        for (int64_t i = 0; i < N; ++i) eps[i] = eps[i] + mu;                       /* :177 */
        mu = norm_rng(d, ORC_S_MU, it, 0, vsum(eps, N) / (double)N, sigmaE / (double)N);   /* :178 */
        for (int64_t i = 0; i < N; ++i) eps[i] = eps[i] - mu;                       /* :179 */
        d->shuffle(d->ctx, ORC_S_PERM, it, markerI, M);                             /* :182 */
        for (int k = 0; k < K; ++k) v[k] = 0;                                       /* :185 */
        for (int64_t j = 0; j < M; ++j) {                                           /* :186 */
            const int64_t marker = markerI[j];
            const double *x = a->X + marker * N;
            const double bo = beta[marker];
            for (int64_t i = 0; i < N; ++i) yt[i] = eps[i] + x[i] * bo;             /* :191 */
            double num = vdot(x, yt, N);                                            /* :201 */
            double p = d->uniform(d->ctx, ORC_S_MARK_U, it, j);                     /* :213 */
            int k = mixture_pick(K, pi, cVa, cVaI, xsq[marker], num, sigmaE, sigmaG, p, logL, muk, denom, tmp);
            if (k == 0) beta[marker] = 0;                                           /* :226 */
            else if (k > 0) beta[marker] = norm_rng(d, ORC_S_MARK_Z, it, j, muk[k], sigmaE / denom[k - 1]);  /* :228 */
            if (k >= 0) { v[k] += 1.0; comp[marker] = k; }                          /* :230-231 */
            const double bn = beta[marker];
            for (int64_t i = 0; i < N; ++i) eps[i] = yt[i] - x[i] * bn;             /* :243 */
        }
        int m0 = (int)(M - v[0]);                                                   /* :247 */
        sigmaG = inv_scaled_chisq_rng(d, ORC_S_GAMMA, it, 0, a->v0G + m0,
                                      (vsqnorm(beta, M) * m0 + a->v0G * a->s02G) / (a->v0G + m0));  /* :248 (Q6) */
        sigmaE = inv_scaled_chisq_rng(d, ORC_S_GAMMA, it, 1, a->v0E + N,
                                      (vsqnorm(eps, N) + a->v0E * a->s02E) / (a->v0E + N));         /* :251 */
        for (int k = 0; k < K; ++k) alpha[k] = v[k] + 1.0;
        dirichlet_rng(d, it, 2, alpha, K, pi);                                      /* :255 */
        if (a->pi_trace) memcpy(a->pi_trace + (int64_t)it * K, pi, K * 8);
        if (keep(it, a->burn_in, a->thinning, a->emit_all)) {                       /* :257-261 */
            int64_t o = 0; row[o++] = it; row[o++] = mu;
            memcpy(row + o, beta, M * 8); o += M;
            row[o++] = sigmaE; row[o++] = sigmaG;
            memcpy(row + o, comp, M * 8); o += M;
            memcpy(row + o, eps, N * 8); o += N;
            if (sink) sink(sctx, row, L);
        }
    }
    free(pi); free(cVa); free(cVaI); free(logL); free(muk); free(denom); free(v); free(tmp); free(alpha);
    free(beta); free(yt); free(eps); free(xsq); free(comp); free(row); free(markerI);
    return ORC_OK;
}

/* shared sweep of Groups / Grstart: src/BayesRv2Groups.cpp:232-298 == src/BRv2Grstart.cpp:183-250 */
typedef struct {
    int64_t N, M; int K, G; const double *X; const double *cva; const int32_t *gAssign;
    double *beta, *eps, *yt, *comp, *xsq, *v, *betaAcum, *pi, *sigmaGG;
    double *cVa, *cVaI, *logL, *muk, *denom, *tmp;
} grp_state;

static void groups_sweep(grp_state *s, const orc_draws *d, int it, const int32_t *markerI, double sigmaE)
{
    const int64_t N = s->N; const int K = s->K, G = s->G;
    for (int i = 0; i < G * K; ++i) s->v[i] = 0;                                    /* Groups:230 */
    for (int g = 0; g < G; ++g) s->betaAcum[g] = 0;                                 /* :231 */
    for (int64_t j = 0; j < s->M; ++j) {                                            /* :232 */
        const int64_t marker = markerI[j];
        const int g = s->gAssign[marker];
        const double sigmaG = s->sigmaGG[g];                                        /* :235 */
        s->cVa[0] = 0; s->cVaI[0] = 0;                                              /* :237-238 */
        for (int k = 1; k < K; ++k) { s->cVa[k] = s->cva[g + (int64_t)(k - 1) * G]; s->cVaI[k] = 1.0 / s->cVa[k]; }  /* :239-240 */
        const double *x = s->X + marker * N;
        const double bo = s->beta[marker];
        for (int64_t i = 0; i < N; ++i) s->yt[i] = s->eps[i] + x[i] * bo;           /* :243 */
        double num = vdot(x, s->yt, N);                                             /* :254 */
        double p = d->uniform(d->ctx, ORC_S_MARK_U, it, j);                         /* :266 runif / Grstart:217 rbeta(1,1) */
        int k = mixture_pick(K, s->pi + g * K, s->cVa, s->cVaI, s->xsq[marker], num, sigmaE, sigmaG, p,
                             s->logL, s->muk, s->denom, s->tmp);
        if (k == 0) s->beta[marker] = 0;                                            /* :277 */
        else if (k > 0) {
            s->beta[marker] = norm_rng(d, ORC_S_MARK_Z, it, j, s->muk[k], sigmaE / s->denom[k - 1]);  /* :279 */
            s->betaAcum[g] += pow(s->beta[marker], 2);                              /* :280 */
        }
        if (k >= 0) { s->v[g * K + k] += 1.0; s->comp[marker] = k; }                /* :283-284 */
        const double bn = s->beta[marker];
        for (int64_t i = 0; i < N; ++i) s->eps[i] = s->yt[i] - x[i] * bn;           /* :295 */
    }
}
/* per-group variance + pi draws: Groups:307-312 == Grstart:257-262 */
static void groups_hyper(grp_state *s, const orc_draws *d, int it, double v0G, double s02G, double *alpha)
{
    const int K = s->K;
    for (int g = 0; g < s->G; ++g) {
        int m0 = (int)(vsum(s->v + g * K, K) - s->v[g * K]);                        /* :308 */
        s->sigmaGG[g] = inv_scaled_chisq_rng(d, ORC_S_GAMMA, it, 2 + (int64_t)g * (K + 1), v0G + m0,
                                             (s->betaAcum[g] * m0 + v0G * s02G) / (v0G + m0));   /* :309 */
        for (int k = 0; k < K; ++k) alpha[k] = s->v[g * K + k] + 1.0;
        dirichlet_rng(d, it, 2 + (int64_t)g * (K + 1) + 1, alpha, K, s->pi + g * K);            /* :310 */
    }
}
static void grp_alloc(grp_state *s)
{
    const int K = s->K;
    s->yt = calloc(s->N, 8); s->xsq = calloc(s->M, 8); s->v = calloc((size_t)s->G * K, 8);
    s->betaAcum = calloc(s->G, 8); s->pi = calloc((size_t)s->G * K, 8);
    s->cVa = calloc(K, 8); s->cVaI = calloc(K, 8); s->logL = calloc(K + 1, 8); s->muk = calloc(K, 8);
    s->denom = calloc(K, 8); s->tmp = calloc(K, 8);
}
static void grp_free(grp_state *s)
{
    free(s->yt); free(s->xsq); free(s->v); free(s->betaAcum); free(s->pi);
    free(s->cVa); free(s->cVaI); free(s->logL); free(s->muk); free(s->denom); free(s->tmp);
}

/* =====================================================================================
 * BayesRSamplerV2Groups -- src/BayesRv2Groups.cpp:75-361
 * ===================================================================================== */
int orc_groups_run(const orc_groups_args *a, const orc_draws *d, orc_row_fn sink, void *sctx)
{
    int rc = check_iters(a->max_iterations, a->burn_in, a->thinning);
    if (rc) return rc;
    const int64_t N = a->N, M = a->M, F = a->F; const int K = a->ncva + 1, G = a->groups;
    const int64_t L = orc_groups_rowlen(N, M, G, F);
    grp_state s; memset(&s, 0, sizeof s);
    s.N = N; s.M = M; s.K = K; s.G = G; s.X = a->X; s.cva = a->cva; s.gAssign = a->gAssign;
    grp_alloc(&s);
    s.beta = calloc(M, 8); s.eps = calloc(N, 8); s.comp = calloc(M, 8); s.sigmaGG = calloc(G, 8);
    double *alphaF = calloc(F ? F : 1, 8), *row = calloc(L, 8), *dal = calloc(K, 8);
    int32_t *markerI = malloc(M * sizeof(int32_t)), *fixedI = malloc((F ? F : 1) * sizeof(int32_t));
    for (int64_t i = 0; i < M; ++i) markerI[i] = (int32_t)i;
    for (int64_t i = 0; i < F; ++i) fixedI[i] = (int32_t)i;
    for (int g = 0; g < G; ++g) {                                                   /* :170-175 */
        s.pi[g * K] = 0.5;
        for (int k = 1; k < K; ++k) s.pi[g * K + k] = 0.5 / K;
    }
    double mu = 0;                                                                  /* :188 */
    for (int g = 0; g < G; ++g) s.sigmaGG[g] = d->uniform(d->ctx, ORC_S_INIT_U, -1, g);   /* :194-195 */
    double sigmaF = d->uniform(d->ctx, ORC_S_INIT_U, -1, G);                        /* :197 */
    for (int64_t i = 0; i < N; ++i) s.eps[i] = a->Y[i] - mu;                        /* :203 */
    double sigmaE = vsqnorm(s.eps, N) / N * 0.5;                                    /* :204 */
    for (int64_t j = 0; j < M; ++j) s.xsq[j] = vsqnorm(a->X + j * N, N);            /* :205 */

    for (int it = 0; it < a->max_iterations; ++it) {
        for (int64_t i = 0; i < N; ++i) s.eps[i] = s.eps[i] + mu;                   /* :212 */
        mu = norm_rng(d, ORC_S_MU, it, 0, vsum(s.eps, N) / (double)N, sigmaE / (double)N);   /* :213 */
        for (int64_t i = 0; i < N; ++i) s.eps[i] = s.eps[i] - mu;                   /* :214 */
        d->shuffle(d->ctx, ORC_S_FIXPERM, it, fixedI, F);                           /* :216 */
        for (int64_t cf = 0; cf < F; ++cf) {                                        /* :217-225 */
            const int64_t cur = fixedI[cf];
            const double *f = a->fixed + cur * N;
            const double ca = alphaF[cur];
            for (int64_t i = 0; i < N; ++i) s.yt[i] = s.eps[i] + f[i] * ca;         /* :220 */
            double denom_f = (double)(N - 1) + (sigmaE / sigmaF);                   /* :221 (Q8) */
            double num_f = vdot(f, s.yt, N);                                        /* :222 */
            alphaF[cur] = norm_rng(d, ORC_S_FIX_Z, it, cf, num_f / denom_f, sigmaE / denom_f);   /* :223 */
            const double na = alphaF[cur];
            for (int64_t i = 0; i < N; ++i) s.eps[i] = s.yt[i] - f[i] * na;         /* :224 */
        }
        d->shuffle(d->ctx, ORC_S_PERM, it, markerI, M);                             /* :227 */
        groups_sweep(&s, d, it, markerI, sigmaE);
        sigmaF = inv_scaled_chisq_rng(d, ORC_S_GAMMA, it, 0, a->v0E + F,
                                      (vsqnorm(alphaF, F) + a->v0E * a->s02E) / (a->v0E + F));   /* :301 (Q8) */
        sigmaE = inv_scaled_chisq_rng(d, ORC_S_GAMMA, it, 1, a->v0E + N,
                                      (vsqnorm(s.eps, N) + a->v0E * a->s02E) / (a->v0E + N));     /* :304 */
        groups_hyper(&s, d, it, a->v0G, a->s02G, dal);                              /* :307-312 */
        if (a->pi_trace) memcpy(a->pi_trace + (int64_t)it * G * K, s.pi, (size_t)G * K * 8);
        if (keep(it, a->burn_in, a->thinning, a->emit_all)) {                       /* :314-317 */
            int64_t o = 0; row[o++] = it; row[o++] = mu;
            memcpy(row + o, s.beta, M * 8); o += M;
            row[o++] = sigmaE;
            memcpy(row + o, s.comp, M * 8); o += M;
            memcpy(row + o, s.sigmaGG, G * 8); o += G;
            memcpy(row + o, s.eps, N * 8); o += N;
            memcpy(row + o, alphaF, F * 8); o += F;
            row[o++] = sigmaF;
            if (sink) sink(sctx, row, L);
        }
    }
    grp_free(&s); free(s.beta); free(s.eps); free(s.comp); free(s.sigmaGG);
    free(alphaF); free(row); free(dal); free(markerI); free(fixedI);
    return ORC_OK;
}

/* =====================================================================================
 * BRV2Grstart -- src/BRv2Grstart.cpp:77-306
 * ===================================================================================== */
int orc_grstart_run(const orc_grstart_args *a, const orc_draws *d, orc_row_fn sink, void *sctx)
{
    int rc = check_iters(a->max_iterations, a->burn_in, a->thinning);
    if (rc) return rc;
    const int64_t N = a->N, M = a->M; const int K = a->ncva + 1, G = a->groups;
    const int64_t L = orc_grstart_rowlen(N, M, G);
    grp_state s; memset(&s, 0, sizeof s);
    s.N = N; s.M = M; s.K = K; s.G = G; s.X = a->X; s.cva = a->cva; s.gAssign = a->gAssign;
    grp_alloc(&s);
    s.beta = malloc(M * 8); memcpy(s.beta, a->beta, M * 8);            /* by-value arguments, :77 */
    s.eps = malloc(N * 8); memcpy(s.eps, a->epsilon, N * 8);
    s.comp = malloc(M * 8); memcpy(s.comp, a->components, M * 8);
    s.sigmaGG = malloc(G * 8); memcpy(s.sigmaGG, a->sigmaGG, G * 8);
    double *row = calloc(L, 8), *dal = calloc(K, 8);
    int32_t *markerI = malloc(M * sizeof(int32_t));
    for (int64_t i = 0; i < M; ++i) markerI[i] = (int32_t)i;
    double mu = a->mu, sigmaE = a->sigmaE;
    for (int64_t j = 0; j < M; ++j) s.xsq[j] = vsqnorm(a->X + j * N, N);            /* :156 */
    for (int64_t i = 0; i < M; ++i) s.v[s.gAssign[i] * K + (int)s.comp[i]] += 1.0;  /* :159-162 */
    for (int g = 0; g < G; ++g) {                                                   /* :163-165 */
        for (int k = 0; k < K; ++k) dal[k] = s.v[g * K + k] + 1.0;
        dirichlet_rng_st(d, ORC_S_INIT_G, -1, (int64_t)g * (K + 1) + 1, dal, K, s.pi + g * K);
    }
    for (int it = 0; it < a->max_iterations; ++it) {
        for (int64_t i = 0; i < N; ++i) s.eps[i] = s.eps[i] + mu;                   /* :173 */
        mu = norm_rng(d, ORC_S_MU, it, 0, vsum(s.eps, N) / (double)N, sigmaE / (double)N);   /* :174 */
        for (int64_t i = 0; i < N; ++i) s.eps[i] = s.eps[i] - mu;                   /* :175 */
        d->shuffle(d->ctx, ORC_S_PERM, it, markerI, M);                             /* :178 */
        groups_sweep(&s, d, it, markerI, sigmaE);                                   /* :180-250 */
        sigmaE = inv_scaled_chisq_rng(d, ORC_S_GAMMA, it, 1, a->v0E + N,
                                      (vsqnorm(s.eps, N) + a->v0E * a->s02E) / (a->v0E + N));     /* :254 */
        groups_hyper(&s, d, it, a->v0G, a->s02G, dal);                              /* :257-262 */
        if (a->pi_trace) memcpy(a->pi_trace + (int64_t)it * G * K, s.pi, (size_t)G * K * 8);
        if (keep(it, a->burn_in, a->thinning, a->emit_all)) {                       /* :264-268 */
            int64_t o = 0; row[o++] = it; row[o++] = mu;
            memcpy(row + o, s.beta, M * 8); o += M;
            row[o++] = sigmaE;
            memcpy(row + o, s.comp, M * 8); o += M;
            memcpy(row + o, s.sigmaGG, G * 8); o += G;
            memcpy(row + o, s.eps, N * 8); o += N;
            if (sink) sink(sctx, row, L);
        }
    }
    grp_free(&s); free(s.beta); free(s.eps); free(s.comp); free(s.sigmaGG);
    free(row); free(dal); free(markerI);
    return ORC_OK;
}

/* =====================================================================================
 * HorseshoeR -- src/HorseshoeR.cpp:109-302
 * ===================================================================================== */
int orc_horseshoe_run(const orc_hs_args *a, const orc_draws *d, orc_row_fn sink, void *sctx)
{
    int rc = check_iters(a->max_iterations, a->burn_in, a->thinning);               /* :119-123 */
    if (rc) return rc;
    const int64_t N = a->N, M = a->M; const int64_t L = orc_hs_rowlen(N, M);
    const double A = a->A, vL = a->vL, vT = a->vT, vC = a->vC, sC = a->sC;
    double c2 = a->c2;
    double *lambda = calloc(M, 8), *v = calloc(M, 8), *beta = calloc(M, 8), *yt = calloc(N, 8), *eps = calloc(N, 8);
    double *row = calloc(L, 8), *tmp = calloc(M, 8);
    int32_t *markerI = malloc(M * sizeof(int32_t));
    for (int64_t i = 0; i < M; ++i) markerI[i] = (int32_t)i;
    double tau = d->uniform(d->ctx, ORC_S_INIT_U, -1, 0);                           /* :171 (overwritten at :192) */
    double mu = 0;                                                                  /* :173 */
    for (int64_t j = 0; j < M; ++j) v[j] = inv_gamma_rate_rng(d, ORC_S_INIT_G, -1, j, 0.5 + 0.5 * 0.0, 1.0);   /* :176 */
    for (int64_t j = 0; j < M; ++j) lambda[j] = inv_gamma_rate_rng(d, ORC_S_INIT_G, -1, M + j, 0.5 * vL, vL * 1.0);
    /* ^ :179 consumes draws from v BEFORE v.setOnes() took effect?  No: :177 v.setOnes() precedes :179, so x = 1. */
    for (int64_t j = 0; j < M; ++j) { v[j] = 1.0; lambda[j] = 1.0; }                /* :177,180 */
    for (int64_t i = 0; i < N; ++i) eps[i] = a->Y[i] - mu - 0.0;                    /* :186 */
    double sigmaE = vsqnorm(eps, N) / N * 0.5;                                      /* :187 */
    double eta = inv_gamma_rate_rng(d, ORC_S_INIT_G, -1, 2 * M, 0.5, 1 / (sigmaE * pow(A, 2)));   /* :189 */
    tau = (1.0 / eta) * inv_gamma_rate_rng(d, ORC_S_INIT_G, -1, 2 * M + 1, 0.5 * vT, vT);         /* :192 */

    for (int it = 0; it < a->max_iterations; ++it) {
        for (int64_t i = 0; i < N; ++i) eps[i] = eps[i] + mu;                       /* :210 */
        mu = norm_rng(d, ORC_S_MU, it, 0, vsum(eps, N) / (double)N, sigmaE / (double)N);   /* :211 */
        for (int64_t i = 0; i < N; ++i) eps[i] = eps[i] - mu;                       /* :212 */
        d->shuffle(d->ctx, ORC_S_PERM, it, markerI, M);                             /* :215 */
        eta = inv_gamma_rate_rng(d, ORC_S_GAMMA, it, 0, 0.5 + 0.5 * vT, (1.0 / (sigmaE * A * A)) + vT / tau);   /* :217 */
        for (int64_t j = 0; j < M; ++j)                                             /* :218 */
            v[j] = inv_gamma_rate_rng(d, ORC_S_HS_NU, it, j, 0.5 + 0.5 * vL, vL / lambda[j] + 1.0);
        for (int64_t j = 0; j < M; ++j) {                                           /* :219 */
            const int64_t marker = markerI[j];
            const double *x = a->X + marker * N;
            const double bo = beta[marker];
            for (int64_t i = 0; i < N; ++i) yt[i] = eps[i] + x[i] * bo;             /* :224 */
            /* :234 -- one expression; squaredNorm() evaluated twice */
            double s = tau * c2 * lambda[marker] / (tau * lambda[marker] + c2);
            double num = vdot(x, yt, N);
            double d1 = vsqnorm(x, N) + (sigmaE / s);
            double d2 = vsqnorm(x, N) + (sigmaE / s);
            beta[marker] = num / d1 + sqrt(sigmaE / d2) * norm_rng(d, ORC_S_MARK_Z, it, j, 0, 1);
            const double bn = beta[marker];
            for (int64_t i = 0; i < N; ++i) eps[i] = yt[i] - x[i] * bn;             /* :238 */
        }
        for (int64_t j = 0; j < M; ++j)                                             /* :242 */
            lambda[j] = inv_gamma_rate_rng(d, ORC_S_HS_LAM, it, j, 0.5 + 0.5 * vL,
                                           vL * (1.0 / v[j]) + (0.5 * (beta[j] * beta[j])) * (1.0 / tau));
        for (int64_t j = 0; j < M; ++j) tmp[j] = pow(beta[j], 2) / lambda[j];
        tau = inv_gamma_rate_rng(d, ORC_S_GAMMA, it, 1, 0.5 * (M + vT), vT / eta + ((0.5) * vsum(tmp, M)));   /* :245 */
        c2 = inv_gamma_rate_rng(d, ORC_S_GAMMA, it, 2, 0.5 * vC + 0.5 * M, vC * sC * 0.5 + 0.5 * vsqnorm(beta, M));   /* :248 */
        sigmaE = inv_scaled_chisq_rng(d, ORC_S_GAMMA, it, 3, a->v0E + N,
                                      (vsqnorm(eps, N) + a->v0E * a->s02E) / (a->v0E + N));   /* :253 */
        if (a->hyper_trace) { a->hyper_trace[3 * it] = eta; a->hyper_trace[3 * it + 1] = tau; a->hyper_trace[3 * it + 2] = c2; }
        if (keep(it, a->burn_in, a->thinning, a->emit_all)) {                       /* :255-259 */
            int64_t o = 0; row[o++] = it; row[o++] = mu;
            memcpy(row + o, beta, M * 8); o += M;
            row[o++] = sigmaE; row[o++] = tau;
            memcpy(row + o, lambda, M * 8); o += M;
            memcpy(row + o, eps, N * 8); o += N;
            if (sink) sink(sctx, row, L);
        }
    }
    free(lambda); free(v); free(beta); free(yt); free(eps); free(row); free(tmp); free(markerI);
    return ORC_OK;
}

/* =====================================================================================
 * CSV text
 * ===================================================================================== */
typedef struct { char *buf; size_t cap, n; } sbuf;
static void sput(sbuf *s, const char *t)
{
    size_t l = strlen(t);
    if (s->buf && s->n + l <= s->cap) memcpy(s->buf + s->n, t, l);
    s->n += l;
}
static void sputi(sbuf *s, const char *name, long i, const char *tail)
{
    char t[64]; snprintf(t, sizeof t, "%s[%ld]%s", name, i, tail); sput(s, t);
}
size_t orc_format_header(int kind, int64_t N, int64_t M, int G, int64_t F, char *buf, size_t cap)
{
    sbuf s = { buf, cap, 0 };
    sput(&s, "iteration,"); sput(&s, "mu,");
    for (int64_t i = 0; i < M; ++i) sputi(&s, "beta", i + 1, ",");
    if (kind == ORC_KIND_V2) {                                   /* src/BayesRv2.cpp:21-36 */
        sput(&s, "sigmaE,"); sput(&s, "sigmaG,");
        for (int64_t i = 0; i < M; ++i) sputi(&s, "comp", i + 1, ",");
        for (int64_t i = 0; i < N - 1; ++i) sputi(&s, "epsilon", i + 1, ",");
        sputi(&s, "epsilon", N, "");
    } else if (kind == ORC_KIND_GROUPS || kind == ORC_KIND_GRSTART) {   /* Groups:25-54, Grstart:26-50 */
        sput(&s, "sigmaE,");
        for (int64_t i = 0; i < M; ++i) sputi(&s, "comp", i + 1, ",");
        for (int i = 0; i < G; ++i) sputi(&s, "sigmaG", i + 1, ",");
        for (int64_t i = 0; i < N - 1; ++i) sputi(&s, "epsilon", i + 1, ",");
        if (kind == ORC_KIND_GROUPS) {
            sputi(&s, "epsilon", N, ",");
            for (int64_t i = 0; i < F; ++i) sputi(&s, "alpha", i + 1, ",");
            sput(&s, "sigmaF");
        } else sputi(&s, "epsilon", N, "");
    } else {                                                      /* HorseshoeR.cpp:279-291 (trailing comma) */
        sput(&s, "sigmaE,"); sput(&s, "tau,");
        for (int64_t i = 0; i < M; ++i) sputi(&s, "lambda", i + 1, ",");
        for (int64_t i = 0; i < N; ++i) sputi(&s, "epsilon", i + 1, ",");
    }
    sput(&s, "\n");
    return s.n;
}
/* Eigen IOFormat(StreamPrecision, DontAlignCols, ", ", ...) on a default ofstream: %g with 6
 * significant digits, ", " between coefficients, "\n" by the caller (src/BayesRv2.cpp:72,266) */
size_t orc_format_row(const double *row, int64_t len, char *buf, size_t cap)
{
    sbuf s = { buf, cap, 0 };
    char t[40];
    for (int64_t i = 0; i < len; ++i) {
        snprintf(t, sizeof t, "%g", row[i]);
        if (i) sput(&s, ", ");
        sput(&s, t);
    }
    sput(&s, "\n");
    return s.n;
}
