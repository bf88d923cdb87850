/*
 * bayesrr_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99, no Eigen / no R) of the four Gibbs samplers of
 * medical-genomics-group/BayesRRcpp.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * shipped CUDA path never links, includes or calls anything under oracle/.
 *
 * Pinning status: the reference holds no tests, golden vectors or fixtures
 * (SURVEY.md section 4), so the restatement is pinned against the reference's OWN
 * SOURCE FILES compiled here against a minimal Eigen/Rcpp/R shim (oracle/shim/,
 * recipe oracle/Makefile target `ref`, output oracle/_ref/libbayesrr_ref.so) under
 * a shared sequential draw stream -- see tests/test_oracle_vs_ref.py.  Where
 * /root/reference is absent (GPU box) the committed fixtures in tests/golden/
 * generated from that build are used instead.
 *
 * Reference sites restated (file:line relative to /root/reference):
 *   src/BayesRv2.cpp:146-274        -> orc_v2_run
 *   src/BayesRv2Groups.cpp:170-333  -> orc_groups_run
 *   src/BRv2Grstart.cpp:155-282     -> orc_grstart_run
 *   src/HorseshoeR.cpp:168-264      -> orc_horseshoe_run
 *   src/distributions.cpp:12-39,60-65 -> static helpers in bayesrr_oracle.c
 *   headers src/BayesRv2.cpp:16-37, Groups:25-54, Grstart:26-50, Horseshoe:279-291
 *                                   -> orc_write_header / orc_format_row
 */
#ifndef BAYESRR_ORACLE_H
#define BAYESRR_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- draw source: every random number the samplers consume, at the
 * distribution-output level: u ~ U(0,1), z ~ N(0,1),
 * g ~ Gamma(shape, 1), and the marker shuffle.  (stream, it, idx) identify the
 * draw; sequential sources ignore them, keyed sources (Philox, replay tables)
 * use them.  `it` is -1 for draws made before the first iteration.          */
enum {
    ORC_S_INIT_U  = 0,  /* uniform: V2 idx0 = sigmaG; Groups idx g = sigmaG[g], idx G = sigmaF; HS idx0 = discarded tau */
    ORC_S_MU      = 1,  /* normal, idx 0 */
    ORC_S_MARK_U  = 2,  /* uniform, idx = sweep position j */
    ORC_S_MARK_Z  = 3,  /* normal,  idx = sweep position j (drawn only when a non-zero component is chosen; HS: always) */
    ORC_S_GAMMA   = 4,  /* gamma, idx = slot (see each sampler) */
    ORC_S_PERM    = 5,  /* marker shuffle */
    ORC_S_FIX_Z   = 6,  /* normal, idx = position in the fixed-effect order */
    ORC_S_FIXPERM = 7,  /* fixed-effect shuffle */
    ORC_S_HS_NU   = 8,  /* gamma, idx = marker index */
    ORC_S_HS_LAM  = 9,  /* gamma, idx = marker index */
    ORC_S_INIT_G  = 10, /* gamma before the first iteration (HS: 0..M-1 v, M..2M-1 lambda, 2M eta, 2M+1 tau; Grstart: pi slots) */
    ORC_N_STREAMS = 11
};

typedef struct orc_draws {
    void *ctx;
    double (*uniform)(void *ctx, int stream, int64_t it, int64_t idx);
    double (*normal)(void *ctx, int stream, int64_t it, int64_t idx);
    double (*gamma)(void *ctx, int stream, int64_t it, int64_t idx, double shape);
    /* in-place shuffle of order[0..n) exactly as std::random_shuffle does it:
     * for i = 1..n-1: swap(order[i], order[r_i % (i+1)]) -- the source supplies r_i */
    void (*shuffle)(void *ctx, int stream, int64_t it, int32_t *order, int64_t n);
} orc_draws;

/* --- Philox4x32-10 keyed source (same (seed,stream,it,idx) -> draw map as the CUDA path) */
typedef struct { uint32_t key[2]; } orc_philox;
void orc_philox_init(orc_philox *p, uint64_t seed);
void orc_philox_raw(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4]);
orc_draws orc_philox_source(orc_philox *p);

/* --- sequential source: own counter per kind; shuffle uses libc rand() like libstdc++'s
 * std::random_shuffle.  Used to pin the restatement against oracle/_ref (same streams fed
 * to the shimmed R:: calls of the real reference sources). */
typedef struct { orc_philox px; uint64_t n_u, n_z, n_g; } orc_seq;
void orc_seq_init(orc_seq *s, uint64_t seed);
orc_draws orc_seq_source(orc_seq *s);
double orc_seq_next_uniform(orc_seq *s);
double orc_seq_next_normal(orc_seq *s);
double orc_seq_next_gamma(orc_seq *s, double shape);

/* --- recording / replay tables.  Layout per iteration t in [0,n_iter):
 *   mark_u[t*M+j], mark_z[t*M+j] (NaN when not drawn), mu_z[t], perm[t*M+j] (marker visited at
 *   position j), gam[t*n_gam+slot], fix_z[t*F+j], fixperm[t*F+j], hs_nu[t*M+j], hs_lam[t*M+j];
 *   init_u[n_init_u], init_g[n_init_g] for the pre-iteration draws.                          */
typedef struct orc_tables {
    int64_t n_iter, M, F, n_gam, n_init_u, n_init_g;
    double *mark_u, *mark_z, *mu_z, *gam, *fix_z, *hs_nu, *hs_lam, *init_u, *init_g;
    int32_t *perm, *fixperm;
} orc_tables;
typedef struct { orc_draws inner; orc_tables *t; } orc_recorder;
orc_draws orc_record_source(orc_recorder *r);      /* draws from r->inner, stores into r->t */
orc_draws orc_replay_source(orc_tables *t);        /* reads t (perm: sets order = perm row)   */

/* ---- output rows */
typedef void (*orc_row_fn)(void *ctx, const double *row, int64_t len);

enum { ORC_OK = 0, ORC_ERR_ITER = 1, ORC_ERR_ARG = 2 };
enum { ORC_KIND_V2 = 0, ORC_KIND_GROUPS = 1, ORC_KIND_GRSTART = 2, ORC_KIND_HORSESHOE = 3 };

/* BayesRSamplerV2 -- src/BayesRv2.cpp:60.  X column-major N x M.  pi_init (length K=ncva+1):
 * explicit because the reference reads uninitialised memory for it.
 * emit_all != 0: emit a row after EVERY iteration (state trace for tests) instead of the
 * burn_in/thinning rule of :257-259.  pi_trace (optional, n_iter*K) receives pi after each iteration.
 * gamma slots per iteration: 0 sigmaG, 1 sigmaE, 2+k pi[k].                                  */
typedef struct {
    int max_iterations, burn_in, thinning;
    int64_t N, M; const double *X, *Y;
    double sigma0, v0E, s02E, v0G, s02G;
    const double *cva; int ncva;
    const double *pi_init;
    int emit_all; double *pi_trace;
} orc_v2_args;
int orc_v2_run(const orc_v2_args *a, const orc_draws *d, orc_row_fn sink, void *sink_ctx);
static inline int64_t orc_v2_rowlen(int64_t N, int64_t M) { return 2 * M + 4 + N; }

/* BayesRSamplerV2Groups -- src/BayesRv2Groups.cpp:75.  cva: groups x (K-1) column-major (Eigen).
 * gamma slots: 0 sigmaF, 1 sigmaE, 2+g*(K+1) sigmaG[g], 2+g*(K+1)+1+k pi[g][k].  pi_trace: n_iter*G*K (row-major g,k) */
typedef struct {
    int max_iterations, burn_in, thinning;
    int64_t N, M; const double *X, *Y;
    double sigma0, v0E, s02E, v0G, s02G;
    const double *cva; int ncva; int groups; const int32_t *gAssign;
    const double *fixed; int64_t F;
    int emit_all; double *pi_trace;
} orc_groups_args;
int orc_groups_run(const orc_groups_args *a, const orc_draws *d, orc_row_fn sink, void *sink_ctx);
static inline int64_t orc_groups_rowlen(int64_t N, int64_t M, int G, int64_t F) { return 2 * M + 3 + G + N + F + 1; }

/* BRV2Grstart -- src/BRv2Grstart.cpp:77.  beta, sigmaGG, epsilon, components are copied (by-value in the reference).
 * init gamma slots (ORC_S_INIT_G): g*(K+1)+1+k pi[g][k]; per-iteration slots as Groups (slot 0 unused). */
typedef struct {
    int max_iterations, burn_in, thinning;
    double mu; const double *beta; double sigmaE; const double *sigmaGG;
    int64_t N, M; const double *X; const double *epsilon; const double *components;
    double sigma0, v0E, s02E, v0G, s02G;
    const double *cva; int ncva; int groups; const int32_t *gAssign;
    int emit_all; double *pi_trace;
} orc_grstart_args;
int orc_grstart_run(const orc_grstart_args *a, const orc_draws *d, orc_row_fn sink, void *sink_ctx);
static inline int64_t orc_grstart_rowlen(int64_t N, int64_t M, int G) { return 2 * M + 3 + G + N; }

/* HorseshoeR -- src/HorseshoeR.cpp:109.  gamma slots per iteration: 0 eta, 1 tau, 2 c2, 3 sigmaE. */
typedef struct {
    int max_iterations, burn_in, thinning;
    int64_t N, M; const double *X, *Y;
    double A, v0E, s02E, vL, vT, c2, vC, sC;
    int emit_all; double *hyper_trace; /* optional n_iter*3: eta, tau, c2 */
} orc_hs_args;
int orc_horseshoe_run(const orc_hs_args *a, const orc_draws *d, orc_row_fn sink, void *sink_ctx);
static inline int64_t orc_hs_rowlen(int64_t N, int64_t M) { return 2 * M + 4 + N; }

/* CSV text exactly as the reference writes it: returns bytes written into buf
 * (or needed if buf==NULL). */
size_t orc_format_header(int kind, int64_t N, int64_t M, int G, int64_t F, char *buf, size_t cap);
size_t orc_format_row(const double *row, int64_t len, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
