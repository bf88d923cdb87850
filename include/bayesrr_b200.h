/*
 * bayesrr_b200.h -- C ABI of the B200-native BayesR / BayesRR / Horseshoe Gibbs sampler.
 *
 * Plain pointers and sizes only (no torch / Eigen / Rcpp types).  Every entry point returns 0 on
 * success or a BRR_E_* code; brr_last_error() returns the message of the last failure on the calling
 * thread.  There is no CPU fallback: every call that computes requires an sm_100 device and fails
 * with BRR_E_CUDA otherwise.
 *
 * Reference interface replaced (file:line relative to the reference repository):
 *   brr_BayesRSamplerV2        <- void BayesRSamplerV2(...)        src/BayesRv2.cpp:60        (glue src/RcppExports.cpp:39-59)
 *   brr_BayesRSamplerV2Groups  <- void BayesRSamplerV2Groups(...)  src/BayesRv2Groups.cpp:75  (glue src/RcppExports.cpp:61-84)
 *   brr_BRV2Grstart            <- void BRV2Grstart(...)            src/BRv2Grstart.cpp:77     (glue src/RcppExports.cpp:10-37)
 *   brr_HorseshoeR             <- void HorseshoeR(...)             src/HorseshoeR.cpp:109     (glue src/RcppExports.cpp:86-108)
 * Arguments keep the reference's order and meaning; every Eigen matrix/vector becomes (pointer, sizes),
 * column-major like Eigen.  Results are delivered through `outputFile` exactly like the reference
 * (header + one ", "-separated row per kept iteration, src/BayesRv2.cpp:16-37,72,260-267).
 * INTEGRATION.md shows the Rcpp-side binding a maintainer adds.
 */
#ifndef BAYESRR_B200_H
#define BAYESRR_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
    BRR_OK = 0,
    BRR_E_ITER = 1,     /* max_iterations < burn_in || max_iterations < 1 || burn_in < 1 (src/BayesRv2.cpp:76-80);
                           the output file has been truncated (and, for V2, the header written) like the reference */
    BRR_E_ARG = 2,      /* invalid argument (null pointer, size mismatch, thinning < 1, ...) */
    BRR_E_GENO = 3,     /* packed genotypes hold code 3 / a .bed file holds missing genotypes and imputation was not asked for */
    BRR_E_CUDA = 4,     /* CUDA failure or no sm_100 device: there is no CPU fallback */
    BRR_E_IO = 5,       /* output file cannot be opened / written */
    BRR_E_SIZE = 6      /* problem does not fit this build's limits (K, rows per device) */
};
enum { BRR_V2 = 0, BRR_GROUPS = 1, BRR_GRSTART = 2, BRR_HORSESHOE = 3 };

const char *brr_last_error(void);
int brr_abi_version(void);

/* The reference's console messages -- "iteration: <n>\n" whenever iteration % max(1, max_iterations / 10) == 0 and
 * "duration: <s>s\n" at the end (src/BayesRv2.cpp:173-175,276-278, Rcpp::Rcout there; HorseshoeR also prints "initial eta",
 * "initial tau" and tau / eta / sigmaE with every progress line, src/HorseshoeR.cpp:191,195,203-206) -- are handed to this
 * call-back by the four entry points, on the calling thread.  NULL (the default): the messages are dropped.  Process-wide. */
typedef void (*brr_message_fn)(void *ctx, const char *text);
void brr_set_message_handler(brr_message_fn fn, void *ctx);

/* ------------------------------------------------------------------------------------------------
 * The four reference entry points.  X is N x M column-major fp64 (what Rcpp hands over as
 * Eigen::MatrixXd); genotype-like columns are packed to 2 bits, any other column stays dense fp64
 * (brr_geno_from_dense).
 * ------------------------------------------------------------------------------------------------ */
int brr_BayesRSamplerV2(const char *outputFile, int seed, int max_iterations, int burn_in, int thinning,
                        const double *X, int64_t N, int64_t M, const double *Y,
                        double sigma0, double v0E, double s02E, double v0G, double s02G,
                        const double *cva, int ncva);

/* cva: groups x ncva column-major (Eigen::MatrixXd); gAssign: M zero-based group ids; fixed: N x F column-major (F may be 0) */
int brr_BayesRSamplerV2Groups(const char *outputFile, int seed, int max_iterations, int burn_in, int thinning,
                              const double *X, int64_t N, int64_t M, const double *Y,
                              double sigma0, double v0E, double s02E, double v0G, double s02G,
                              const double *cva, int ncva, int groups, const int32_t *gAssign,
                              const double *fixed, int64_t F);

/* beta: M; sigmaGG: groups; epsilon: N; components: M doubles (as in the reference, src/BRv2Grstart.cpp:77) */
int brr_BRV2Grstart(const char *outputFile, int seed, int max_iterations, int burn_in, int thinning,
                    double mu, const double *beta, double sigmaE, const double *sigmaGG,
                    const double *X, int64_t N, int64_t M, const double *epsilon, const double *components,
                    double sigma0, double v0E, double s02E, double v0G, double s02G,
                    const double *cva, int ncva, int groups, const int32_t *gAssign);

int brr_HorseshoeR(const char *outputFile, int seed, int max_iterations, int burn_in, int thinning,
                   const double *X, int64_t N, int64_t M, const double *Y,
                   double A, double v0E, double s02E, double vL, double vT, double c2, double vC, double sC);

/* ------------------------------------------------------------------------------------------------
 * Genotype storage / packing layer: 2-bit codes, column-major, 16-byte aligned column stride,
 * resident in HBM; per-SNP affine map x = a + d*code (a = -mean/sd, d = 1/sd for scale()d columns).
 * ------------------------------------------------------------------------------------------------ */
typedef struct brr_geno brr_geno;

/* Ingest a dense column-major fp64 matrix (host memory).  A column with at most three equally spaced values (relative tolerance
 * 1e-9) -- a genotype column, raw or centred / scaled -- is packed to 2-bit codes; any other column (a continuous covariate, e.g.
 * the scale()d methylation probes the reference's vignette binds to the genotypes, vignettes/BayesRR.Rmd:150-167) is kept as a
 * dense fp64 column and takes the same sweep through the same interfaces. */
int brr_geno_from_dense(const double *X, int64_t N, int64_t M, int device, brr_geno **out);
/* Turn the markers cols[0..n_cols) (ascending) of a store into dense fp64 columns holding values (host, N x n_cols column-major,
 * taken as given: centre / scale them beforehand): how a store built from packed codes or a .bed file gets its continuous
 * covariates, and how the ranks of a row-sharded chain pass their rows of them (before brr_geno_shard_stats). */
int brr_geno_set_dense_columns(brr_geno *g, const int32_t *cols, int64_t n_cols, const double *values);
/* number of dense columns; dense_idx (may be NULL, M entries): index of each marker's dense column or -1 */
int brr_geno_dense_columns(const brr_geno *g, int64_t *n_dense, int32_t *dense_idx);
/* Adopt host 2-bit codes: column j starts at packed + j*col_stride_bytes, individual i is bits
 * 2*(i%4).. of byte i/4, code 3 is rejected (missing data is not a concept of the reference).
 * mean/sd NULL -> computed from the codes (sd with the N-1 denominator, like R scale()).       */
int brr_geno_from_packed(const uint8_t *packed, int64_t col_stride_bytes, int64_t N, int64_t M,
                         const double *mean, const double *sd, int device, brr_geno **out);
/* PLINK 1 .bed file (SNP-major) -> store, without a dense detour: N_total individuals and M markers as counted from the .fam / .bim
 * files; rows [row0, row0 + N) of the file (row0 a multiple of 4; N <= 0: all rows) -- a row shard reads only its own bytes.
 * Codes count A1 alleles (00 -> 2, 10 -> 1, 11 -> 0).  Missing genotypes (01): BRR_E_GENO unless impute_missing != 0, in which
 * case they take the integer code nearest to the column mean of the observed genotypes; *n_missing (may be NULL) = how many (in
 * these rows).  A ROW SHARD read with impute_missing keeps its missing genotypes until brr_geno_shard_stats, which sums the
 * per-column counts over the ranks first -- every rank fills with the same value, that of the unsharded file -- and the store
 * can be used only after that call. */
int brr_geno_from_bed(const char *bed_path, int64_t N_total, int64_t M, int64_t row0, int64_t N, int impute_missing,
                      int device, brr_geno **out, int64_t *n_missing);
/* Device-side synthetic generator: g_ij ~ Binomial(2, p_j), p_j ~ U(0.05, 0.5), standardised.
 * Row sharding: this store holds rows [row0, row0+N) of a virtual N_total-row matrix (statistics are
 * computed over the local rows only until brr_geno_shard_stats is called).                      */
int brr_geno_synthetic(int64_t N, int64_t M, uint64_t seed, int64_t row0, int device, brr_geno **out);
int brr_geno_dims(const brr_geno *g, int64_t *N, int64_t *M, int64_t *col_stride_bytes);
/* host copies of the per-SNP statistics (any pointer may be NULL): code mean, code sd, a, d, ||x||^2 */
int brr_geno_stats(const brr_geno *g, double *mean, double *sd, double *a, double *d, double *xsq);
/* host copy of the packed codes (M * col_stride_bytes) */
int brr_geno_codes(const brr_geno *g, uint8_t *packed_out);
/* y = X * b for host vectors b (M) -> y (N): builds phenotypes for configurations whose dense X cannot exist */
int brr_geno_matvec(const brr_geno *g, const double *b, double *y);
void brr_geno_free(brr_geno *g);

/* ------------------------------------------------------------------------------------------------
 * Chain objects: what the four entry points are built from; used directly by tests, the benchmark and
 * configurations that cannot pass through a dense X.
 * ------------------------------------------------------------------------------------------------ */
typedef struct brr_chain brr_chain;

typedef struct brr_config {
    int kind;                       /* BRR_V2 ... BRR_HORSESHOE */
    uint64_t seed;                  /* Philox key (the reference ignores its seed) */
    int max_iterations, burn_in, thinning;
    const double *Y;                /* N (not used by BRR_GRSTART) */
    double sigma0, v0E, s02E, v0G, s02G;
    const double *cva; int ncva;    /* V2: ncva values; Groups/Grstart: groups x ncva column-major */
    int groups; const int32_t *gAssign;
    const double *fixed; int64_t F; /* Groups only */
    const double *pi_init;          /* V2 only, K = ncva+1 values; NULL -> {0.5, 0.5*cva/sum(cva)} */
    /* BRR_GRSTART state (src/BRv2Grstart.cpp:61-67) */
    double mu0; const double *beta0; double sigmaE0; const double *sigmaGG0;
    const double *epsilon0; const double *components0;
    /* BRR_HORSESHOE (src/HorseshoeR.cpp:109) */
    double A, vL, vT, c2, vC, sC;
    /* engine options; 0 = default */
    int block;                      /* markers per Gibbs block: 32, 64 or 128 (default 128) */
    int gram_impl;                  /* 0 = tcgen05 int8 tensor cores, 1 = dp4a CUDA cores (validation) */
    int workers;                    /* worker CTAs of the persistent sweep kernel (default: SMs - 1 - the SMs left to the Gram
                                       kernel of the next iteration, which runs beside the sweep) */
} brr_config;

/* Draw-replay tables (host memory, copied at set time).  Layout per iteration t in [0, n_iter):
 *   perm[t*M + j]   marker visited at sweep position j
 *   mark_u[t*M + j] uniform used at position j;  mark_z[t*M + j] standard normal (NaN where unused)
 *   mu_z[t]; gam[t*n_gam + slot] unit-scale gamma variates; fix_z[t*F + j], fixperm[t*F + j];
 *   hs_nu[t*M + marker], hs_lam[t*M + marker];  init_u / init_g: draws made before iteration 0.
 * Slot numbering per sampler is documented in DESIGN.md.                                            */
typedef struct brr_replay {
    int64_t n_iter, M, F, n_gam, n_init_u, n_init_g;
    const double *mark_u, *mark_z, *mu_z, *gam, *fix_z, *hs_nu, *hs_lam, *init_u, *init_g;
    const int32_t *perm, *fixperm;
} brr_replay;

int brr_chain_create(const brr_config *cfg, brr_geno *g, brr_chain **out);
int brr_chain_set_replay(brr_chain *c, const brr_replay *r);
/* Append sample rows to `path` exactly like the reference writer (header written immediately). */
int brr_chain_open_output(brr_chain *c, const char *path);
/* Lossless sink beside (or instead of) the CSV: 64-byte header ("BRRSMP1\0", int32 kind, int32 groups, int64 N, M, F, row_len, zero
 * padding) followed by the sample rows as raw little-endian fp64, reference row layout. */
int brr_chain_open_binary_output(brr_chain *c, const char *path);
/* Lossless checkpoint / resume for all four samplers: brr_chain_save writes the complete state after the last
 * completed iteration; brr_chain_load, applied to a freshly created chain of the same configuration (same data, seed, shape)
 * before its first run, makes it continue that chain bit for bit.  Row-sharded chains: one file per rank. */
int brr_chain_save(brr_chain *c, const char *path);
int brr_chain_load(brr_chain *c, const char *path);
int64_t brr_chain_row_len(const brr_chain *c);
/* Run n_iter further iterations.  rows (may be NULL): receives up to max_rows sample rows (reference row
 * layout) for kept iterations, or for every iteration when emit_all != 0; *n_rows = rows produced. */
int brr_chain_run(brr_chain *c, int n_iter, int emit_all, double *rows, int64_t max_rows, int64_t *n_rows);
/* mixture proportions after the last completed iteration: groups x K row-major (V2: K values) */
int brr_chain_get_pi(brr_chain *c, double *pi);
/* Horseshoe: eta, tau, c2 after the last completed iteration */
int brr_chain_get_hyper(brr_chain *c, double *eta_tau_c2);
/* residual variance after the last completed iteration */
int brr_chain_get_sigmaE(brr_chain *c, double *sigmaE);
/* device time (ms, CUDA events on the chain's stream) and kernel launches of the last brr_chain_run */
int brr_chain_last_timing(const brr_chain *c, double *ms, int64_t *launches);
/* device time (ms, CUDA events) of the last brr_chain_run split by kernel: [0] block-Gram kernel (on its own stream, beside the
 * previous iteration's sweep: not on the critical path), [1] tables + persistent sweep kernel, [2] hyper-parameter kernel(s);
 * summed over the iterations of that run */
int brr_chain_kernel_ms(const brr_chain *c, double *gram_sweep_hyper_ms);
/* SM-clock cycle accounting over the last brr_chain_run (16 values).  Sampler CTA: [0] waiting for the workers' dots ([1] of it for a block's first chunk of 32, [7] for its last),
 * [2] the whole serial in-block pass, [3] block set-up before it, [4] rounds of the walk, [5] state-changing marker steps,
 * [6] blocks, [11] bookkeeping, [12] receiving a block's dots; [9] draws that took the fp64 evaluation (K = 3, 4); with -DBRR_ROUND_PROFILE=1 instead [9] threshold tests, [14] full
 * draws + corrections, [15] sub-window prologues + waiting for the look-ahead correction (warp 1).  First worker CTA: [8] waiting for / applying deltas, [10] dot stage + send,
 * [13] forming column totals */
int brr_chain_sweep_profile(const brr_chain *c, double *out16);
/* launch geometry chosen for the sweep kernel */
int brr_chain_geometry(const brr_chain *c, int *block, int *workers, int *rows_per_worker_max, int *smem_bytes);
/* flush the writer and close the output file */
int brr_chain_close_output(brr_chain *c);
void brr_chain_destroy(brr_chain *c);

/* ------------------------------------------------------------------------------------------------
 * Row-sharded chains: the individuals (rows of X, Y, fixed, epsilon) are partitioned over `world` devices,
 * one rank per device (one process per GPU, or several ranks as threads of one process).  Every rank runs
 * the whole chain (beta, components, pi, sigma replicated and bit-identical); per Gibbs block the ranks'
 * partial X_b^T eps vectors are exchanged device-to-device over NVLink peer memory inside the persistent
 * sweep kernel and summed in rank order, and the per-iteration Gram partials are summed by a peer-memory
 * kernel.  The host supplies two collectives, used at set-up time and at the
 * run boundaries only (never inside the iteration loop):
 * ------------------------------------------------------------------------------------------------ */
enum { BRR_MAX_WORLD = 8 };
typedef struct brr_comm {
    int rank, world;
    /* in-place sum over ranks of buf[0..n); every rank must end with identical bits; returns 0 on success */
    int (*allreduce_sum)(void *ctx, double *buf, int64_t n);
    /* recv (world * nbytes) = concatenation of every rank's send (nbytes) in rank order; returns 0 on success */
    int (*allgather)(void *ctx, const void *send, void *recv, int64_t nbytes);
    void *ctx;                      /* must stay valid for the life of the objects created with this comm */
} brr_comm;

/* Replace the per-SNP statistics (mean, sd, a, d, ||x||^2) of a row shard by those of the whole matrix
 * (code sums allreduced over the ranks; sd with the N_total - 1 denominator).  Collective. */
int brr_geno_shard_stats(brr_geno *g, const brr_comm *comm);
/* Like brr_chain_create for a row shard: cfg->Y / fixed / epsilon0 hold this rank's rows, everything indexed by
 * marker is replicated.  `g` must have had brr_geno_shard_stats applied.  Sample rows carry the residuals of
 * ALL ranks (reference row layout with N = N_total) on every rank.  Collective; so are brr_chain_run and
 * brr_chain_destroy of a sharded chain. */
int brr_chain_create_sharded(const brr_config *cfg, brr_geno *g, const brr_comm *comm, brr_chain **out);
/* exercises the two call-backs (CPU only: no device needed): buf[0..n) is sum-allreduced, gathered[world] receives
 * every rank's `token` */
int brr_comm_selftest(const brr_comm *comm, double *buf, int64_t n, int64_t token, int64_t *gathered);

/* Text of one sample row as the writer emits it: "%g" (6 significant digits) of every value joined by ", " (reference
 * src/BayesRv2.cpp:72,266).  Returns the length without the terminating NUL (the text is truncated to cap - 1 characters); CPU only. */
int64_t brr_format_row(const double *row, int64_t len, char *out, int64_t cap);

/* Stand-alone kernels exposed for parity tests and roofline measurement.
 * gram: G[b][i][j] = sum_n code[n, order[b*B+i]] * code[n, order[b*B+j]]  (int32, nb x B x B; order index -1 = padding) */
int brr_gram_blocks(const brr_geno *g, const int32_t *order, int64_t n_order, int block, int impl,
                    int32_t *G_out, double *ms);
/* same, plus X[b][jl][k] = sum_n code[n, order[b*B - LA + jl]] * code[n, order[b*B + k]] (int32, nb x LA x B; block 0: zeros),
 * LA = brr_lookahead(B): the products with the last LA markers of the previous block that the sweep's look-ahead correction
 * uses (X_out may be NULL) */
int brr_gram_cross_blocks(const brr_geno *g, const int32_t *order, int64_t n_order, int block, int impl,
                          int32_t *G_out, int32_t *X_out, double *ms);
/* look-ahead depth of the sweep for Gibbs blocks of `block` markers (the whole block: 128 for 128, 64 for 64, 32 for 32; a build-time constant of the
 * library): the deltas of a block's last LA markers reach the next block's dots through the cross-Gram correction; 0 = unsupported block */
int brr_lookahead(int block);
/* r[j] = sum_n x[n, j] * eps[n] for every marker (host eps N -> host r M), the streaming X^T eps kernel */
int brr_xt_eps(const brr_geno *g, const double *eps, double *r, double *ms);
/* measured fp64 FMA throughput of the device (thread-level DFMAs per second, independent chains on every SM): the roofline the
 * workers' dot stage is reported against (bench.py) */
int brr_peak_fp64(int device, double *dfma_per_s);
/* counter-based draws exactly as the device code makes them (for generator parity tests) */
int brr_draws_sample(uint64_t seed, int stream, int64_t it, int64_t idx0, int64_t n, int kind /*0 u, 1 z, 2 gamma*/,
                     double shape, double *out);
int brr_shuffle_host(uint64_t seed, int stream, int64_t it, int32_t *order, int64_t n);

/* test hook, no device needed: `nrows` copies of `row` (value 0 replaced by the row index) through the product's queue-backed writer into
 * `path` -- CSV text as brr_format_row gives it, or raw fp64 when `binary` */
int brr_writer_selftest(const char *path, const double *row, int64_t len, int64_t nrows, int binary);

#ifdef __cplusplus
}
#endif
#endif
