// Parameters of the persistent per-iteration sweep kernel (sweep.cu) and of the hyper-parameter kernels (hyper.cu).
#pragma once
#include "common.cuh"

namespace brr {

constexpr int SWEEP_THREADS = 256;
constexpr int KMAX = 16;          // mixture components supported by the in-block sampler
constexpr int MAXR = BRR_MAX_WORLD;   // ranks of a row-sharded chain
constexpr int SWEEP_REDUCERS = 4;     // reducer CTAs of the sweep kernel: 32 warps, one column total each per chunk of 32 markers

// Scalars of the chain that live on the device between kernels.
struct IterScalars {
    double mu;        // intercept of the iteration whose sweep ran last
    double mu_next;   // intercept drawn for the next iteration (reference src/BayesRv2.cpp:178)
    double shift;     // mu - mu_next: added to every residual when the next sweep loads it (:177,:179)
    double eps_sum;   // sum of residuals after the shift (kept analytically inside the sweep)
    double eps_sq;    // ||eps||^2 after the last sweep
    double sigmaE, sigmaF;
    double tau, c2, eta, eta_next;   // horseshoe (eta_next: drawn for the next iteration, reference HorseshoeR.cpp:217)
    double beta_sq;        // ||beta||^2 after the last sweep
    int64_t it_done;       // iterations completed
};

struct SweepParams {
    // genotypes (local rows)
    const uint8_t *packed; int64_t stride; int64_t N;
    const double *colA, *colD, *colS, *colXsq, *colCsum;
    double n_total;
    // iteration inputs
    const int32_t *perm;          // M markers in visiting order
    const int32_t *gram;          // nb x gram_tile_entries(B) int32 (codes), rows/cols in visiting order, block-upper trapezoid of every tile (common.cuh)
    const int32_t *xgram;         // nb x lookahead(B) x B int32: products with the last lookahead(B) markers of the previous block
    // stores with dense fp64 columns: the same tiles as fp64 (gram.cu, gram_dense_kernel) and the dense columns
    const double *gramd, *xgramd;
    const double *dense; const int32_t *denseIdx; int64_t Npad;   // Npad x Md column-major; per-marker dense column index or -1; null when none
    const uint8_t *gtab;          // nb x table bytes: per-marker tables of this iteration (tables_kernel), sampler smem layout
    int64_t M; int nb;
    int64_t it;
    // chain state
    double *eps;                  // Npad residuals (local rows)
    double *beta;                 // M
    double *comp;                 // M (component ids stored as doubles like the reference's sample row)
    IterScalars *sc;
    // mixture model (kind 0)
    int K, G;
    const int32_t *gAssign;       // M or null (single group)
    const double *cva;            // G x (K-1) column-major
    const double *sigmaG;         // G
    const double *pi;             // G x K row-major
    double *vcount;               // G x K  (output: component counts of this sweep)
    double *betaAcum;             // G      (output: sum of squared non-zero draws per group, sweep order)
    // horseshoe (kind 1)
    const double *lambda;         // M
    // draws
    PhiloxKey key;
    const double *tbl_u, *tbl_z;  // replay tables of this iteration (M each) or null
    // fixed effects (Groups)
    int F; const double *fixed;   // N x F column-major (local rows), or null
    const int32_t *fixperm;       // F
    const double *fixG;           // F x F Gram of the fixed columns (fp64)
    double *alpha;                // F
    const double *tbl_fix_z;      // F or null
    // grid protocol
    uint64_t *ll_part;            // 2 x PS x nW flagged-word slots (2 x u64 each; two phase parities): workers' partial dots; zeroed before launch
    // cross-rank exchange over peer memory (row-sharded chains; R == 1: this device's own window).  Never zeroed between
    // launches: flags are global phase numbers, monotone over the life of the chain.
    int rank, R;
    uint64_t *xred[MAXR];         // every rank's window of column totals: [phase mod 4][PS][R] slots; this rank writes [.][.][rank]
    uint64_t *xfin[MAXR];         // every rank's window of end-of-sweep sums: [R][2] slots (sum eps, sum eps^2 over that rank's rows)
    uint32_t xphase0;             // global number of this launch's first phase
    uint64_t *ll_fin;             // nW x 2 slots: workers' sum eps, sum eps^2 -> sampler; zeroed before launch
    uint64_t *ll_delta;           // 2 x PS slots (two phase parities): the sampler's per-marker deltas, streamed as they are decided
    uint64_t *ll_bcast;           // (3 x PS + 1) slots: sampler -> workers delta, a*delta, d*delta (+ sentinel); zeroed before launch
    long long *prof;              // optional cycle accounting of the sampler CTA: wait, reduce, pass, publish, windows, full steps, blocks
    int *abort_flag;              // set by the in-kernel watchdog (1: hand-over timed out, 2: bulk copy timed out)
    double *fin;                  // 2: sum eps, sum eps^2 over ALL rows (all ranks), written by the sampler CTA
    int nW; int PS;
    int nR;                       // reducer CTAs after the workers (column totals of the partial dots)
    const int32_t *unit0;         // nW + 1: first 64-row unit of every worker
    int seg_bytes;                // bytes reserved per staged column segment (max units * 16)
};

// Geometry helpers shared by host and device
struct SweepGeom { int B, TW, nW, seg_bytes; size_t smem_bytes; };

void launch_sweep(int kind, int B, int TW, const SweepParams &p, size_t smem, cudaStream_t stream);
// per-marker tables of the iteration described by `p` into gtab (nb x sweep_table_bytes); kind as for launch_sweep
void launch_tables(int kind, int B, const SweepParams &p, uint8_t *gtab, cudaStream_t stream);
size_t sweep_table_bytes(int kind, int B, int K, int G, int F);
// load every kernel an iteration launches (see preload_kernel)
void preload_tables(int kind);
void preload_gram(int B, int impl);
void preload_hyper(int kind);
void preload_allsum();
size_t sweep_smem_bytes(int kind, int B, int TW, int K, int G, int F, int seg_bytes, bool dense = false);
int sweep_max_coresident(int kind, int B, int TW, size_t smem, bool dense = false);

// d_G: nb x gram_tile_entries(B) self products (a tile is stored as its block-upper trapezoid, common.cuh); d_X (nullable): nb x lookahead(B) x B products with the last lookahead(B) markers of the previous block
// max_ctas > 0: persistent grid of at most that many CTAs (the tensor-core kernel loops over the blocks)
// abort_flag (nullable): the chain's sticky watchdog flag (a lost bulk copy raises code 4)
void launch_gram(const brr_geno *g, const int32_t *d_order, int64_t n_order, int B, int impl, int32_t *d_G, int32_t *d_X, cudaStream_t stream,
                 int max_ctas = 0, int *abort_flag = nullptr);
// Stores with dense columns: the same tiles as fp64 -- the exact int32 counts of the packed pairs (d_Gi, d_Xi: launch_gram's output) converted,
// every pair with a dense column as a fp64 dot over the local rows in a fixed order.  d_Gd: nb x gram_tile_entries(B), d_Xd: nb x lookahead(B) x B.
void launch_gram_dense(const brr_geno *g, const int32_t *d_order, int64_t n_order, int B, const int32_t *d_Gi, const int32_t *d_Xi,
                       double *d_Gd, double *d_Xd, cudaStream_t stream, int max_ctas = 0);
void preload_gram_dense();

}  // namespace brr
