// Host drivers of the four samplers: chain state in HBM, per-iteration launch sequence
//   chain stream:  tables kernel -> persistent sweep kernel -> hyper-parameter kernel(s)  [-> row snapshot D2H]
//   Gram stream:   [host shuffle -> H2D] -> block-Gram kernel (+ sum over ranks), one iteration ahead, beside the sweep
// and the C ABI on top of it (entry points, sharded chains, sinks, checkpoints).  Mirrors the driver loops of the reference (src/BayesRv2.cpp:146-274,
// src/BayesRv2Groups.cpp:170-333, src/BRv2Grstart.cpp:155-282, src/HorseshoeR.cpp:168-264); nothing here computes on
// the CPU except initial scalars, the O(M) marker shuffle (std::random_shuffle in the reference, :182) and row packing.
#include "sweep.cuh"
#include "hyper.cuh"
#include "shard.cuh"
#include "writer.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>

using namespace brr;

namespace {

constexpr int PERM_RING = 3, ROW_RING = 2, KEV_RING = 16;

// the reference's console messages (header: brr_set_message_handler)
std::atomic<brr_message_fn> g_msg_fn{nullptr};
std::atomic<void *> g_msg_ctx{nullptr};

// chain_init carves the chain's buffers out of one device and one page-locked allocation: every cudaMalloc / cudaMallocHost
// is a driver call of milliseconds (and a varying number of them), and a chain needs about thirty.  A buffer that does not
// fit (or is allocated outside chain_init) falls back to its own allocation.
struct Arena {
    uint8_t *base = nullptr; size_t cap = 0, used = 0;
    void *take(size_t bytes)
    {
        const size_t o = (used + 255) / 256 * 256;
        if (!base || o + bytes > cap) return nullptr;
        used = o + bytes;
        return base + o;
    }
};
thread_local Arena *t_dev_arena = nullptr, *t_pin_arena = nullptr;
struct ArenaScope {
    ArenaScope(Arena *d, Arena *h) { t_dev_arena = d; t_pin_arena = h; }
    ~ArenaScope() { t_dev_arena = nullptr; t_pin_arena = nullptr; }
};

template <class T> struct DevBuf {
    T *p = nullptr; size_t n = 0; bool owned = true;
    void alloc(size_t count)
    {
        release(); n = count;
        if (!count) return;
        if (t_dev_arena) if (void *q = t_dev_arena->take(count * sizeof(T))) { p = static_cast<T *>(q); owned = false; return; }
        BRR_CUDA(cudaMalloc(&p, count * sizeof(T))); owned = true;
    }
    void zero(cudaStream_t s = 0) { if (p) BRR_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void upload(const T *h, size_t count) { if (count) BRR_CUDA(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice)); }
    void from(const std::vector<T> &v) { alloc(v.size()); upload(v.data(), v.size()); }
    void release() { if (p && owned) cudaFree(p); p = nullptr; n = 0; owned = true; }
    ~DevBuf() { release(); }
};
template <class T> struct PinBuf {
    T *p = nullptr; size_t n = 0; bool owned = true;
    void alloc(size_t count)
    {
        release(); n = count;
        if (!count) return;
        if (t_pin_arena) if (void *q = t_pin_arena->take(count * sizeof(T))) { p = static_cast<T *>(q); owned = false; return; }
        BRR_CUDA(cudaMallocHost(&p, count * sizeof(T))); owned = true;
    }
    void release() { if (p && owned) cudaFreeHost(p); p = nullptr; n = 0; owned = true; }
    ~PinBuf() { release(); }
};

struct RowSnap {            // one in-flight sample row
    PinBuf<double> row;     // full row, reference layout
    PinBuf<double> scal;    // IterScalars (as bytes) followed by sigmaG[G]
    PinBuf<int> abort;      // the in-kernel watchdog flag as of this snapshot: a row taken after it fired is never delivered
    cudaEvent_t ready = nullptr;
    bool pending = false;
    int64_t it = -1;
};

}  // namespace

struct brr_chain {
    brr_geno *g = nullptr;
    int device = 0;                                               // g->device, kept here: the store may be freed before the chain
    int kind = 0, K = 0, G = 1; int64_t N = 0, M = 0, F = 0;     // N: rows of this rank
    int64_t N_total = 0;                                          // rows of all ranks (== N unless row-sharded)
    brr_comm comm{0, 1, nullptr, nullptr, nullptr};
    Window win;                                                   // exchange window (holds eps; peers write / read it when sharded)
    double *d_eps = nullptr;                                      // = win.eps(rank)
    DevBuf<uint8_t> gtab;                                         // per-marker tables of the current iteration (tables_kernel)
    uint64_t seed = 0; PhiloxKey key{0, 0};
    int max_iterations = 0, burn_in = 0, thinning = 1;
    double sigma0 = 0, v0E = 0, s02E = 0, v0G = 0, s02G = 0;
    double A = 0, vL = 0, vT = 0, c2 = 0, vC = 0, sC = 0;
    std::vector<double> Y, cva, pi_init, fixed, beta0, sigmaGG0, eps0, comp0;
    std::vector<int32_t> gAssign;
    double mu0 = 0, sigmaE0 = 0;
    int gram_impl = 0;
    // geometry
    int B = 128, TW = 1, nW = 1, seg_bytes = 16, rows_per_worker = 64, PS = 128, nb = 0; size_t smem = 0;
    // device state
    DevBuf<double> beta, comp, sigmaG, pi, vcount, betaAcum, d_cva, alpha, d_fixed, fixG, lambda, nu, hs_part;
    DevBuf<double> fin;
    DevBuf<uint64_t> ll;                     // flagged-word hand-over buffers inside the device: [part nW*PS | bcast PS | delta 2*PS+1 | fin 2*nW] slots
    DevBuf<int32_t> d_gAssign, unit0, gram[2];      // block Gram + cross tiles of iterations it (it & 1) and it + 1
    bool dense = false;                             // the store holds dense fp64 columns: fp64 Gram tiles (gramd), 64-marker blocks
    DevBuf<double> gramd[2];
    // Gram pipeline: the Gram of iteration it + 1 depends only on that iteration's marker order, so it runs on its own stream,
    // on the SMs the sweep kernel of iteration it leaves free (a persistent grid of gram_ctas CTAs)
    cudaStream_t gstream = nullptr;
    cudaEvent_t ev_gram0[2] = {}, ev_gram1[2] = {}, ev_sweep_done[2] = {};
    bool sweep_recorded[2] = {false, false};
    int64_t prepared_upto = -1;                      // marker order + Gram are in place for iterations <= this
    int gram_ctas = 0;
    Arena dev_arena, pin_arena;                      // backing store of the buffers chain_init allocates
    DevBuf<IterScalars> sc;
    DevBuf<int> abort_flag;
    DevBuf<long long> prof;
    PinBuf<int32_t> h_perm[PERM_RING]; DevBuf<int32_t> d_perm[PERM_RING]; cudaEvent_t perm_free[PERM_RING] = {}; bool perm_used[PERM_RING] = {};
    std::vector<int32_t> markerI, fixedI;
    // replay
    bool replay = false; int64_t rp_iters = 0, rp_ngam = 0;
    DevBuf<double> rp_u, rp_z, rp_mu, rp_gam, rp_fixz, rp_nu, rp_lam;
    std::vector<int32_t> rp_perm, rp_fixperm; std::vector<double> rp_init_u, rp_init_g, rp_mu_h, rp_gam_h, rp_nu_h;
    // rows
    RowSnap snaps[ROW_RING]; int64_t snap_seq = 0, deliver_seq = 0;
    std::unique_ptr<SampleWriter> writer, bwriter;   // CSV (reference format) and binary (lossless fp64) sinks
    int64_t restored_perm = -1;                      // brr_chain_load: markerI / fixedI already hold the order of this iteration
    int64_t it = 0; bool initialised = false;
    cudaStream_t stream = nullptr; cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0; int64_t last_launches = 0;
    // kernel timing: a fixed ring of KEV_RING iterations x 4 events (set-up | tables + sweep | hyper boundaries); a slot's
    // times are read when the ring comes round to it again (the host never runs more than PERM_RING iterations ahead)
    std::vector<cudaEvent_t> kev;
    double last_kernel_ms[3] = {0, 0, 0};

    int64_t row_len() const
    {
        const int64_t N = N_total;
        switch (kind) {
        case BRR_V2: return 2 * M + 4 + N;                   // reference src/BayesRv2.cpp:136
        case BRR_GROUPS: return 2 * M + 3 + G + N + F + 1;   // src/BayesRv2Groups.cpp:152
        case BRR_GRSTART: return 2 * M + 3 + G + N;          // src/BRv2Grstart.cpp:140
        default: return 2 * M + 4 + N;                       // src/HorseshoeR.cpp:157
        }
    }
    ~brr_chain()
    {
        cudaSetDevice(device);
        for (auto &e : perm_free) if (e) cudaEventDestroy(e);
        for (auto &s : snaps) if (s.ready) cudaEventDestroy(s.ready);
        for (auto &e : kev) if (e) cudaEventDestroy(e);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
        if (gstream) cudaStreamDestroy(gstream);
        for (int i = 0; i < 2; ++i) { if (ev_gram0[i]) cudaEventDestroy(ev_gram0[i]); if (ev_gram1[i]) cudaEventDestroy(ev_gram1[i]); if (ev_sweep_done[i]) cudaEventDestroy(ev_sweep_done[i]); }
        win.release();
        if (dev_arena.base) cudaFree(dev_arena.base);
        if (pin_arena.base) cudaFreeHost(pin_arena.base);
    }
};

namespace {

double host_sum(const double *x, int64_t n) { long double s = 0; for (int64_t i = 0; i < n; ++i) s += x[i]; return (double)s; }
double host_sqnorm(const double *x, int64_t n) { long double s = 0; for (int64_t i = 0; i < n; ++i) s += (long double)x[i] * x[i]; return (double)s; }

double init_uniform(brr_chain *c, int64_t idx)
{
    if (c->replay) { BRR_REQUIRE(idx < (int64_t)c->rp_init_u.size(), BRR_E_ARG, "replay: init_u too short"); return c->rp_init_u[idx]; }
    return draw_uniform(c->key, S_INIT_U, -1, idx);
}
double init_gamma(brr_chain *c, int64_t idx, double shape)
{
    if (c->replay) { BRR_REQUIRE(idx < (int64_t)c->rp_init_g.size(), BRR_E_ARG, "replay: init_g too short"); return c->rp_init_g[idx]; }
    return draw_gamma(c->key, S_INIT_G, -1, idx, shape);
}

void choose_geometry(brr_chain *c, int want_block, int want_workers)
{
    int dev = 0, sms = 0;
    BRR_CUDA(cudaGetDevice(&dev));
    BRR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int kidx = c->kind == BRR_HORSESHOE ? 1 : c->dense ? 0 : c->K == 4 ? 2 : c->K == 3 ? 3 : 0;   // sweep-kernel variant
    BRR_REQUIRE(c->kind == BRR_HORSESHOE || (c->K >= 2 && c->K <= KMAX), BRR_E_SIZE,
                "number of mixture components must be in [2, " + std::to_string(KMAX) + "]");
    const int64_t units = (c->N + 63) / 64;
    int B = want_block ? want_block : 128;
    BRR_REQUIRE(B == 32 || B == 64 || B == 128, BRR_E_ARG, "block must be 32, 64 or 128");
    if (c->dense) B = 64;       // fp64 Gram tiles: 128-marker tiles do not fit beside the tables, and one geometry keeps the instantiations few
    // SMs set aside for the Gram kernel of the next iteration, which runs beside the sweep (none when the caller fixes the workers)
    // The sweep's pace is set by the sampler CTA and the hand-over latencies, not by the workers' throughput (measured flat from 98 to 124
    // workers at every BASELINE shape; it grows mildly with the rows per worker), while the Gram kernel is bound by shared-memory
    // bandwidth per SM: 3.3e-8 ms x markers x local rows / SMs with a 64-marker look-ahead (192 marker rows per operand tile; the
    // traffic per tile is 576 bytes per marker row + 32 KB for the A operand), against ~62 ns (mixture) or ~100 ns (horseshoe) per
    // marker for the sweep.  The Gram kernel gets the SMs that keep it at ~85 % of the sweep's time; the workers take the rest,
    // trimmed to the fewest that reach the same rows per worker, and kept at <= 1024 rows each (TW <= 2: 512-row operand tiles)
    // while that leaves the Gram kernel at least a ninth of the device.
    int gram_sms = 0;
    if (want_workers <= 0 && sms >= 64) {
        const double la_cost = B == 128 ? (576.0 * (128 + lookahead(128)) + 32768.0) / (576.0 * 192 + 32768.0) : 1.0;
        const double per_row = (c->kind == BRR_HORSESHOE ? 4.3e-4 : 6.9e-4) * la_cost;
        gram_sms = (int)std::min<double>(sms / 2, std::max<double>(16.0, std::ceil((double)c->N * per_row)));
    }
    int nW = want_workers > 0 ? want_workers : sms - 1 - SWEEP_REDUCERS - gram_sms;
    if (want_workers <= 0 && units > 0) {
        const int64_t nw_max = std::min<int64_t>(units, nW + 2);
        int64_t maxu = (units + nw_max - 1) / nw_max;
        if (maxu > 16 && (units + 15) / 16 <= sms - 1 - SWEEP_REDUCERS - sms / 9) maxu = 16;
        nW = (int)((units + maxu - 1) / maxu);
    }
    nW = (int)std::max<int64_t>(1, std::min<int64_t>(nW, units));
    while (true) {
        const int64_t maxu = (units + nW - 1) / nW;
        const int words = (int)maxu * 4;
        const int TW = words <= 32 ? 1 : words <= 64 ? 2 : words <= 128 ? 4 : 0;
        BRR_REQUIRE(TW != 0, BRR_E_SIZE, "more than 2048 rows per worker CTA (" + std::to_string(maxu * 64) +
                    "): shard the individuals over more devices");
        const int seg = ((int)maxu | 1) * 16;   // an odd number of 16-byte units per staged column: the unpack's 128-bit reads of consecutive columns are conflict-free
        const size_t smem = sweep_smem_bytes(kidx, B, TW, c->K, c->G, (int)c->F, seg, c->dense);
        if (smem > 227 * 1024) {
            BRR_REQUIRE(B > 32 && !c->dense, BRR_E_SIZE, "sweep kernel does not fit shared memory (reduce K or groups)");
            B /= 2; continue;
        }
        const int cores = sweep_max_coresident(kidx, B, TW, smem, c->dense);
        BRR_REQUIRE(cores >= 2, BRR_E_CUDA, "sweep kernel cannot be made co-resident on this device");
        BRR_REQUIRE(cores >= 2 + SWEEP_REDUCERS, BRR_E_CUDA, "sweep kernel cannot be made co-resident on this device");
        if (nW + 1 + SWEEP_REDUCERS > cores) { nW = cores - 1 - SWEEP_REDUCERS; continue; }
        c->B = B; c->TW = TW; c->nW = nW; c->seg_bytes = seg; c->rows_per_worker = (int)maxu * 64; c->smem = smem;
        c->gram_ctas = std::max(1, sms - (nW + 1 + SWEEP_REDUCERS));
        break;
    }
    c->PS = (int)std::max<int64_t>(c->B, c->F);
    c->nb = (int)((c->M + c->B - 1) / c->B);
    std::vector<int32_t> u0(c->nW + 1);
    for (int w = 0; w <= c->nW; ++w) u0[w] = (int32_t)(units * w / c->nW);
    c->unit0.from(u0);
}

// ---- lazy initialisation: everything the reference does before `for (iteration ...)`
void chain_init(brr_chain *c)
{
    const int64_t N = c->N, M = c->M, F = c->F; const int K = c->K, G = c->G;
    BRR_CUDA(cudaSetDevice(c->g->device));
    SetupTrace tr("chain_init");
    {   // one device and one page-locked allocation for everything below (sizes: the large buffers exactly, the rest bounded)
        const size_t gram_ints = (size_t)c->nb * (gram_tile_entries(c->B) + lookahead(c->B) * c->B);
        const size_t pn = (size_t)c->nb * c->B + (size_t)std::max<int64_t>(F, 1);
        const size_t tab = (size_t)c->nb * sweep_table_bytes(c->kind == BRR_HORSESHOE ? 1 : 0, c->B, K, G, (int)F);
        const size_t ll_words = (2 * (size_t)c->nW * c->PS + (3 * (size_t)c->PS + 1) + 2 * (size_t)c->PS + 2 * (size_t)c->nW) * 2;
        const size_t dev_bytes = 2 * gram_ints * 4 + (c->dense ? 2 * gram_ints * 8 : 0) + tab + PERM_RING * pn * 4 + 8 * (size_t)M * 8 + (size_t)N * F * 8 + ll_words * 8 +
                                 ((size_t)G * (K + 2) * 3 + (size_t)F * (F + 3)) * 8 + 64 * 256 + ((size_t)1 << 16);
        const size_t pin_bytes = PERM_RING * pn * 4 + ROW_RING * ((size_t)c->row_len() + sizeof(IterScalars) / 8 + 1 + (size_t)G) * 8 + 24 * 256;
        if (!c->dev_arena.base) { BRR_CUDA(cudaMalloc(&c->dev_arena.base, dev_bytes)); c->dev_arena.cap = dev_bytes; }
        if (!c->pin_arena.base) { BRR_CUDA(cudaMallocHost(&c->pin_arena.base, pin_bytes)); c->pin_arena.cap = pin_bytes; }
        c->dev_arena.used = 0; c->pin_arena.used = 0;
    }
    ArenaScope arena_scope(&c->dev_arena, &c->pin_arena);
    tr.mark("arenas (one cudaMalloc, one cudaMallocHost)");
    IterScalars sc; memset(&sc, 0, sizeof sc);
    std::vector<double> eps(c->g->Npad, 0.0), beta(M, 0.0), comp(M, 0.0), sigG(G, 0.0), pi((size_t)G * std::max(K, 1), 0.0);
    double mu = 0.0;
    const double n_all = (double)c->N_total;
    // ||v||^2 and sum(v) over the rows of ALL ranks (row-sharded chains: host allreduce at set-up time)
    auto all_sqnorm = [&](const double *x) { double v = host_sqnorm(x, N); comm_allreduce(c->comm, &v, 1); return v; };
    auto all_sum = [&](const double *x) { double v = host_sum(x, N); comm_allreduce(c->comm, &v, 1); return v; };
    if (c->kind == BRR_V2) {
        sigG[0] = init_uniform(c, 0);                                                   // src/BayesRv2.cpp:162
        for (int k = 0; k < K; ++k) pi[k] = c->pi_init[k];                              // :150,:164 (Q1)
        for (int64_t i = 0; i < N; ++i) eps[i] = c->Y[i] - mu - 0.0;                    // :168 (beta == 0)
        sc.sigmaE = all_sqnorm(eps.data()) / n_all * 0.5;                               // :169
    } else if (c->kind == BRR_GROUPS) {
        for (int g = 0; g < G; ++g) { pi[(size_t)g * K] = 0.5; for (int k = 1; k < K; ++k) pi[(size_t)g * K + k] = 0.5 / K; }   // Groups:170-175
        for (int g = 0; g < G; ++g) sigG[g] = init_uniform(c, g);                       // :194-195
        sc.sigmaF = init_uniform(c, G);                                                 // :197
        for (int64_t i = 0; i < N; ++i) eps[i] = c->Y[i] - mu;                          // :203
        sc.sigmaE = all_sqnorm(eps.data()) / n_all * 0.5;                               // :204
    } else if (c->kind == BRR_GRSTART) {
        mu = c->mu0; sc.sigmaE = c->sigmaE0;
        beta = c->beta0; comp = c->comp0; sigG = c->sigmaGG0;
        std::copy(c->eps0.begin(), c->eps0.end(), eps.begin());
        std::vector<double> v((size_t)G * K, 0.0);
        for (int64_t i = 0; i < M; ++i) {                                               // Grstart:159-162
            const int k = (int)comp[i], g = c->gAssign[i];
            BRR_REQUIRE(k >= 0 && k < K, BRR_E_ARG, "components entry outside [0, K)");
            v[(size_t)g * K + k] += 1.0;
        }
        for (int g = 0; g < G; ++g) {                                                   // :163-165
            double s = 0; std::vector<double> gg(K);
            for (int k = 0; k < K; ++k) gg[k] = 1.0 * init_gamma(c, (int64_t)g * (K + 1) + 1 + k, v[(size_t)g * K + k] + 1.0);
            { double s0 = 0, s1 = 0, s2 = 0, s3 = 0; int i = 0;
              for (; i + 4 <= K; i += 4) { s0 += gg[i]; s1 += gg[i + 1]; s2 += gg[i + 2]; s3 += gg[i + 3]; }
              s = (s0 + s2) + (s1 + s3); for (; i < K; ++i) s += gg[i]; }
            for (int k = 0; k < K; ++k) pi[(size_t)g * K + k] = gg[k] / s;
        }
    } else {
        for (int64_t i = 0; i < N; ++i) eps[i] = c->Y[i] - mu - 0.0;                    // HorseshoeR.cpp:186
        sc.sigmaE = all_sqnorm(eps.data()) / n_all * 0.5;                               // :187
        // :171,:176,:179 draw and discard (tau, v, lambda are overwritten at :177,:180,:192); keyed draws need not be consumed
        const double eta0 = 1.0 / ((1.0 / (1 / (sc.sigmaE * std::pow(c->A, 2)))) * init_gamma(c, 2 * M, 0.5));       // :189
        sc.eta = eta0;
        sc.tau = (1.0 / eta0) * (1.0 / ((1.0 / c->vT) * init_gamma(c, 2 * M + 1, 0.5 * c->vT)));                      // :192
        sc.c2 = c->c2;
        std::vector<double> lam(M, 1.0), nu(M);
        // eta and nu of iteration 0 (:217-218) -- the per-iteration kernels draw them for iteration it+1 afterwards
        const double ge = c->replay ? c->rp_gam_h[0] : draw_gamma(c->key, S_GAMMA, 0, 0, 0.5 + 0.5 * c->vT);
        sc.eta_next = 1.0 / ((1.0 / ((1.0 / (sc.sigmaE * c->A * c->A)) + c->vT / sc.tau)) * ge);
        for (int64_t j = 0; j < M; ++j) {
            const double gn = c->replay ? c->rp_nu_h[j] : draw_gamma(c->key, S_HS_NU, 0, j, 0.5 + 0.5 * c->vL);
            nu[j] = 1.0 / ((1.0 / (c->vL / lam[j] + 1.0)) * gn);
        }
        c->lambda.from(lam); c->nu.from(nu);
        c->hs_part.alloc((size_t)((M + 255) / 256) * 2);
    }
    // first intercept draw (:177-179 of iteration 0)
    {
        const double es = all_sum(eps.data()), n = n_all;
        const double z = c->replay ? c->rp_mu_h[0] : draw_normal(c->key, S_MU, 0, 0);
        sc.mu = mu;
        sc.mu_next = (es + n * mu) / n + std::sqrt(sc.sigmaE / n) * z;
        sc.shift = mu - sc.mu_next;
        sc.eps_sum = es + n * sc.shift;
    }
    BRR_CUDA(cudaMemcpy(c->d_eps, eps.data(), eps.size() * 8, cudaMemcpyHostToDevice));
    c->beta.from(beta); c->comp.from(comp); c->sigmaG.from(sigG); c->pi.from(pi);
    c->vcount.alloc((size_t)G * std::max(K, 1)); c->vcount.zero(); c->betaAcum.alloc(G); c->betaAcum.zero();
    if (c->kind != BRR_HORSESHOE) c->d_cva.from(c->cva);
    if (!c->gAssign.empty()) c->d_gAssign.from(c->gAssign);
    if (F > 0) {
        c->d_fixed.from(c->fixed);
        std::vector<double> fg((size_t)F * F + F, 0.0), al(F, 0.0);
        for (int64_t a = 0; a < F; ++a) {
            for (int64_t b = 0; b < F; ++b) {
                long double s = 0; const double *fa = &c->fixed[a * N], *fb = &c->fixed[b * N];
                for (int64_t i = 0; i < N; ++i) s += (long double)fa[i] * fb[i];
                fg[a * F + b] = (double)s;
            }
            fg[(size_t)F * F + a] = host_sum(&c->fixed[a * N], N);
        }
        comm_allreduce(c->comm, fg.data(), (int64_t)fg.size());
        c->fixG.from(fg); c->alpha.from(al);
    }
    tr.mark("initial state on the host and its upload");
    std::vector<IterScalars> scv(1, sc); c->sc.from(scv);
    c->ll.alloc((2 * (size_t)c->nW * c->PS + (3 * (size_t)c->PS + 1) + 2 * (size_t)c->PS + 2 * (size_t)c->nW) * 2); c->ll.zero();
    c->fin.alloc(2); c->fin.zero();
    c->abort_flag.alloc(1); c->abort_flag.zero();
    c->prof.alloc(16); c->prof.zero();
    for (auto &gb : c->gram) gb.alloc((size_t)c->nb * (gram_tile_entries(c->B) + lookahead(c->B) * c->B));      // self tiles (block-upper trapezoids, common.cuh), then the look-ahead cross tiles
    if (c->dense) for (auto &gb : c->gramd) gb.alloc((size_t)c->nb * (gram_tile_entries(c->B) + lookahead(c->B) * c->B));
    c->gtab.alloc((size_t)c->nb * sweep_table_bytes(c->kind == BRR_HORSESHOE ? 1 : 0, c->B, K, G, (int)F));
    tr.mark("device buffers (hand-over words, Gram x2, tables)");
    const size_t pn = (size_t)c->nb * c->B + (size_t)std::max<int64_t>(F, 1);
    for (int i = 0; i < PERM_RING; ++i) {
        c->h_perm[i].alloc(pn); c->d_perm[i].alloc(pn);
        BRR_CUDA(cudaEventCreateWithFlags(&c->perm_free[i], cudaEventDisableTiming));
    }
    c->markerI.resize(M); for (int64_t i = 0; i < M; ++i) c->markerI[i] = (int32_t)i;   // :137-140
    c->fixedI.resize(F); for (int64_t i = 0; i < F; ++i) c->fixedI[i] = (int32_t)i;
    for (auto &s : c->snaps) {
        s.row.alloc((size_t)c->row_len()); s.scal.alloc(sizeof(IterScalars) / 8 + 1 + (size_t)G); s.abort.alloc(1);
        BRR_CUDA(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
    }
    tr.mark("pinned rings (marker order, sample rows)");
    BRR_CUDA(cudaDeviceSynchronize());
    tr.mark("device synchronize");
    c->initialised = true;
}

std::string watchdog_message(int flag)
{
    return "in-kernel watchdog fired (code " + std::to_string(flag) + ": 2 bulk copy, 3 Gram sum over ranks, 4 Gram bulk copy, 10-19 sweep hand-overs: "
           "11 partial dots, 12 deltas, 13 flagged word, 14 fixed-effect totals, 15 block dots, 16 end-of-sweep sums, 17 fixed-point range, "
           "18 / 19 look-ahead correction)";
}

void deliver_row(brr_chain *c, RowSnap &s, double *rows, int64_t max_rows, int64_t *n_rows)
{
    BRR_CUDA(cudaEventSynchronize(s.ready));
    BRR_REQUIRE(*s.abort.p == 0, BRR_E_CUDA, watchdog_message(*s.abort.p));       // no row of a chain whose sweep was aborted reaches a sink
    const int64_t N = c->N_total, M = c->M, F = c->F; const int G = c->G;
    IterScalars sc; memcpy(&sc, s.scal.p, sizeof sc);
    const double *sg = s.scal.p + sizeof(IterScalars) / 8 + 1;
    double *r = s.row.p;
    r[0] = (double)s.it; r[1] = sc.mu;
    switch (c->kind) {
    case BRR_V2: r[2 + M] = sc.sigmaE; r[3 + M] = sg[0]; break;                         // src/BayesRv2.cpp:260
    case BRR_GROUPS:                                                                     // src/BayesRv2Groups.cpp:317
        r[2 + M] = sc.sigmaE; for (int g = 0; g < G; ++g) r[3 + 2 * M + g] = sg[g];
        r[3 + 2 * M + G + N + F] = sc.sigmaF; break;
    case BRR_GRSTART: r[2 + M] = sc.sigmaE; for (int g = 0; g < G; ++g) r[3 + 2 * M + g] = sg[g]; break;   // src/BRv2Grstart.cpp:267
    default: r[2 + M] = sc.sigmaE; r[3 + M] = sc.tau; break;                             // src/HorseshoeR.cpp:258
    }
    const int64_t L = c->row_len();
    if (rows && *n_rows < max_rows) memcpy(rows + *n_rows * L, r, (size_t)L * 8);
    if (c->writer) c->writer->enqueue(r, (size_t)L);                                     // q.enqueue(sample)  :261
    if (c->bwriter) c->bwriter->enqueue(r, (size_t)L);
    ++*n_rows;
    s.pending = false;
}

void snapshot_row(brr_chain *c, int64_t it, double *rows, int64_t max_rows, int64_t *n_rows)
{
    RowSnap &s = c->snaps[c->snap_seq % ROW_RING];
    if (s.pending) { deliver_row(c, s, rows, max_rows, n_rows); ++c->deliver_seq; }
    const int64_t N = c->N_total, M = c->M, F = c->F; const int G = c->G;
    double *r = s.row.p; cudaStream_t st = c->stream;
    auto d2h = [&](double *dst, const double *src, int64_t n) { if (n) BRR_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 8, cudaMemcpyDeviceToHost, st)); };
    // residuals of every rank, straight from the peers' windows: complete once this rank's sweep has ended (it waited for
    // every rank's end-of-sweep sums) and stable until this rank takes part in the next sweep (stream order)
    auto eps_all = [&](double *dst) { for (int q = 0; q < c->win.R; ++q) d2h(dst + c->win.row0[q], c->win.eps(q), c->win.n_rows[q]); };
    d2h(r + 2, c->beta.p, M);
    switch (c->kind) {
    case BRR_V2: d2h(r + 4 + M, c->comp.p, M); eps_all(r + 4 + 2 * M); break;
    case BRR_GROUPS: d2h(r + 3 + M, c->comp.p, M); eps_all(r + 3 + 2 * M + G); d2h(r + 3 + 2 * M + G + N, c->alpha.p, F); break;
    case BRR_GRSTART: d2h(r + 3 + M, c->comp.p, M); eps_all(r + 3 + 2 * M + G); break;
    default: d2h(r + 4 + M, c->lambda.p, M); eps_all(r + 4 + 2 * M); break;
    }
    BRR_CUDA(cudaMemcpyAsync(s.scal.p, c->sc.p, sizeof(IterScalars), cudaMemcpyDeviceToHost, st));
    d2h(s.scal.p + sizeof(IterScalars) / 8 + 1, c->sigmaG.p, G);
    BRR_CUDA(cudaMemcpyAsync(s.abort.p, c->abort_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    BRR_CUDA(cudaEventRecord(s.ready, st));
    s.pending = true; s.it = it;
    ++c->snap_seq;
}

void run_iterations_body(brr_chain *c, int n_iter, int emit_all, double *rows, int64_t max_rows, int64_t *n_rows);

void run_iterations(brr_chain *c, int n_iter, int emit_all, double *rows, int64_t max_rows, int64_t *n_rows)
{
    try { run_iterations_body(c, n_iter, emit_all, rows, max_rows, n_rows); }
    catch (...) {   // nothing of this call may still be in flight when the error reaches the caller (who may destroy the chain)
        if (c->stream) cudaStreamSynchronize(c->stream);
        if (c->gstream) cudaStreamSynchronize(c->gstream);
        for (auto &s : c->snaps) s.pending = false;
        c->deliver_seq = c->snap_seq;
        throw;
    }
}

void run_iterations_body(brr_chain *c, int n_iter, int emit_all, double *rows, int64_t max_rows, int64_t *n_rows)
{
    BRR_CUDA(cudaSetDevice(c->g->device));
    if (!c->initialised) chain_init(c);
    const bool sharded = c->win.R > 1;
    if (sharded) { double token = 1.0; comm_allreduce(c->comm, &token, 1); }    // ranks enter the launch sequence together
    const int64_t M = c->M, F = c->F; const int K = c->K, G = c->G;
    const int kk = c->kind == BRR_HORSESHOE ? 1 : c->dense ? 0 : K == 4 ? 2 : K == 3 ? 3 : 0;   // sweep-kernel variant
    int64_t launches = 0;
    c->prof.zero(c->stream);
    while (c->kev.size() < (size_t)4 * KEV_RING) { cudaEvent_t e; BRR_CUDA(cudaEventCreate(&e)); c->kev.push_back(e); }
    c->last_kernel_ms[0] = c->last_kernel_ms[1] = c->last_kernel_ms[2] = 0;
    auto harvest = [&](int slot) {      // add the kernel times of the iteration that used ring slot `slot`
        BRR_CUDA(cudaEventSynchronize(c->kev[4 * slot + 3]));
        for (int k = 1; k < 3; ++k) {
            float t = 0; BRR_CUDA(cudaEventElapsedTime(&t, c->kev[4 * slot + k], c->kev[4 * slot + k + 1]));
            c->last_kernel_ms[k] += t;
        }
    };
    BRR_CUDA(cudaEventRecord(c->ev0, c->stream));
    const size_t self_ints = (size_t)c->nb * gram_tile_entries(c->B), all_ints = self_ints + (size_t)c->nb * lookahead(c->B) * c->B;
    const size_t fo = (size_t)c->nb * c->B;
    // Marker order (host shuffle, reference :182) + block Gram of iteration j, on the Gram stream.  Called once per iteration,
    // in order, one iteration ahead of the sweep.
    auto prepare = [&](int64_t j) {
        const bool rp = c->replay;
        const int slot = (int)(j % PERM_RING), gb = (int)(j & 1);
        if (c->perm_used[slot]) BRR_CUDA(cudaEventSynchronize(c->perm_free[slot]));
        int32_t *hp = c->h_perm[slot].p;
        const bool reshuffle = j != c->restored_perm;   // a loaded checkpoint already holds this iteration's order
        if (rp) memcpy(hp, &c->rp_perm[(size_t)j * M], (size_t)M * 4);
        else { if (reshuffle) shuffle_host(c->key, S_PERM, j, c->markerI.data(), M); memcpy(hp, c->markerI.data(), (size_t)M * 4); }   // :182
        if (F > 0) {
            if (rp) memcpy(hp + fo, &c->rp_fixperm[(size_t)j * F], (size_t)F * 4);
            else { if (reshuffle) shuffle_host(c->key, S_FIXPERM, j, c->fixedI.data(), F); memcpy(hp + fo, c->fixedI.data(), (size_t)F * 4); }   // Groups:216
        }
        // the Gram buffer j & 1 was last read by the sweep of iteration j - 2
        if (c->sweep_recorded[gb]) BRR_CUDA(cudaStreamWaitEvent(c->gstream, c->ev_sweep_done[gb], 0));
        BRR_CUDA(cudaMemcpyAsync(c->d_perm[slot].p, hp, (size_t)(M) * 4, cudaMemcpyHostToDevice, c->gstream));
        if (F > 0) BRR_CUDA(cudaMemcpyAsync(c->d_perm[slot].p + fo, hp + fo, (size_t)F * 4, cudaMemcpyHostToDevice, c->gstream));
        c->perm_used[slot] = true;
        BRR_CUDA(cudaEventRecord(c->ev_gram0[gb], c->gstream));
        if (c->dense) {   // exact int32 counts of the packed pairs, then the fp64 tiles (pairs with a dense column: fp64 dots)
            int32_t *gi = c->gram[gb].p;
            launch_gram(c->g, c->d_perm[slot].p, M, c->B, c->gram_impl, gi, gi + self_ints, c->gstream, c->gram_ctas, c->abort_flag.p);
            double *out = sharded ? reinterpret_cast<double *>(c->win.gram(c->win.rank, gb)) : c->gramd[gb].p;
            launch_gram_dense(c->g, c->d_perm[slot].p, M, c->B, gi, gi + self_ints, out, out + self_ints, c->gstream, c->gram_ctas);
            ++launches;
            if (sharded) { launch_gram_allsum(c->win, gb, (uint32_t)j + 1u, c->gramd[gb].p, all_ints, true, c->abort_flag.p, c->gstream, c->gram_ctas); ++launches; }
        } else if (sharded) {   // partial Gram over this rank's rows into the window, then the exact sum over all ranks (peer reads)
            int32_t *part = c->win.gram(c->win.rank, gb);
            launch_gram(c->g, c->d_perm[slot].p, M, c->B, c->gram_impl, part, part + self_ints, c->gstream, c->gram_ctas, c->abort_flag.p);
            launch_gram_allsum(c->win, gb, (uint32_t)j + 1u, c->gram[gb].p, all_ints, false, c->abort_flag.p, c->gstream, c->gram_ctas);
            ++launches;
        } else launch_gram(c->g, c->d_perm[slot].p, M, c->B, c->gram_impl, c->gram[gb].p, c->gram[gb].p + self_ints, c->gstream, c->gram_ctas, c->abort_flag.p);
        BRR_CUDA(cudaEventRecord(c->ev_gram1[gb], c->gstream));
        c->prepared_upto = j;
        ++launches;
    };
    auto can_prepare = [&](int64_t j) { return !c->replay || j < c->rp_iters; };
    for (int n = 0; n < n_iter; ++n) {
        const int64_t it = c->it;
        const bool rp = c->replay;
        if (rp) BRR_REQUIRE(it < c->rp_iters, BRR_E_ARG, "replay tables exhausted");
        const int slot = (int)(it % PERM_RING), gb = (int)(it & 1);
        if (c->prepared_upto < it) prepare(it);
        const int ks = n % KEV_RING;
        if (n >= KEV_RING) harvest(ks);
        BRR_CUDA(cudaEventRecord(c->kev[4 * ks], c->stream));
        BRR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_gram1[gb], 0));
        c->ll.zero(c->stream);
        BRR_CUDA(cudaEventRecord(c->kev[4 * ks + 1], c->stream));

        SweepParams p; memset(&p, 0, sizeof p);
        const brr_geno *g = c->g;
        p.packed = g->d_packed; p.stride = g->stride; p.N = g->N;
        p.colA = g->d_a; p.colD = g->d_d; p.colS = g->d_S; p.colXsq = g->d_xsq; p.colCsum = g->d_csum; p.n_total = g->n_total;
        p.perm = c->d_perm[slot].p; p.gram = c->gram[gb].p; p.M = M; p.nb = c->nb; p.it = it;
        p.eps = c->d_eps; p.beta = c->beta.p; p.comp = c->comp.p; p.sc = c->sc.p;
        p.K = K; p.G = G; p.gAssign = c->d_gAssign.p; p.cva = c->d_cva.p; p.sigmaG = c->sigmaG.p; p.pi = c->pi.p;
        p.vcount = c->vcount.p; p.betaAcum = c->betaAcum.p; p.lambda = c->lambda.p;
        p.key = c->key;
        p.tbl_u = rp && c->rp_u.p ? c->rp_u.p + (size_t)it * M : nullptr;
        p.tbl_z = rp && c->rp_z.p ? c->rp_z.p + (size_t)it * M : nullptr;
        p.F = (int)F; p.fixed = c->d_fixed.p; p.fixperm = c->d_perm[slot].p + fo; p.fixG = c->fixG.p; p.alpha = c->alpha.p;
        p.tbl_fix_z = rp && F > 0 && c->rp_fixz.p ? c->rp_fixz.p + (size_t)it * F : nullptr;
        p.ll_part = c->ll.p; p.ll_bcast = p.ll_part + 2 * (size_t)c->nW * c->PS * 2; p.ll_delta = p.ll_bcast + (3 * (size_t)c->PS + 1) * 2;
        p.ll_fin = p.ll_delta + 2 * (size_t)c->PS * 2;
        p.xgram = c->gram[gb].p + self_ints;
        if (c->dense) {
            p.gramd = c->gramd[gb].p; p.xgramd = c->gramd[gb].p + self_ints;
            p.dense = g->d_dense; p.denseIdx = g->d_dense_idx; p.Npad = g->Npad;
        }
        p.abort_flag = c->abort_flag.p; p.prof = c->prof.p; p.fin = c->fin.p;
        p.rank = c->win.rank; p.R = c->win.R;
        for (int q = 0; q < c->win.R; ++q) { p.xred[q] = c->win.xred(q); p.xfin[q] = c->win.xfin(q); }
        p.xphase0 = (uint32_t)((uint64_t)it * (uint64_t)(c->nb + (F > 0 ? 1 : 0)));
        p.nW = c->nW; p.nR = SWEEP_REDUCERS; p.PS = c->PS; p.unit0 = c->unit0.p; p.seg_bytes = c->seg_bytes;
        p.gtab = c->gtab.p;
        launch_tables(kk, c->B, p, c->gtab.p, c->stream);
        launch_sweep(kk, c->B, c->TW, p, c->smem, c->stream);
        BRR_CUDA(cudaEventRecord(c->kev[4 * ks + 2], c->stream));
        BRR_CUDA(cudaEventRecord(c->perm_free[slot], c->stream));
        BRR_CUDA(cudaEventRecord(c->ev_sweep_done[gb], c->stream));
        c->sweep_recorded[gb] = true;
        // next iteration's marker order and Gram: beside this sweep, on the SMs it leaves free
        if (c->prepared_upto < it + 1 && can_prepare(it + 1)) prepare(it + 1);

        HyperParams h; memset(&h, 0, sizeof h);
        h.kind = c->kind; h.it = it; h.n_total = g->n_total; h.M = M; h.K = K; h.G = G; h.F = F;
        h.v0E = c->v0E; h.s02E = c->s02E; h.v0G = c->v0G; h.s02G = c->s02G;
        h.sc = c->sc.p; h.sigmaG = c->sigmaG.p; h.pi = c->pi.p; h.vcount = c->vcount.p; h.betaAcum = c->betaAcum.p;
        h.beta = c->beta.p; h.alpha = c->alpha.p; h.fin = c->fin.p; h.nW = 1; h.key = c->key;
        const bool next_in = rp && it + 1 < c->rp_iters;
        h.tbl_gam = rp && c->rp_gam.p ? c->rp_gam.p + (size_t)it * c->rp_ngam : nullptr;
        h.tbl_gam_next = next_in && c->rp_gam.p ? c->rp_gam.p + (size_t)(it + 1) * c->rp_ngam : nullptr;
        h.tbl_mu_z_next = next_in && c->rp_mu.p ? c->rp_mu.p + (it + 1) : nullptr;
        h.A = c->A; h.vL = c->vL; h.vT = c->vT; h.vC = c->vC; h.sC = c->sC;
        h.lambda = c->lambda.p; h.nu = c->nu.p; h.hs_part = c->hs_part.p;
        h.tbl_hs_lam = rp && c->rp_lam.p ? c->rp_lam.p + (size_t)it * M : nullptr;
        h.tbl_hs_nu_next = next_in && c->rp_nu.p ? c->rp_nu.p + (size_t)(it + 1) * M : nullptr;
        launch_hyper(h, c->stream);
        BRR_CUDA(cudaEventRecord(c->kev[4 * ks + 3], c->stream));
        launches += 2 + hyper_launch_count(c->kind);

        if (emit_all || (it >= c->burn_in && it % c->thinning == 0))                     // :257-259
            snapshot_row(c, it, rows, max_rows, n_rows);
        ++c->it;
    }
    BRR_CUDA(cudaEventRecord(c->ev1, c->stream));
    BRR_CUDA(cudaStreamSynchronize(c->stream));
    BRR_CUDA(cudaStreamSynchronize(c->gstream));
    if (sharded) { double token = 1.0; comm_allreduce(c->comm, &token, 1); }    // no rank leaves while a peer may still be copying its residuals
    {
        int flag = 0;
        BRR_CUDA(cudaMemcpy(&flag, c->abort_flag.p, sizeof(int), cudaMemcpyDeviceToHost));
        BRR_REQUIRE(flag == 0, BRR_E_CUDA, watchdog_message(flag));
    }
    while (c->deliver_seq < c->snap_seq) {
        RowSnap &s = c->snaps[c->deliver_seq % ROW_RING];
        if (s.pending) deliver_row(c, s, rows, max_rows, n_rows);
        ++c->deliver_seq;
    }
    float ms = 0; BRR_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_ms = ms; c->last_launches = launches;
    for (int n = std::max(0, n_iter - KEV_RING); n < n_iter; ++n) harvest(n % KEV_RING);
    // the Gram kernel runs beside the sweep of the previous iteration: its own duration is only known for the two most recent
    // preparations (their events are still in place); scale to the run
    {
        double sum = 0; int cnt = 0;
        for (int gb = 0; gb < 2; ++gb) {
            float t = 0;
            if (cudaEventElapsedTime(&t, c->ev_gram0[gb], c->ev_gram1[gb]) == cudaSuccess) { sum += t; ++cnt; }
        }
        (void)cudaGetLastError();
        c->last_kernel_ms[0] = cnt ? sum / cnt * n_iter : 0.0;
    }
}

void check_iters(int max_iterations, int burn_in, int thinning)
{
    // the only validation of the reference that returns (src/BayesRv2.cpp:76-80)
    BRR_REQUIRE(!(max_iterations < burn_in || max_iterations < 1 || burn_in < 1), BRR_E_ITER,
                "error: burn_in has to be a positive integer and smaller than the maximum number of iterations");
    BRR_REQUIRE(thinning >= 1, BRR_E_ARG, "thinning must be >= 1 (the reference divides by it, src/BayesRv2.cpp:259)");
}

}  // namespace

// =================================================================================================
static int chain_create_impl(const brr_config *cfg, brr_geno *g, const brr_comm *comm, brr_chain **out)
{
    return guarded([&] {
        BRR_REQUIRE(cfg && g && out, BRR_E_ARG, "brr_chain_create: null pointer");
        if (comm) {
            BRR_REQUIRE(comm->world >= 1 && comm->world <= BRR_MAX_WORLD && comm->rank >= 0 && comm->rank < comm->world, BRR_E_SIZE,
                        "brr_comm: world must be in [1, " + std::to_string(BRR_MAX_WORLD) + "] and rank in [0, world)");
            BRR_REQUIRE(comm->world == 1 || (comm->allreduce_sum && comm->allgather), BRR_E_ARG, "brr_comm: call-backs missing");
        }
        BRR_REQUIRE(cfg->kind >= BRR_V2 && cfg->kind <= BRR_HORSESHOE, BRR_E_ARG, "unknown sampler kind");
        check_iters(cfg->max_iterations, cfg->burn_in, cfg->thinning);
        BRR_REQUIRE(!g->pending_impute, BRR_E_ARG, "this .bed row shard still holds missing genotypes: brr_geno_shard_stats decides their fill "
                    "value from the counts of all ranks and must be called first");
        require_device(g->device);
        SetupTrace tr("chain_create");
        std::unique_ptr<brr_chain> c(new brr_chain());
        c->g = g; c->device = g->device; c->kind = cfg->kind; c->N = g->N; c->M = g->M;
        c->dense = g->Md > 0;
        if (comm) c->comm = *comm;
        c->seed = cfg->seed; c->key = PhiloxKey{ (uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32) };
        c->max_iterations = cfg->max_iterations; c->burn_in = cfg->burn_in; c->thinning = cfg->thinning;
        c->sigma0 = cfg->sigma0; c->v0E = cfg->v0E; c->s02E = cfg->s02E; c->v0G = cfg->v0G; c->s02G = cfg->s02G;
        c->gram_impl = cfg->gram_impl;
        const int64_t N = c->N, M = c->M;
        if (cfg->kind != BRR_GRSTART) { BRR_REQUIRE(cfg->Y, BRR_E_ARG, "Y is null"); c->Y.assign(cfg->Y, cfg->Y + N); }
        if (cfg->kind == BRR_HORSESHOE) {
            c->K = 0; c->G = 1;
            c->A = cfg->A; c->vL = cfg->vL; c->vT = cfg->vT; c->c2 = cfg->c2; c->vC = cfg->vC; c->sC = cfg->sC;
        } else {
            BRR_REQUIRE(cfg->cva && cfg->ncva >= 1, BRR_E_ARG, "cva is null or empty");
            c->K = cfg->ncva + 1;
            c->G = cfg->kind == BRR_V2 ? 1 : cfg->groups;
            BRR_REQUIRE(c->G >= 1, BRR_E_ARG, "groups must be >= 1");
            BRR_REQUIRE((int64_t)c->G * (c->K + 1) <= 4096, BRR_E_SIZE, "groups x components too large for the in-kernel tables");
            c->cva.assign(cfg->cva, cfg->cva + (size_t)c->G * cfg->ncva);
            if (cfg->kind != BRR_V2) {
                BRR_REQUIRE(cfg->gAssign, BRR_E_ARG, "gAssign is null");
                c->gAssign.assign(cfg->gAssign, cfg->gAssign + M);
                for (int64_t i = 0; i < M; ++i) BRR_REQUIRE(c->gAssign[i] >= 0 && c->gAssign[i] < c->G, BRR_E_ARG, "gAssign entry outside [0, groups)");
            }
            if (cfg->kind == BRR_V2) {
                c->pi_init.resize(c->K);
                if (cfg->pi_init) std::copy(cfg->pi_init, cfg->pi_init + c->K, c->pi_init.begin());
                else {   // evident intent of src/BayesRv2.cpp:148-150
                    double s = 0; for (int k = 0; k < cfg->ncva; ++k) s += cfg->cva[k];
                    c->pi_init[0] = 0.5; for (int k = 1; k < c->K; ++k) c->pi_init[k] = 0.5 * cfg->cva[k - 1] / s;
                }
            }
            if (cfg->kind == BRR_GROUPS && cfg->F > 0) {
                BRR_REQUIRE(cfg->fixed, BRR_E_ARG, "fixed is null but F > 0");
                c->F = cfg->F; c->fixed.assign(cfg->fixed, cfg->fixed + (size_t)N * cfg->F);
            }
            if (cfg->kind == BRR_GRSTART) {
                BRR_REQUIRE(cfg->beta0 && cfg->sigmaGG0 && cfg->epsilon0 && cfg->components0, BRR_E_ARG, "restart state has a null pointer");
                c->mu0 = cfg->mu0; c->sigmaE0 = cfg->sigmaE0;
                c->beta0.assign(cfg->beta0, cfg->beta0 + M); c->sigmaGG0.assign(cfg->sigmaGG0, cfg->sigmaGG0 + c->G);
                c->eps0.assign(cfg->epsilon0, cfg->epsilon0 + N); c->comp0.assign(cfg->components0, cfg->components0 + M);
            }
        }
        tr.mark("argument copies");
        choose_geometry(c.get(), cfg->block, cfg->workers);
        tr.mark("geometry (occupancy queries)");
        const int R = c->comm.world;
        if (R > 1) {   // every rank must cut the chain into the same Gibbs blocks: agree on the smallest block any rank chose
            std::vector<double> bs(R, 0.0);
            bs[c->comm.rank] = (double)c->B;
            comm_allreduce(c->comm, bs.data(), R);
            const int bmin = (int)*std::min_element(bs.begin(), bs.end());
            if (bmin != c->B) choose_geometry(c.get(), bmin, cfg->workers);
            BRR_REQUIRE(c->B == bmin, BRR_E_SIZE, "ranks of a sharded chain cannot agree on a Gibbs block size");
        }
        {   // every kernel of the iteration loop is loaded now, not on its first launch (common.cuh, preload_kernel)
            const int kidx = c->kind == BRR_HORSESHOE ? 1 : 0;
            preload_tables(kidx); preload_gram(c->B, c->gram_impl); preload_hyper(c->kind);
            if (R > 1) preload_allsum();
            if (c->dense) preload_gram_dense();
        }
        tr.mark("kernel preload");
        c->win.rank = c->comm.rank; c->win.R = R;
        c->win.layout(c->PS, c->nb, c->B, g->Npad, c->dense ? 8 : 4);
        c->win.allocate();
        c->win.connect(c->comm, g->device, g->N, c->B, c->kind, c->M);
        c->N_total = c->win.n_total;
        c->d_eps = c->win.eps(c->win.rank);
        if (c->win.colocated > 1) {   // ranks sharing one device: every rank's sweep and Gram grids must fit on it side by side
            int sms = 0;
            BRR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->device));
            c->gram_ctas = std::max(1, (sms - c->win.colocated * (c->nW + 1 + SWEEP_REDUCERS)) / c->win.colocated);
        }
        tr.mark("exchange window");
        if (R > 1) BRR_REQUIRE((double)c->N_total == g->n_total, BRR_E_ARG,
                               "the genotype store of a sharded chain needs brr_geno_shard_stats first (its statistics must cover all ranks' rows)");
        BRR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        BRR_CUDA(cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            BRR_CUDA(cudaEventCreate(&c->ev_gram0[i])); BRR_CUDA(cudaEventCreate(&c->ev_gram1[i]));
            BRR_CUDA(cudaEventCreateWithFlags(&c->ev_sweep_done[i], cudaEventDisableTiming));
        }
        BRR_CUDA(cudaEventCreate(&c->ev0)); BRR_CUDA(cudaEventCreate(&c->ev1));
        tr.mark("streams and events");
        *out = c.release();
    });
}

extern "C" int brr_chain_create(const brr_config *cfg, brr_geno *g, brr_chain **out) { return chain_create_impl(cfg, g, nullptr, out); }
extern "C" int brr_chain_create_sharded(const brr_config *cfg, brr_geno *g, const brr_comm *comm, brr_chain **out)
{
    if (!comm) { set_last_error("brr_chain_create_sharded: comm is null"); return BRR_E_ARG; }
    return chain_create_impl(cfg, g, comm, out);
}

extern "C" int brr_chain_set_replay(brr_chain *c, const brr_replay *r)
{
    return guarded([&] {
        BRR_REQUIRE(c && r, BRR_E_ARG, "null pointer");
        BRR_REQUIRE(!c->initialised, BRR_E_ARG, "replay tables must be set before the first brr_chain_run");
        BRR_REQUIRE(r->M == c->M && r->n_iter >= 1 && r->perm && r->mu_z, BRR_E_ARG, "replay tables do not match the chain");
        BRR_CUDA(cudaSetDevice(c->g->device));
        const size_t T = (size_t)r->n_iter, M = (size_t)c->M, F = (size_t)c->F;
        c->rp_iters = r->n_iter; c->rp_ngam = r->n_gam;
        c->rp_perm.assign(r->perm, r->perm + T * M);
        for (size_t i = 0; i < T * M; ++i) BRR_REQUIRE(c->rp_perm[i] >= 0 && c->rp_perm[i] < c->M, BRR_E_ARG, "replay perm entry out of range");
        if (F) { BRR_REQUIRE(r->fixperm && r->fix_z && r->F == c->F, BRR_E_ARG, "replay: fixed-effect tables missing");
                 c->rp_fixperm.assign(r->fixperm, r->fixperm + T * F); c->rp_fixz.alloc(T * F); c->rp_fixz.upload(r->fix_z, T * F); }
        c->rp_mu_h.assign(r->mu_z, r->mu_z + T); c->rp_mu.alloc(T); c->rp_mu.upload(r->mu_z, T);
        if (r->mark_u) { c->rp_u.alloc(T * M); c->rp_u.upload(r->mark_u, T * M); }
        if (r->mark_z) { c->rp_z.alloc(T * M); c->rp_z.upload(r->mark_z, T * M); }
        BRR_REQUIRE(r->gam && r->n_gam >= 1, BRR_E_ARG, "replay: gamma table missing");
        c->rp_gam_h.assign(r->gam, r->gam + T * (size_t)r->n_gam);
        c->rp_gam.alloc(T * (size_t)r->n_gam); c->rp_gam.upload(r->gam, T * (size_t)r->n_gam);
        if (c->kind == BRR_HORSESHOE) {
            BRR_REQUIRE(r->hs_nu && r->hs_lam, BRR_E_ARG, "replay: horseshoe tables missing");
            c->rp_nu_h.assign(r->hs_nu, r->hs_nu + M);
            c->rp_nu.alloc(T * M); c->rp_nu.upload(r->hs_nu, T * M);
            c->rp_lam.alloc(T * M); c->rp_lam.upload(r->hs_lam, T * M);
        } else BRR_REQUIRE(r->mark_u, BRR_E_ARG, "replay: mark_u missing");
        if (r->init_u && r->n_init_u > 0) c->rp_init_u.assign(r->init_u, r->init_u + r->n_init_u);
        if (r->init_g && r->n_init_g > 0) c->rp_init_g.assign(r->init_g, r->init_g + r->n_init_g);
        c->replay = true;
    });
}

extern "C" int brr_chain_open_output(brr_chain *c, const char *path)
{
    return guarded([&] {
        BRR_REQUIRE(c && path, BRR_E_ARG, "null pointer");
        c->writer.reset(new SampleWriter(path, sample_header(c->kind, c->N_total, c->M, c->G, c->F), c->kind != BRR_HORSESHOE));
    });
}
extern "C" int brr_chain_open_binary_output(brr_chain *c, const char *path)
{
    return guarded([&] {
        BRR_REQUIRE(c && path, BRR_E_ARG, "null pointer");
        c->bwriter.reset(new SampleWriter(path, binary_sample_header(c->kind, c->N_total, c->M, c->G, c->F, c->row_len()), true, true));
    });
}
extern "C" int brr_chain_close_output(brr_chain *c)
{
    return guarded([&] {
        BRR_REQUIRE(c, BRR_E_ARG, "null pointer");
        if (c->writer) { std::unique_ptr<SampleWriter> w(std::move(c->writer)); w->finish(); }
        if (c->bwriter) { std::unique_ptr<SampleWriter> w(std::move(c->bwriter)); w->finish(); }
    });
}
extern "C" int64_t brr_chain_row_len(const brr_chain *c) { return c ? c->row_len() : -1; }

extern "C" int brr_chain_run(brr_chain *c, int n_iter, int emit_all, double *rows, int64_t max_rows, int64_t *n_rows)
{
    const int rc = guarded([&] {
        BRR_REQUIRE(c && n_iter >= 0, BRR_E_ARG, "bad arguments");
        int64_t produced = 0;
        run_iterations(c, n_iter, emit_all, rows, max_rows, &produced);
        if (n_rows) *n_rows = produced;
    });
    // A sharded chain that failed keeps its exchange window for the life of the process: its peers may still be running (or have kernels
    // queued) that store into it, and a freed window would turn one rank's error into an illegal address on every device that maps it.
    if (rc != BRR_OK && c && c->win.R > 1) c->win.keep_on_release = true;
    return rc;
}
// ---- lossless checkpoint / resume for all four samplers (the reference restarts only the Groups model, from
// 6-digit CSV text and without its draw state: src/BRv2Grstart.cpp:61-67).  The file holds everything the next iteration reads:
// scalars, beta, components, residuals (this rank's rows), sigmaG, pi, alpha, lambda / nu, the marker order the in-place
// shuffle has reached, and the iteration counter that keys the Philox draws -- a resumed chain continues bit for bit.
namespace {
struct CkptHeader {
    char magic[8]; int32_t kind, K, G, world; int64_t N, N_total, M, F, it, perm_iter; uint64_t seed; int64_t sc_bytes;
};
template <class T> void put(FILE *f, const T *p, size_t n) { BRR_REQUIRE(fwrite(p, sizeof(T), n, f) == n, BRR_E_IO, "checkpoint write failed"); }
template <class T> void get(FILE *f, T *p, size_t n) { BRR_REQUIRE(fread(p, sizeof(T), n, f) == n, BRR_E_IO, "checkpoint file is truncated"); }
void d2h_put(FILE *f, const double *d, size_t n) { std::vector<double> h(n); if (n) BRR_CUDA(cudaMemcpy(h.data(), d, n * 8, cudaMemcpyDeviceToHost)); put(f, h.data(), n); }
void get_h2d(FILE *f, double *d, size_t n) { std::vector<double> h(n); get(f, h.data(), n); if (n) BRR_CUDA(cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice)); }
}  // namespace

extern "C" int brr_chain_save(brr_chain *c, const char *path)
{
    return guarded([&] {
        BRR_REQUIRE(c && path, BRR_E_ARG, "null pointer");
        BRR_REQUIRE(!c->replay, BRR_E_ARG, "a chain driven by replay tables has no draw state to checkpoint");
        BRR_CUDA(cudaSetDevice(c->g->device));
        if (!c->initialised) chain_init(c);
        BRR_CUDA(cudaStreamSynchronize(c->stream)); BRR_CUDA(cudaStreamSynchronize(c->gstream));
        FILE *f = fopen(path, "wb");
        BRR_REQUIRE(f, BRR_E_IO, std::string("cannot open checkpoint file '") + path + "'");
        try {
            CkptHeader h; memset(&h, 0, sizeof h);
            memcpy(h.magic, "BRRCKP1", 7);
            h.kind = c->kind; h.K = c->K; h.G = c->G; h.world = c->win.R; h.N = c->N; h.N_total = c->N_total; h.M = c->M; h.F = c->F;
            h.it = c->it; h.perm_iter = c->prepared_upto; h.seed = c->seed; h.sc_bytes = (int64_t)sizeof(IterScalars);
            put(f, &h, 1);
            IterScalars sc; BRR_CUDA(cudaMemcpy(&sc, c->sc.p, sizeof sc, cudaMemcpyDeviceToHost));
            put(f, &sc, 1);
            const size_t M = (size_t)c->M, GK = (size_t)c->G * std::max(c->K, 1);
            d2h_put(f, c->d_eps, (size_t)c->N); d2h_put(f, c->beta.p, M); d2h_put(f, c->comp.p, M);
            d2h_put(f, c->sigmaG.p, (size_t)c->G); d2h_put(f, c->pi.p, GK); d2h_put(f, c->alpha.p, (size_t)c->F);
            if (c->kind == BRR_HORSESHOE) { d2h_put(f, c->lambda.p, M); d2h_put(f, c->nu.p, M); }
            put(f, c->markerI.data(), M); put(f, c->fixedI.data(), (size_t)c->F);
        } catch (...) { fclose(f); throw; }
        BRR_REQUIRE(fclose(f) == 0, BRR_E_IO, "checkpoint write failed");
    });
}

extern "C" int brr_chain_load(brr_chain *c, const char *path)
{
    return guarded([&] {
        BRR_REQUIRE(c && path, BRR_E_ARG, "null pointer");
        BRR_REQUIRE(!c->replay, BRR_E_ARG, "a chain driven by replay tables cannot resume from a checkpoint");
        BRR_REQUIRE(c->it == 0 && c->prepared_upto < 0, BRR_E_ARG, "load a checkpoint into a freshly created chain, before its first run");
        BRR_CUDA(cudaSetDevice(c->g->device));
        FILE *f = fopen(path, "rb");
        BRR_REQUIRE(f, BRR_E_IO, std::string("cannot open checkpoint file '") + path + "'");
        try {
            CkptHeader h; get(f, &h, 1);
            BRR_REQUIRE(memcmp(h.magic, "BRRCKP1", 7) == 0, BRR_E_IO, "not a checkpoint file");
            BRR_REQUIRE(h.kind == c->kind && h.K == c->K && h.G == c->G && h.N == c->N && h.N_total == c->N_total && h.M == c->M && h.F == c->F &&
                        h.world == c->win.R && h.sc_bytes == (int64_t)sizeof(IterScalars), BRR_E_ARG,
                        "the checkpoint was written by a chain of another shape (sampler, components, groups, rows, markers, fixed effects or ranks)");
            BRR_REQUIRE(h.seed == c->seed, BRR_E_ARG, "the checkpoint was written with another seed: the draws would not continue the chain");
            if (!c->initialised) chain_init(c);
            IterScalars sc; get(f, &sc, 1);
            BRR_CUDA(cudaMemcpy(c->sc.p, &sc, sizeof sc, cudaMemcpyHostToDevice));
            const size_t M = (size_t)c->M, GK = (size_t)c->G * std::max(c->K, 1);
            get_h2d(f, c->d_eps, (size_t)c->N); get_h2d(f, c->beta.p, M); get_h2d(f, c->comp.p, M);
            get_h2d(f, c->sigmaG.p, (size_t)c->G); get_h2d(f, c->pi.p, GK); get_h2d(f, c->alpha.p, (size_t)c->F);
            if (c->kind == BRR_HORSESHOE) { get_h2d(f, c->lambda.p, M); get_h2d(f, c->nu.p, M); }
            get(f, c->markerI.data(), M); get(f, c->fixedI.data(), (size_t)c->F);
            c->it = h.it;
            c->restored_perm = h.perm_iter;                   // markerI is the order OF that iteration: prepare() must not shuffle it again
            c->prepared_upto = std::min<int64_t>(h.it - 1, h.perm_iter - 1);
        } catch (...) { fclose(f); throw; }
        fclose(f);
    });
}

extern "C" int brr_chain_get_pi(brr_chain *c, double *pi)
{
    return guarded([&] {
        BRR_REQUIRE(c && pi && c->initialised && c->kind != BRR_HORSESHOE, BRR_E_ARG, "bad arguments");
        BRR_CUDA(cudaSetDevice(c->g->device));
        BRR_CUDA(cudaMemcpy(pi, c->pi.p, (size_t)c->G * c->K * 8, cudaMemcpyDeviceToHost));
    });
}
extern "C" int brr_chain_get_hyper(brr_chain *c, double *out)
{
    return guarded([&] {
        BRR_REQUIRE(c && out && c->initialised, BRR_E_ARG, "bad arguments");
        BRR_CUDA(cudaSetDevice(c->g->device));
        IterScalars sc; BRR_CUDA(cudaMemcpy(&sc, c->sc.p, sizeof sc, cudaMemcpyDeviceToHost));
        out[0] = sc.eta; out[1] = sc.tau; out[2] = sc.c2;
    });
}
extern "C" int brr_chain_get_sigmaE(brr_chain *c, double *sigmaE)
{
    return guarded([&] {
        BRR_REQUIRE(c && sigmaE && c->initialised, BRR_E_ARG, "bad arguments");
        BRR_CUDA(cudaSetDevice(c->g->device));
        IterScalars sc; BRR_CUDA(cudaMemcpy(&sc, c->sc.p, sizeof sc, cudaMemcpyDeviceToHost));
        *sigmaE = sc.sigmaE;
    });
}
extern "C" void brr_set_message_handler(brr_message_fn fn, void *ctx) { g_msg_ctx.store(ctx); g_msg_fn.store(fn); }
extern "C" int brr_chain_last_timing(const brr_chain *c, double *ms, int64_t *launches)
{
    return guarded([&] {
        BRR_REQUIRE(c, BRR_E_ARG, "null pointer");
        if (ms) *ms = c->last_ms;
        if (launches) *launches = c->last_launches;
    });
}
extern "C" int brr_chain_kernel_ms(const brr_chain *c, double *gram_sweep_hyper_ms)
{
    return guarded([&] {
        BRR_REQUIRE(c && gram_sweep_hyper_ms, BRR_E_ARG, "null pointer");
        for (int k = 0; k < 3; ++k) gram_sweep_hyper_ms[k] = c->last_kernel_ms[k];
    });
}
extern "C" int brr_chain_sweep_profile(const brr_chain *c, double *out16)
{
    return guarded([&] {
        BRR_REQUIRE(c && out16 && c->initialised, BRR_E_ARG, "bad arguments");
        BRR_CUDA(cudaSetDevice(c->g->device));
        long long h[16];
        BRR_CUDA(cudaMemcpy(h, c->prof.p, sizeof h, cudaMemcpyDeviceToHost));
        for (int i = 0; i < 16; ++i) out16[i] = (double)h[i];
    });
}
extern "C" int brr_chain_geometry(const brr_chain *c, int *block, int *workers, int *rows_per_worker_max, int *smem_bytes)
{
    return guarded([&] {
        BRR_REQUIRE(c, BRR_E_ARG, "null pointer");
        if (block) *block = c->B;
        if (workers) *workers = c->nW;
        if (rows_per_worker_max) *rows_per_worker_max = c->rows_per_worker;
        if (smem_bytes) *smem_bytes = (int)c->smem;
    });
}
extern "C" void brr_chain_destroy(brr_chain *c)
{
    if (!c) return;
    try { if (c->writer) c->writer->finish(); } catch (...) { }
    try { if (c->bwriter) c->bwriter->finish(); } catch (...) { }
    delete c;
}

// =================================================================================================
// The four reference entry points
namespace {

struct GenoGuard { brr_geno *g = nullptr; ~GenoGuard() { brr_geno_free(g); } };
struct ChainGuard { brr_chain *c = nullptr; ~ChainGuard() { brr_chain_destroy(c); } };

void check_rc(int rc) { if (rc != BRR_OK) throw Error(rc, brr_last_error()); }

void touch_file(const char *path, const std::string &header)
{
    FILE *f = fopen(path, "w");
    BRR_REQUIRE(f, BRR_E_IO, std::string("cannot open output file '") + path + "'");
    if (!header.empty()) fwrite(header.data(), 1, header.size(), f);
    fclose(f);
}

void say(const std::string &text)
{
    if (brr_message_fn fn = g_msg_fn.load()) fn(g_msg_ctx.load(), text.c_str());
}
std::string gfmt(double v) { char b[40]; snprintf(b, sizeof b, "%g", v); return b; }   // operator<<(double) of an ostream at its default precision

void run_entry(const char *outputFile, const brr_config &cfg, const double *X, int64_t N, int64_t M)
{
    GenoGuard gg; ChainGuard cg;
    const auto t1 = std::chrono::steady_clock::now();
    check_rc(brr_geno_from_dense(X, N, M, 0, &gg.g));
    check_rc(brr_chain_create(&cfg, gg.g, &cg.c));
    check_rc(brr_chain_open_output(cg.c, outputFile));
    double h[3];
    if (cfg.kind == BRR_HORSESHOE) {                                                    // src/HorseshoeR.cpp:191,195
        check_rc(brr_chain_run(cg.c, 0, 0, nullptr, 0, nullptr));                       // initialisation only
        check_rc(brr_chain_get_hyper(cg.c, h));
        say("initial eta " + gfmt(h[0]) + "\n"); say("initial tau " + gfmt(h[1]) + "\n");
    }
    // The chain runs in chunks that end where the reference prints its progress line (src/BayesRv2.cpp:173-175: before iteration
    // `it` when it > 0 and it % (max_iterations / 10) == 0; a period of 0 -- fewer than ten iterations, a division by zero there --
    // counts as 1).  Every chunk boundary also checks the in-kernel watchdog, so an aborted chain stops within one chunk.
    const int period = std::max(1, cfg.max_iterations / 10);
    for (int it = 0; it < cfg.max_iterations; ) {
        if (it > 0) {
            say("iteration: " + std::to_string(it) + "\n");
            if (cfg.kind == BRR_HORSESHOE) {                                            // src/HorseshoeR.cpp:203-206
                double sE = 0;
                check_rc(brr_chain_get_hyper(cg.c, h)); check_rc(brr_chain_get_sigmaE(cg.c, &sE));
                say(" tau " + gfmt(h[1]) + "\n"); say(" eta " + gfmt(h[0]) + "\n"); say("sigmaE" + gfmt(sE) + "\n");
            }
        }
        const int n = std::min(period, cfg.max_iterations - it);
        check_rc(brr_chain_run(cg.c, n, 0, nullptr, 0, nullptr));
        it += n;
    }
    check_rc(brr_chain_close_output(cg.c));
    const auto secs = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t1).count();
    say("duration: " + std::to_string((long long)secs) + "s\n");                        // src/BayesRv2.cpp:276-278
}

}  // namespace

extern "C" int brr_BayesRSamplerV2(const char *outputFile, int seed, int max_iterations, int burn_in, int thinning,
                                   const double *X, int64_t N, int64_t M, const double *Y,
                                   double sigma0, double v0E, double s02E, double v0G, double s02G,
                                   const double *cva, int ncva)
{
    return guarded([&] {
        BRR_REQUIRE(outputFile && X && Y && cva && N > 0 && M > 0 && ncva > 0, BRR_E_ARG, "BayesRSamplerV2: bad arguments");
        touch_file(outputFile, sample_header(BRR_V2, N, M, 1, 0));   // file + header exist even when validation fails (:69-70,76-80)
        check_iters(max_iterations, burn_in, thinning);
        brr_config cfg; memset(&cfg, 0, sizeof cfg);
        cfg.kind = BRR_V2; cfg.seed = (uint64_t)(uint32_t)seed; cfg.max_iterations = max_iterations; cfg.burn_in = burn_in; cfg.thinning = thinning;
        cfg.Y = Y; cfg.sigma0 = sigma0; cfg.v0E = v0E; cfg.s02E = s02E; cfg.v0G = v0G; cfg.s02G = s02G; cfg.cva = cva; cfg.ncva = ncva;
        run_entry(outputFile, cfg, X, N, M);
    });
}

extern "C" int brr_BayesRSamplerV2Groups(const char *outputFile, int seed, int max_iterations, int burn_in, int thinning,
                                         const double *X, int64_t N, int64_t M, const double *Y,
                                         double sigma0, double v0E, double s02E, double v0G, double s02G,
                                         const double *cva, int ncva, int groups, const int32_t *gAssign,
                                         const double *fixed, int64_t F)
{
    return guarded([&] {
        BRR_REQUIRE(outputFile && X && Y && cva && gAssign && N > 0 && M > 0 && ncva > 0 && groups > 0 && F >= 0, BRR_E_ARG,
                    "BayesRSamplerV2Groups: bad arguments");
        touch_file(outputFile, "");                                  // opened before validation, header after it (Groups:86,113)
        check_iters(max_iterations, burn_in, thinning);
        brr_config cfg; memset(&cfg, 0, sizeof cfg);
        cfg.kind = BRR_GROUPS; cfg.seed = (uint64_t)(uint32_t)seed; cfg.max_iterations = max_iterations; cfg.burn_in = burn_in; cfg.thinning = thinning;
        cfg.Y = Y; cfg.sigma0 = sigma0; cfg.v0E = v0E; cfg.s02E = s02E; cfg.v0G = v0G; cfg.s02G = s02G; cfg.cva = cva; cfg.ncva = ncva;
        cfg.groups = groups; cfg.gAssign = gAssign; cfg.fixed = fixed; cfg.F = F;
        run_entry(outputFile, cfg, X, N, M);
    });
}

extern "C" int brr_BRV2Grstart(const char *outputFile, int seed, int max_iterations, int burn_in, int thinning,
                               double mu, const double *beta, double sigmaE, const double *sigmaGG,
                               const double *X, int64_t N, int64_t M, const double *epsilon, const double *components,
                               double sigma0, double v0E, double s02E, double v0G, double s02G,
                               const double *cva, int ncva, int groups, const int32_t *gAssign)
{
    return guarded([&] {
        BRR_REQUIRE(outputFile && X && beta && sigmaGG && epsilon && components && cva && gAssign && N > 0 && M > 0 && ncva > 0 && groups > 0,
                    BRR_E_ARG, "BRV2Grstart: bad arguments");
        touch_file(outputFile, "");                                  // Grstart:84; no header is ever written
        check_iters(max_iterations, burn_in, thinning);
        brr_config cfg; memset(&cfg, 0, sizeof cfg);
        cfg.kind = BRR_GRSTART; cfg.seed = (uint64_t)(uint32_t)seed; cfg.max_iterations = max_iterations; cfg.burn_in = burn_in; cfg.thinning = thinning;
        cfg.sigma0 = sigma0; cfg.v0E = v0E; cfg.s02E = s02E; cfg.v0G = v0G; cfg.s02G = s02G; cfg.cva = cva; cfg.ncva = ncva;
        cfg.groups = groups; cfg.gAssign = gAssign;
        cfg.mu0 = mu; cfg.beta0 = beta; cfg.sigmaE0 = sigmaE; cfg.sigmaGG0 = sigmaGG; cfg.epsilon0 = epsilon; cfg.components0 = components;
        run_entry(outputFile, cfg, X, N, M);
    });
}

extern "C" int brr_HorseshoeR(const char *outputFile, int seed, int max_iterations, int burn_in, int thinning,
                              const double *X, int64_t N, int64_t M, const double *Y,
                              double A, double v0E, double s02E, double vL, double vT, double c2, double vC, double sC)
{
    return guarded([&] {
        BRR_REQUIRE(outputFile && X && Y && N > 0 && M > 0, BRR_E_ARG, "HorseshoeR: bad arguments");
        check_iters(max_iterations, burn_in, thinning);              // before the file is touched (HorseshoeR.cpp:119-123,274)
        brr_config cfg; memset(&cfg, 0, sizeof cfg);
        cfg.kind = BRR_HORSESHOE; cfg.seed = (uint64_t)(uint32_t)seed; cfg.max_iterations = max_iterations; cfg.burn_in = burn_in; cfg.thinning = thinning;
        cfg.Y = Y; cfg.v0E = v0E; cfg.s02E = s02E; cfg.A = A; cfg.vL = vL; cfg.vT = vT; cfg.c2 = c2; cfg.vC = vC; cfg.sC = sC;
        run_entry(outputFile, cfg, X, N, M);
    });
}
