// Row-sharded chains: exchange window (cudaMalloc + CUDA IPC / same-process peer access), the peer-memory Gram sum
// and the collective part of the genotype statistics.  See shard.cuh.
#include "shard.cuh"
#include <unistd.h>
#include <cstring>
#include <random>

namespace brr {

void comm_check(int rc, const char *what)
{
    BRR_REQUIRE(rc == 0, BRR_E_ARG, std::string("brr_comm call-back failed: ") + what + " returned " + std::to_string(rc));
}
void comm_allreduce(const brr_comm &comm, double *buf, int64_t n)
{
    if (comm.world <= 1 || n == 0) return;
    comm_check(comm.allreduce_sum(comm.ctx, buf, n), "allreduce_sum");
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// 64 random bits drawn once per process
static uint64_t process_nonce()
{
    static const uint64_t nonce = [] {
        std::random_device rd;
        uint64_t v = ((uint64_t)rd() << 32) ^ (uint64_t)rd();
        v ^= (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count() * 0x9E3779B97F4A7C15ull;
        v ^= (uint64_t)getpid() << 17;
        return v ? v : 1;
    }();
    return nonce;
}

void Window::layout(int PS, int nb, int B, int64_t Npad, int gram_elem_bytes)
{
    size_t o = 0;
    off_xred = o; o = align_up(o + (size_t)4 * PS * R * 16, 256);
    off_xfin = o; o = align_up(o + (size_t)R * 2 * 16, 256);
    off_ready = o; o = align_up(o + (size_t)R * 4, 256);
    gram_bytes = align_up((size_t)nb * (gram_tile_entries(B) + lookahead(B) * B) * gram_elem_bytes, 256);
    off_gram = o; if (R > 1) o = o + 2 * gram_bytes;
    off_eps = o; o = align_up(o + (size_t)Npad * 8, 256);
    bytes = o;
}

void Window::allocate()
{
    BRR_CUDA(cudaMalloc(&base, bytes));
    BRR_CUDA(cudaMemset(base, 0, bytes));
    peer[rank] = base;
}

void Window::connect(const brr_comm &comm, int device, int64_t n_local, int B, int kind, int64_t M)
{
    PeerBlob mine; memset(&mine, 0, sizeof mine);
    mine.pid = (int32_t)getpid(); mine.nonce = process_nonce(); mine.device = device; mine.base = (uint64_t)(uintptr_t)base;
    mine.n_local = n_local; mine.block = B; mine.kind = kind; mine.M = M;
    if (R > 1) BRR_CUDA(cudaIpcGetMemHandle(&mine.handle, base));
    std::vector<PeerBlob> all(R);
    if (R > 1) comm_check(comm.allgather(comm.ctx, &mine, all.data(), (int64_t)sizeof(PeerBlob)), "allgather");
    else all[0] = mine;
    BRR_REQUIRE(all[rank].nonce == mine.nonce && all[rank].base == mine.base, BRR_E_ARG, "brr_comm::allgather did not return this rank's own entry at index `rank`");
    n_total = 0; colocated = 1;
    for (int r = 0; r < R; ++r) {
        BRR_REQUIRE(all[r].block == B && all[r].kind == kind && all[r].M == M, BRR_E_ARG,
                    "ranks of a sharded chain disagree on the sampler, the number of markers or the Gibbs block size");
        n_rows[r] = all[r].n_local; row0[r] = n_total; n_total += all[r].n_local;
        if (r == rank) continue;
        // same process (ranks are threads): the address is directly usable.  Decided by a random per-process token, not by the pid
        // alone -- ranks in different PID namespaces (one container per rank) can share a pid; anything else takes the IPC handle
        if (all[r].nonce == mine.nonce && all[r].pid == mine.pid) {
            if (all[r].device == device) ++colocated;
            if (all[r].device != device) {
                int can = 0;
                BRR_CUDA(cudaDeviceCanAccessPeer(&can, device, all[r].device));
                BRR_REQUIRE(can, BRR_E_CUDA, "devices of a sharded chain cannot access each other's memory (no NVLink / P2P)");
                const cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) BRR_CUDA(e);
                (void)cudaGetLastError();
            }
            peer[r] = reinterpret_cast<uint8_t *>((uintptr_t)all[r].base);
        } else {
            void *ptr = nullptr;
            BRR_CUDA(cudaIpcOpenMemHandle(&ptr, all[r].handle, cudaIpcMemLazyEnablePeerAccess));
            peer[r] = static_cast<uint8_t *>(ptr); ipc_opened[r] = true;
        }
    }
}

void Window::release()
{
    if (keep_on_release) { base = nullptr; return; }      // (the peers' mappings stay open too: their kernels may still be reading this rank's flags)
    for (int r = 0; r < MAXR; ++r) if (ipc_opened[r]) { cudaIpcCloseMemHandle(peer[r]); ipc_opened[r] = false; }
    if (base) cudaFree(base);
    base = nullptr;
}

// ------------------------------------------------------------------------------------------------
namespace {

struct AllSumParams {
    int rank, R;
    uint32_t epoch;
    uint32_t *ready[MAXR];          // every rank's flag array (R entries): ready[r][s] = last epoch rank s has published to r
    const int32_t *part[MAXR];      // every rank's partial Gram (fp64 for stores with dense columns)
    int32_t *sum; size_t n4;        // int4 (fp64: double2) elements
    int *abort_flag;
};

template <bool F64>
__global__ void __launch_bounds__(256) gram_allsum_kernel(const __grid_constant__ AllSumParams q)
{
    __shared__ int s_ok;
    const int tid = threadIdx.x;
    if (tid == 0) s_ok = 1;
    // the partial of this rank is complete (stream order): tell every peer
    if (blockIdx.x == 0 && tid < q.R && tid != q.rank) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(q.ready[tid] + q.rank), "r"(q.epoch) : "memory");
    }
    __syncthreads();
    if (tid < q.R && tid != q.rank) {
        const uint32_t *f = q.ready[q.rank] + tid;
        const long long t0 = clock64();
        while (true) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int32_t)(v - q.epoch) >= 0) break;
            if (clock64() - t0 > 20000000000LL || *reinterpret_cast<volatile int *>(q.abort_flag) != 0) { atomicCAS(q.abort_flag, 0, 3); s_ok = 0; break; }
        }
    }
    __syncthreads();
    if (!s_ok) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (F64) {      // partial tiles of a store with dense columns: summed in rank order on every rank (identical bits everywhere)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < q.n4; i += stride) {
            double2 v[MAXR];
#pragma unroll
            for (int r = 0; r < MAXR; ++r) if (r < q.R) v[r] = __ldcg(reinterpret_cast<const double2 *>(q.part[r]) + i);
            double2 acc = v[0];
#pragma unroll
            for (int r = 1; r < MAXR; ++r) if (r < q.R) { acc.x += v[r].x; acc.y += v[r].y; }
            reinterpret_cast<double2 *>(q.sum)[i] = acc;
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < q.n4; i += stride) {
        int4 acc = make_int4(0, 0, 0, 0);
        int4 v[MAXR];
#pragma unroll
        for (int r = 0; r < MAXR; ++r) if (r < q.R) v[r] = __ldcg(reinterpret_cast<const int4 *>(q.part[r]) + i);   // all loads in flight
#pragma unroll
        for (int r = 0; r < MAXR; ++r) if (r < q.R) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        reinterpret_cast<int4 *>(q.sum)[i] = acc;
    }
}

}  // namespace

void preload_allsum() { preload_kernel(gram_allsum_kernel<false>); preload_kernel(gram_allsum_kernel<true>); }

void launch_gram_allsum(const Window &w, int buf, uint32_t epoch, void *d_sum, size_t n_elems, bool f64, int *abort_flag, cudaStream_t stream, int max_ctas)
{
    AllSumParams q; memset(&q, 0, sizeof q);
    q.rank = w.rank; q.R = w.R; q.epoch = epoch; q.sum = static_cast<int32_t *>(d_sum); q.n4 = f64 ? n_elems / 2 : n_elems / 4; q.abort_flag = abort_flag;
    for (int r = 0; r < w.R; ++r) { q.ready[r] = w.ready(r); q.part[r] = w.gram(r, buf); }
    const unsigned blocks = (unsigned)std::min<size_t>((q.n4 + 255) / 256, (size_t)(max_ctas > 0 ? max_ctas : 148) * 8);
    if (f64) gram_allsum_kernel<true><<<blocks ? blocks : 1, 256, 0, stream>>>(q);
    else gram_allsum_kernel<false><<<blocks ? blocks : 1, 256, 0, stream>>>(q);
    BRR_CUDA(cudaGetLastError());
}

}  // namespace brr

using namespace brr;

extern "C" int brr_comm_selftest(const brr_comm *comm, double *buf, int64_t n, int64_t token, int64_t *gathered)
{
    return guarded([&] {
        BRR_REQUIRE(comm && comm->world >= 1 && comm->world <= BRR_MAX_WORLD && comm->rank >= 0 && comm->rank < comm->world, BRR_E_ARG,
                    "brr_comm: world must be in [1, " + std::to_string(BRR_MAX_WORLD) + "] and rank in [0, world)");
        BRR_REQUIRE(comm->world == 1 || (comm->allreduce_sum && comm->allgather), BRR_E_ARG, "brr_comm: call-backs missing");
        if (buf && n > 0) comm_allreduce(*comm, buf, n);
        if (gathered) {
            if (comm->world > 1) comm_check(comm->allgather(comm->ctx, &token, gathered, 8), "allgather");
            else gathered[0] = token;
        }
    });
}

extern "C" int brr_geno_shard_stats(brr_geno *g, const brr_comm *comm)
{
    return guarded([&] {
        BRR_REQUIRE(g && comm, BRR_E_ARG, "null pointer");
        BRR_REQUIRE(comm->world >= 1 && comm->world <= BRR_MAX_WORLD, BRR_E_SIZE, "world size outside [1, " + std::to_string(BRR_MAX_WORLD) + "]");
        BRR_CUDA(cudaSetDevice(g->device));
        const int64_t M = g->M;
        {   // missing genotypes of a .bed row shard: one fill value per column for all ranks (collective: every rank takes part)
            double pending = g->pending_impute ? 1.0 : 0.0;
            comm_allreduce(*comm, &pending, 1);
            if (pending > 0.0) {
                BRR_REQUIRE(g->pending_impute, BRR_E_ARG, "some ranks read their .bed row shard with impute_missing and some without");
                std::vector<double> cnt = g->pending_cnt;
                comm_allreduce(*comm, cnt.data(), (int64_t)cnt.size());      // integers below 2^53: exact in any order
                geno_impute_pending(g, cnt);
            }
        }
        std::vector<double> buf((size_t)2 * M + 1);
        BRR_CUDA(cudaMemcpy(buf.data(), g->d_S, M * 8, cudaMemcpyDeviceToHost));
        BRR_CUDA(cudaMemcpy(buf.data() + M, g->d_Q, M * 8, cudaMemcpyDeviceToHost));
        buf[2 * M] = (double)g->N;
        comm_allreduce(*comm, buf.data(), (int64_t)buf.size());          // integers below 2^53: exact in any order (dense columns: fp64 sums,
                                                                         // identical on every rank by the call-back's contract)
        BRR_CUDA(cudaMemcpy(g->d_S, buf.data(), M * 8, cudaMemcpyHostToDevice));
        BRR_CUDA(cudaMemcpy(g->d_Q, buf.data() + M, M * 8, cudaMemcpyHostToDevice));
        g->n_total = buf[2 * M];
        geno_affine_from_stats(g);
    });
}
