// Block Gram kernels: G_b = C_b^T C_b over the raw 0/1/2 codes of the B markers of Gibbs block b
// (markers taken in the iteration's permuted order), exact in int32.  The reference has no such
// operation: it is what replaces the marker-to-marker dependency through the N-vector residual
// (reference src/BayesRv2.cpp:191 -> :243) by a B x B correction ("block-Gram identity").
//
//   gram_tc_kernel   : tcgen05.mma kind::i8, accumulators in TMEM, operands unpacked 2-bit -> int8 into the
//                      canonical K-major (no-swizzle) shared-memory layout, packed columns staged by
//                      cp.async.bulk (TMA) with mbarrier completion.  One CTA per block.
//   gram_dp4a_kernel : CUDA-core reference of the same integers (validation / debugging only).
#include "common.cuh"

namespace brr {

constexpr int GRAM_KC = 256;                       // rows (K) per operand tile (one commit group of MMAs)
// rows per load stage: ONE bulk copy per marker and stage (bulk copies have a fixed issue cost of tens of cycles each; 128-byte
// copies left the kernel copy-issue bound), as many rows as fit beside the two operand tiles of R = B + lookahead(B) marker rows
__host__ __device__ constexpr int gram_ldr(int R) { return R > 224 ? 512 : R > 192 ? 768 : 1024; }
__host__ __device__ constexpr int gram_stage_row(int R) { return gram_ldr(R) / 4 + 16; }   // bytes per staged packed column (padded to an odd number of 16-byte units: conflict-free 128-bit reads)
__host__ __device__ constexpr int gram_tile_bytes(int R) { return (R > 128 ? R : 128) * GRAM_KC; }   // int8 operand tile: R marker rows x KC (the A operand always reads 128: rows >= R stay zero)
constexpr int GRAM_SBO = (GRAM_KC / 16) * 128;     // byte stride between 8-marker groups
constexpr int GRAM_LBO = 128;                      // byte stride between K-adjacent 8x16B core matrices

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait (~2 s): a lost completion must not hang the GPU.  It raises the chain's sticky abort flag (code 4; the host turns
// it into BRR_E_CUDA like every other in-kernel watchdog) and the kernel runs on to its end with a result nobody reads.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int *abort_flag)
{
    const long long t0 = clock64();
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > 4000000000LL) { if (abort_flag) atomicCAS(abort_flag, 0, 4); break; }
    }
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (tcgen05): start address, LBO, SBO in 16-byte units, version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(GRAM_LBO >> 4) << 16) | ((uint64_t)(GRAM_SBO >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// 16 2-bit codes (one packed word) -> 16 bytes of the K dimension.  The rows land in the order 0,4,8,12, 1,5,9,13, ... inside
// their group of 16: a sum over rows does not care, and both operands of every product are read from the same tile, so
// they see the same order -- seven ALU operations per word instead of the twenty-eight of an in-order expansion (the
// kernel is bound by this unpack, not by the tensor pipe).
__device__ __forceinline__ uint4 expand16(uint32_t w)
{
    constexpr uint32_t M = 0x03030303u;
    return make_uint4(w & M, (w >> 2) & M, (w >> 4) & M, (w >> 6) & M);
}

// CROSS: additionally X[blk][jl][k] = sum_n code[n, order[blk*B - LA + jl]] * code[n, order[blk*B + k]], LA = lookahead(B) -- the
// products of the block's markers with the last LA markers of the previous block (the look-ahead correction of the sweep,
// DESIGN.md 3.1): those markers are LA more rows of the B operand (N = B + LA columns of the accumulator), A is unchanged.
template <int B, bool CROSS>
__global__ void __launch_bounds__(256, 1)
gram_tc_kernel(const uint8_t *__restrict__ packed, int64_t stride, int64_t Npad,
               const int32_t *__restrict__ order, int64_t n_order, int32_t *__restrict__ G, int32_t *__restrict__ X, int *abort_flag)
{
    static_assert(B == 32 || B == 64 || B == 128, "block size");
    constexpr int LA = lookahead(B);
    constexpr int R = CROSS ? B + LA : B;                   // marker rows of the operand tile
    static_assert(R <= 256 && R % 16 == 0, "N of the MMA (one accumulator column per marker row of the tile)");
    constexpr int GRAM_LDR = gram_ldr(R), GRAM_STAGE_ROW = gram_stage_row(R), GRAM_TILE_BYTES = gram_tile_bytes(R);
    constexpr int TE = gram_tile_entries(B);                // entries of a stored self tile (block-upper trapezoid, common.cuh)
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *tile0 = smem;                                  // 2 operand tiles of max(R, 128) x KC bytes
    uint8_t *stage0 = smem + 2 * GRAM_TILE_BYTES;           // 2 x R x GRAM_STAGE_ROW staged packed columns
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage0 + 2 * R * GRAM_STAGE_ROW);   // full[2], free[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    int32_t *cols = reinterpret_cast<int32_t *>(tmem_slot + 2);                        // R marker ids

    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t *full = bars, *freeb = bars + 2;
    const int64_t nblocks = (n_order + B - 1) / B;

    if (warp == 0) {   // TMEM: 128 lanes x 256 int32 columns (R = B + lookahead(B) <= 256 used)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // Persistent over blocks: the grid may be smaller than the number of blocks -- the chain runs this kernel on the SMs the
    // sweep kernel of the previous iteration leaves free (chain.cu), a grid of that many CTAs.
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    if (tid < R) {
        const int64_t o = tid < B ? blk * B + tid : blk * B - LA + (tid - B);   // rows B.. : tail of the previous block
        cols[tid] = (o >= 0 && o < n_order && (tid < B || blk > 0)) ? order[o] : -1;
    }
    // zero both operand tiles (rows of padding markers and rows >= R must read as 0) and both stages
    for (int i = tid; i < (2 * GRAM_TILE_BYTES + 2 * R * GRAM_STAGE_ROW) / 16; i += 256)
        reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        if (blk != (int64_t)blockIdx.x)
            for (int i = 0; i < 4; ++i) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[i])) : "memory");
        mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&freeb[0], 1); mbar_init(&freeb[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int nloads = (int)((Npad + GRAM_LDR - 1) / GRAM_LDR);
    int nvalid = 0;
    for (int c = 0; c < R; ++c) nvalid += cols[c] >= 0;

    auto load_rows = [&](int L) { const int64_t left = Npad - (int64_t)L * GRAM_LDR; return (int)(left < GRAM_LDR ? left : GRAM_LDR); };
    auto issue_loads = [&](int L) {   // one bulk copy per marker: its rows [L * LDR, L * LDR + load_rows)
        const int s = L & 1;
        const uint32_t bytes = (uint32_t)load_rows(L) / 4;
        if (tid == 0) mbar_expect_tx(&full[s], (uint32_t)nvalid * bytes);
        if (tid < R && cols[tid] >= 0)
            bulk_g2s(stage0 + (s * R + tid) * GRAM_STAGE_ROW, packed + (int64_t)cols[tid] * stride + (int64_t)L * (GRAM_LDR / 4), bytes, &full[s]);
    };
    if (nvalid == 0) {   // nothing to do but keep the protocol simple: write zeros
        for (int i = tid; i < TE; i += 256) G[blk * TE + i] = 0;
        if (CROSS) for (int i = tid; i < LA * B; i += 256) X[blk * LA * B + i] = 0;
    } else {
        issue_loads(0);
        if (nloads > 1) issue_loads(1);
        // instruction descriptor: D = S32, A = B = unsigned 8-bit, both K-major, N = B, M = 128
        constexpr uint32_t idesc = (2u << 4) | ((uint32_t)(R >> 3) << 17) | ((128u >> 4) << 24);
        int g = 0;                                   // running tile count: tile g uses operand buffer g & 1
        for (int L = 0; L < nloads; ++L) {
            const int s = L & 1;
            mbar_wait(&full[s], (uint32_t)((L >> 1) & 1), abort_flag);
            const uint8_t *stage = stage0 + s * R * GRAM_STAGE_ROW;
            const int nsub = load_rows(L) / GRAM_KC;
            for (int sub = 0; sub < nsub; ++sub, ++g) {
                const int ts = g & 1;
                if (g >= 2) mbar_wait(&freeb[ts], (uint32_t)(((g >> 1) - 1) & 1), abort_flag);     // the MMAs that read this buffer are done
                uint8_t *tile = tile0 + ts * GRAM_TILE_BYTES;
                // unpack: item = (v, c): 64 rows of marker c -> four 16-byte core-matrix rows
                constexpr int ITEMS = R * (GRAM_KC / 64);
#pragma unroll
                for (int it = 0; it < (ITEMS + 255) / 256; ++it) {
                    const int item = tid + it * 256;
                    if (ITEMS % 256 != 0 && item >= ITEMS) break;
                    const int c = item % R, v = item / R;
                    const uint4 q = *reinterpret_cast<const uint4 *>(stage + c * GRAM_STAGE_ROW + (sub * (GRAM_KC / 64) + v) * 16);
                    uint8_t *dst = tile + (c >> 3) * GRAM_SBO + (c & 7) * 16 + (v * 4) * GRAM_LBO;
                    *reinterpret_cast<uint4 *>(dst) = expand16(q.x);
                    *reinterpret_cast<uint4 *>(dst + GRAM_LBO) = expand16(q.y);
                    *reinterpret_cast<uint4 *>(dst + 2 * GRAM_LBO) = expand16(q.z);
                    *reinterpret_cast<uint4 *>(dst + 3 * GRAM_LBO) = expand16(q.w);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                __syncthreads();
                if (sub == nsub - 1 && L + 2 < nloads) issue_loads(L + 2);      // every thread is past its last read of stage s
                if (tid == 0) {   // (issuing from an elected lane of a warp-uniform branch makes the eight MMAs leave back to back instead of ~100 cycles
                                  // apart -- measured SLOWER here, 0.64 instead of 0.575 ms at config 2: the kernel is bound by shared-memory bandwidth,
                                  // 48 KB of unpack stores + 80 KB of operand reads per 256-row tile, and the spaced-out issue interleaves the two better)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t base = smem_u32(tile);
#pragma unroll
                    for (int kk = 0; kk < GRAM_KC / 32; ++kk) {
                        const uint64_t da = umma_desc(base + kk * 2 * GRAM_LBO);
                        const uint32_t acc = (g > 0 || kk > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(tmem), "l"(da), "l"(da), "r"(idesc), "r"(acc) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(smem_u32(&freeb[ts])) : "memory");
                }
            }
        }
        {   // all MMAs done?  (commit groups complete in order: the last two cover both buffers)
            const int last = g - 1;
            mbar_wait(&freeb[last & 1], (uint32_t)((last >> 1) & 1), abort_flag);
            if (g > 1) { const int l2 = last - 1; mbar_wait(&freeb[l2 & 1], (uint32_t)((l2 >> 1) & 1), abort_flag); }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: TMEM lane = marker row i, column = marker j.  Of the self tile only the columns of the row's own sub-window and the
        // later ones are stored (gram_tile_index, common.cuh): warp w holds sub-window w's rows
        if (warp < 4) {
            const int row = warp * 32 + (tid & 31);
#pragma unroll
            for (int c0 = 0; c0 < B; c0 += 32) {
                if (c0 < warp * 32) continue;                 // (warp-uniform)
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < B) {
                    int4 *dst = reinterpret_cast<int4 *>(G + blk * TE + gram_tile_index(B, row, c0));
#pragma unroll
                    for (int q = 0; q < 8; ++q) dst[q] = make_int4((int)v[4 * q], (int)v[4 * q + 1], (int)v[4 * q + 2], (int)v[4 * q + 3]);
                }
            }
            if (CROSS)     // accumulator columns B .. B + LA - 1: products with the previous block's tail, stored [jl][k]
#pragma unroll
            for (int x0 = 0; x0 < LA; x0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(B + x0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < B) {
#pragma unroll
                    for (int jl = 0; jl < 32; ++jl) X[(blk * LA + x0 + jl) * B + row] = (int)v[jl];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();            // the accumulator has been read out and every thread is done with this block's barriers
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---- CUDA-core reference: 8 x 8 register tile per thread, dp4a over int8-expanded codes.
template <int B>
__global__ void __launch_bounds__(256) gram_dp4a_kernel(const uint8_t *__restrict__ packed, int64_t stride, int64_t Npad,
                                                        const int32_t *__restrict__ order, int64_t n_order,
                                                        int32_t *__restrict__ G)
{
    constexpr int KR = 64;                     // rows per smem tile
    constexpr int TS = (B + 15) / 16;          // per-thread tile side: 256 threads = 16 x 16
    __shared__ uint32_t tile[B][KR / 4 + 1];   // int8 x 4 words
    __shared__ int32_t cols[B];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    const int64_t blk = blockIdx.x;
    if (tid < B) { const int64_t o = blk * B + tid; cols[tid] = o < n_order ? order[o] : -1; }
    __syncthreads();
    int acc[TS][TS];
#pragma unroll
    for (int a = 0; a < TS; ++a)
#pragma unroll
        for (int b = 0; b < TS; ++b) acc[a][b] = 0;
    for (int64_t r0 = 0; r0 < Npad; r0 += KR) {
        for (int item = tid; item < B * (KR / 16); item += 256) {
            const int c = item / (KR / 16), w = item % (KR / 16);
            uint4 e = make_uint4(0, 0, 0, 0);
            if (cols[c] >= 0) e = expand16(*reinterpret_cast<const uint32_t *>(packed + (int64_t)cols[c] * stride + r0 / 4 + w * 4));
            tile[c][w * 4] = e.x; tile[c][w * 4 + 1] = e.y; tile[c][w * 4 + 2] = e.z; tile[c][w * 4 + 3] = e.w;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < KR / 4; ++k) {
            uint32_t av[TS], bv[TS];
#pragma unroll
            for (int a = 0; a < TS; ++a) { av[a] = tile[ti + 16 * a][k]; bv[a] = tile[tj + 16 * a][k]; }
#pragma unroll
            for (int a = 0; a < TS; ++a)
#pragma unroll
                for (int b = 0; b < TS; ++b) acc[a][b] = __dp4a((int)av[a], (int)bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < TS; ++a)
#pragma unroll
        for (int b = 0; b < TS; ++b) {
            const int i = ti + 16 * a, j = tj + 16 * b;
            if (i < B && j < B && gram_tile_index(B, i, j) >= 0) G[blk * gram_tile_entries(B) + gram_tile_index(B, i, j)] = acc[a][b];
        }
}

// ---- CUDA-core reference of the cross products (validation): bit-sliced, a*b = lo_a lo_b + 2 lo_a hi_b + 2 hi_a lo_b + 4 hi_a hi_b
__global__ void __launch_bounds__(256) cross_bits_kernel(const uint8_t *__restrict__ packed, int64_t stride, int64_t Npad,
                                                         const int32_t *__restrict__ order, int64_t n_order, int B,
                                                         int32_t *__restrict__ X)
{
    const int64_t blk = blockIdx.x;
    const int64_t nwords = Npad / 16;
    const int LA = lookahead(B);
    for (int pair = threadIdx.x; pair < LA * B; pair += blockDim.x) {
        const int jl = pair / B, k = pair % B;
        const int64_t oj = blk * B - LA + jl, ok = blk * B + k;
        int32_t acc = 0;
        if (blk > 0 && oj >= 0 && ok < n_order && order[oj] >= 0 && order[ok] >= 0) {
            const uint32_t *a = reinterpret_cast<const uint32_t *>(packed + (int64_t)order[oj] * stride);
            const uint32_t *b = reinterpret_cast<const uint32_t *>(packed + (int64_t)order[ok] * stride);
            for (int64_t w = 0; w < nwords; ++w) {
                const uint32_t x = a[w], y = b[w];
                const uint32_t xl = x & 0x55555555u, xh = (x >> 1) & 0x55555555u, yl = y & 0x55555555u, yh = (y >> 1) & 0x55555555u;
                acc += __popc(xl & yl) + 2 * (__popc(xl & yh) + __popc(xh & yl)) + 4 * __popc(xh & yh);
            }
        }
        X[(blk * LA + jl) * B + k] = acc;
    }
}

// ---- Stores with dense fp64 columns (the reference's X is any Eigen::MatrixXd, src/BayesRv2Groups.cpp:75).  The sweep
// then reads its Gram and cross tiles as fp64: entry (k, j) = sum over the local rows of c_k c_j with c the 2-bit code of a packed
// column or the value of a dense one.  Pairs of packed columns are the exact int32 counts of the tensor-core kernel, converted;
// every pair with a dense column is a fp64 dot in a fixed order (lane-strided 16-row words, xor tree), so runs and ranks reproduce
// it bit for bit.  One CTA per Gibbs block (persistent over blocks); a warp per pair.
__device__ __forceinline__ void load16(const uint8_t *__restrict__ packed_col, const double *__restrict__ dense_col, int64_t w, double (&out)[16])
{
    if (dense_col != nullptr) {
        const double2 *x2 = reinterpret_cast<const double2 *>(dense_col + w * 16);
#pragma unroll
        for (int q = 0; q < 8; ++q) { const double2 v = x2[q]; out[2 * q] = v.x; out[2 * q + 1] = v.y; }
    } else {
        const uint32_t word = reinterpret_cast<const uint32_t *>(packed_col)[w];
#pragma unroll
        for (int q = 0; q < 16; ++q) out[q] = (double)((word >> (2 * q)) & 3u);
    }
}

template <int B>
__global__ void __launch_bounds__(256) gram_dense_kernel(const uint8_t *__restrict__ packed, int64_t stride, int64_t Npad,
                                                         const double *__restrict__ dense, const int32_t *__restrict__ dense_idx,
                                                         const int32_t *__restrict__ order, int64_t n_order,
                                                         const int32_t *__restrict__ Gi, const int32_t *__restrict__ Xi,
                                                         double *__restrict__ Gd, double *__restrict__ Xd)
{
    constexpr int LA = lookahead(B), R = B + LA, TE = gram_tile_entries(B);
    __shared__ int32_t cols[R], didx[R];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nblocks = (n_order + B - 1) / B, nwords = Npad / 16;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        __syncthreads();
        if (tid < R) {
            const int64_t o = tid < B ? blk * B + tid : blk * B - LA + (tid - B);   // rows B.. : tail of the previous block
            const int32_t m = (o >= 0 && o < n_order && (tid < B || blk > 0)) ? order[o] : -1;
            cols[tid] = m; didx[tid] = m >= 0 ? dense_idx[m] : -1;
        }
        __syncthreads();
        for (int i = tid; i < TE; i += 256) Gd[blk * TE + i] = (double)Gi[blk * TE + i];
        for (int i = tid; i < LA * B; i += 256) Xd[blk * LA * B + i] = (double)Xi[blk * LA * B + i];
        __syncthreads();
        // pairs (r, j): r any of the R staged markers, j one of the block's own B, at least one of them dense; r < B is the self
        // tile, of which the upper triangle is computed and mirrored
        for (int pair = warp; pair < R * B; pair += 8) {
            const int r = pair / B, j = pair % B;
            if (cols[r] < 0 || cols[j] < 0 || (didx[r] < 0 && didx[j] < 0) || (r < B && j < r)) continue;
            const uint8_t *pr = packed + (int64_t)cols[r] * stride, *pj = packed + (int64_t)cols[j] * stride;
            const double *dr = didx[r] >= 0 ? dense + (int64_t)didx[r] * Npad : nullptr, *dj = didx[j] >= 0 ? dense + (int64_t)didx[j] * Npad : nullptr;
            double acc = 0.0;
            for (int64_t w = lane; w < nwords; w += 32) {
                double a[16], b[16];
                load16(pr, dr, w, a); load16(pj, dj, w, b);
#pragma unroll
                for (int q = 0; q < 16; ++q) acc = fma(a[q], b[q], acc);
            }
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                if (r < B) {     // (r, j) with j >= r is always stored; its mirror only inside r's own sub-window
                    Gd[blk * TE + gram_tile_index(B, r, j)] = acc;
                    if (gram_tile_index(B, j, r) >= 0) Gd[blk * TE + gram_tile_index(B, j, r)] = acc;
                }
                else Xd[(blk * LA + (r - B)) * B + j] = acc;
            }
        }
    }
}

void preload_gram_dense() { preload_kernel(gram_dense_kernel<64>); }

void launch_gram_dense(const brr_geno *g, const int32_t *d_order, int64_t n_order, int B, const int32_t *d_Gi, const int32_t *d_Xi,
                       double *d_Gd, double *d_Xd, cudaStream_t stream, int max_ctas)
{
    const int64_t nb = (n_order + B - 1) / B;
    if (nb == 0) return;
    BRR_REQUIRE(B == 64, BRR_E_SIZE, "stores with dense columns run 64-marker Gibbs blocks");
    const unsigned grid = (unsigned)(max_ctas > 0 && max_ctas < nb ? max_ctas : nb);
    gram_dense_kernel<64><<<grid, 256, 0, stream>>>(g->d_packed, g->stride, g->Npad, g->d_dense, g->d_dense_idx, d_order, n_order, d_Gi, d_Xi, d_Gd, d_Xd);
    BRR_CUDA(cudaGetLastError());
}

template <int B, bool CROSS> static size_t gram_tc_smem()
{
    constexpr int R = CROSS ? B + lookahead(B) : B;
    return 2 * gram_tile_bytes(R) + 2 * R * gram_stage_row(R) + 4 * 8 + 8 + R * 4 + 64;
}

void preload_gram(int B, int impl)
{
    if (impl == 0) {
        if (B == 32) preload_kernel(gram_tc_kernel<32, true>);
        else if (B == 64) preload_kernel(gram_tc_kernel<64, true>);
        else preload_kernel(gram_tc_kernel<128, true>);
    } else {
        if (B == 32) preload_kernel(gram_dp4a_kernel<32>);
        else if (B == 64) preload_kernel(gram_dp4a_kernel<64>);
        else preload_kernel(gram_dp4a_kernel<128>);
        preload_kernel(cross_bits_kernel);
    }
}

// launch on `stream`; G must hold nblocks * gram_tile_entries(B) int32 (tiles stored as block-upper trapezoids, common.cuh)
void launch_gram(const brr_geno *g, const int32_t *d_order, int64_t n_order, int B, int impl, int32_t *d_G, int32_t *d_X, cudaStream_t stream,
                 int max_ctas, int *abort_flag)
{

    const int64_t nb = (n_order + B - 1) / B;
    if (nb == 0) return;
    BRR_REQUIRE(B == 32 || B == 64 || B == 128, BRR_E_ARG, "block must be 32, 64 or 128");
#define BRR_GRAM_CASE(BB)                                                                                                   \
    if (B == BB) {                                                                                                          \
        if (impl == 0) {                                                                                                    \
            /* raised once per device (common.cuh, ensure_dynamic_smem) */       \
            if (d_X) ensure_dynamic_smem((const void *)gram_tc_kernel<BB, true>, gram_tc_smem<BB, true>()); \
            else ensure_dynamic_smem((const void *)gram_tc_kernel<BB, false>, gram_tc_smem<BB, false>()); \
            const unsigned grid = (unsigned)(max_ctas > 0 && max_ctas < nb ? max_ctas : nb);                                \
            if (d_X) gram_tc_kernel<BB, true><<<grid, 256, gram_tc_smem<BB, true>(), stream>>>(g->d_packed, g->stride, g->Npad, d_order, n_order, d_G, d_X, abort_flag); \
            else gram_tc_kernel<BB, false><<<grid, 256, gram_tc_smem<BB, false>(), stream>>>(g->d_packed, g->stride, g->Npad, d_order, n_order, d_G, nullptr, abort_flag); \
        } else {                                                                                                            \
            gram_dp4a_kernel<BB><<<(unsigned)nb, 256, 0, stream>>>(g->d_packed, g->stride, g->Npad, d_order, n_order, d_G);      \
        }                                                                                                                   \
    }
    BRR_GRAM_CASE(32) BRR_GRAM_CASE(64) BRR_GRAM_CASE(128)
#undef BRR_GRAM_CASE
    BRR_CUDA(cudaGetLastError());
    if (impl != 0 && d_X) {
        cross_bits_kernel<<<(unsigned)nb, 256, 0, stream>>>(g->d_packed, g->stride, g->Npad, d_order, n_order, B, d_X);
        BRR_CUDA(cudaGetLastError());
    }
}

}  // namespace brr

using namespace brr;

extern "C" int brr_gram_cross_blocks(const brr_geno *g, const int32_t *order, int64_t n_order, int block, int impl,
                                     int32_t *G_out, int32_t *X_out, double *ms)
{
    return guarded([&] {
        BRR_REQUIRE(g && order && G_out && n_order > 0, BRR_E_ARG, "bad arguments");
        BRR_REQUIRE(block == 32 || block == 64 || block == 128, BRR_E_ARG, "block must be 32, 64 or 128");
        BRR_CUDA(cudaSetDevice(g->device));
        for (int64_t i = 0; i < n_order; ++i)
            BRR_REQUIRE(order[i] >= -1 && order[i] < g->M, BRR_E_ARG, "order entry out of range");
        const int64_t nb = (n_order + block - 1) / block;
        int32_t *d_order = nullptr, *d_G = nullptr, *d_X = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
        const int TE = gram_tile_entries(block);
        std::vector<int32_t> trimmed((size_t)nb * TE);
        try {
            BRR_CUDA(cudaMalloc(&d_order, n_order * 4));
            BRR_CUDA(cudaMalloc(&d_G, (size_t)nb * TE * 4));
            if (X_out) BRR_CUDA(cudaMalloc(&d_X, (size_t)nb * lookahead(block) * block * 4));
            BRR_CUDA(cudaMemcpy(d_order, order, n_order * 4, cudaMemcpyHostToDevice));
            BRR_CUDA(cudaEventCreate(&e0)); BRR_CUDA(cudaEventCreate(&e1));
            launch_gram(g, d_order, n_order, block, impl, d_G, d_X, 0, 0, nullptr);   // warm-up (module load, attribute)
            BRR_CUDA(cudaEventRecord(e0));
            launch_gram(g, d_order, n_order, block, impl, d_G, d_X, 0, 0, nullptr);
            BRR_CUDA(cudaEventRecord(e1));
            BRR_CUDA(cudaEventSynchronize(e1));
            float t = 0; BRR_CUDA(cudaEventElapsedTime(&t, e0, e1)); if (ms) *ms = t;
            BRR_CUDA(cudaMemcpy(trimmed.data(), d_G, (size_t)nb * TE * 4, cudaMemcpyDeviceToHost));
            if (X_out) BRR_CUDA(cudaMemcpy(X_out, d_X, (size_t)nb * lookahead(block) * block * 4, cudaMemcpyDeviceToHost));
        } catch (...) { cudaFree(d_order); cudaFree(d_G); cudaFree(d_X); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); throw; }
        cudaFree(d_order); cudaFree(d_G); cudaFree(d_X); cudaEventDestroy(e0); cudaEventDestroy(e1);
        // the kernels store a tile as its block-upper trapezoid (common.cuh); the caller gets whole B x B tiles: an entry that is not
        // stored is its mirror image (the Gram is symmetric)
        for (int64_t b = 0; b < nb; ++b)
            for (int i = 0; i < block; ++i)
                for (int j = 0; j < block; ++j) {
                    const int idx = gram_tile_index(block, i, j) >= 0 ? gram_tile_index(block, i, j) : gram_tile_index(block, j, i);
                    G_out[((size_t)b * block + i) * block + j] = trimmed[(size_t)b * TE + idx];
                }
    });
}

extern "C" int brr_lookahead(int block) { return block == 32 || block == 64 || block == 128 ? lookahead(block) : 0; }

extern "C" int brr_gram_blocks(const brr_geno *g, const int32_t *order, int64_t n_order, int block, int impl,
                               int32_t *G_out, double *ms)
{
    return brr_gram_cross_blocks(g, order, n_order, block, impl, G_out, nullptr, ms);
}
