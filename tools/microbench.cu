// Latency microbenchmarks that ground the design of the serial Gibbs pass (results in DESIGN.md):
// dependent fp64 FMA chain, exp(), shuffle, shared-memory load, integer division, and a global-memory flag hop
// between two co-resident CTAs (what one grid hand-over costs).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void lat_kernel(double *out, long long *cyc, double x0, int divisor)
{
    __shared__ double sh[64];
    const int lane = threadIdx.x;
    sh[lane] = x0 + lane; sh[lane + 32] = x0;
    __syncthreads();
    double x = x0;
    long long t0, t1;
    // dependent DFMA
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, 1.0000001, 1e-9);
    t1 = clock64();
    if (lane == 0) cyc[0] = (t1 - t0) / 256;
    // dependent exp
    double y = x * 1e-3;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) y = exp(y * 1e-3);
    t1 = clock64();
    if (lane == 0) cyc[1] = (t1 - t0) / 64;
    // dependent 64-bit shuffle + add
    double z = y;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) z += __shfl_xor_sync(0xffffffffu, z, 1);
    t1 = clock64();
    if (lane == 0) cyc[2] = (t1 - t0) / 64;
    // dependent LDS.64 (pointer chase through values)
    int idx = lane & 31;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) idx = ((int)sh[idx & 63]) & 31;
    t1 = clock64();
    if (lane == 0) cyc[3] = (t1 - t0) / 64;
    // integer division by a runtime value
    int q = lane + 1000003;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) q = q / divisor + 1000003;
    t1 = clock64();
    if (lane == 0) cyc[4] = (t1 - t0) / 64;
    // reciprocal (1/x) fp64
    double r = x0 + 1.5;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) r = 1.0 / (r + 0.5);
    t1 = clock64();
    if (lane == 0) cyc[5] = (t1 - t0) / 64;
    // ballot + ffs
    unsigned m = 0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) m = __ballot_sync(0xffffffffu, (lane + m) & 1) + __ffs(m);
    t1 = clock64();
    if (lane == 0) cyc[6] = (t1 - t0) / 64;
    // independent DFMA throughput (8 chains)
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = x0 + k;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], 1.0000001, 1e-9);
    t1 = clock64();
    if (lane == 0) cyc[7] = (t1 - t0);     // 512 warp-DFMAs
    out[lane] = x + y + z + idx + q + r + m + a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7];
}

// ping-pong between CTA 0 and CTA 1 through 8-byte flagged words in global memory
__global__ void hop_kernel(volatile unsigned long long *buf, long long *cyc, int iters)
{
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            buf[0] = (unsigned long long)i;
            while (buf[16] != (unsigned long long)i) { }
        } else {
            while (buf[0] != (unsigned long long)i) { }
            buf[16] = (unsigned long long)i;
        }
    }
    long long t1 = clock64();
    if (me == 0) cyc[0] = (t1 - t0) / iters;    // round trip = 2 hops
}

// same with release/acquire + a data word (the fence-based protocol)
__global__ void hop_fence_kernel(unsigned *flag, double *data, long long *cyc, int iters)
{
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            data[0] = i; __threadfence(); atomicExch(&flag[0], (unsigned)i);
            unsigned v; do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag + 32) : "memory"); } while (v != (unsigned)i);
        } else {
            unsigned v; do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v != (unsigned)i);
            data[64] = data[0]; __threadfence(); atomicExch(&flag[32], (unsigned)i);
        }
    }
    long long t1 = clock64();
    if (me == 0) cyc[0] = (t1 - t0) / iters;
}

// one SM reading n 16-byte slots written earlier (L2 resident): bandwidth of the sampler's gather
__global__ void gather_kernel(const ulonglong2 *slots, int n, double *out, long long *cyc)
{
    long long t0 = clock64();
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        ulonglong2 v;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(slots + i));
        s += v.x ^ v.y;
    }
    out[threadIdx.x] = (double)s;
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 64);
    lat_kernel<<<1, 32>>>(out, cyc, 1.0, 4);
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("dependent DFMA          : %lld cycles\n", h[0]);
    printf("dependent exp()         : %lld cycles\n", h[1]);
    printf("dependent shfl64 + DADD : %lld cycles\n", h[2]);
    printf("dependent LDS.64 + cvt  : %lld cycles\n", h[3]);
    printf("int division (runtime)  : %lld cycles\n", h[4]);
    printf("fp64 reciprocal + add   : %lld cycles\n", h[5]);
    printf("ballot + ffs            : %lld cycles\n", h[6]);
    printf("512 independent DFMAs   : %lld cycles (1 warp)\n", h[7]);
    unsigned long long *buf; cudaMalloc(&buf, 4096); cudaMemset(buf, 0, 4096);
    hop_kernel<<<2, 32>>>(buf, cyc, 2000);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("flag ping-pong (volatile 8B words)  : %lld cycles per round trip\n", h[0]);
    unsigned *flag; double *data; cudaMalloc(&flag, 4096); cudaMalloc(&data, 4096); cudaMemset(flag, 0, 4096);
    hop_fence_kernel<<<2, 32>>>(flag, data, cyc, 2000);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("flag ping-pong (fence + atomic)     : %lld cycles per round trip\n", h[0]);
    ulonglong2 *slots; cudaMalloc(&slots, 147 * 128 * 16); cudaMemset(slots, 1, 147 * 128 * 16);
    for (int threads : {256, 512, 1024}) {
        gather_kernel<<<1, threads>>>(slots, 147 * 128, out, cyc);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("gather 147x128 16-byte slots, %4d threads: %lld cycles\n", threads, h[0]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
