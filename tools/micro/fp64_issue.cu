// Micro-benchmark: fp64 latency and issue interval seen by ONE warp (the sampler CTA's serial walk is a single warp).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double *out, long long *cyc, int iters, int active_warps_mask)
{
    const int warp = threadIdx.x >> 5;
    if (!((active_warps_mask >> warp) & 1)) return;
    double a[CHAINS];
    for (int c = 0; c < CHAINS; ++c) a[c] = 1.0 + threadIdx.x * 1e-3 + c;
    const double m = 1.0000001, b = 1e-9;
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) a[c] = fma(a[c], m, b);
    }
    const long long t1 = clock64();
    double s = 0; for (int c = 0; c < CHAINS; ++c) s += a[c];
    out[threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
}

__global__ void shfl_kernel(double *out, long long *cyc, int iters)
{
    double a = threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = __shfl_sync(0xffffffffu, a, (i + 1) & 31);
    const long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_kernel(double *out, long long *cyc, int iters)   // shfl -> dfma -> shfl -> dfma: the horseshoe chain
{
    double a = threadIdx.x * 1e-3, c = 0.5;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { const double d = __shfl_sync(0xffffffffu, fma(a, 0.999, c), i & 31); a = fma(d, -1e-3, a); }
    const long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS> void run(const char *what, int threads, int mask, double *d_out, long long *d_cyc)
{
    const int iters = 4096;
    dfma_kernel<CHAINS><<<1, threads>>>(d_out, d_cyc, iters, mask);
    dfma_kernel<CHAINS><<<1, threads>>>(d_out, d_cyc, iters, mask);
    cudaDeviceSynchronize();
    long long h[8] = {}; cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int w = 0; w < 8; ++w) if ((mask >> w) & 1) mx = h[w] > mx ? h[w] : mx;
    printf("%-58s %7.2f cycles per DFMA per warp (%d chains)\n", what, (double)mx / ((double)iters * CHAINS), CHAINS);
}

int main()
{
    double *d_out; long long *d_cyc;
    cudaMalloc(&d_out, 256 * 8); cudaMalloc(&d_cyc, 8 * 8); cudaMemset(d_cyc, 0, 64);
    run<1>("1 warp, dependent chain (latency)", 32, 1, d_out, d_cyc);
    run<2>("1 warp, 2 independent chains", 32, 1, d_out, d_cyc);
    run<4>("1 warp, 4 independent chains", 32, 1, d_out, d_cyc);
    run<8>("1 warp, 8 independent chains (issue interval)", 32, 1, d_out, d_cyc);
    run<16>("1 warp, 16 independent chains", 32, 1, d_out, d_cyc);
    run<8>("warps 0 and 4 (same scheduler), 8 chains each", 256, 0x11, d_out, d_cyc);
    run<8>("warps 0 and 1 (different schedulers), 8 chains each", 256, 0x03, d_out, d_cyc);
    run<8>("warps 0-3 (four schedulers), 8 chains each", 256, 0x0f, d_out, d_cyc);
    run<8>("8 warps, 8 chains each", 256, 0xff, d_out, d_cyc);
    const int iters = 4096; long long h = 0;
    shfl_kernel<<<1, 32>>>(d_out, d_cyc, iters); shfl_kernel<<<1, 32>>>(d_out, d_cyc, iters); cudaDeviceSynchronize();
    cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost); printf("dependent 64-bit __shfl_sync: %.2f cycles\n", (double)h / iters);
    chain_kernel<<<1, 32>>>(d_out, d_cyc, iters); chain_kernel<<<1, 32>>>(d_out, d_cyc, iters); cudaDeviceSynchronize();
    cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost); printf("dfma -> shfl -> dfma round: %.2f cycles\n", (double)h / iters);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
