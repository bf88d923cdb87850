// Variants of the packed-code x fp64 dot inner loop (16 two-bit codes per word), cycles per (warp, word, column).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double code2d_sel(uint32_t c) { return __hiloint2double(c ? (int)(0x3FE00000u + (c << 20)) : 0, 0); }
__constant__ double c_lut[4] = {0.0, 1.0, 2.0, 0.0};

template <int V>
__global__ void __launch_bounds__(256, 1) dot_kernel(const uint32_t *x, const double *eps, double *out, long long *cyc, int ncols)
{
    __shared__ uint32_t xs[128 * 32];
    __shared__ double lut_s[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 128 * 32; i += 256) xs[i] = x[i];
    if (tid < 4) lut_s[tid] = tid == 3 ? 0.0 : (double)tid;
    double e[16];
    for (int q = 0; q < 16; ++q) e[q] = eps[lane * 16 + q];
    __syncthreads();
    const long long t0 = clock64();
    double total = 0.0;
    for (int c = warp; c < ncols; c += 8) {
        const uint32_t w = xs[c * 32 + lane];
        double acc = 0.0, acc2 = 0.0;
        if (V == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) acc = fma(code2d_sel((w >> (2 * q)) & 3u), e[q], acc);
        } else if (V == 1) {      // int -> fp64 conversion
#pragma unroll
            for (int q = 0; q < 16; ++q) acc = fma((double)((w >> (2 * q)) & 3u), e[q], acc);
        } else if (V == 2) {      // bit planes, predicated adds
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (w & (1u << (2 * q))) acc += e[q];
                if (w & (2u << (2 * q))) acc2 += e[q];
            }
            acc = fma(2.0, acc2, acc);
        } else if (V == 3) {      // shared-memory LUT
#pragma unroll
            for (int q = 0; q < 16; ++q) acc = fma(lut_s[(w >> (2 * q)) & 3u], e[q], acc);
        } else if (V == 4) {      // two accumulators, select form
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
                acc = fma(code2d_sel((w >> (2 * q)) & 3u), e[q], acc);
                acc2 = fma(code2d_sel((w >> (2 * q + 2)) & 3u), e[q + 1], acc2);
            }
            acc += acc2;
        } else if (V == 5) {      // multiply-free: masks on the high/low words of e
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const uint32_t c = (w >> (2 * q)) & 3u;
                const int m1 = -(int)(c & 1u), m2 = -(int)(c >> 1);
                const double a = __hiloint2double(__double2hiint(e[q]) & m1, __double2loint(e[q]) & m1);
                const double b = __hiloint2double(__double2hiint(e[q]) & m2, __double2loint(e[q]) & m2);
                acc += a; acc2 += b;
            }
            acc = fma(2.0, acc2, acc);
        }
        total += acc;
    }
    const long long t1 = clock64();
    out[tid] = total;
    if (tid == 0) cyc[0] = t1 - t0;
}

int main()
{
    uint32_t *x; double *eps, *out; long long *cyc, h;
    cudaMalloc(&x, 128 * 32 * 4); cudaMalloc(&eps, 512 * 8); cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8);
    uint32_t hx[128 * 32]; for (int i = 0; i < 128 * 32; ++i) hx[i] = (uint32_t)(i * 2654435761u) & 0xAAAAAAAAu ? ((uint32_t)(i * 2654435761u) & 0x66666666u) : 0x11111111u;
    double he[512]; for (int i = 0; i < 512; ++i) he[i] = 0.001 * i - 0.2;
    cudaMemcpy(x, hx, sizeof hx, cudaMemcpyHostToDevice); cudaMemcpy(eps, he, sizeof he, cudaMemcpyHostToDevice);
    const char *names[] = {"select hi-word", "int->fp64 cvt", "bit-plane predicated add", "smem LUT", "select, 2 accumulators", "mask-and, no multiply"};
#define RUN(V) dot_kernel<V><<<1, 256>>>(x, eps, out, cyc, 128); dot_kernel<V><<<1, 256>>>(x, eps, out, cyc, 128); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-28s: %6lld cycles for 128 columns x 32 words/col on one SM (%.1f per column)\n", names[V], h, h / 128.0);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
