// Micro-benchmark: latencies of the instructions on the serial walk's dependent chain, as ONE warp sees them (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
#define BENCH(NAME, INIT, FIN, ...)                                                          \
    __global__ void NAME(double *out, long long *cyc, uint64_t *gmem)                        \
    {                                                                                        \
        const int lane = threadIdx.x & 31;                                                   \
        (void)lane;                                                                          \
        __shared__ int sm[256];                                                              \
        for (int i = threadIdx.x; i < 256; i += 32) sm[i] = ((i * 37 + 11) & 255) * 4;       \
        __syncwarp();                                                                        \
        INIT;                                                                                \
        const long long t0 = clock64();                                                      \
        _Pragma("unroll 8") for (int i = 0; i < ITERS; ++i) { __VA_ARGS__; }                        \
        const long long t1 = clock64();                                                      \
        FIN;                                                                                 \
        if (threadIdx.x == 0) cyc[0] = t1 - t0;                                              \
    }

BENCH(k_dadd, double a = lane * 1e-3, out[lane] = a, a = a + 1.000001)
BENCH(k_dfma, double a = lane * 1e-3, out[lane] = a, a = fma(a, 0.999, 1e-3))
BENCH(k_fadd, float a = lane * 1e-3f, out[lane] = a, a = a + 1.000001f)
BENCH(k_ffma, float a = lane * 1e-3f, out[lane] = a, a = fmaf(a, 0.999f, 1e-3f))
// DSETP -> select of both halves -> DADD
BENCH(k_dsetp_sel, double a = lane * 1e-3; double b = 100.0, out[lane] = a, a = (a > b ? 0.5 : a) + 1.0)
// F2F.F32.F64 -> F2F.F64.F32 -> DADD
BENCH(k_f2f_round, double a = lane * 1e-3, out[lane] = a, { float f = (float)a; a = (double)f + 1.0; })
// F2F.F32.F64 -> integer op -> back into the double's low word
BENCH(k_f2f_down, double a = 1.0 + lane * 1e-3, out[lane] = a, { float f = (float)a; a = __hiloint2double(__double2hiint(a), __float_as_int(f) & 0xff); a += 1.0; })
BENCH(k_ex2, float a = lane * 1e-3f, out[lane] = a, { float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a)); a = e * 0.25f; })
// vote -> ffs (BREV + FLO) -> compare
BENCH(k_vote_ffs, int x = lane, out[lane] = x, { unsigned m = __ballot_sync(0xffffffffu, x > 3); int j = __ffs(m); x = j + lane; })
BENCH(k_vote_only, int x = lane, out[lane] = x, { unsigned m = __ballot_sync(0xffffffffu, x > 3); x = (int)(m & 7u) + lane; })
BENCH(k_vote_popc, int x = lane, out[lane] = x, { unsigned m = __ballot_sync(0xffffffffu, x > 3); int j = __popc((m - 1u) & ~m); x = j + lane; })
BENCH(k_shfl32, int x = lane, out[lane] = x, x = __shfl_sync(0xffffffffu, x, (x + 1) & 31))
BENCH(k_shfl64, double a = lane, out[lane] = a, a = __shfl_sync(0xffffffffu, a, (i + 1) & 31) + 1.0)
BENCH(k_lds, int x = lane * 4, out[lane] = x, x = *reinterpret_cast<volatile int *>(reinterpret_cast<char *>(sm) + x))
BENCH(k_lds_i2d_dfma, int x = lane * 4; double acc = 0.0,
      out[lane] = acc,
      { int v = *reinterpret_cast<volatile int *>(reinterpret_cast<char *>(sm) + x); acc = fma(__hiloint2double(0x43300000, v) - 4503599627370496.0, 1e-9, acc); x = ((int)acc + lane * 4) & 1020; })
// does a strong store per iteration slow a dependent ALU chain?
BENCH(k_stg_strong, int x = lane,
      out[lane] = x,
      { x = x * 3 + 1; asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(gmem + 2 * lane), "l"((uint64_t)x), "l"((uint64_t)i) : "memory"); })
BENCH(k_alu_only, int x = lane, out[lane] = x, { x = x * 3 + 1; asm volatile("" ::: "memory"); })
// branch per iteration (uniform, not taken / taken alternately)
BENCH(k_branch, int x = lane, out[lane] = x, { if ((i & 1) == 0) x = x * 3 + 1; else x = x * 5 + 2; asm volatile("" ::: "memory"); })
// the shape of one round of the walk: dadd -> dsetp -> vote -> ffs -> lds -> i2d -> dfma x3 ; beside it f2f -> ffma -> ex2 -> fadd x3 -> fsetp -> vote -> shfl64 -> dfma
BENCH(k_round, double corr = lane * 1e-3; double r0 = 0.3; int acc = 0,
      out[lane] = corr + acc,
      {
          const double num0 = r0 + corr;
          const unsigned cm = __ballot_sync(0xffffffffu, fabs(num0) > 0.1 + lane * 1e-4);
          const int js = (__ffs(cm) - 1) & 31;
          const int v = *reinterpret_cast<volatile int *>(reinterpret_cast<char *>(sm) + ((js * 8 + lane) & 255) * 4);
          const double g = fma(1e-3, __hiloint2double(0x43300000, v) - 4503599627370496.0, 1e-4) * 1e-3 + 1e-6;
          const float nf = (float)num0, n2 = nf * nf;
          float e1, e2, e3;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(0.1f, n2, -1.f)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(0.2f, n2, -2.f)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fmaf(0.3f, n2, -3.f)));
          const float c1 = 1.f + e1, c2 = c1 + e2, S = c2 + e3, t = 0.37f * S;
          const unsigned um = __ballot_sync(0xffffffffu, fabsf(t - c1) < 1e-3f * S);
          double bn = t > 1.f ? num0 * 0.5 : 0.0;
          bn = t > c1 ? num0 * 0.25 : bn;
          const double dl = __shfl_sync(0xffffffffu, bn - 1e-3, js);
          corr = fma(-g, dl, corr);
          acc += (int)um;
      })

template <typename K> void run(const char *what, K k, double *d_out, long long *d_cyc, uint64_t *d_g)
{
    k<<<1, 32>>>(d_out, d_cyc, d_g); k<<<1, 32>>>(d_out, d_cyc, d_g);
    cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %8.2f cycles per iteration\n", what, (double)h / ITERS);
}

int main()
{
    double *d_out; long long *d_cyc; uint64_t *d_g;
    cudaMalloc(&d_out, 256 * 8); cudaMalloc(&d_cyc, 64); cudaMalloc(&d_g, 4096);
    run("DADD", k_dadd, d_out, d_cyc, d_g);
    run("DFMA", k_dfma, d_out, d_cyc, d_g);
    run("FADD", k_fadd, d_out, d_cyc, d_g);
    run("FFMA", k_ffma, d_out, d_cyc, d_g);
    run("DSETP + 2 SEL + DADD", k_dsetp_sel, d_out, d_cyc, d_g);
    run("F2F.F32.F64 + F2F.F64.F32 + DADD", k_f2f_round, d_out, d_cyc, d_g);
    run("F2F.F32.F64 + LOP + DADD", k_f2f_down, d_out, d_cyc, d_g);
    run("MUFU.EX2 + FMUL", k_ex2, d_out, d_cyc, d_g);
    run("ISETP + VOTE + LOP + IADD", k_vote_only, d_out, d_cyc, d_g);
    run("ISETP + VOTE + BREV + FLO + IADD", k_vote_ffs, d_out, d_cyc, d_g);
    run("ISETP + VOTE + IADD + LOP + POPC + IADD", k_vote_popc, d_out, d_cyc, d_g);
    run("SHFL.IDX (32-bit) + IADD + LOP", k_shfl32, d_out, d_cyc, d_g);
    run("SHFL.IDX x2 (64-bit) + DADD", k_shfl64, d_out, d_cyc, d_g);
    run("LDS (pointer chase)", k_lds, d_out, d_cyc, d_g);
    run("LDS + i2d + DFMA + F2I + IADD + LOP", k_lds_i2d_dfma, d_out, d_cyc, d_g);
    run("IMAD chain + st.relaxed.gpu.v2.u64", k_stg_strong, d_out, d_cyc, d_g);
    run("IMAD chain alone", k_alu_only, d_out, d_cyc, d_g);
    run("IMAD chain + alternating uniform branch", k_branch, d_out, d_cyc, d_g);
    run("model of one round of the walk", k_round, d_out, d_cyc, d_g);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
