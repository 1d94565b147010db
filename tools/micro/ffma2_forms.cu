// Micro-benchmark: throughput of the FFMA2 operand forms on sm_100a (register pairs, 32-bit immediate broadcast,
// scalar-register broadcast), FMUL2, scalar FFMA (reg / imm), and an FFMA2 + FFMA mix.  One number per form:
// warp-instructions per cycle per SM sub-partition, from clock64 deltas over a long dependent-free stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ffma2_forms tools/micro/ffma2_forms.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP8(X) X X X X X X X X

template <int FORM>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b, unsigned long long *cyc)
{
    const float t = threadIdx.x;
    float2 x0 = make_float2(t, t + 1.f), x1 = make_float2(t + 2.f, t + 3.f), x2 = make_float2(t + 4.f, t + 5.f), x3 = make_float2(t + 6.f, t + 7.f);
    float2 x4 = make_float2(t + 8.f, t + 9.f), x5 = make_float2(t + 10.f, t + 11.f), x6 = make_float2(t + 12.f, t + 13.f), x7 = make_float2(t + 14.f, t + 15.f);
    const float2 a2 = make_float2(a, a + 1.f), b2 = make_float2(b, b + 1.f);   // distinct halves: true register pairs
    const float2 as = make_float2(a, a);                                        // equal halves: scalar broadcast form
    const float2 im = make_float2(0.13083174824714660645f, 0.13083174824714660645f);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (FORM == 0) {        // reg, reg, reg
                x0 = __ffma2_rn(x0, a2, b2); x1 = __ffma2_rn(x1, a2, b2); x2 = __ffma2_rn(x2, a2, b2); x3 = __ffma2_rn(x3, a2, b2);
                x4 = __ffma2_rn(x4, a2, b2); x5 = __ffma2_rn(x5, a2, b2); x6 = __ffma2_rn(x6, a2, b2); x7 = __ffma2_rn(x7, a2, b2);
            } else if (FORM == 1) { // reg, reg, imm
                x0 = __ffma2_rn(x0, a2, im); x1 = __ffma2_rn(x1, a2, im); x2 = __ffma2_rn(x2, a2, im); x3 = __ffma2_rn(x3, a2, im);
                x4 = __ffma2_rn(x4, a2, im); x5 = __ffma2_rn(x5, a2, im); x6 = __ffma2_rn(x6, a2, im); x7 = __ffma2_rn(x7, a2, im);
            } else if (FORM == 2) { // reg, scalar broadcast, reg
                x0 = __ffma2_rn(x0, as, b2); x1 = __ffma2_rn(x1, as, b2); x2 = __ffma2_rn(x2, as, b2); x3 = __ffma2_rn(x3, as, b2);
                x4 = __ffma2_rn(x4, as, b2); x5 = __ffma2_rn(x5, as, b2); x6 = __ffma2_rn(x6, as, b2); x7 = __ffma2_rn(x7, as, b2);
            } else if (FORM == 3) { // FMUL2 reg, reg
                x0 = __fmul2_rn(x0, a2); x1 = __fmul2_rn(x1, a2); x2 = __fmul2_rn(x2, a2); x3 = __fmul2_rn(x3, a2);
                x4 = __fmul2_rn(x4, a2); x5 = __fmul2_rn(x5, a2); x6 = __fmul2_rn(x6, a2); x7 = __fmul2_rn(x7, a2);
            } else if (FORM == 4) { // scalar FFMA reg, reg, reg (16 per group)
                x0.x = fmaf(x0.x, a, b); x1.x = fmaf(x1.x, a, b); x2.x = fmaf(x2.x, a, b); x3.x = fmaf(x3.x, a, b);
                x4.x = fmaf(x4.x, a, b); x5.x = fmaf(x5.x, a, b); x6.x = fmaf(x6.x, a, b); x7.x = fmaf(x7.x, a, b);
                x0.y = fmaf(x0.y, a, b); x1.y = fmaf(x1.y, a, b); x2.y = fmaf(x2.y, a, b); x3.y = fmaf(x3.y, a, b);
                x4.y = fmaf(x4.y, a, b); x5.y = fmaf(x5.y, a, b); x6.y = fmaf(x6.y, a, b); x7.y = fmaf(x7.y, a, b);
            } else if (FORM == 5) { // scalar FFMA reg, reg, imm (16 per group)
                x0.x = fmaf(x0.x, a, 0.1308f); x1.x = fmaf(x1.x, a, 0.1308f); x2.x = fmaf(x2.x, a, 0.1308f); x3.x = fmaf(x3.x, a, 0.1308f);
                x4.x = fmaf(x4.x, a, 0.1308f); x5.x = fmaf(x5.x, a, 0.1308f); x6.x = fmaf(x6.x, a, 0.1308f); x7.x = fmaf(x7.x, a, 0.1308f);
                x0.y = fmaf(x0.y, a, 0.1308f); x1.y = fmaf(x1.y, a, 0.1308f); x2.y = fmaf(x2.y, a, 0.1308f); x3.y = fmaf(x3.y, a, 0.1308f);
                x4.y = fmaf(x4.y, a, 0.1308f); x5.y = fmaf(x5.y, a, 0.1308f); x6.y = fmaf(x6.y, a, 0.1308f); x7.y = fmaf(x7.y, a, 0.1308f);
            } else if (FORM == 6) { // mix: 4 FFMA2 (reg) + 8 scalar FFMA
                x0 = __ffma2_rn(x0, a2, b2); x1 = __ffma2_rn(x1, a2, b2); x2 = __ffma2_rn(x2, a2, b2); x3 = __ffma2_rn(x3, a2, b2);
                x4.x = fmaf(x4.x, a, b); x5.x = fmaf(x5.x, a, b); x6.x = fmaf(x6.x, a, b); x7.x = fmaf(x7.x, a, b);
                x4.y = fmaf(x4.y, a, b); x5.y = fmaf(x5.y, a, b); x6.y = fmaf(x6.y, a, b); x7.y = fmaf(x7.y, a, b);
            } else if (FORM == 8) { // FFMA2 with three distinct register pairs, rotating operands (no operand reuse)
                x0 = __ffma2_rn(x1, x2, x0); x1 = __ffma2_rn(x2, x3, x1); x2 = __ffma2_rn(x3, x4, x2); x3 = __ffma2_rn(x4, x5, x3);
                x4 = __ffma2_rn(x5, x6, x4); x5 = __ffma2_rn(x6, x7, x5); x6 = __ffma2_rn(x7, x0, x6); x7 = __ffma2_rn(x0, x1, x7);
            } else if (FORM == 9) { // FFMA2 pair * R.F32 scalar + pair with distinct registers everywhere (fc1 / fc2 shape)
                x0 = __ffma2_rn(x1, make_float2(x2.x, x2.x), x0); x1 = __ffma2_rn(x2, make_float2(x3.y, x3.y), x1);
                x2 = __ffma2_rn(x3, make_float2(x4.x, x4.x), x2); x3 = __ffma2_rn(x4, make_float2(x5.y, x5.y), x3);
                x4 = __ffma2_rn(x5, make_float2(x6.x, x6.x), x4); x5 = __ffma2_rn(x6, make_float2(x7.y, x7.y), x5);
                x6 = __ffma2_rn(x7, make_float2(x0.x, x0.x), x6); x7 = __ffma2_rn(x0, make_float2(x1.y, x1.y), x7);
            } else if (FORM == 10) { // scalar FFMA with three distinct registers, rotating (16 per group)
                x0.x = fmaf(x1.x, x2.x, x0.x); x1.x = fmaf(x2.x, x3.x, x1.x); x2.x = fmaf(x3.x, x4.x, x2.x); x3.x = fmaf(x4.x, x5.x, x3.x);
                x4.x = fmaf(x5.x, x6.x, x4.x); x5.x = fmaf(x6.x, x7.x, x5.x); x6.x = fmaf(x7.x, x0.x, x6.x); x7.x = fmaf(x0.x, x1.x, x7.x);
                x0.y = fmaf(x1.y, x2.y, x0.y); x1.y = fmaf(x2.y, x3.y, x1.y); x2.y = fmaf(x3.y, x4.y, x2.y); x3.y = fmaf(x4.y, x5.y, x3.y);
                x4.y = fmaf(x5.y, x6.y, x4.y); x5.y = fmaf(x6.y, x7.y, x5.y); x6.y = fmaf(x7.y, x0.y, x6.y); x7.y = fmaf(x0.y, x1.y, x7.y);
            } else if (FORM == 7) { // FFMA2 pair * pair(self) + imm: the tanh Horner step  p = p*u + c
                x0 = __ffma2_rn(x0, x7, im); x1 = __ffma2_rn(x1, x7, im); x2 = __ffma2_rn(x2, x7, im); x3 = __ffma2_rn(x3, x7, im);
                x4 = __ffma2_rn(x4, x7, im); x5 = __ffma2_rn(x5, x7, im); x6 = __ffma2_rn(x6, x7, im); x0 = __ffma2_rn(x0, x7, im);
            }
        }
    }
    const long long t1 = clock64();
    const float s = ((x0.x + x1.x) + (x2.x + x3.x)) + ((x4.x + x5.x) + (x6.x + x7.x)) + ((x0.y + x1.y) + (x2.y + x3.y)) + ((x4.y + x5.y) + (x6.y + x7.y));
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (unsigned long long)(t1 - t0);
}

template <int FORM>
void run(const char *name, int per_group)
{
    float *out; unsigned long long *cyc, h = 0;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    const int iters = 2048, grid = 148 * 2, threads = 256;   // 8 warps per SM sub-partition... 2 CTAs x 8 warps = 4 per SMSP
    k<FORM><<<grid, threads>>>(out, iters, 0.999f, 0.001f, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<FORM><<<grid, threads>>>(out, iters, 0.999f, 0.001f, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst_per_warp = (double)iters * 8 * per_group;
    const double warps_per_smsp = 2.0 * 8 / 4;
    (void)h;
    printf("%-44s %8.3f ms  (SM-clock cycles per warp-instruction per SMSP at 1965 MHz: %.3f)\n", name, ms,
           ms * 1e-3 * 1.965e9 / (inst_per_warp * warps_per_smsp));
    cudaFree(out); cudaFree(cyc);
}

template <int FORM>
__global__ void __launch_bounds__(256) kd(double *out, int iters, double a, double b)
{
    const double t = threadIdx.x;
    double x0 = t, x1 = t + 1., x2 = t + 2., x3 = t + 3., x4 = t + 4., x5 = t + 5., x6 = t + 6., x7 = t + 7.;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (FORM == 0) {
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
                x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
            } else if (FORM == 1) {
                x0 = fma(x1, x2, x0); x1 = fma(x2, x3, x1); x2 = fma(x3, x4, x2); x3 = fma(x4, x5, x3);
                x4 = fma(x5, x6, x4); x5 = fma(x6, x7, x5); x6 = fma(x7, x0, x6); x7 = fma(x0, x1, x7);
            } else {
                x0 = fma(x0, x1, 0.0083333333333333332); x1 = fma(x1, x2, 0.0083333333333333332); x2 = fma(x2, x3, 0.0083333333333333332); x3 = fma(x3, x4, 0.0083333333333333332);
                x4 = fma(x4, x5, 0.0083333333333333332); x5 = fma(x5, x6, 0.0083333333333333332); x6 = fma(x6, x7, 0.0083333333333333332); x7 = fma(x7, x0, 0.0083333333333333332);
            }
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;
}

template <int FORM>
void rund(const char *name, int per_group)
{
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 2048, grid = 148 * 2, threads = 256;
    kd<FORM><<<grid, threads>>>(out, iters, 0.999, 0.001);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kd<FORM><<<grid, threads>>>(out, iters, 0.999, 0.001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms  (SM-clock cycles per warp-instruction per SMSP at 1965 MHz: %.3f)\n", name, ms,
           ms * 1e-3 * 1.965e9 / ((double)iters * 8 * per_group * 4.0));
    cudaFree(out);
}

int main()
{
    run<0>("FFMA2 reg,reg,reg", 8);
    run<1>("FFMA2 reg,reg,imm", 8);
    run<2>("FFMA2 reg,scalar-bcast,reg", 8);
    run<3>("FMUL2 reg,reg", 8);
    run<4>("FFMA reg,reg,reg", 16);
    run<5>("FFMA reg,reg,imm", 16);
    run<6>("mix 4 FFMA2 + 8 FFMA", 12);
    run<7>("FFMA2 pair,pair,imm (Horner step)", 8);
    run<8>("FFMA2 3 distinct pairs, no reuse", 8);
    run<9>("FFMA2 pair, R.F32 scalar, pair, all distinct", 8);
    run<10>("FFMA 3 distinct regs, no reuse", 16);
    rund<0>("DFMA reg,reg,reg (2 reused)", 8);
    rund<1>("DFMA 3 distinct pairs, no reuse", 8);
    rund<2>("DFMA pair,pair,const", 8);
    return 0;
}
