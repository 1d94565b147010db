// Micro-benchmark: do the other pipes of an sm_100a sub-partition co-issue with the packed FP32 FMA stream?
// Every case runs 8 independent FFMA2 chains (pair, pair, immediate: 2.03 cycles per warp-instruction when alone) plus
// N instructions of another kind per group; if the other pipe dispatches independently the time per group stays at
// 8 x 2.03 cycles until the other pipe itself saturates, if it shares the FMA pipe's dispatch the times add.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o mix_pipes tools/micro/mix_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float clampx(float x) { float y; asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(9.02f)); return y; }

// KIND: 0 none, 1 DFMA (pair,pair,const), 2 MUFU.RCP, 3 FMNMX, 4 LDS.128, 5 F2F f64->f32, 6 DFMA three distinct pairs, 7 scalar FFMA imm,
//       8 DMUL, 9 DADD
template <int KIND, int N>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, double da)
{
    __shared__ float4 sm[256];
    const float t = threadIdx.x;
    sm[threadIdx.x] = make_float4(t, t, t, t);
    __syncthreads();
    float2 x0 = make_float2(t, t + 1.f), x1 = make_float2(t + 2.f, t + 3.f), x2 = make_float2(t + 4.f, t + 5.f), x3 = make_float2(t + 6.f, t + 7.f);
    float2 x4 = make_float2(t + 8.f, t + 9.f), x5 = make_float2(t + 10.f, t + 11.f), x6 = make_float2(t + 12.f, t + 13.f), x7 = make_float2(t + 14.f, t + 15.f);
    const float2 a2 = make_float2(a, a + 1.f);
    const float2 im = make_float2(0.13083174824714660645f, 0.13083174824714660645f);
    double d[8]; float f[8]; float4 q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = da + i; f[i] = a + i; q[i] = make_float4(0, 0, 0, 0); }
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __ffma2_rn(x0, a2, im); x1 = __ffma2_rn(x1, a2, im); x2 = __ffma2_rn(x2, a2, im); x3 = __ffma2_rn(x3, a2, im);
            x4 = __ffma2_rn(x4, a2, im); x5 = __ffma2_rn(x5, a2, im); x6 = __ffma2_rn(x6, a2, im); x7 = __ffma2_rn(x7, a2, im);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (KIND == 1) d[j & 7] = fma(d[j & 7], da, 0.0083333333333333332);
                if (KIND == 2) f[j & 7] = rcp_approx(f[j & 7]);
                if (KIND == 3) f[j & 7] = clampx(f[j & 7] );
                if (KIND == 4) { const float4 v = sm[(idx + 8 * j) & 255]; q[j & 7].x += v.x; idx = (idx + 1) & 255; }
                if (KIND == 5) f[j & 7] += (float)d[j & 7];
                if (KIND == 6) d[j & 7] = fma(d[(j + 1) & 7], d[(j + 2) & 7], d[j & 7]);
                if (KIND == 7) f[j & 7] = fmaf(f[j & 7], a, 0.1308f);
                if (KIND == 8) d[j & 7] = __dmul_rn(d[j & 7], da);
                if (KIND == 9) d[j & 7] = __dadd_rn(d[j & 7], da);
            }
        }
    }
    float s = ((x0.x + x1.x) + (x2.x + x3.x)) + ((x4.x + x5.x) + (x6.x + x7.x)) + ((x0.y + x1.y) + (x2.y + x3.y)) + ((x4.y + x5.y) + (x6.y + x7.y));
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (float)d[i] + f[i] + q[i].x;
    if (s == 123.456f) out[0] = s;
}

template <int KIND, int N>
void run(const char *name, int ctas_per_sm)
{
    float *out;
    cudaMalloc(&out, 4);
    const int iters = 1024, grid = 148 * ctas_per_sm, threads = 256;      // ctas_per_sm x 2 warps per sub-partition
    k<KIND, N><<<grid, threads>>>(out, iters, 0.999f, 0.999);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND, N><<<grid, threads>>>(out, iters, 0.999f, 0.999);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double groups_per_smsp = (double)iters * 8 * (2.0 * ctas_per_sm);
    printf("%-52s warps/SMSP %d  cycles per group (8 FFMA2 + %2d other): %7.2f\n", name, 2 * ctas_per_sm, N, ms * 1e-3 * 1.965e9 / groups_per_smsp);
    cudaFree(out);
}

int main()
{
    run<0, 0>("FFMA2 alone", 2);
    run<1, 2>("+ DFMA pair,pair,const", 2);
    run<1, 4>("+ DFMA pair,pair,const", 2);
    run<1, 8>("+ DFMA pair,pair,const", 2);
    run<6, 4>("+ DFMA three distinct pairs", 2);
    run<8, 4>("+ DMUL", 2);
    run<9, 4>("+ DADD", 2);
    run<2, 1>("+ MUFU.RCP", 2);
    run<2, 2>("+ MUFU.RCP", 2);
    run<2, 4>("+ MUFU.RCP", 2);
    run<3, 4>("+ FMNMX.XORSIGN", 2);
    run<3, 8>("+ FMNMX.XORSIGN", 2);
    run<4, 2>("+ LDS.128", 2);
    run<4, 4>("+ LDS.128", 2);
    run<5, 2>("+ F2F.F32.F64 (+FADD)", 2);
    run<7, 4>("+ scalar FFMA imm", 2);
    run<7, 8>("+ scalar FFMA imm", 2);
    run<0, 0>("FFMA2 alone", 1);
    run<1, 4>("+ DFMA pair,pair,const", 1);
    run<2, 2>("+ MUFU.RCP", 1);
    return 0;
}
