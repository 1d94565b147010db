// Micro-benchmark: dependent-issue latency (cycles) of the instructions K1's per-step critical path is made of, one warp per SM
// sub-partition so that nothing hides it.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o latency tools/micro/latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float rcp_approx(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double rcp64(double x) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }

template <int KIND>
__global__ void k(double *out, int iters, double a, double b, long long *cyc)
{
    __shared__ float4 sm[64];
    sm[threadIdx.x & 63] = make_float4(1.f, 2.f, 3.f, (float)(threadIdx.x & 63));
    __syncthreads();
    double x = a + threadIdx.x; float f = (float)a + threadIdx.x; float2 f2 = make_float2(f, f + 1.f); int idx = threadIdx.x & 63;
    const float2 c2 = make_float2((float)b, (float)b);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (KIND == 0) x = fma(x, b, 0.5);
            if (KIND == 1) x = __dmul_rn(x, b);
            if (KIND == 2) x = __dadd_rn(x, b);
            if (KIND == 3) f2 = __ffma2_rn(f2, c2, make_float2(0.5f, 0.5f));
            if (KIND == 4) f = fmaf(f, (float)b, 0.5f);
            if (KIND == 5) f = rcp_approx(f);
            if (KIND == 6) f = __shfl_sync(0xffffffffu, f, (threadIdx.x + 1) & 31);
            if (KIND == 7) { const float4 v = sm[idx]; idx = (int)v.w; f += v.x; }
            if (KIND == 8) f = (float)((double)f * b);                  // F2F f32->f64, DMUL, F2F f64->f32
            if (KIND == 9) x = rcp64(x);
            if (KIND == 10) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
            if (KIND == 11) { float y; asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(y) : "f"(f), "f"(9.02f)); f = y; }
        }
    }
    const long long t1 = clock64();
    if (x + f + f2.x + f2.y + idx == 123.456) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND> void run(const char *name)
{
    double *out; long long *cyc, h = 0;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<KIND><<<148, 32>>>(out, iters, 1.0001, 0.99991, cyc);
    k<KIND><<<148, 32>>>(out, iters, 1.0001, 0.99991, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.2f cycles per dependent instruction\n", name, (double)h / (iters * 16.0));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("DFMA"); run<1>("DMUL"); run<2>("DADD"); run<3>("FFMA2"); run<4>("FFMA"); run<5>("MUFU.RCP"); run<6>("SHFL.IDX (32-bit)");
    run<7>("LDS.128 pointer chase (+FADD)"); run<8>("F2F.F64.F32 + DMUL + F2F.F32.F64"); run<9>("MUFU.RCP64H"); run<10>("SHFL.IDX x2 (64-bit)");
    run<11>("FMNMX.XORSIGN");
    return 0;
}
