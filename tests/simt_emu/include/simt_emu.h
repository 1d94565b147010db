// simt_emu.h -- TEST INFRASTRUCTURE ONLY: a minimal single-threaded SIMT emulator that lets g++ compile the
// engine's CUDA sources (simple-es_b200/csrc/*.cu, *.cuh) for the HOST, so that the kernels' logic -- warp
// schedulers, ballots / shuffles / match_any, shared-memory layouts, Philox streams, the numerical contract's
// operation order -- can be checked bit for bit against the CPU oracle in a container without a GPU
// (tests/test_simt_emu.py).  It is NOT a CPU fallback of the product: the package never loads this library,
// simple_es_b200._lib.load() knows only libses_b200.so, and every product entry point fails without a CUDA device
// (tests/test_host_logic.py::test_no_cpu_fallback).  It says nothing about performance, races between warps or the
// memory model: every CUDA thread is a ucontext fiber, a warp's lanes run one after the other up to their next warp
// collective, CTAs run one after the other.  All emulator state is thread_local: a multi-GPU test runs one emulated
// rank per OS thread of one process (peer "IPC" mappings are plain pointers, the flag barrier spins across threads).
//
// What is emulated: __global__ functions called through simt::launch (the build script rewrites <<< >>>), threadIdx /
// blockIdx / blockDim / gridDim (x only), static and dynamic __shared__, __syncthreads, __syncwarp, __ballot_sync,
// __shfl_sync, __shfl_xor_sync, __shfl_up_sync, __match_any_sync, atomicAdd / atomicExch, the *_rn arithmetic intrinsics (plain IEEE
// operations: the translation unit is compiled with -ffp-contract=off), packed float2 intrinsics, bit casts, and the part
// of the runtime API ses_abi.cu uses (device memory == host memory, one synchronous stream).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <ucontext.h>

#include <vector>

#define SES_SIMT_EMU 1
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static const
#define __align__(n) alignas(n)

// ------------------------------------------------------------------------------------------------ vector types
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

// ------------------------------------------------------------------------------------------------ operation counters
// Optional (-DSIMT_EMU_COUNT, tools/emu_opcount.py): every arithmetic intrinsic a lane executes is counted, which gives the
// exact per-lane operation mix of a kernel's SOURCE (the algorithmic work behind the roofline arithmetic of DESIGN.md)
// without a profiler.  Plain C operators are not counted (the sources use the _rn intrinsics and fma()/fmaf() throughout).
namespace simt {
enum { C_F32_FMA, C_F32_MUL, C_F32_ADD, C_F32X2_FMA, C_F32X2_MUL, C_F32X2_ADD, C_F32_DIV, C_F32_SQRT, C_F32_RCP, C_F32_MINMAX,
       C_F64_FMA, C_F64_MUL, C_F64_ADD, C_F64_DIV, C_F64_SQRT, C_F64_RCP, C_SHFL, C_VOTE, C_MATCH, C_SYNCWARP, C_SYNCTHREADS, C_ATOMIC,
       C_COUNT };
#ifdef SIMT_EMU_COUNT
inline unsigned long long g_counts[C_COUNT] = {0};          // global on purpose: summed over every emulated thread of the process
#define SIMT_CNT(k) (++simt::g_counts[simt::k])
#else
#define SIMT_CNT(k) ((void)0)
#endif
}  // namespace simt

// ------------------------------------------------------------------------------------------------ fibers
namespace simt {

enum { READY = 0, WAIT_WARP = 1, WAIT_CTA = 2, DONE = 3, WAIT_NAMED = 4 };
enum { K_SYNCWARP = 1, K_BALLOT, K_SHFL, K_MATCH };

struct Warp {
    uint64_t opnd[2][32];
    uint32_t arrived[2];
    int kind[2];
    int alive, waiting;
};

struct Fiber {
    ucontext_t ctx;
    uint3 tid;
    int state, lane;
    unsigned wseq;          // number of warp collectives this lane has completed
    int nbar;               // named barrier this thread waits at (WAIT_NAMED)
    Warp *warp;
};

inline thread_local Fiber *g_cur = nullptr;
inline thread_local ucontext_t g_sched;
inline thread_local uint3 g_blockIdx = {0, 0, 0};
inline thread_local dim3 g_blockDim, g_gridDim;
inline thread_local unsigned char *g_dyn_smem = nullptr;
inline thread_local int g_cta_waiting = 0;
inline thread_local int g_named_waiting[16], g_named_count[16];     // bar.sync id, count (ids 1..15)
inline thread_local unsigned long long g_switches = 0, g_launches = 0;

inline unsigned char *dyn_smem() { return g_dyn_smem; }

[[noreturn]] inline void die(const char *msg)
{
    fprintf(stderr, "simt_emu: %s (block %u, thread %u)\n", msg, g_blockIdx.x, g_cur ? g_cur->tid.x : 0u);
    abort();
}

inline void yield()
{
    Fiber *f = g_cur;
    ++g_switches;
    swapcontext(&f->ctx, &g_sched);
}

// a lane arrives at a warp collective with its operand; returns the parity of the exchange buffer holding the results
inline int warp_collective(int kind, uint64_t v)
{
    Fiber *f = g_cur;
    if (!f) die("warp collective outside a kernel");
    Warp &w = *f->warp;
    const int par = (int)(f->wseq & 1u);
    if (w.arrived[par] && w.kind[par] != kind) die("lanes of one warp wait at different warp collectives");
    w.kind[par] = kind;
    w.opnd[par][f->lane] = v;
    w.arrived[par] |= 1u << f->lane;
    w.waiting += 1;
    f->state = WAIT_WARP;
    yield();
    f->wseq += 1;
    return par;
}

template <class T>
inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "warp collectives move at most 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

struct Stacks {
    char *base = nullptr;
    size_t n = 0;
    static constexpr size_t SZ = 256 << 10;
    char *get(size_t nthreads)
    {
        if (nthreads > n) {
            if (base) munmap(base, n * SZ);
            base = (char *)mmap(nullptr, nthreads * SZ, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (base == MAP_FAILED) die("mmap of fiber stacks failed");
            n = nthreads;
        }
        return base;
    }
};
inline thread_local Stacks g_stacks;

template <class F>
void fiber_entry(unsigned lo, unsigned hi)
{
    F *body = reinterpret_cast<F *>(((uintptr_t)hi << 32) | (uintptr_t)lo);
    (*body)();
    g_cur->state = DONE;
    // returning switches to uc_link (the scheduler)
}

// run one CTA: fibers round robin, warp by warp; a warp collective resolves when every live lane of the warp waits at it,
// __syncthreads when every live thread of the CTA does (exited threads count as arrived, as on the hardware)
template <class F>
void run_cta(unsigned nthreads, F &body)
{
    const unsigned nwarps = (nthreads + 31) / 32;
    std::vector<Fiber> fibers(nthreads);
    std::vector<Warp> warps(nwarps);
    char *stacks = g_stacks.get(nthreads);
    for (unsigned w = 0; w < nwarps; ++w) {
        memset(&warps[w], 0, sizeof(Warp));
        warps[w].alive = (int)((w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32);
    }
    const uintptr_t bp = reinterpret_cast<uintptr_t>(&body);
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber &f = fibers[t];
        f.tid = uint3{t, 0, 0};
        f.state = READY;
        f.lane = (int)(t & 31);
        f.wseq = 0;
        f.warp = &warps[t >> 5];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = stacks + (size_t)t * Stacks::SZ;
        f.ctx.uc_stack.ss_size = Stacks::SZ;
        f.ctx.uc_link = &g_sched;
        makecontext(&f.ctx, (void (*)())fiber_entry<F>, 2, (unsigned)(bp & 0xffffffffu), (unsigned)(bp >> 32));
    }
    int alive = (int)nthreads;
    g_cta_waiting = 0;
    memset(g_named_waiting, 0, sizeof(g_named_waiting));
    while (alive > 0) {
        bool progress = false;
        for (unsigned w = 0; w < nwarps; ++w) {
            Warp &wp = warps[w];
            const unsigned t0 = w * 32, t1 = t0 + 32 < nthreads ? t0 + 32 : nthreads;
            for (unsigned t = t0; t < t1; ++t) {
                Fiber &f = fibers[t];
                if (f.state != READY) continue;
                g_cur = &f;
                swapcontext(&g_sched, &f.ctx);
                g_cur = nullptr;
                progress = true;
                if (f.state == DONE) { alive -= 1; wp.alive -= 1; }
            }
            if (wp.alive > 0 && wp.waiting == wp.alive) {
                // resolve: the results stay readable in opnd[par] / arrived[par]; the other buffer is recycled
                int par = -1;
                for (unsigned t = t0; t < t1; ++t)
                    if (fibers[t].state == WAIT_WARP) {
                        const int pp = (int)(fibers[t].wseq & 1u);
                        if (par >= 0 && pp != par) die("lanes of one warp are at different collective counts");
                        par = pp;
                        fibers[t].state = READY;
                    }
                wp.arrived[par ^ 1] = 0;
                wp.waiting = 0;
                progress = true;
            }
        }
        if (alive > 0 && g_cta_waiting == alive) {
            for (unsigned t = 0; t < nthreads; ++t)
                if (fibers[t].state == WAIT_CTA) fibers[t].state = READY;
            g_cta_waiting = 0;
            progress = true;
        }
        for (int b = 1; b < 16; ++b)
            if (g_named_waiting[b] && g_named_waiting[b] == g_named_count[b]) {     // exited threads never arrive, as on the hardware
                for (unsigned t = 0; t < nthreads; ++t)
                    if (fibers[t].state == WAIT_NAMED && fibers[t].nbar == b) fibers[t].state = READY;
                g_named_waiting[b] = 0;
                progress = true;
            }
        if (!progress) die("deadlock: no thread of the CTA can run (divergent barrier?)");
    }
}

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F &&body)
{
    if (grid.y != 1 || grid.z != 1 || block.y != 1 || block.z != 1) die("only 1-D launches are emulated");
    if (block.x < 1 || block.x > 1024) die("bad block size");
    ++g_launches;
    g_gridDim = grid;
    g_blockDim = block;
    void *dyn = nullptr;
    if (smem && posix_memalign(&dyn, 128, smem)) die("out of memory (dynamic shared memory)");
    g_dyn_smem = static_cast<unsigned char *>(dyn);
    for (unsigned b = 0; b < grid.x; ++b) {
        g_blockIdx = uint3{b, 0, 0};
        if (dyn) memset(dyn, 0xA5, smem);          // shared memory starts undefined on the device
        run_cta(block.x, body);
    }
    g_dyn_smem = nullptr;
    free(dyn);
}

inline float rcp_approx(float x) { SIMT_CNT(C_F32_RCP); return 1.0f / x; }      // MUFU.RCP stand-in (the device value is within 1 ulp of this)
inline double rcp_approx_f64(double x)
{
    // MUFU.RCP64H stand-in: a reciprocal good to ~20 bits whose low word is zero
    SIMT_CNT(C_F64_RCP);
    return from_bits<double>(to_bits(1.0 / x) & 0xFFFFFFFF00000000ull);
}
inline float min_xorsign_abs(float a, float b)
{
    // min(|a|, |b|) with sign(a) ^ sign(b)
    SIMT_CNT(C_F32_MINMAX);
    const float m = fminf(fabsf(a), fabsf(b));
    return (signbit(a) != signbit(b)) ? -m : m;
}
inline unsigned lanemask_lt() { return (1u << g_cur->lane) - 1u; }
// %smid stand-in: blocks are dealt round robin over the emulated SMs (SES_SIMT_EMU_SMS, default 2)
inline unsigned smid()
{
    const char *v = getenv("SES_SIMT_EMU_SMS");
    const unsigned n = v && *v ? (unsigned)atoi(v) : 2u;
    return g_blockIdx.x % (n ? n : 1u);
}

}  // namespace simt

#define threadIdx (simt::g_cur->tid)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)

// ------------------------------------------------------------------------------------------------ barriers, collectives
inline void __syncthreads()
{
    SIMT_CNT(C_SYNCTHREADS);
    simt::Fiber *f = simt::g_cur;
    simt::g_cta_waiting += 1;
    f->state = simt::WAIT_CTA;
    simt::yield();
}
namespace simt {
// bar.sync id, count: `count` threads of the CTA meet at barrier `id` (1..15)
inline void named_barrier(int id, int count)
{
    SIMT_CNT(C_SYNCTHREADS);
    if (id < 1 || id > 15 || count < 1) die("named barrier id out of range");
    Fiber *f = g_cur;
    if (g_named_waiting[id] && g_named_count[id] != count) die("threads disagree on the count of a named barrier");
    g_named_count[id] = count;
    g_named_waiting[id] += 1;
    if (g_named_waiting[id] > count) die("more threads than its count arrived at a named barrier");
    f->nbar = id;
    f->state = WAIT_NAMED;
    yield();
}
}  // namespace simt
inline void __syncwarp(unsigned = 0xffffffffu) { SIMT_CNT(C_SYNCWARP); simt::warp_collective(simt::K_SYNCWARP, 0); }
inline void __threadfence_system() {}
inline void __threadfence() {}

inline unsigned __ballot_sync(unsigned mask, int pred)
{
    SIMT_CNT(C_VOTE);
    const int par = simt::warp_collective(simt::K_BALLOT, pred ? 1u : 0u);
    const simt::Warp &w = *simt::g_cur->warp;
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((w.arrived[par] >> l) & 1u) && w.opnd[par][l]) r |= 1u << l;
    return r & mask;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    SIMT_CNT(C_SHFL);
    const int par = simt::warp_collective(simt::K_SHFL, simt::to_bits(v));
    const simt::Warp &w = *simt::g_cur->warp;
    const int lane = simt::g_cur->lane;
    const int s = (lane & ~(width - 1)) | (src & (width - 1));
    if (!((w.arrived[par] >> s) & 1u)) return v;
    return simt::from_bits<T>(w.opnd[par][s]);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lanemask, int width = 32)
{
    SIMT_CNT(C_SHFL);
    const int par = simt::warp_collective(simt::K_SHFL, simt::to_bits(v));
    const simt::Warp &w = *simt::g_cur->warp;
    const int lane = simt::g_cur->lane;
    const int s = lane ^ lanemask;
    if ((s & ~(width - 1)) != (lane & ~(width - 1)) || !((w.arrived[par] >> s) & 1u)) return v;
    return simt::from_bits<T>(w.opnd[par][s]);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32)
{
    SIMT_CNT(C_SHFL);
    const int par = simt::warp_collective(simt::K_SHFL, simt::to_bits(v));
    const simt::Warp &w = *simt::g_cur->warp;
    const int lane = simt::g_cur->lane;
    const int s = lane - (int)delta;
    if (s < (lane & ~(width - 1)) || !((w.arrived[par] >> s) & 1u)) return v;      // below the segment: the lane keeps its own value
    return simt::from_bits<T>(w.opnd[par][s]);
}
template <class T>
inline unsigned __match_any_sync(unsigned mask, T v)
{
    SIMT_CNT(C_MATCH);
    const uint64_t mine = simt::to_bits(v);
    const int par = simt::warp_collective(simt::K_MATCH, mine);
    const simt::Warp &w = *simt::g_cur->warp;
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((w.arrived[par] >> l) & 1u) && w.opnd[par][l] == mine) r |= 1u << l;
    return r & mask;
}

template <class T, class U>
inline T atomicAdd(T *p, U v) { SIMT_CNT(C_ATOMIC); const T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U>
inline T atomicExch(T *p, U v) { const T o = *p; *p = (T)v; return o; }
template <class T, class U>
inline T atomicMax(T *p, U v) { const T o = *p; if ((T)v > o) *p = (T)v; return o; }

inline void __nanosleep(unsigned) {}
template <class T>
inline T __ldcg(const T *p) { return *p; }
inline long long clock64()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}

// ------------------------------------------------------------------------------------------------ integer intrinsics
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline unsigned __fns(unsigned mask, unsigned base, int offset)
{
    // position of the offset-th set bit of mask at or above `base` (offset >= 1), 0xffffffff if there is none
    if (offset == 0) return ((mask >> base) & 1u) ? base : 0xffffffffu;
    if (offset > 0) {
        for (unsigned b = base; b < 32; ++b)
            if ((mask >> b) & 1u) { if (--offset == 0) return b; }
        return 0xffffffffu;
    }
    for (int b = (int)base; b >= 0; --b)
        if ((mask >> b) & 1u) { if (++offset == 0) return (unsigned)b; }
    return 0xffffffffu;
}
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------------ floating point
// Separately rounded IEEE operations; the translation unit is compiled with -ffp-contract=off -mfma, so fmaf / fma are
// the only fused operations -- the numerical contract of DESIGN.md section 4.
inline float __uint_as_float(unsigned b) { return simt::from_bits<float>(b); }
inline unsigned __float_as_uint(float f) { return (unsigned)simt::to_bits(f); }
inline int __float_as_int(float f) { return (int)simt::to_bits(f); }
inline float __int_as_float(int b) { return simt::from_bits<float>((uint64_t)(unsigned)b); }
inline double __longlong_as_double(long long b) { return simt::from_bits<double>((uint64_t)b); }
inline long long __double_as_longlong(double d) { return (long long)simt::to_bits(d); }
inline float __fadd_rn(float a, float b) { SIMT_CNT(C_F32_ADD); return a + b; }
inline float __fsub_rn(float a, float b) { SIMT_CNT(C_F32_ADD); return a - b; }
inline float __fmul_rn(float a, float b) { SIMT_CNT(C_F32_MUL); return a * b; }
inline float __fdiv_rn(float a, float b) { SIMT_CNT(C_F32_DIV); return a / b; }
inline float __fmaf_rn(float a, float b, float c) { SIMT_CNT(C_F32_FMA); return __builtin_fmaf(a, b, c); }
inline float __fsqrt_rn(float a) { SIMT_CNT(C_F32_SQRT); return sqrtf(a); }
inline double __dadd_rn(double a, double b) { SIMT_CNT(C_F64_ADD); return a + b; }
inline double __dsub_rn(double a, double b) { SIMT_CNT(C_F64_ADD); return a - b; }
inline double __dmul_rn(double a, double b) { SIMT_CNT(C_F64_MUL); return a * b; }
inline double __ddiv_rn(double a, double b) { SIMT_CNT(C_F64_DIV); return a / b; }
inline double __fma_rn(double a, double b, double c) { SIMT_CNT(C_F64_FMA); return __builtin_fma(a, b, c); }
inline double __dsqrt_rn(double a) { SIMT_CNT(C_F64_SQRT); return sqrt(a); }
inline float2 __fadd2_rn(float2 a, float2 b) { SIMT_CNT(C_F32X2_ADD); return float2{a.x + b.x, a.y + b.y}; }
inline float2 __fmul2_rn(float2 a, float2 b) { SIMT_CNT(C_F32X2_MUL); return float2{a.x * b.x, a.y * b.y}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { SIMT_CNT(C_F32X2_FMA); return float2{__builtin_fmaf(a.x, b.x, c.x), __builtin_fmaf(a.y, b.y, c.y)}; }
inline long long __double2ll_rn(double a) { return llrint(a); }       // default rounding mode: to nearest even
inline int __double2int_rn(double a) { return (int)lrint(a); }
inline int __double2int_rz(double a) { return (int)a; }
inline int __float2int_rn(float a) { return (int)lrintf(a); }
inline int __float2int_rz(float a) { return (int)a; }

// ------------------------------------------------------------------------------------------------ runtime API subset
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801, cudaErrorMemoryAllocation = 2 };
typedef struct simt_stream *cudaStream_t;
typedef struct simt_event { double t; } *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp { int major, minor, multiProcessorCount, clockRate; char name[64]; };

inline const char *cudaGetErrorString(cudaError_t e)
{
    return e == cudaSuccess ? "no error" : e == cudaErrorNotSupported ? "not supported by the SIMT emulator" : "emulator error";
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof(*p));
    p->major = 10; p->minor = 0;
    p->clockRate = 1000000;          // kHz; the emulated clock64() counts nanoseconds
    const char *v = getenv("SES_SIMT_EMU_SMS");
    p->multiProcessorCount = v && *v ? atoi(v) : 2;
    snprintf(p->name, sizeof(p->name), "SIMT emulator (host)");
    return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMalloc(T **p, size_t n)
{
    void *q = nullptr;
    if (posix_memalign(&q, 256, n ? n : 1)) return cudaErrorMemoryAllocation;
    memset(q, 0xCD, n);                        // device memory starts undefined
    *p = static_cast<T *>(q);
    return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t)
{
    const char *v = getenv("SES_SIMT_EMU_CTAS_PER_SM");            // resident CTAs per emulated SM (default 2)
    *n = v && *v && atoi(v) > 0 ? atoi(v) : 2;
    return cudaSuccess;
}
// "IPC": the emulated ranks of a multi-GPU test live in ONE process (one OS thread per rank), so a memory handle is just
// the pointer and a peer mapping is the buffer itself
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new simt_event{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
constexpr unsigned cudaEventDisableTiming = 2;
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }          // emulated launches are synchronous
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(1, n); return *p ? cudaSuccess : cudaErrorNotSupported; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = (double)clock64() * 1e-6; return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }

#ifdef SIMT_EMU_COUNT
// fma() / fmaf() / fmaxf() / fminf() are called by name in the kernel sources: route them through counting wrappers
inline float simt_cnt_fmaf(float a, float b, float c) { SIMT_CNT(C_F32_FMA); return __builtin_fmaf(a, b, c); }
inline double simt_cnt_fma(double a, double b, double c) { SIMT_CNT(C_F64_FMA); return __builtin_fma(a, b, c); }
inline float simt_cnt_fmaxf(float a, float b) { SIMT_CNT(C_F32_MINMAX); return __builtin_fmaxf(a, b); }
inline float simt_cnt_fminf(float a, float b) { SIMT_CNT(C_F32_MINMAX); return __builtin_fminf(a, b); }
#define fmaf simt_cnt_fmaf
#define fma simt_cnt_fma
#define fmaxf simt_cnt_fmaxf
#define fminf simt_cnt_fminf
extern "C" inline __attribute__((used, visibility("default"))) void simt_emu_counters(unsigned long long *out, int reset)
{
    for (int i = 0; i < simt::C_COUNT; ++i) { out[i] = simt::g_counts[i]; if (reset) simt::g_counts[i] = 0; }
}
#endif
