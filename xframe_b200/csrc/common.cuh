// Shared device/host helpers for the xfb200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>

#define XFB_WARP 32

extern thread_local std::string g_xfb_err;

#define XFB_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            char buf__[512];                                                               \
            snprintf(buf__, sizeof buf__, "%s:%d %s -> %s", __FILE__, __LINE__, #call,     \
                     cudaGetErrorString(e__));                                             \
            g_xfb_err = buf__;                                                             \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

#define XFB_FAIL(...)                                                                      \
    do {                                                                                   \
        char buf__[512];                                                                   \
        snprintf(buf__, sizeof buf__, __VA_ARGS__);                                        \
        g_xfb_err = buf__;                                                                 \
        return 1;                                                                          \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// FP64 tensor-core MMA, native sm_100a shape DMMA.8x8x4.
// Fragment ownership (PTX ISA, mma.m8n8k4 .f64):
//   A (8x4): lane holds A[lane>>2][lane&3]
//   B (4x8): lane holds B[lane&3][lane>>2]
//   C (8x8): lane holds C[lane>>2][2*(lane&3)+{0,1}]
// ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit accesses (complex128 elements)
__device__ __forceinline__ double2 ldg2(const double2* p) { return __ldg(p); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// 1 / sqrt(x) for x > 0 (normal range): MUFU.RSQ64H seed (2^-22) + two Newton steps, 1-2 ulp, no slow-path call -- the
// library rsqrt() branches to a subroutine for special operands, which serialises unrolled epilogues.
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double xh = 0.5 * x;
    double r = fma(-(xh * y), y, 0.5);
    y = fma(y, r, y);
    r = fma(-(xh * y), y, 0.5);
    return fma(y, r, y);
}

// sqrt(ni / sq) restricted to (sq >= 0) & (ni >= 0), 0 elsewhere -- the multiplier of project_to_modified_intensity
// (fxs_Projections.py:899-909) with the reference's IEEE corner cases (sq = 0: inf, or NaN for ni = 0).  Two independent
// branch-free rsqrt sequences instead of a division followed by a square root: the fused FFT epilogue evaluates 8 of
// them per thread and needs the instruction-level parallelism.  2-3 ulp; subnormal operands are treated as zero.
__device__ __forceinline__ double mod_intensity_multiplier(double ni, double sq) {
    const double tiny = 2.2250738585072014e-308;              // DBL_MIN
    const double a = (sq >= tiny) ? rsqrt_pos(sq) : __longlong_as_double(0x7ff0000000000000LL);     // 1 / sqrt(sq), inf at 0
    const double b = (ni >= tiny) ? ni * rsqrt_pos(ni) : 0.0;                                        // sqrt(ni)
    return (sq >= 0.0 && ni >= 0.0) ? a * b : 0.0;
}

// 256-bit global store of two consecutive complex values (sm_100: STG.E.256); p must be 32-byte aligned
__device__ __forceinline__ void st_global_256(double2* p, double2 a, double2 b) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}
// 256-bit global loads of two consecutive complex values (LDG.E.256); p must be 32-byte aligned.  _nc: read-only data
__device__ __forceinline__ void ld_global_256_nc(const double2* p, double2& a, double2& b) {
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
__device__ __forceinline__ void ld_global_256(const double2* p, double2& a, double2& b) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p) : "memory");
}

// cudaFuncSetAttribute is per device: a process may hold plans on several GPUs (Plan(device=...)), so the "already set"
// flags are kept per device index.
struct XfbPerDeviceOnce { bool done[64] = {}; };
static inline bool xfb_first_on_device(XfbPerDeviceOnce& f) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (f.done[dev]) return false;
    f.done[dev] = true;
    return true;
}

