// TEST INFRASTRUCTURE ONLY.  A stand-in for <cuda_runtime.h> that lets g++ compile the kernel files of
// geosplatting_b200/csrc for the host, so that the "-m 'not gpu'" suite can execute the REAL kernel source -- index
// arithmetic and gradient formulas included -- against the golden fixtures in this GPU-less container.
// tests/emu/build.py rewrites `kernel<<<grid, block, smem, stream>>>(args)` into gsb_emu::launch(grid, block, smem, ...)
// (args).  Two execution modes:
//   sequential (default): every thread of every block runs to completion, one after another -- for kernels without
//       shared memory or warp collectives.  GSB_HOST_EMULATION is defined and the few warp reductions of such kernels
//       route around themselves (atomics);
//   SIMT (-DGSB_EMU_SIMT, simt.h): the threads of a block are fibers that rendezvous at __syncthreads and the *_sync
//       warp primitives; shared memory is real.  The kernels compile their true code paths (GSB_HOST_EMULATION is NOT
//       defined); only inline PTX is replaced (GSB_NO_INLINE_PTX).
// Nothing in the product imports, links or ships this; the shipped library is built by nvcc for sm_100a only.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GSB_NO_INLINE_PTX 1
#ifndef GSB_EMU_SIMT
#define GSB_HOST_EMULATION 1
#endif
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static dim3 blockIdx, blockDim, threadIdx, gridDim;

struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fdividef(float a, float b) { return a / b; }
#define __expf(x) expf(x)
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fsqrt_rn(float x) { return sqrtf(x); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline void __threadfence_system() {}
static inline void __threadfence() {}

typedef void *cudaStream_t;
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
// two 'SMs': persistent kernels get a grid smaller than their job count and have to loop
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 2; return cudaSuccess; }
typedef void *cudaEvent_t;
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
// the host build is one stream in program order: events and cross-stream waits are no-ops
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }

template <class T> static inline T atomicAdd(T *p, T v) { T old = *p; *p = old + v; return old; }
static inline float2 atomicAdd(float2 *p, float2 v) {   // red.global.add.v2.f32
    float2 old = *p;
    p->x += v.x; p->y += v.y;
    return old;
}
static inline float4 atomicAdd(float4 *p, float4 v) {   // sm_90+ vector reduction (red.global.add.v4.f32)
    float4 old = *p;
    p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w;
    return old;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }

#ifdef GSB_EMU_SIMT
#define __shared__ static
#include "simt.h"
#else
// warp intrinsics have no sequential meaning: kernels route around them under GSB_HOST_EMULATION; reaching one is a bug
template <class T> static inline T __shfl_xor_sync(unsigned, T, int) { abort(); }

namespace gsb_emu {
template <class F> struct Launcher {
    dim3 grid, block;
    F f;
    template <class... A> void operator()(A... a) const {
        gridDim = grid;
        blockDim = block;
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx)
                for (unsigned t = 0; t < block.x; ++t) {
                    blockIdx = dim3(bx, by, 0);
                    threadIdx = dim3(t, 0, 0);
                    f(a...);
                }
    }
};
template <class F> static inline Launcher<F> launch(dim3 grid, dim3 block, size_t, F f) { return Launcher<F>{grid, block, f}; }
}  // namespace gsb_emu
#endif
