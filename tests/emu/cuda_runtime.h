// TEST INFRASTRUCTURE ONLY.  A stand-in for <cuda_runtime.h> that lets g++ compile the simple streaming kernels of
// geosplatting_b200/csrc (those without shared memory or warp intrinsics) for the host, so that the "-m 'not gpu'" suite
// can execute the REAL kernel source -- index arithmetic and gradient formulas included -- against the golden fixtures
// in this GPU-less container.  tests/emu/build.py rewrites `kernel<<<grid, block, 0, stream>>>(args)` into
// gsb_emu::launch(grid, block, ...)(args), which runs every thread of every block one after another.  Nothing in the
// product imports, links or ships this; the shipped library is built by nvcc for sm_100a only.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define GSB_HOST_EMULATION 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct gsb_emu_dim3 { unsigned x, y, z; };
static thread_local gsb_emu_dim3 blockIdx, blockDim, threadIdx, gridDim;

struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

typedef void *cudaStream_t;
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }

template <class T> static inline T atomicAdd(T *p, T v) { T old = *p; *p = old + v; return old; }
template <class T> static inline T __ldg(const T *p) { return *p; }

namespace gsb_emu {
template <class F> struct Launcher {
    int grid, block;
    F f;
    template <class... A> void operator()(A... a) const {
        gridDim = {(unsigned)grid, 1, 1};
        blockDim = {(unsigned)block, 1, 1};
        for (int b = 0; b < grid; ++b)
            for (int t = 0; t < block; ++t) {
                blockIdx = {(unsigned)b, 0, 0};
                threadIdx = {(unsigned)t, 0, 0};
                f(a...);
            }
    }
};
template <class F> static inline Launcher<F> launch(int grid, int block, F f) { return Launcher<F>{grid, block, f}; }
}  // namespace gsb_emu
