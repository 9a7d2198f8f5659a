// TEST INFRASTRUCTURE ONLY: small kernels with known answers that exercise the SIMT emulation itself (tests/emu/simt.h)
// -- warp shuffles, votes, partial masks, barriers with exited threads, static and dynamic shared memory, 2-D grids --
// so that a green parity test over the product kernels cannot be an artefact of the emulator.  Valid CUDA: the same file
// compiles with nvcc.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void warp_scan_kernel(int n, const int *__restrict__ in, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int v = i < n ? in[i] : 0;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (i < n) out[i] = v;                       // inclusive prefix sum inside each warp
}

__global__ void votes_kernel(int n, const int *__restrict__ in, unsigned *__restrict__ ballot, int *__restrict__ any_odd,
                             int *__restrict__ all_pos, int *__restrict__ xor_partner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;                          // n is a multiple of 32 here: whole warps exit together
    const int v = in[i];
    ballot[i] = __ballot_sync(0xffffffffu, v & 1);
    any_odd[i] = __any_sync(0xffffffffu, v & 1);
    all_pos[i] = __all_sync(0xffffffffu, v > 0);
    xor_partner[i] = __shfl_xor_sync(0xffffffffu, v, 5);
}

// lanes 0..15 and 16..31 of every warp run DIFFERENT collectives on disjoint masks at the same time
__global__ void half_warp_kernel(const int *__restrict__ in, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int v = in[i];
    if (lane < 16) {
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0x0000ffffu, v, o);        // sum of the low half
    } else {
        for (int o = 8; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffff0000u, v, o)); // max of the high half
        v = __shfl_sync(0xffff0000u, v, 16) * 2;                                         // broadcast from lane 16
    }
    out[i] = v;
}

// block sum through static shared memory; WARPS that lie entirely beyond n exit BEFORE the barrier (allowed: exited
// threads do not count), the lanes of a partial warp stay and contribute zero (a shuffle must not name an exited lane)
__global__ void block_sum_kernel(int n, const float *__restrict__ in, float *__restrict__ block_sums) {
    __shared__ float s[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ((i & ~31) >= n) return;
    float v = i < n ? in[i] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        const int warps = (min(n - (int)(blockIdx.x * blockDim.x), (int)blockDim.x) + 31) / 32;
        for (int w = 0; w < warps; ++w) t += s[w];
        block_sums[blockIdx.x] = t;
    }
}

// reverse every block's slice through dynamic shared memory; 2-D grid (blockIdx.y selects the row)
__global__ void reverse_rows_kernel(int cols, const int *__restrict__ in, int *__restrict__ out) {
    extern __shared__ int tile[];
    const int row = blockIdx.y, base = blockIdx.x * blockDim.x;
    const int x = base + threadIdx.x;
    if (x < cols) tile[threadIdx.x] = in[row * cols + x];
    __syncthreads();
    const int width = min((int)blockDim.x, cols - base);
    if ((int)threadIdx.x < width) out[row * cols + base + threadIdx.x] = tile[width - 1 - threadIdx.x];
}

// every thread writes, barrier, every thread reads its neighbour: wrong without the barrier in ANY single-thread order
__global__ void neighbour_kernel(int *__restrict__ out) {
    __shared__ int s[256];
    s[threadIdx.x] = threadIdx.x * 3 + blockIdx.x;
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s[(threadIdx.x + 1) % blockDim.x] + s[(threadIdx.x + blockDim.x - 1) % blockDim.x];
}

// NEGATIVE CONTROL: lanes beyond n exit, the survivors of the partial warp then shuffle from them -- undefined in CUDA;
// the emulation must refuse it rather than invent a value
__global__ void bad_shuffle_kernel(int n, const float *__restrict__ in, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = in[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    out[i] = v;
}

// NEGATIVE CONTROL: half of a warp waits at the block barrier, the other half at a warp collective that names them
__global__ void bad_barrier_kernel(int *__restrict__ out) {
    if ((threadIdx.x & 31) < 16) __syncthreads();
    else __syncwarp(0xffffffffu);
    out[threadIdx.x] = 1;
}

}  // namespace

#define EMU_API extern "C" __attribute__((visibility("default")))

EMU_API int emu_warp_scan(int n, const int *in, int *out, void *stream) {
    warp_scan_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, in, out);
    return 0;
}
EMU_API int emu_votes(int n, const int *in, unsigned *ballot, int *any_odd, int *all_pos, int *xor_partner, void *stream) {
    votes_kernel<<<(n + 95) / 96, 96, 0, (cudaStream_t)stream>>>(n, in, ballot, any_odd, all_pos, xor_partner);
    return 0;
}
EMU_API int emu_half_warp(int n, const int *in, int *out, void *stream) {
    half_warp_kernel<<<n / 64, 64, 0, (cudaStream_t)stream>>>(in, out);
    return 0;
}
EMU_API int emu_block_sum(int n, const float *in, float *block_sums, void *stream) {
    block_sum_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, in, block_sums);
    return 0;
}
EMU_API int emu_reverse_rows(int rows, int cols, const int *in, int *out, void *stream) {
    reverse_rows_kernel<<<dim3((cols + 63) / 64, rows), 64, 64 * sizeof(int), (cudaStream_t)stream>>>(cols, in, out);
    return 0;
}
EMU_API int emu_neighbour(int blocks, int *out, void *stream) {
    neighbour_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out);
    return 0;
}
EMU_API int emu_bad_shuffle(int n, const float *in, float *out, void *stream) {
    bad_shuffle_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(n, in, out);
    return 0;
}
EMU_API int emu_bad_barrier(int *out, void *stream) {
    bad_barrier_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(out);
    return 0;
}
