// Shared helpers for libgeosplat_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/geosplat_b200.h"

// Named constants of the rasterizer contract (gsplat 1.4.0 semantics).
#define GSB_ALPHA_CLAMP 0.999f
#define GSB_ALPHA_MIN (1.0f / 255.0f)
#define GSB_T_STOP 1e-4f
#define GSB_RADIUS_DET_FLOOR 0.01f
#define GSB_FOV_MARGIN 0.3f
#define GSB_COMP_EPS 1e-6f

void gsb_set_error(const char *fmt, ...);

#define GSB_CHECK_ARG(cond)                                                         \
    do {                                                                            \
        if (!(cond)) {                                                              \
            gsb_set_error("%s: invalid argument: %s", __func__, #cond);             \
            return GSB_EINVAL;                                                      \
        }                                                                           \
    } while (0)

#define GSB_CHECK_CUDA(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            gsb_set_error("%s: CUDA error %s at %s:%d", __func__, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return GSB_ECUDA;                                                       \
        }                                                                           \
    } while (0)

#define GSB_CHECK_LAUNCH() GSB_CHECK_CUDA(cudaGetLastError())

static inline int gsb_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// SMs of the current device (148 on a B200): the grid unit of the persistent kernels.
static inline int gsb_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        n = 148;
    return n;
}

// floor(log2(n_tiles)) + 1, the number of key bits the tile id occupies.
static inline __host__ __device__ int gsb_tile_bits(int n_tiles) {
    int b = 0;
    while ((1 << b) <= n_tiles) ++b;
    return b;
}

// Kernel-side camera (passed by value in the launch parameters).
struct CamK {
    float r[9];
    float t[3];
    float fx, fy, cx, cy;
    int W, H;
    float near_plane, far_plane, eps2d, radius_clip;
    int antialiased;
    int camera_id;
    int tile_w, tile_h;
};

static inline CamK gsb_make_cam(const gsb_camera *c) {
    CamK k;
    const float *vm = c->viewmat;
    k.r[0] = vm[0]; k.r[1] = vm[1]; k.r[2] = vm[2];
    k.r[3] = vm[4]; k.r[4] = vm[5]; k.r[5] = vm[6];
    k.r[6] = vm[8]; k.r[7] = vm[9]; k.r[8] = vm[10];
    k.t[0] = vm[3]; k.t[1] = vm[7]; k.t[2] = vm[11];
    k.fx = c->fx; k.fy = c->fy; k.cx = c->cx; k.cy = c->cy;
    k.W = c->width; k.H = c->height;
    k.near_plane = c->near_plane; k.far_plane = c->far_plane;
    k.eps2d = c->eps2d; k.radius_clip = c->radius_clip;
    k.antialiased = c->antialiased;
    k.camera_id = c->camera_id;
    k.tile_w = (c->width + GSB_TILE - 1) / GSB_TILE;
    k.tile_h = (c->height + GSB_TILE - 1) / GSB_TILE;
    return k;
}
