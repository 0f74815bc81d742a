// Block-level building blocks on top of fft_core.cuh: phase runner, 2-D forward/inverse drivers, reductions.
#pragma once
#include <math.h>
#include "fft_core.cuh"
#include "gd_common.cuh"

namespace gd {
using namespace gdfft;

#define GD_PHASE(ITEMS, ...)                                                   \
    do {                                                                       \
        for (int w = threadIdx.x; w < (ITEMS); w += blockDim.x) { __VA_ARGS__; } \
        __syncthreads();                                                       \
    } while (0)

template <int N> __device__ __forceinline__ void fill_twiddles(float2* tw) {
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
        float s, c;
        sincospif(2.0f * (float)m / (float)N, &s, &c);
        tw[m] = make_float2(c, -s);
    }
}

// Z (packed row pairs) -> S (half spectrum).  Leading barrier: Z and tw were just written by other threads.
template <class F, int SP> __device__ __forceinline__ void fwd2d(float2* Z, float2* S, const float2* tw) {
    __syncthreads();
    GD_PHASE(F::F1_ITEMS, F::F1(w, Z, tw));
    GD_PHASE(F::F2_ITEMS, F::F2(w, Z));
    GD_PHASE(F::F3_ITEMS, F::F3(w, Z, S, SP));
    GD_PHASE(F::F4_ITEMS, F::F4(w, S, SP, tw));
    GD_PHASE(F::F5_ITEMS, F::F5(w, S, SP));
}
// S (half spectrum, destroyed) -> Z (packed row pairs of the NIN x NIN corner, unscaled by 1/N^2)
template <class F, int SP> __device__ __forceinline__ void inv2d(float2* S, float2* Z, const float2* tw) {
    __syncthreads();
    GD_PHASE(F::I1_ITEMS, F::I1(w, S, SP, tw));
    GD_PHASE(F::I2_ITEMS, F::I2(w, S, SP));
    GD_PHASE(F::I3_ITEMS, F::I3(w, S, SP, Z));
    GD_PHASE(F::I4_ITEMS, F::I4(w, Z, tw));
    GD_PHASE(F::I5_ITEMS, F::I5(w, Z));
}

__device__ __forceinline__ float block_max(float v, float* red) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
    return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r += red[i];
    return r;
}

// packed-row-pair addressing of a 48x48 stamp: pixel p = r*48+c lives in Z[(r>>1)*N + c].{x|y}
template <int N> __device__ __forceinline__ float& zpix(float2* Z, int r, int c) {
    float2& q = Z[(r >> 1) * N + c];
    return (r & 1) ? q.y : q.x;
}

// Denoiser-input scaling: the ResUNet is bias-free with ReLU only (models/ResUNet.py:11-24), hence positively
// homogeneous; each stamp is scaled by an exact power of two so that max|t| is in [0.5, 1) and fp16 operands
// stay far from overflow, and the tail multiplies the scale back ("hard parts").
__device__ __forceinline__ float pow2_scale(float amax, float* inv) {
    if (!(amax > 0.f) || !isfinite(amax)) { *inv = 1.f; return 1.f; }
    int e;
    frexpf(amax, &e);
    *inv = ldexpf(1.f, -e);
    return ldexpf(1.f, e);
}

}  // namespace gd
