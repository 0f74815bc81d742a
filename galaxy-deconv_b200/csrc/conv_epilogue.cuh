// Fused conv epilogue shared by the tcgen05 kernel (conv_umma.cu) and the CUDA-core kernel (conv_simt.cu):
// ReLU / residual add (models/resnet_basicblock.py:69-71) / U-Net skip add (models/ResUNet.py:37-40) /
// operand-precision copy for the next layer / space-to-depth copy for the following strided conv /
// pixel-shuffle scatter of the k2s2 transposed conv (models/resnet_basicblock.py:81-87).
#pragma once
#include "gd_common.cuh"

namespace gd {

template <typename T> struct ActT;
template <> struct ActT<__half> { static constexpr int CH = 8; };
template <> struct ActT<float> { static constexpr int CH = 4; };

struct RowCtx {
    bool valid;
    int row;        // absolute row on the GEMM level
    int crow, ctap; // space-to-depth target row / tap on the coarse level
    int frow0;      // transposed conv: fine row of sub-pixel (0,0)
};

__device__ __forceinline__ RowCtx make_row_ctx(const ConvParams& p, int m) {
    RowCtx c;
    const Geom& g = p.g;
    c.valid = false; c.row = g.base0 + m; c.crow = 0; c.ctap = 0; c.frow0 = 0;
    if (m >= g.M) return c;
    int b = m / g.S, r = m - b * g.S;
    int y = r / g.Wp, x = r - y * g.Wp;
    if (y >= g.H || x >= g.W) return c;
    c.valid = true;
    if (p.s2d) {
        c.crow = p.gc.base0 + b * p.gc.S + (y >> 1) * p.gc.Wp + (x >> 1);
        c.ctap = ((y & 1) << 1) | (x & 1);
    }
    if (p.mode == 1) c.frow0 = p.gf.base0 + b * p.gf.S + (2 * y) * p.gf.Wp + 2 * x;
    return c;
}

// Store 16 accumulator columns [n0, n0+16) of one valid row.
template <typename T>
__device__ __forceinline__ void epilogue_store16(const ConvParams& p, const RowCtx& rc, int n0, float* v) {
    constexpr int CH = ActT<T>::CH;
    if (p.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    int row, Ptot, c0;
    if (p.mode == 1) {
        int tap = n0 / p.Cf;
        c0 = n0 - tap * p.Cf;
        row = rc.frow0 + (tap >> 1) * p.gf.Wp + (tap & 1);
        Ptot = p.gf.Ptot;
    } else {
        c0 = n0; row = rc.row; Ptot = p.g.Ptot;
    }
    if (p.res32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 r = *reinterpret_cast<const float4*>(p.res32 + ((size_t)(c0 / 4 + q) * Ptot + row) * 4);
            v[4 * q] += r.x; v[4 * q + 1] += r.y; v[4 * q + 2] += r.z; v[4 * q + 3] += r.w;
        }
    }
    if (p.skip32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 r = *reinterpret_cast<const float4*>(p.skip32 + ((size_t)(c0 / 4 + q) * Ptot + row) * 4);
            v[4 * q] += r.x; v[4 * q + 1] += r.y; v[4 * q + 2] += r.z; v[4 * q + 3] += r.w;
        }
    }
    if (p.out32) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(p.out32 + ((size_t)(c0 / 4 + q) * Ptot + row) * 4) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    // operand-precision copies
    auto put = [&](T* base, int cbase, int P, int r) {
        if constexpr (CH == 8) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                __half2 a = __floats2half2_rn(v[8 * h], v[8 * h + 1]), b = __floats2half2_rn(v[8 * h + 2], v[8 * h + 3]);
                __half2 c = __floats2half2_rn(v[8 * h + 4], v[8 * h + 5]), d = __floats2half2_rn(v[8 * h + 6], v[8 * h + 7]);
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
                u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
                *reinterpret_cast<uint4*>(base + ((size_t)(cbase / 8 + h) * P + r) * 8) = u;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(base + ((size_t)(cbase / 4 + q) * P + r) * 4) =
                    make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    };
    if (p.out16) put(reinterpret_cast<T*>(p.out16), c0, Ptot, row);
    if (p.s2d) put(reinterpret_cast<T*>(p.s2d), rc.ctap * p.N + n0, p.gc.Ptot, rc.crow);
}

}  // namespace gd
