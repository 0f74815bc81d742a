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
    int b = (int)div_by_magic((uint32_t)m, g.magS, g.shS), r = m - b * g.S;
    int y = (int)div_by_magic((uint32_t)r, g.magW, g.shW), x = r - y * g.Wp;
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


// ---- split form used by the tcgen05 kernel: the residual / skip loads are issued BEFORE the accumulator is waited
// for, so their L2 latency overlaps the TMEM load and the previous block's stores ----
struct EpiAddr { int row, Ptot, c0; };

__device__ __forceinline__ EpiAddr epi_addr(const ConvParams& p, const RowCtx& rc, int n0) {
    EpiAddr a;
    if (p.mode == 1) {
        int tap = n0 >> p.Cf_log2;              // Cf is a power of two (32 .. 256)
        a.c0 = n0 & (p.Cf - 1);
        a.row = rc.frow0 + (tap >> 1) * p.gf.Wp + (tap & 1);
        a.Ptot = p.gf.Ptot;
    } else {
        a.c0 = n0; a.row = rc.row; a.Ptot = p.g.Ptot;
    }
    return a;
}

// r[16] = res32 + skip32 (zeros when absent)
__device__ __forceinline__ void epi_load16(const ConvParams& p, const EpiAddr& a, float* r) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = 0.f;
    if (p.res32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 t = *reinterpret_cast<const float4*>(p.res32 + ((size_t)(a.c0 / 4 + q) * a.Ptot + a.row) * 4);
            r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w;
        }
    }
    if (p.skip32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 t = *reinterpret_cast<const float4*>(p.skip32 + ((size_t)(a.c0 / 4 + q) * a.Ptot + a.row) * 4);
            r[4 * q] += t.x; r[4 * q + 1] += t.y; r[4 * q + 2] += t.z; r[4 * q + 3] += t.w;
        }
    }
}

// one 16-column group of an fp32 stream tensor
__device__ __forceinline__ void epi_load16_one(const float* src, const EpiAddr& a, float* r) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 t = __ldg(reinterpret_cast<const float4*>(src + ((size_t)(a.c0 / 4 + q) * a.Ptot + a.row) * 4));
        r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w;
    }
}

__device__ __forceinline__ uint4 pack8_half(const float* v) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    return u;
}

// hi/lo split of 8 fp32 values: hi = rn16(v), lo = rn16(v - hi)
__device__ __forceinline__ void split8_hilo(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __half2 hh = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
        h[k] = *reinterpret_cast<const uint32_t*>(&hh); l[k] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// out[0..1] = float(hi) + float(lo) of one packed pair
__device__ __forceinline__ float2 hilo_pair(uint32_t hi, uint32_t lo) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi)), b = __half22float2(*reinterpret_cast<const __half2*>(&lo));
    return make_float2(a.x + b.x, a.y + b.y);
}

// v[16] = relu?(acc) + r  ->  out32 / out16 / s2d   (same arithmetic order as epilogue_store16: (acc + res) + skip
// differs only by association of res + skip, which the fp32 validation mode does not use)
__device__ __forceinline__ void epi_store16_half(const ConvParams& p, const RowCtx& rc, const EpiAddr& a, int n0, float* v,
                                                 const float* r) {
    if (p.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (p.res32 || p.skip32) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += r[i];
    }
    if (p.out32) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(p.out32 + ((size_t)(a.c0 / 4 + q) * a.Ptot + a.row) * 4) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    if (p.out16) {
        __half* base = reinterpret_cast<__half*>(p.out16);
        *reinterpret_cast<uint4*>(base + ((size_t)(a.c0 / 8) * a.Ptot + a.row) * 8) = pack8_half(v);
        *reinterpret_cast<uint4*>(base + ((size_t)(a.c0 / 8 + 1) * a.Ptot + a.row) * 8) = pack8_half(v + 8);
    }
    if (p.s2d) {
        __half* base = reinterpret_cast<__half*>(p.s2d);
        const int cb = rc.ctap * p.N + n0;
        *reinterpret_cast<uint4*>(base + ((size_t)(cb / 8) * p.gc.Ptot + rc.crow) * 8) = pack8_half(v);
        *reinterpret_cast<uint4*>(base + ((size_t)(cb / 8 + 1) * p.gc.Ptot + rc.crow) * 8) = pack8_half(v + 8);
    }
}


// stores only (all arithmetic already applied): out32 / out16 / s2d of one 16-column group
__device__ __forceinline__ void epi_out16(const ConvParams& p, const RowCtx& rc, const EpiAddr& a, int n0, const float* v) {
    if (p.out_lo) {           // fp16 hi/lo stream: hi -> out16 (the operand copy), lo -> out_lo
        uint4 h0, l0, h1, l1;
        split8_hilo(v, h0, l0); split8_hilo(v + 8, h1, l1);
        uint4* oh = reinterpret_cast<uint4*>(p.out16) + (size_t)(a.c0 / 8) * a.Ptot + a.row;
        uint4* ol = reinterpret_cast<uint4*>(p.out_lo) + (size_t)(a.c0 / 8) * a.Ptot + a.row;
        oh[0] = h0; oh[(size_t)a.Ptot] = h1; ol[0] = l0; ol[(size_t)a.Ptot] = l1;
        if (p.s2d) {
            uint4* base = reinterpret_cast<uint4*>(p.s2d);
            const int cb = rc.ctap * p.N + n0;
            base[(size_t)(cb / 8) * p.gc.Ptot + rc.crow] = h0;
            base[(size_t)(cb / 8 + 1) * p.gc.Ptot + rc.crow] = h1;
        }
        return;
    }
    if (p.out32) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(p.out32 + ((size_t)(a.c0 / 4 + q) * a.Ptot + a.row) * 4) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    if (p.out16) {
        __half* base = reinterpret_cast<__half*>(p.out16);
        *reinterpret_cast<uint4*>(base + ((size_t)(a.c0 / 8) * a.Ptot + a.row) * 8) = pack8_half(v);
        *reinterpret_cast<uint4*>(base + ((size_t)(a.c0 / 8 + 1) * a.Ptot + a.row) * 8) = pack8_half(v + 8);
    }
    if (p.s2d) {
        __half* base = reinterpret_cast<__half*>(p.s2d);
        const int cb = rc.ctap * p.N + n0;
        *reinterpret_cast<uint4*>(base + ((size_t)(cb / 8) * p.gc.Ptot + rc.crow) * 8) = pack8_half(v);
        *reinterpret_cast<uint4*>(base + ((size_t)(cb / 8 + 1) * p.gc.Ptot + rc.crow) * 8) = pack8_half(v + 8);
    }
}

// ---- k2s2 transposed conv (mode 1) on the tcgen05 path: sector-complete pixel-shuffle stores ----
// Column order of the packed weights (api.cu pack_up, PREC_FP16_UMMA):
//     n = dy*(2*Cf) + (c/16)*32 + ((c%16)/4)*8 + dx*4 + (c%4)
// so one 32-column accumulator unit holds 16 channels [c0, c0+16) of BOTH sub-pixels dx = 0,1 of fine row 2Y+dy.
// A lane owns one coarse pixel, i.e. the fine pixels o = 2*lane + dx of the warp's run of 64; stored directly, every warp
// store would put 16 bytes at a 32-byte stride (half-filled sectors: measured ~2 TB/s and DRAM read-modify-write fills).
// The unit is therefore transposed through a warp-private 4 KB shared-memory stage (16-byte pieces XOR-swizzled, both
// directions conflict-free): afterwards lane l of store j holds fine pixel 32*j + l, and a warp store covers 512
// contiguous bytes.  Rows and validity travel with the data (shuffles), so coarse-row wrap-arounds inside the warp are
// handled exactly.  Must be called by all 32 lanes.
__device__ __forceinline__ void epi_up_unit(const ConvParams& p, bool valid, int frow0, int n0, const float* v, float4* stage) {
    const int lane = threadIdx.x & 31;
    const int dy = n0 >> (p.Cf_log2 + 1), c0 = ((n0 & (2 * p.Cf - 1)) >> 5) * 16;
    const int R0 = frow0 + dy * p.gf.Wp;
#pragma unroll
    for (int pc = 0; pc < 8; ++pc) {             // piece pc = dx*4 + g: channels 4g..4g+3 of sub-pixel dx
        const int g = pc & 3, dx = pc >> 2;
        stage[8 * lane + (pc ^ (lane & 7))] = make_float4(v[g * 8 + dx * 4], v[g * 8 + dx * 4 + 1], v[g * 8 + dx * 4 + 2], v[g * 8 + dx * 4 + 3]);
    }
    __syncwarp();
    const size_t Pf = (size_t)p.gf.Ptot;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int src = 16 * j + (lane >> 1), dx = lane & 1;
        const int row = __shfl_sync(0xffffffffu, R0, src) + dx;
        const bool ok = __shfl_sync(0xffffffffu, (int)valid, src) != 0;
        float4 x[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) x[g] = stage[8 * src + ((dx * 4 + g) ^ (src & 7))];
        if (ok) {
            if (p.out_lo) {                      // fp16 hi/lo stream
                uint4* oh = reinterpret_cast<uint4*>(p.out16) + (size_t)(c0 >> 3) * Pf + row;
                uint4* ol = reinterpret_cast<uint4*>(p.out_lo) + (size_t)(c0 >> 3) * Pf + row;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float a[8] = {x[2 * h].x, x[2 * h].y, x[2 * h].z, x[2 * h].w, x[2 * h + 1].x, x[2 * h + 1].y, x[2 * h + 1].z, x[2 * h + 1].w};
                    uint4 hi, lo;
                    split8_hilo(a, hi, lo);
                    oh[(size_t)h * Pf] = hi; ol[(size_t)h * Pf] = lo;
                }
                continue;
            }
            if (p.out32) {
                float4* o = reinterpret_cast<float4*>(p.out32) + (size_t)(c0 >> 2) * Pf + row;
#pragma unroll
                for (int g = 0; g < 4; ++g) o[(size_t)g * Pf] = x[g];
            }
            if (p.out16) {
                uint4* o = reinterpret_cast<uint4*>(p.out16) + (size_t)(c0 >> 3) * Pf + row;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float a[8] = {x[2 * h].x, x[2 * h].y, x[2 * h].z, x[2 * h].w, x[2 * h + 1].x, x[2 * h + 1].y, x[2 * h + 1].z, x[2 * h + 1].w};
                    o[(size_t)h * Pf] = pack8_half(a);
                }
            }
        }
    }
    __syncwarp();                                // the stage is reused by the warp's next unit
}

}  // namespace gd
