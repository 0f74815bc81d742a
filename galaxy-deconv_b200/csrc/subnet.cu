// SubNet rho predictor, one CTA per stamp, everything in shared memory.
// Replaces SubNet.forward of models/unrolled_admm_gaussian.py:61-71 (n outputs) and models/Unrolled_ADMM.py:77-90 /
// InitNet :295-308 (2n outputs): pad the PSF to 128x128, |FFT|^2, 4 x [MaxPool2 + 2 x (conv3x3 + BN(eval) + ReLU)]
// 1->4->8->16->16, flatten 1024 (+ alpha) -> Linear 1025->64 -> ReLU -> 64->64 -> ReLU -> 64->n_out -> Softplus, + 1e-6.
// |FFT|^2 is invariant under the G variant's ifftshift and under where the 48x48 PSF sits in the 128 grid, so the PSF
// is transformed at the corner of the grid.  The BatchNorms are folded into the conv weights at pack time (pack.cu).
#include "fft_block.cuh"
#include "launch.cuh"
#include "subnet.cuh"

namespace gd {

using F128 = Real2D<128, 8, 16, 48>;
constexpr int SP128 = 66;
constexpr int SN_THREADS = 256;
constexpr int SN_Z = 24 * 128;                 // float2
constexpr int SN_S = 128 * SP128;              // float2
// conv-phase buffers: feature maps are stored with a one-pixel zero border, [C][H+2][H+2], so the 3x3 taps need no bounds
// checks.  R0 (the FFT's Z+S region, 92 KB) and R1 (70 KB) ping-pong; the pooled maps (<= 4*34*34 floats) live in R2.
constexpr int SN_MAP_MAX = 4 * 66 * 66;        // floats: largest padded map (4 channels at 64x64)
constexpr int SN_POOL_MAX = 4 * 34 * 34;       // floats: largest pooled (conv-input) map: 4x34^2 = 4624 >= 1x66^2, 8x18^2, 16x10^2
constexpr int SN_WMAX = 16 * 16 * 9 + 16;      // floats
constexpr size_t SN_R0 = (size_t)(SN_Z + SN_S) * sizeof(float2);
static_assert(SN_R0 >= (size_t)SN_MAP_MAX * sizeof(float), "R0 must hold the largest padded map");
constexpr size_t SN_SMEM = SN_R0 + 128 * sizeof(float2) + (size_t)(SN_MAP_MAX + SN_POOL_MAX + SN_WMAX + 1088 + 64 + 64) * sizeof(float);

__device__ __forceinline__ float abs2(float2 a) { return a.x * a.x + a.y * a.y; }

// |H|^2 at natural frequency (k1, k2) of the 128 grid from the half spectrum (Hermitian symmetry of a real input)
__device__ __forceinline__ float hth128(const float2* S, int k1, int k2) {
    if (k2 > 64) { k1 = (128 - k1) & 127; k2 = 128 - k2; }
    return abs2(S[F128::L::slot(k1) * SP128 + k2]);
}

__device__ __forceinline__ void zero_fill(float* buf, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = 0.f;
}

// 3x3 conv + folded BN + ReLU on zero-bordered maps.  in [Cin][H+2][H+2] -> out [Cout][H+2][H+2] (border pre-zeroed).
// One work item = 4 adjacent pixels of a row x 4 output channels (16 accumulators): per (ci, tap row) three LDS.64 fetch the
// 6 inputs the 4 pixels share and three broadcast LDS.128 the weights of the 3 taps -> 48 FMAs per 6 shared-memory loads
// (the one-pixel version was bound by its 2 loads per 4 FMAs).  Every output still sums bias, then ci-major / tap-minor.
// wg: [Cin][9][Cout] weights followed by [Cout] biases (pack.cu: api.cu::gd_pack_weights).  H is a multiple of 4.
__device__ void conv3x3_relu(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ wg,
                             float* wsm, int Cin, int Cout, int H) {
    const int nw = Cout * Cin * 9 + Cout;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) wsm[i] = wg[i];
    const int Hp = H + 2, Hq = H >> 2, Q = H * Hq, ngrp = Cout >> 2;
    __syncthreads();
    const float* bias = wsm + Cout * Cin * 9;
    for (int item = threadIdx.x; item < Q * ngrp; item += blockDim.x) {
        const int cg = item / Q, p = item - cg * Q, y = p / Hq, x0 = (p - y * Hq) << 2;
        float a[4][4];
#pragma unroll
        for (int px = 0; px < 4; ++px) { a[px][0] = bias[4 * cg]; a[px][1] = bias[4 * cg + 1]; a[px][2] = bias[4 * cg + 2]; a[px][3] = bias[4 * cg + 3]; }
        const float* ip = in + y * Hp + x0;                   // top-left tap of pixel (y, x0) in the padded map (8-byte aligned)
        const float* wp = wsm + 4 * cg;
        for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                float v[6];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float2 t = *reinterpret_cast<const float2*>(ip + ky * Hp + 2 * k);
                    v[2 * k] = t.x; v[2 * k + 1] = t.y;
                }
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 w = *reinterpret_cast<const float4*>(wp + (ci * 9 + ky * 3 + kx) * Cout);
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        a[px][0] = fmaf(v[px + kx], w.x, a[px][0]); a[px][1] = fmaf(v[px + kx], w.y, a[px][1]);
                        a[px][2] = fmaf(v[px + kx], w.z, a[px][2]); a[px][3] = fmaf(v[px + kx], w.w, a[px][3]);
                    }
                }
            }
            ip += Hp * Hp;
        }
        float* op = out + (4 * cg) * Hp * Hp + (y + 1) * Hp + (x0 + 1);
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            op[px] = fmaxf(a[px][0], 0.f); op[Hp * Hp + px] = fmaxf(a[px][1], 0.f);
            op[2 * Hp * Hp + px] = fmaxf(a[px][2], 0.f); op[3 * Hp * Hp + px] = fmaxf(a[px][3], 0.f);
        }
    }
    __syncthreads();
}

// MaxPool2d(2): in [C][H+2][H+2] (bordered) -> out [C][H/2+2][H/2+2] (border pre-zeroed)
__device__ void maxpool2(const float* __restrict__ in, float* __restrict__ out, int C, int H) {
    const int Ho = H / 2, Hp = H + 2, Hop = Ho + 2;
    for (int item = threadIdx.x; item < C * Ho * Ho; item += blockDim.x) {
        int c = item / (Ho * Ho), p = item - c * Ho * Ho, y = p / Ho, x = p - y * Ho;
        const float* ip = in + c * Hp * Hp + (2 * y + 1) * Hp + 2 * x + 1;
        out[c * Hop * Hop + (y + 1) * Hop + x + 1] = fmaxf(fmaxf(ip[0], ip[1]), fmaxf(ip[Hp], ip[Hp + 1]));
    }
    __syncthreads();
}

// out[o] = act(b[o] + sum_i w[o][i] in[i]);  each warp owns n_out/nwarps outputs and keeps up to 8 of them in flight per
// pass over the inputs (8 independent coalesced weight loads per lane and step)
__device__ void linear(const float* __restrict__ w, const float* __restrict__ bvec, const float* in, float* out, int n_in,
                       int n_out, int act) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int o0 = warp * 8; o0 < n_out; o0 += nwarps * 8) {
        float s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = 0.f;
        for (int i = lane; i < n_in; i += 32) {
            const float f = in[i];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (o0 + k < n_out) s[k] = fmaf(__ldg(w + (size_t)(o0 + k) * n_in + i), f, s[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = s[k];
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0 && o0 + k < n_out) {
                v += bvec[o0 + k];
                if (act == 1) v = fmaxf(v, 0.f);
                else if (act == 2) v = (v > 20.f ? v : log1pf(expf(v))) + 1e-6f;       // nn.Softplus() + 1e-6 (:70)
                out[o0 + k] = v;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SN_THREADS) k_subnet(SubnetParams P, const float* __restrict__ psf,
                                                       const float* __restrict__ alpha, float* __restrict__ rho) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Z = reinterpret_cast<float2*>(smem_raw);
    float2* S = Z + SN_Z;
    float2* tw = S + SN_S;
    float* R0 = reinterpret_cast<float*>(smem_raw);          // aliases Z+S once the spectrum has been pooled
    float* R1 = reinterpret_cast<float*>(tw + 128);
    float* R2 = R1 + SN_MAP_MAX;
    float* wsm = R2 + SN_POOL_MAX;
    float* feat = wsm + SN_WMAX;                             // 1025 features
    float* h1 = feat + 1088;
    float* h2 = h1 + 64;
    const int b = blockIdx.x;
    const float* kb = psf + (size_t)b * NPIX;
    fill_twiddles<128>(tw);
    for (int i = threadIdx.x; i < 24 * 48; i += blockDim.x) {
        int j = i / 48, c = i - j * 48;
        Z[j * 128 + c] = make_float2(kb[(2 * j) * 48 + c], kb[(2 * j + 1) * 48 + c]);
    }
    fwd2d<F128, SP128>(Z, S, tw);
    // MaxPool2d(2) of |H|^2 in natural frequency order -> R2 [1][66][66] (zero border)
    zero_fill(R2, 66 * 66);
    __syncthreads();
    for (int item = threadIdx.x; item < 64 * 64; item += blockDim.x) {
        int i = item >> 6, j = item & 63;
        R2[(i + 1) * 66 + j + 1] = fmaxf(fmaxf(hth128(S, 2 * i, 2 * j), hth128(S, 2 * i, 2 * j + 1)),
                                         fmaxf(hth128(S, 2 * i + 1, 2 * j), hth128(S, 2 * i + 1, 2 * j + 1)));
    }
    __syncthreads();
    const int cin[4] = {1, 4, 8, 16}, cout[4] = {4, 8, 16, 16};
    int H = 64;
    for (int s = 0; s < 4; ++s) {
        const int np = cout[s] * (H + 2) * (H + 2);
        zero_fill(R0, np);
        zero_fill(R1, np);
        __syncthreads();
        conv3x3_relu(R2, R0, P.conv[2 * s], wsm, cin[s], cout[s], H);
        conv3x3_relu(R0, R1, P.conv[2 * s + 1], wsm, cout[s], cout[s], H);
        if (s < 3) {
            zero_fill(R2, cout[s] * (H / 2 + 2) * (H / 2 + 2));
            __syncthreads();
            maxpool2(R1, R2, cout[s], H);
            H >>= 1;
        }
    }
    // features: [16][8][8] flattened channel-major (x.view(N,1,1024), :68) followed by alpha
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        int c = i >> 6, y = (i >> 3) & 7, x = i & 7;
        feat[i] = R1[c * 100 + (y + 1) * 10 + x + 1];
    }
    if (threadIdx.x == 0) feat[1024] = alpha[b];
    __syncthreads();
    linear(P.l1w, P.l1b, feat, h1, 1025, 64, 1);
    linear(P.l2w, P.l2b, h1, h2, 64, 64, 1);
    linear(P.l3w, P.l3b, h2, rho + (size_t)b * P.n_out, 64, P.n_out, 2);
}

int subnet_init() {
    GD_CUDA_CHECK(cudaFuncSetAttribute(k_subnet, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SN_SMEM));
    return GD_OK;
}

int launch_subnet(const SubnetParams& P, const float* psf, const float* alpha, float* rho, int batch, cudaStream_t st) {
    if (batch <= 0) return GD_OK;
    k_subnet<<<batch, SN_THREADS, SN_SMEM, st>>>(P, psf, alpha, rho);
    GD_LAUNCHED();
    return GD_OK;
}

}  // namespace gd
