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
constexpr int SN_BUFB = 4 * 64 * 64;           // floats
constexpr int SN_WMAX = 16 * 16 * 9 + 16;      // floats
constexpr size_t SN_SMEM = (size_t)(SN_Z + SN_S + 128) * sizeof(float2) + (size_t)(SN_BUFB + SN_WMAX + 64 + 64) * sizeof(float);

__device__ __forceinline__ float abs2(float2 a) { return a.x * a.x + a.y * a.y; }

// |H|^2 at natural frequency (k1, k2) of the 128 grid from the half spectrum (Hermitian symmetry of a real input)
__device__ __forceinline__ float hth128(const float2* S, int k1, int k2) {
    if (k2 > 64) { k1 = (128 - k1) & 127; k2 = 128 - k2; }
    return abs2(S[F128::L::slot(k1) * SP128 + k2]);
}

// out[co][y][x] = relu(b[co] + sum_ci sum_tap w[co][ci][tap] in[ci][y+dy][x+dx]), zero padding
__device__ void conv3x3_relu(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ wg,
                             float* wsm, int Cin, int Cout, int H) {
    const int nw = Cout * Cin * 9 + Cout;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) wsm[i] = wg[i];
    __syncthreads();
    const int HW = H * H;
    const float* bias = wsm + Cout * Cin * 9;
    for (int item = threadIdx.x; item < Cout * HW; item += blockDim.x) {
        int co = item / HW, p = item - co * HW, y = p / H, x = p - y * H;
        float acc = bias[co];
        const float* w = wsm + co * Cin * 9;
        for (int ci = 0; ci < Cin; ++ci) {
            const float* ip = in + ci * HW;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    int xx = x + dx;
                    if (xx < 0 || xx >= H) continue;
                    acc = fmaf(w[ci * 9 + (dy + 1) * 3 + dx + 1], ip[yy * H + xx], acc);
                }
            }
        }
        out[item] = fmaxf(acc, 0.f);
    }
    __syncthreads();
}

__device__ void maxpool2(const float* __restrict__ in, float* __restrict__ out, int C, int H) {
    const int Ho = H / 2;
    for (int item = threadIdx.x; item < C * Ho * Ho; item += blockDim.x) {
        int c = item / (Ho * Ho), p = item - c * Ho * Ho, y = p / Ho, x = p - y * Ho;
        const float* ip = in + c * H * H + (2 * y) * H + 2 * x;
        out[item] = fmaxf(fmaxf(ip[0], ip[1]), fmaxf(ip[H], ip[H + 1]));
    }
    __syncthreads();
}

// out[o] = act(b[o] + sum_i w[o][i] in[i]);  one warp per output, lanes stride the inputs
__device__ void linear(const float* __restrict__ w, const float* __restrict__ bvec, const float* in, float* out, int n_in,
                       int n_out, int act) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int o = warp; o < n_out; o += nwarps) {
        float s = 0.f;
        for (int i = lane; i < n_in; i += 32) s = fmaf(w[(size_t)o * n_in + i], in[i], s);
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) {
            s += bvec[o];
            if (act == 1) s = fmaxf(s, 0.f);
            else if (act == 2) s = (s > 20.f ? s : log1pf(expf(s))) + 1e-6f;       // nn.Softplus() + 1e-6 (:70)
            out[o] = s;
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
    float* bufB = reinterpret_cast<float*>(tw + 128);
    float* wsm = bufB + SN_BUFB;
    float* h1 = wsm + SN_WMAX;
    float* h2 = h1 + 64;
    float* bufP = reinterpret_cast<float*>(Z);     // pooled maps / features (<= 4096 floats, fits the 24.5 KB of Z)
    float* bufA = reinterpret_cast<float*>(S);     // 67.6 KB >= 4*64*64 floats
    const int b = blockIdx.x;
    const float* kb = psf + (size_t)b * NPIX;
    fill_twiddles<128>(tw);
    for (int i = threadIdx.x; i < 24 * 48; i += blockDim.x) {
        int j = i / 48, c = i - j * 48;
        Z[j * 128 + c] = make_float2(kb[(2 * j) * 48 + c], kb[(2 * j + 1) * 48 + c]);
    }
    fwd2d<F128, SP128>(Z, S, tw);
    // MaxPool2d(2) of |H|^2 in natural frequency order -> bufP [64][64]
    for (int item = threadIdx.x; item < 64 * 64; item += blockDim.x) {
        int i = item >> 6, j = item & 63;
        float m = fmaxf(fmaxf(hth128(S, 2 * i, 2 * j), hth128(S, 2 * i, 2 * j + 1)),
                        fmaxf(hth128(S, 2 * i + 1, 2 * j), hth128(S, 2 * i + 1, 2 * j + 1)));
        bufP[item] = m;
    }
    __syncthreads();
    const int cin[4] = {1, 4, 8, 16}, cout[4] = {4, 8, 16, 16};
    int H = 64;
    for (int s = 0; s < 4; ++s) {
        conv3x3_relu(bufP, bufA, P.conv[2 * s], wsm, cin[s], cout[s], H);
        conv3x3_relu(bufA, bufB, P.conv[2 * s + 1], wsm, cout[s], cout[s], H);
        if (s < 3) { maxpool2(bufB, bufP, cout[s], H); H >>= 1; }
    }
    // features: bufB [16][8][8] flattened channel-major (x.view(N,1,1024), :68) followed by alpha
    if (threadIdx.x == 0) bufB[1024] = alpha[b];
    __syncthreads();
    linear(P.l1w, P.l1b, bufB, h1, 1025, 64, 1);
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
