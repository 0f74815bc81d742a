// Fused ResBlock kernel for the 32-channel, 48x48 level of the ResUNet denoiser (models/resnet_basicblock.py:69-71,
// mode 'CRC':  x + conv3x3(ReLU(conv3x3(x))) ), tcgen05 path.
//
// Why: at this level a single conv is HBM-bound -- per pixel and channel the two launches of a ResBlock move
// 2 (x16 in) + 2 (t16 out) + 2 (t16 in) + 4 (stream in) + 4 (stream out) + 2 (x16 out) = 16 bytes, 5.9 GB per 5000
// stamps, while their MMAs (N = 32: capped at 40 % of the tensor peak by the A-operand fetch, profiles/) need only
// 0.64 ms.  Here the intermediate t = ReLU(conv1(x)) never leaves the SM: 12 bytes per pixel and channel.
//
// Work item = R = 384 consecutive rows of the padded-linear layout (gd_common.cuh).  conv2 needs t on those rows plus
// a halo of Wp+1 = 50 rows on each side, i.e. 484 rows, computed as 4 tiles of 128 starting 64 rows before the item
// (MMA overhead (4+3)/(3+3) = 1.17x); conv1 in turn needs x on 612 rows, ONE bulk-copied window that feeds all 9 taps
// of all 4 tiles through shifted UMMA descriptors (as in conv_umma.cu).
//
//   phase 1 (MMA):  D1[b] (4 x 32 TMEM columns) = conv1 over the 4 t-tiles                 72 tcgen05.mma (M128 N32 K16)
//   epilogue 1:     D1[b] -> ReLU -> zero the halo pixels -> fp16 -> T[b] in shared memory, in the K-major
//                   no-swizzle operand layout [4 chunks][512 rows][8 halves]; fence.proxy.async
//   phase 2 (MMA):  D2[b] (3 x 32 columns) = conv2 over the 3 output tiles, A operand = T[b]  54 tcgen05.mma
//   epilogue 2:     D2[b] + residual (fp32 stream or recomputed m_head) -> fp32 stream / fp16 operand copy /
//                   space-to-depth copy / m_tail partial sums (the EPI_FULL / EPI_HT epilogues of conv_umma.cu)
//
// Everything is double-buffered (D1, T, D2: 512 TMEM columns, 64 KB of T) and the MMA warps run phase 1 one item ahead
// of phase 2, so the tensor pipe works on P1(i+1) while epilogue 1 converts item i, and on P2(i) while epilogue 2
// drains item i-1.  Warp roles (416 threads, one persistent CTA per SM): warp 0 producer (x windows), warps 1-4 MMA issue
// (warp 1+j owns tile j of both phases) AND epilogue 1, warps 5-12 epilogue 2 (two per TMEM lane quarter, 16 channels each).
// m_tail partial sums are written per 16-channel half: tail_part[h*9 + tap][row] (k_tail_gather sums 2 units).
#include "conv_epilogue.cuh"
#include "kernels.cuh"
#include "launch.cuh"
#include "umma_ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace gd {

constexpr int RB_C = 32;                          // channels (K = N = 32)
constexpr int RB_J1 = 4, RB_J2 = 3;               // t tiles / output tiles per item
constexpr int RB_R = RB_J2 * MTILE;               // 384 output rows per item
constexpr int RB_LEAD = 64;                       // the t window starts 64 rows before the item
constexpr int RB_TROWS = RB_J1 * MTILE;           // 512
constexpr int RB_HALO = 50;                       // Wp + 1 at 48x48
constexpr int RB_WIN = RB_TROWS + 2 * RB_HALO;    // 612 rows of x per item
constexpr int RB_A_STAGE = RB_WIN * RB_C * 2;     // 39168 B
constexpr int RB_A_STAGES = 3;
constexpr int RB_T_BYTES = RB_TROWS * RB_C * 2;   // 32768 B
constexpr int RB_W_BYTES = 9 * RB_C * RB_C * 2;   // 18432 B
constexpr int RB_I_BYTES = RB_C * RB_C * 2;       // 2048 B: fp16 identity in the B-operand layout (residual-by-MMA)
constexpr int RB_SMEM = RB_A_STAGES * RB_A_STAGE + 2 * RB_T_BYTES + 2 * RB_W_BYTES + RB_I_BYTES;   // 221,952 B

struct RbHt { float head[9 * RB_C], tail[9 * RB_C]; };

// v[0..15] += m_head(t)[C0 .. C0+15]: weights are immediate constant-bank operands (kernel parameters)
template <int C0>
__device__ __forceinline__ void rb_head(const float* __restrict__ hw, const float* hcur, float* v) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = fmaf(hcur[t], hw[t * RB_C + C0 + k], v[k]);
}
// 9 per-tap partial sums of m_tail over channels C0 .. C0+15
template <int C0>
__device__ __forceinline__ void rb_tail(const float* __restrict__ tw, const float* v, float* dst, size_t Ptot) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 16; ++k) s[k & 3] = fmaf(v[k], tw[t * RB_C + C0 + k], s[k & 3]);
        dst[(size_t)t * Ptot] = (s[0] + s[1]) + (s[2] + s[3]);
    }
}

// p1: first conv (a = x16, w = W1; its out16 is NOT written), p2: second conv with the full epilogue description.
template <int HT>
__global__ void __launch_bounds__(UMMA_THREADS, 1) k_rb_umma(const ConvParams p1, const ConvParams p2, const int total_items, const int res_mma,
                                                             const __grid_constant__ RbHt htw) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[2 * RB_A_STAGES + 1 + 12];
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    unsigned char* a_smem = smem;
    unsigned char* t_smem = smem + RB_A_STAGES * RB_A_STAGE;
    unsigned char* w_smem = t_smem + 2 * RB_T_BYTES;
    const uint32_t bar0 = smem_u32(bars);
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (RB_A_STAGES + s); };
    const uint32_t w_full = bar0 + 8u * (2 * RB_A_STAGES);
    auto d1_full = [&](int b) { return w_full + 8u * (1 + b); };
    auto d1_empty = [&](int b) { return w_full + 8u * (3 + b); };
    auto t_full = [&](int b) { return w_full + 8u * (5 + b); };
    auto t_empty = [&](int b) { return w_full + 8u * (7 + b); };
    auto d2_full = [&](int b) { return w_full + 8u * (9 + b); };
    auto d2_empty = [&](int b) { return w_full + 8u * (11 + b); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < RB_A_STAGES; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), RB_J1); }
        mbar_init(w_full, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(d1_full(b), RB_J1); mbar_init(d1_empty(b), 4);
            mbar_init(t_full(b), 4); mbar_init(t_empty(b), RB_J2);
            mbar_init(d2_full(b), RB_J2); mbar_init(d2_empty(b), EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // identity B operand [K/8][N][8]: element (k, n) = (k == n).  With it the tensor pipe adds the hi half of the residual
        // (= the ResBlock's own fp16 input, already resident in the x window) to D2: one more "tap", no LSU traffic.
        __half* I = reinterpret_cast<__half*>(w_smem + 2 * RB_W_BYTES);
        for (int i = threadIdx.x; i < RB_C * RB_C; i += blockDim.x) {
            const int kc = i / (RB_C * 8), n = (i / 8) % RB_C, kk = i % 8;
            I[i] = __float2half(kc * 8 + kk == n ? 1.f : 0.f);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const Geom& g = p2.g;
    const int n_my = total_items > (int)blockIdx.x ? (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0) {
        // ===== producer: one x window per item =====
        if (lane == 0) {
            mbar_expect_tx(w_full, 2u * RB_W_BYTES);
            bulk_g2s(smem_u32(w_smem), p1.w, RB_W_BYTES, w_full);
            bulk_g2s(smem_u32(w_smem) + RB_W_BYTES, p2.w, RB_W_BYTES, w_full);
            const unsigned char* act = reinterpret_cast<const unsigned char*>(p1.a);
            int as = 0, aph = 0;
            for (int s = 0; s < n_my; ++s) {
                const int im = (int)blockIdx.x + s * (int)gridDim.x;
                const size_t win0 = (size_t)g.base0 + (size_t)im * RB_R - RB_LEAD - RB_HALO;
                mbar_wait(a_empty(as), aph ^ 1);
                mbar_expect_tx(a_full(as), (uint32_t)RB_A_STAGE);
                const uint32_t adst = smem_u32(a_smem + (size_t)as * RB_A_STAGE);
#pragma unroll
                for (int ch = 0; ch < RB_C / 8; ++ch)
                    bulk_g2s(adst + (uint32_t)ch * RB_WIN * 16, act + ((size_t)ch * g.Ptot + win0) * 16, (uint32_t)RB_WIN * 16, a_full(as));
                if (++as == RB_A_STAGES) { as = 0; aph ^= 1; }
            }
        }
    } else if (warp <= MMA_WARPS) {
        // ===== MMA issue + epilogue 1.  Warp 1+jw issues tile jw of phase 1 and (jw < 3) of phase 2 from one elected lane;
        // phase 1 runs one item ahead of phase 2.  The same four warps (TMEM lane quarters 1,2,3,0) then convert D1 into the
        // shared-memory operand T while the tensor pipe works through the queued phase 2 of the previous item. =====
        const int jw = warp - 1, q = warp & 3;
        const uint32_t leader = lane == 0;
        const uint32_t idesc = instr_desc_f16(MTILE, RB_C);
        const uint64_t a_desc0 = smem_desc(smem_u32(a_smem), RB_WIN * 16, 128);
        const uint64_t t_desc0 = smem_desc(smem_u32(t_smem), RB_TROWS * 16, 128);
        const uint64_t w1_desc0 = smem_desc(smem_u32(w_smem), RB_C * 16, 128);
        const uint64_t w2_desc0 = smem_desc(smem_u32(w_smem) + RB_W_BYTES, RB_C * 16, 128);
        const uint64_t i_desc0 = smem_desc(smem_u32(w_smem) + 2 * RB_W_BYTES, RB_C * 16, 128);
        constexpr uint32_t A_KK = (2 * RB_WIN * 16) >> 4, T_KK = (2 * RB_TROWS * 16) >> 4, W_KK = (2 * RB_C * 16) >> 4;
        constexpr uint32_t W_TAP = (RB_C / 8) * RB_C;       // 16-byte units per tap of the packed weights
        mbar_wait(w_full, 0);
        int as = 0, aph = 0;
        for (int s = 0; s <= n_my; ++s) {
            const int b = s & 1, ph = (s >> 1) & 1;
            if (s < n_my) {                                  // ---- phase 1 of item s ----
                mbar_wait(d1_empty(b), ph ^ 1);
                mbar_wait(a_full(as), aph);
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)(b * 128 + jw * RB_C);
                const uint64_t ad = a_desc0 + (uint64_t)((uint32_t)as * (RB_A_STAGE >> 4) + (uint32_t)(RB_HALO + jw * MTILE));
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint64_t at = ad + (uint64_t)(int64_t)p1.off[tap];
                        const uint64_t bt = w1_desc0 + (uint64_t)(tap * W_TAP);
                        if (tap == 0) tc_mma_f16(d, at, bt, idesc, 0u); else tc_mma_f16_acc(d, at, bt, idesc);
                        tc_mma_f16_acc(d, at + A_KK, bt + W_KK, idesc);
                    }
                }
                __syncwarp();
                if (!res_mma || jw >= RB_J2) tc_commit_pred(a_empty(as), leader);     // res_mma: warps 0-2 release the window after phase 2
                tc_commit_pred(d1_full(b), leader);
                if (++as == RB_A_STAGES) { as = 0; aph ^= 1; }
            }
            if (s >= 1 && jw < RB_J2) {                      // ---- phase 2 of item s-1 ----
                const int i = s - 1, b2 = i & 1, ph2 = (i >> 1) & 1;
                mbar_wait(d2_empty(b2), ph2 ^ 1);
                mbar_wait(t_full(b2), ph2);
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)(256 + b2 * 128 + jw * RB_C);
                const uint64_t ad = t_desc0 + (uint64_t)((uint32_t)b2 * (RB_T_BYTES >> 4) + (uint32_t)(RB_LEAD + jw * MTILE));
                if (elect_one()) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint64_t at = ad + (uint64_t)(int64_t)p2.off[tap];
                        const uint64_t bt = w2_desc0 + (uint64_t)(tap * W_TAP);
                        if (tap == 0) tc_mma_f16(d, at, bt, idesc, 0u); else tc_mma_f16_acc(d, at, bt, idesc);
                        tc_mma_f16_acc(d, at + T_KK, bt + W_KK, idesc);
                    }
                    if (res_mma) {                           // D2 += x_hi * I  (output rows of tile jw inside the x window of item i)
                        const uint32_t as2 = (uint32_t)(i % RB_A_STAGES);
                        const uint64_t ax = a_desc0 + (uint64_t)(as2 * (RB_A_STAGE >> 4) + (uint32_t)(RB_HALO + RB_LEAD + jw * MTILE));
                        tc_mma_f16_acc(d, ax, i_desc0, idesc);
                        tc_mma_f16_acc(d, ax + A_KK, i_desc0 + W_KK, idesc);
                    }
                }
                __syncwarp();
                if (res_mma) tc_commit_pred(a_empty(i % RB_A_STAGES), leader);
                tc_commit_pred(t_empty(b2), leader);
                tc_commit_pred(d2_full(b2), leader);
            }
            if (s < n_my) {                                  // ---- epilogue 1 of item s: D1 -> ReLU -> fp16 -> T; halo pixels = 0 ----
                const int im = (int)blockIdx.x + s * (int)gridDim.x;
                mbar_wait(t_empty(b), ph ^ 1);               // phase 2 of item s-2 has finished reading T[b]
                mbar_wait(d1_full(b), ph);
                tc_fence_after();
                uint4* T = reinterpret_cast<uint4*>(t_smem + (size_t)b * RB_T_BYTES);
#pragma unroll
                for (int j = 0; j < RB_J1; ++j) {
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 128 + j * RB_C);
                    uint32_t r0[16], r1[16];
                    tc_ld16_nowait(taddr, r0);
                    tc_ld16_nowait(taddr + 16, r1);
                    const int tr = j * MTILE + q * 32 + lane;
                    const int m = im * RB_R - RB_LEAD + tr;  // GEMM row of this t row (may lie outside [0, M))
                    bool valid = m >= 0 && m < g.M;
                    if (valid) {
                        const uint32_t bq = div_by_magic((uint32_t)m, g.magS, g.shS), r = (uint32_t)m - bq * (uint32_t)g.S;
                        const uint32_t y = div_by_magic(r, g.magW, g.shW), x = r - y * (uint32_t)g.Wp;
                        valid = (int)y < g.H && (int)x < g.W;
                    }
                    tc_ld_wait16(r0);
                    tc_ld_wait16(r1);
                    float v[32];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        v[k] = valid ? fmaxf(__uint_as_float(r0[k]), 0.f) : 0.f;
                        v[16 + k] = valid ? fmaxf(__uint_as_float(r1[k]), 0.f) : 0.f;
                    }
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) T[ch * RB_TROWS + tr] = pack8_half(v + 8 * ch);
                }
                tc_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA's async proxy
                __syncwarp();
                if (lane == 0) { mbar_arrive(d1_empty(b)); mbar_arrive(t_full(b)); }
            }
        }
    } else {
        // ===== epilogue 2: D2 + residual -> outputs.  Warp (q, h) owns TMEM lanes 32q..32q+31 and channels 16h..16h+15 of the
        // three output tiles; the fp32 residual (and head input) of the NEXT item is requested at the end of this item. =====
        const int q = warp & 3, h = (warp - EPI_WARP0) >> 2, c0 = 16 * h;
        const uint32_t Ptot = (uint32_t)g.Ptot;
        float add[RB_J2][16];
        auto issue_res = [&](int im, int j) {
            if (!p2.res32 && !p2.res_hi) return;
            const int m = im * RB_R + j * MTILE + q * 32 + lane;
            if (m >= g.M) return;
            if (p2.res_hi) {                     // fp16 hi/lo stream: raw packed halves (add[j][0..7] = hi, [8..15] = lo), decoded at use
                const size_t o = (size_t)(c0 >> 3) * Ptot + (g.base0 + m);
                const uint4* sh = reinterpret_cast<const uint4*>(p2.res_hi) + o;
                const uint4* sl = reinterpret_cast<const uint4*>(p2.res_lo) + o;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint4 u = __ldg(sl + (size_t)k * Ptot);
                    const uint4 t = res_mma ? make_uint4(0u, 0u, 0u, 0u) : __ldg(sh + (size_t)k * Ptot);   // res_mma: hi was added by the tensor pipe
                    add[j][4 * k] = __uint_as_float(t.x); add[j][4 * k + 1] = __uint_as_float(t.y); add[j][4 * k + 2] = __uint_as_float(t.z); add[j][4 * k + 3] = __uint_as_float(t.w);
                    add[j][8 + 4 * k] = __uint_as_float(u.x); add[j][9 + 4 * k] = __uint_as_float(u.y); add[j][10 + 4 * k] = __uint_as_float(u.z); add[j][11 + 4 * k] = __uint_as_float(u.w);
                }
                return;
            }
            const float4* src = reinterpret_cast<const float4*>(p2.res32) + (size_t)(c0 >> 2) * Ptot + (g.base0 + m);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 t = __ldg(src + (size_t)k * Ptot);
                add[j][4 * k] = t.x; add[j][4 * k + 1] = t.y; add[j][4 * k + 2] = t.z; add[j][4 * k + 3] = t.w;
            }
        };
        // HT: 3x3 neighbourhood of the 1-channel denoiser input of every row, also requested one item ahead
        float hin[HT ? RB_J2 : 1][9];
        auto issue_head = [&](int im, int j) {
            if (!HT || !p2.head_t) return;
            const int m = im * RB_R + j * MTILE + q * 32 + lane;
            if (m >= g.M) return;
            const float* tp = p2.head_t + (g.base0 + m);
#pragma unroll
            for (int t = 0; t < 9; ++t) hin[HT ? j : 0][t] = __ldg(tp + p2.off[t]);
        };
        if (n_my > 0) {
#pragma unroll
            for (int j = 0; j < RB_J2; ++j) { issue_res((int)blockIdx.x, j); issue_head((int)blockIdx.x, j); }
        }
        for (int s = 0; s < n_my; ++s) {
            const int im = (int)blockIdx.x + s * (int)gridDim.x;
            const int b = s & 1, ph = (s >> 1) & 1;
            mbar_wait(d2_full(b), ph);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < RB_J2; ++j) {
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 + b * 128 + j * RB_C + c0);
                uint32_t r0[16];
                tc_ld16_nowait(taddr, r0);
                const int m = im * RB_R + j * MTILE + q * 32 + lane;
                const RowCtx rc = make_row_ctx(p2, m);
                tc_ld_wait16(r0);
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r0[k]);
                if (p2.relu) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
                }
                if (p2.res_hi && m < g.M) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float2 f = hilo_pair(__float_as_uint(add[j][k]), __float_as_uint(add[j][8 + k]));
                        v[2 * k] += f.x; v[2 * k + 1] += f.y;
                    }
                } else if (p2.res32 && m < g.M) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] += add[j][k];
                }
                if constexpr (HT) {
                    if (p2.head_t) {
                        float hcur[9];
#pragma unroll
                        for (int t = 0; t < 9; ++t) hcur[t] = hin[HT ? j : 0][t];
                        if (h == 0) rb_head<0>(htw.head, hcur, v); else rb_head<16>(htw.head, hcur, v);
                    }
                    if (p2.tail_part) {
                        if (rc.valid) {
                            float* dst = p2.tail_part + (size_t)(h * 9) * Ptot + rc.row;
                            if (h == 0) rb_tail<0>(htw.tail, v, dst, Ptot); else rb_tail<16>(htw.tail, v, dst, Ptot);
                        }
                        continue;
                    }
                }
                if (rc.valid) {
                    const EpiAddr a0 = epi_addr(p2, rc, c0);
                    if (p2.skip32) {
                        float sk[16];
                        epi_load16_one(p2.skip32, a0, sk);
#pragma unroll
                        for (int k = 0; k < 16; ++k) v[k] += sk[k];
                    }
                    epi_out16(p2, rc, a0, c0, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_empty(b));
            // All loads of the NEXT item are issued together, after every register of this item has been consumed: a warp
            // has only six load scoreboards, so a load issued between two uses makes the second use wait for it as well.
            if (s + 1 < n_my) {
#pragma unroll
                for (int j = 0; j < RB_J2; ++j) { issue_res(im + (int)gridDim.x, j); issue_head(im + (int)gridDim.x, j); }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

static int g_rb_sms = 0;

int conv_rb_init() {
    int dev;
    GD_CUDA_CHECK(cudaGetDevice(&dev));
    GD_CUDA_CHECK(cudaDeviceGetAttribute(&g_rb_sms, cudaDevAttrMultiProcessorCount, dev));
    GD_CUDA_CHECK(cudaFuncSetAttribute(k_rb_umma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM));
    GD_CUDA_CHECK(cudaFuncSetAttribute(k_rb_umma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM));
    return GD_OK;
}

bool conv_rb_supported(const ConvParams& p1, const ConvParams& p2) {
    return p1.ntaps == 9 && p2.ntaps == 9 && p1.Kt == RB_C && p1.N == RB_C && p2.Kt == RB_C && p2.N == RB_C && p1.mode == 0 && p2.mode == 0 &&
           p1.g.Wp + 1 == RB_HALO && p1.g.base0 >= RB_LEAD + RB_HALO && p1.relu && !p1.res32 && !p1.skip32 && !p1.s2d && !p1.out32 &&
           p1.out16 == p2.a && p1.g.M == p2.g.M && p2.out16 != p1.a;     // in-place fp16 output would race with neighbours' halo reads
}

int launch_conv_rb(const ConvParams& p1, const ConvParams& p2, cudaStream_t st) {
    if (p2.g.M <= 0) return GD_OK;
    if (!conv_rb_supported(p1, p2)) { set_error("conv_rb: unsupported ResBlock shape"); return GD_EUNSUPPORTED; }
    if (!g_rb_sms) { set_error("conv_rb: library not initialised"); return GD_ECUDA; }
    const int items = (p2.g.M + RB_R - 1) / RB_R;
    const int grid = items < g_rb_sms ? items : g_rb_sms;
    cudaEvent_t e1 = nullptr;
    const double flops = 2.0 * 2.0 * (double)(p2.g.M / p2.g.S) * p2.g.H * p2.g.W * (double)RB_C * RB_C * 9;   // both convs, valid pixels
    { int rc = conv_profile_mark(flops, st, &e1); if (rc != GD_OK) return rc; }
    // residual-by-MMA: the hi half of an hi/lo residual that IS this ResBlock's fp16 input is added by an identity "tap"
    static int res_mma_env = -1;
    if (res_mma_env < 0) { const char* e = getenv("GDECONV_RESMMA"); res_mma_env = e ? atoi(e) != 0 : 0; }   // validated (parity, determinism); off: < 1 % (profiles/README)
    const int res_mma = res_mma_env && p2.res_hi && p2.res_hi == p1.a;
    RbHt hw;
    memset(&hw, 0, sizeof(hw));
    if (p2.head_t || p2.tail_part) {
        if (p2.head_t) memcpy(hw.head, p2.head_w, sizeof(hw.head));
        if (p2.tail_part) memcpy(hw.tail, p2.tail_w, sizeof(hw.tail));
        k_rb_umma<1><<<grid, UMMA_THREADS, RB_SMEM, st>>>(p1, p2, items, res_mma, hw);
    } else {
        k_rb_umma<0><<<grid, UMMA_THREADS, RB_SMEM, st>>>(p1, p2, items, res_mma, hw);
    }
    GD_LAUNCHED();
    if (e1) GD_CUDA_CHECK(cudaEventRecord(e1, st));
    return GD_OK;
}

}  // namespace gd
