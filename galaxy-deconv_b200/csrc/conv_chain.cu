// Layer-chained tcgen05 convolution: up to four consecutive 3x3 C->C layers of one U-Net level (two ResBlocks:
// conv-ReLU-conv + residual, twice; models/resnet_basicblock.py:59-71) in ONE persistent launch.
//
// Why: at the wide levels every layer is HBM-bound when run as its own launch over a large chunk -- the fp16 operand
// tensors between the layers (64..128 B/pixel each way) go out to HBM and come back.  Here the layers are software
// pipelined over the 128*J-row items of the chunk: CTA c owns items k = c, c+G, c+2G, ... for ALL layers and, at step s,
// runs layer l on its item j = s - l*Dj.  An item of layer l needs items k-1, k, k+1 of layer l-1 (3x3 halo); those were
// finished Dj steps earlier by the neighbouring CTAs, so they are still in the 126 MB L2 when the bulk copies fetch
// them -- reads of intermediate tensors never touch HBM, and there is one launch (one ramp-up, one tail) instead of four.
//
// Cross-CTA dependencies are published per (layer, item) in global memory: every epilogue warp, after its stores,
// does  __syncwarp; __threadfence; atomicAdd(flag, 1)  and the producer thread of a consumer acquires
// flag == #epilogue warps (ld.acquire.gpu) for items k-1, k, k+1, then issues fence.proxy.async before the bulk
// copies.  The same dependencies also cover the write-after-read hazards of the ping-pong buffers (a layer may only
// overwrite rows whose readers -- the three neighbouring items of the previous layer -- are complete).  The fp32
// residual of an item was written by the SAME thread of the SAME CTA two layers earlier (same item, same row/column
// mapping), so it is ordered by program order and read with plain (coherent) loads.  Waits are bounded and trap.
//
// Everything else (resident weights, A ring with tap-shifted descriptors, four MMA-issuing warps, double-buffered TMEM
// accumulators, lean epilogues) is the machinery of conv_umma.cu.
#include <cstdlib>
#include <cstring>

#include "conv_epilogue.cuh"
#include "kernels.cuh"
#include "launch.cuh"
#include "umma_ptx.cuh"

namespace gd {

constexpr int CHAIN_MAX_LAYERS = 4;
constexpr int CHAIN_BARS = 2 * MAX_A_STAGES + 1 + 4;

struct ChainParams {
    ConvParams L[CHAIN_MAX_LAYERS];   // per layer: a, w, relu, res32, skip32, out32, out16, s2d, gc (shared g, off, Kt = N = C)
    int plain[CHAIN_MAX_LAYERS];      // 1: epilogue = ReLU? + fp16 copy only
    int nl, Dj, J, C, BK, halo, win_rows, a_stage_bytes, a_stages, layer_bytes, items;
    unsigned int* flags;              // [nl][items], zeroed before the launch
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned int* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_flag(const unsigned int* p, uint32_t target) {
    if (ld_acquire_gpu(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu(p) < target)
        if (clock64() - t0 > 4000000000ll) __trap();
}

template <int KK>
__global__ void __launch_bounds__(UMMA_THREADS, 1) k_conv_chain(const __grid_constant__ ChainParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[CHAIN_BARS];
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    unsigned char* a_smem = smem;
    unsigned char* b_smem = smem + (size_t)P.a_stages * P.a_stage_bytes;
    const uint32_t bar0 = smem_u32(bars);
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (MAX_A_STAGES + s); };
    const uint32_t w_full = bar0 + 8u * (2 * MAX_A_STAGES);
    auto acc_full = [&](int s) { return w_full + 8u * (1 + s); };
    auto acc_empty = [&](int s) { return w_full + 8u * (3 + s); };
    const int J = P.J, C = P.C;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_A_STAGES; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), J); }
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), J); mbar_init(acc_empty(s), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const Geom& g = P.L[0].g;
    const int nslabs = C / P.BK, chunks = P.BK / 8, KC = C / 8;
    const int G = (int)gridDim.x, c0 = (int)blockIdx.x;
    const int nj = P.items > c0 ? (P.items - c0 + G - 1) / G : 0;          // items owned by this CTA (per layer)
    const int nsteps = nj > 0 ? nj + (P.nl - 1) * P.Dj : 0;
    // the CTA's work sequence: for s: for l: j = s - l*Dj (if 0 <= j < nj) -> item k = c0 + j*G of layer l

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            int as = 0, aph = 0;
            mbar_expect_tx(w_full, (uint32_t)(P.nl * P.layer_bytes));
            for (int l = 0; l < P.nl; ++l) {
                const unsigned char* wts = reinterpret_cast<const unsigned char*>(P.L[l].w);
                const uint32_t bdst = smem_u32(b_smem) + (uint32_t)(l * P.layer_bytes);
                for (int off = 0; off < P.layer_bytes; off += 16384) {
                    const int n = P.layer_bytes - off < 16384 ? P.layer_bytes - off : 16384;
                    bulk_g2s(bdst + off, wts + off, (uint32_t)n, w_full);
                }
            }
            for (int s = 0; s < nsteps; ++s)
                for (int l = 0; l < P.nl; ++l) {
                    const int j = s - l * P.Dj;
                    if (j < 0 || j >= nj) continue;
                    const int k = c0 + j * G;
                    if (l > 0) {
                        const unsigned int* f = P.flags + (size_t)(l - 1) * P.items;
                        if (k > 0) wait_flag(f + k - 1, EPI_WARPS);
                        wait_flag(f + k, EPI_WARPS);
                        if (k + 1 < P.items) wait_flag(f + k + 1, EPI_WARPS);
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    const unsigned char* act = reinterpret_cast<const unsigned char*>(P.L[l].a);
                    const size_t row0 = (size_t)g.base0 + (size_t)k * J * MTILE - P.halo;
                    for (int sl = 0; sl < nslabs; ++sl) {
                        mbar_wait(a_empty(as), aph ^ 1);
                        mbar_expect_tx(a_full(as), (uint32_t)P.a_stage_bytes);
                        const uint32_t adst = smem_u32(a_smem + (size_t)as * P.a_stage_bytes);
                        for (int ch = 0; ch < chunks; ++ch)
                            bulk_g2s(adst + (uint32_t)ch * P.win_rows * 16,
                                     act + ((size_t)(sl * chunks + ch) * g.Ptot + row0) * 16, (uint32_t)P.win_rows * 16, a_full(as));
                        if (++as == P.a_stages) { as = 0; aph ^= 1; }
                    }
                }
        }
    } else if (warp <= MMA_WARPS) {
        // ===== MMA issuers: warp 1+jw owns tile jw of every item =====
        const int jw = warp - 1;
        if (jw < J) {
            int as = 0, aph = 0, acs = 0, accph = 0;
            const uint32_t idesc = instr_desc_f16(MTILE, C);
            const uint32_t a_lbo = (uint32_t)P.win_rows * 16, b_lbo = (uint32_t)C * 16;
            const uint64_t a_desc0 = smem_desc(smem_u32(a_smem), a_lbo, 128);
            const uint64_t b_desc0 = smem_desc(smem_u32(b_smem), b_lbo, 128);
            const uint32_t a_stage_u = (uint32_t)P.a_stage_bytes >> 4, layer_u = (uint32_t)P.layer_bytes >> 4;
            const uint32_t a_kk = (2 * a_lbo) >> 4, b_kk = (2 * b_lbo) >> 4;
            mbar_wait(w_full, 0);
            for (int s = 0; s < nsteps; ++s)
                for (int l = 0; l < P.nl; ++l) {
                    const int j = s - l * P.Dj;
                    if (j < 0 || j >= nj) continue;
                    mbar_wait(acc_empty(acs), accph ^ 1);
                    tc_fence_after();
                    const uint32_t dcol = tmem + (uint32_t)(acs * ACC_STAGE_COLS + jw * C);
                    for (int sl = 0; sl < nslabs; ++sl) {
                        mbar_wait(a_full(as), aph);
                        tc_fence_after();
                        const uint64_t ad_s = a_desc0 + (uint64_t)((uint32_t)as * a_stage_u + (uint32_t)(P.halo + jw * MTILE));
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint64_t bd = b_desc0 + (uint64_t)((uint32_t)l * layer_u + (uint32_t)((tap * KC + sl * chunks) * C));
                            const uint64_t ad_t = ad_s + (uint64_t)(int64_t)P.L[0].off[tap];
                            tc_mma_tap<1, KK>(dcol, (uint32_t)C, ad_t, bd, a_kk, b_kk, idesc, (uint32_t)((sl | tap) != 0));
                        }
                        tc_commit_pred(a_empty(as), 0);
                        if (++as == P.a_stages) { as = 0; aph ^= 1; }
                    }
                    tc_commit_pred(acc_full(acs), 0);
                    if (++acs == 2) { acs = 0; accph ^= 1; }
                }
        }
    } else {
        // ===== epilogue: 8 warps, nu = J*C/32 = 4 units per item -> 2 units per warp (i = 0, 1) =====
        const int e = warp - EPI_WARP0, q = warp & 3, half = e >> 2;
        const int nb32 = C / 32, nb32_log2 = nb32 == 1 ? 0 : 1;
        const uint32_t Ptot = (uint32_t)g.Ptot;
        float add[2][32];
        int acs = 0, accph = 0;
        // residual of (layer l, item k, unit i) -> add[i]; plain coherent loads (written earlier by this very thread)
        auto issue_res = [&](int l, int k, int i) {
            const ConvParams& p = P.L[l];
            if (P.plain[l] || !p.res32) return;
            const int uu = half + 2 * i, jt = uu >> nb32_log2, b = uu & (nb32 - 1);
            const int m = (k * J + jt) * MTILE + q * 32 + lane;
            if (m >= g.M) return;
            const float4* src = reinterpret_cast<const float4*>(p.res32) + (size_t)((b * 32) >> 2) * Ptot + (g.base0 + m);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float4 v = src[(size_t)t * Ptot];
                add[i][4 * t] = v.x; add[i][4 * t + 1] = v.y; add[i][4 * t + 2] = v.z; add[i][4 * t + 3] = v.w;
            }
        };
        // first work of this CTA is always (layer 0, item c0) when nj > 0; layer 0 may itself carry a residual
        if (nj > 0) { issue_res(0, c0, 0); issue_res(0, c0, 1); }
        for (int s = 0; s < nsteps; ++s)
            for (int l = 0; l < P.nl; ++l) {
                const int j = s - l * P.Dj;
                if (j < 0 || j >= nj) continue;
                const int k = c0 + j * G;
                // next work of the sequence (for the residual prefetch)
                int nl_ = l, ns_ = s, nk = -1;
                for (int it = 0; it < P.nl * (P.Dj + 2) && nk < 0; ++it) {
                    if (++nl_ == P.nl) { nl_ = 0; ++ns_; }
                    if (ns_ >= nsteps) break;
                    const int jj = ns_ - nl_ * P.Dj;
                    if (jj >= 0 && jj < nj) nk = c0 + jj * G;
                }
                const ConvParams& p = P.L[l];
                const bool plain = P.plain[l] != 0;
                mbar_wait(acc_full(acs), accph);
                tc_fence_after();
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int uu = half + 2 * i, jt = uu >> nb32_log2, b = uu & (nb32 - 1);
                    const int n0 = b * 32;
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acs * ACC_STAGE_COLS + jt * C + b * 32);
                    uint32_t r0[16], r1[16];
                    tc_ld16_nowait(taddr, r0);
                    tc_ld16_nowait(taddr + 16, r1);
                    const int m = (k * J + jt) * MTILE + q * 32 + lane;
                    const RowCtx rc = make_row_ctx(p, m);
                    float v[32];
                    tc_ld_wait16(r0);
                    tc_ld_wait16(r1);
#pragma unroll
                    for (int t = 0; t < 16; ++t) { v[t] = __uint_as_float(r0[t]); v[16 + t] = __uint_as_float(r1[t]); }
                    if (p.relu) {
#pragma unroll
                        for (int t = 0; t < 32; ++t) v[t] = fmaxf(v[t], 0.f);
                    }
                    if (plain) {
                        if (rc.valid) {
                            uint4* dst = reinterpret_cast<uint4*>(p.out16) + (size_t)(n0 >> 3) * Ptot + rc.row;
#pragma unroll
                            for (int t = 0; t < 4; ++t) dst[(size_t)t * Ptot] = pack8_half(v + 8 * t);
                        }
                        if (nk >= 0) issue_res(nl_, nk, i);
                    } else {
                        if (p.res32 && m < g.M) {
#pragma unroll
                            for (int t = 0; t < 32; ++t) v[t] += add[i][t];
                        }
                        if (nk >= 0) issue_res(nl_, nk, i);
                        if (rc.valid) {
                            const EpiAddr a0 = epi_addr(p, rc, n0), a1 = epi_addr(p, rc, n0 + 16);
                            if (p.skip32) {
                                float sk[32];
                                epi_load16_one(p.skip32, a0, sk); epi_load16_one(p.skip32, a1, sk + 16);
#pragma unroll
                                for (int t = 0; t < 32; ++t) v[t] += sk[t];
                            }
                            epi_out16(p, rc, a0, n0, v);
                            epi_out16(p, rc, a1, n0 + 16, v + 16);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(acc_empty(acs));
                    __threadfence();                                        // publish this warp's share of (l, k)
                    atomicAdd(P.flags + (size_t)l * P.items + k, 1u);
                }
                if (++acs == 2) { acs = 0; accph ^= 1; }
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

static int g_chain_sms = 0;
static int g_chain_dj = 3;

int conv_chain_init() {
    int dev;
    GD_CUDA_CHECK(cudaGetDevice(&dev));
    GD_CUDA_CHECK(cudaDeviceGetAttribute(&g_chain_sms, cudaDevAttrMultiProcessorCount, dev));
    GD_CUDA_CHECK(cudaFuncSetAttribute(k_conv_chain<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UMMA_SMEM_MAX));
    if (const char* e = getenv("GDECONV_CHAIN_DJ")) { g_chain_dj = atoi(e); if (g_chain_dj < 1) g_chain_dj = 1; }
    return GD_OK;
}

// flags: device buffer of at least chain_flag_words(...) 32-bit words
size_t chain_flag_words(int max_rows) { return (size_t)CHAIN_MAX_LAYERS * ((max_rows + MTILE - 1) / MTILE + 8); }

// layers[i]: 3x3 C->C layers on the same geometry, consecutive in the network (each one's `a` is produced by the previous
// one's out16 / by the layer before that).  Returns GD_EUNSUPPORTED when the shape does not fit (caller falls back to
// one launch per layer).
int launch_conv_chain(const ConvParams* layers, int nl, unsigned int* flags, cudaStream_t st) {
    if (nl < 2 || nl > CHAIN_MAX_LAYERS) { set_error("conv_chain: %d layers", nl); return GD_EUNSUPPORTED; }
    const ConvParams& p0 = layers[0];
    if (p0.g.M <= 0) return GD_OK;
    ChainParams P;
    memset(&P, 0, sizeof(P));
    const int C = p0.N;
    for (int i = 0; i < nl; ++i) {
        const ConvParams& p = layers[i];
        if (p.ntaps != 9 || p.Kt != C || p.N != C || p.mode != 0 || p.g.H != p0.g.H || p.g.M != p0.g.M) {
            set_error("conv_chain: layer %d is not a 3x3 C->C layer of the segment", i);
            return GD_EUNSUPPORTED;
        }
        P.L[i] = p;
        P.plain[i] = !p.res32 && !p.skip32 && !p.out32 && !p.s2d && p.out16;
    }
    if (C != 32 && C != 64) { set_error("conv_chain: C=%d (supported: 32, 64)", C); return GD_EUNSUPPORTED; }
    P.nl = nl; P.Dj = g_chain_dj; P.C = C; P.BK = 32;
    P.J = 128 / C;                                    // 4 units of 32 columns per item -> 2 per epilogue warp
    P.halo = p0.g.Wp + 1;
    P.win_rows = MTILE * P.J + 2 * P.halo;
    P.a_stage_bytes = P.win_rows * P.BK * 2;
    P.layer_bytes = 9 * C * C * 2;
    const size_t b_region = (size_t)nl * P.layer_bytes;
    if (b_region + 2 * (size_t)P.a_stage_bytes > UMMA_SMEM_MAX) { set_error("conv_chain: weights of %d layers do not fit", nl); return GD_EUNSUPPORTED; }
    P.a_stages = (int)((UMMA_SMEM_MAX - b_region) / P.a_stage_bytes);
    if (P.a_stages > MAX_A_STAGES) P.a_stages = MAX_A_STAGES;
    const int tiles = (p0.g.M + MTILE - 1) / MTILE;
    P.items = (tiles + P.J - 1) / P.J;
    P.flags = flags;
    const size_t smem = (size_t)P.a_stages * P.a_stage_bytes + b_region;
    GD_CUDA_CHECK(cudaMemsetAsync(flags, 0, (size_t)nl * P.items * sizeof(unsigned int), st));
    const int grid = P.items < g_chain_sms ? P.items : g_chain_sms;
    k_conv_chain<2><<<grid, UMMA_THREADS, smem, st>>>(P);
    GD_LAUNCHED();
    return GD_OK;
}

}  // namespace gd
