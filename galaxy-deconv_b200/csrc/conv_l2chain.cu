// Level-1 chain kernel of the ResUNet denoiser (64 channels at 24x24, path G: models/ResUNet.py:33 and :37), tcgen05 path:
// DOWN = the two ResBlocks of m_down2 + its k2s2 strided conv, UP = the k2s2 transposed conv of m_up2 + its two ResBlocks + the
// U-Net skip (four 3x3 convolutions, models/resnet_basicblock.py:69-71, and one 1-tap GEMM) in ONE launch per stage, with every
// intermediate on chip.  One launch per conv moves the fp16 hi/lo stream through HBM around every ResBlock and runs the
// residual-carrying second convs at the HBM roofline; here a work item is ONE WHOLE STAMP (625 padded-linear rows = 5 tiles of 128,
// no halo to recompute), so HBM sees the stage's input and output once.
//
//   * the fp16 operand copies of the stream (X) and of ReLU(conv1) (T) live in shared memory: 2 x 8 chunk planes x 672 rows x 16 B;
//     rows 625..671 of a plane are never written (zero row below the stamp + gap absorbing the tap shifts);
//   * the residual stream is hi + lo: hi = the operand copy in X, lo = rn16(x - hi) packed two per column in TENSOR MEMORY
//     (5 tiles x 32 columns) -- |x - (hi + lo)| <= 2^-22 |x| as in the HBM stream of the other levels; the helper warps bring the lo
//     halves of an item straight from global memory (DOWN) or from this CTA's scratch (UP) into tensor memory;
//   * a conv's weights (74 KB) do not fit beside X and T, so the MMA loop is TAP-OUTER: the five tiles' accumulators (5 x 64 TMEM
//     columns) are live together and each of the nine 8 KB tap slices streams ONCE per conv and stamp through a 7-stage ring;
//   * to let a conv's epilogue overlap MMAs although every tile completes with the last tap, the tiles form two phases on the same
//     ring: phase 0 (tiles 0, 1) runs L2_LAG taps ahead of phase 1 (tiles 2, 3, 4).  Phase 0's epilogue drains under the tail of
//     phase 1, and the next conv's phase 0 starts as soon as the tiles it reads (0..2) are written back, while tiles 3 and 4 still
//     drain.  Streaming the ring once per phase instead (no lag) starves the MMAs: 48 KB in flight cover < 1 us of phase-0 issue.
//     Measured on one box (G(8), 10,000 stamps): no lag 57.3 k gal/s, lag 3 57.9 k, lag 4 58.1 k; without this kernel 55.4 k.
//   * the N = 64 MMA reads 4 KB of A and 2 KB of B from shared memory: 48 cycles at 128 B/cycle against 32 cycles of math, so the
//     tensor pipe cannot exceed 67 % here; the kernel runs at ~58 cycles per MMA while issuing;
//   * DOWN ends with the strided conv (64 -> 128 channels at 12x12): the last epilogue writes its space-to-depth operand
//     [32][169][8] over the T planes (which it fills exactly) instead of HBM, two M128 N128 K256 tiles reuse the accumulator columns,
//     a fifth epilogue stores x3 (fp32 skip + fp16 operand copy); the helper warps restore T's zero rows for the next item.
//     1.00 + 0.29 ms (chain + separate launch) -> 1.08 ms per 5000 stamps, 58.3 k -> 59.6 k gal/s;
//   * UP starts with the transposed conv (128 -> 4 x 64 channels): the coarse map [16][169][8] is bulk-copied into the T planes,
//     two M128 N256 K128 tiles fill ALL 512 tensor-memory columns (the lo stream does not exist yet), the epilogue scatters the hi
//     halves of the four sub-pixels straight into X and passes the lo halves through a per-CTA, L2-resident scratch (a thread can
//     only write its own TMEM lane).  1.02 + 0.29 ms -> 1.27 ms: the drain per item (A load, 16 MMAs, scatter, scratch -> TMEM)
//     costs nearly what the launch did, but 0.74 GB of fine-level hi/lo traffic per call is gone; 59.6 k -> 60.2 k gal/s.
//
// Warp roles (512 threads, one persistent CTA per SM): warp 0 producers (lane 0: tap slices and the k2s2 weights; lane 1: DOWN: hi ->
// X in two row ranges, L2 prefetch of the next item; UP: coarse map -> T), warps 1-3 MMA issue (one thread sustains only ~1
// tcgen05.mma per 57 cycles), warps 4-15 epilogue in three groups of four quadrant warps (unit = tile x 32-channel half, round
// robin); warps 4-7 also move the lo halves of the item's input stream to tensor memory.
#include "conv_epilogue.cuh"
#include "kernels.cuh"
#include "launch.cuh"
#include "umma_ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace gd {

constexpr int L2_C = 64;
constexpr int L2_WP = 25;
constexpr int L2_S = 625;                            // padded-linear rows of one stamp at 24x24
constexpr int L2_PIX = 600;                          // rows of real image rows (y < 24)
constexpr int L2_TILES = 5;
constexpr int L2_GAP = 32;                           // >= Wp + 1
constexpr int L2_PSTRIDE = L2_TILES * MTILE + L2_GAP;            // 672
constexpr int L2_ACT_BYTES = (16 * L2_PSTRIDE + L2_GAP) * 16;    // 172,544: planes 0-7 = X, 8-15 = T
constexpr int L2_WSTAGE = L2_C * L2_C * 2;           // 8,192: one tap of one conv, [8][64][8]
constexpr int L2_WSTAGES = 7;
constexpr int L2_SMEM = L2_ACT_BYTES + L2_WSTAGES * L2_WSTAGE;   // 229,888
constexpr int L2_THREADS = (1 + 3 + 4 + 8) * 32;
constexpr int L2_LO_COL = L2_TILES * L2_C;           // 320: packed lo stream, 32 columns per tile

// tile phases of one conv: the MMAs of phase 1 run while the epilogue of phase 0 drains, and the next conv's phase 0 starts as soon
// as the tiles it reads (0..2) have been written back
constexpr int L2_PH0_TILES = 2;                      // phase 0 = tiles 0, 1; phase 1 = tiles 2, 3, 4
constexpr int L2_EPI_GROUPS = 3;                     // warps 4-15: three groups of four quadrant warps
constexpr int L2_LAG = 4;                            // phase 1 runs this many taps behind phase 0 on the same ring of tap slices
constexpr int L2_X0_ROWS = (L2_PH0_TILES + 1) * MTILE;          // 384: rows of tiles 0..2, everything phase 0 reads

// m_down2's k2s2 strided conv (64 -> 128 channels, 12x12) on the space-to-depth operand the last epilogue leaves in the T planes
constexpr int L2_DS = 169;                           // rows of one coarse stamp = rows per K chunk of the operand [32][169][8]
constexpr int L2_DN = 128;                           // its output channels
constexpr int L2_DSTAGES = 32 * L2_DN * 16 / L2_WSTAGE;         // 8 ring stages of two K16 steps each
static_assert(32 * L2_DS * 16 == (8 * L2_PSTRIDE + L2_GAP) * 16, "the operand fills the T planes and the trailing gap exactly");

// m_up2's k2s2 transposed conv (128 -> 4 x 64 channels) at the head of an UP item: A = the coarse fp16 map [16][169][8] in the T planes,
// N = 256 (column = sub-pixel * 64 + channel), two M128 tiles over all 512 TMEM columns
constexpr int L2_UN = 256;
constexpr int L2_USTAGES = 16 * L2_UN * 16 / L2_WSTAGE;         // 8 ring stages, one K16 step each
constexpr int L2_SCRATCH_BYTES = 8 * L2_S * 16;                 // per CTA: lo halves of the transposed conv's output, [8][625][8] fp16

enum { L2B_X_FULL = 0, L2B_DOWN_FULL = 2, L2B_DOWN_EMPTY, L2B_MMA_DONE, L2B_LO_DONE, L2B_ACC_FULL, L2B_TILE_DONE = L2B_ACC_FULL + 2, L2B_W_FULL = L2B_TILE_DONE + L2_TILES,
       L2B_W_EMPTY = L2B_W_FULL + L2_WSTAGES, L2B_COUNT = L2B_W_EMPTY + L2_WSTAGES };

struct L2ChainParams {
    int nb, mode;                  // stamps; 0 = m_down2's ResBlocks (-> space-to-depth copy), 1 = m_up2's (+ U-Net skip -> fp16 map)
    Geom g1, g2;                   // 24x24 and 12x12 geometry of the chunk
    const void *x_hi, *x_lo;       // fp16 hi / lo planes of the stage's input stream [8][g1.Ptot][8]
    const void* w[4];              // packed 3x3 weights [tap][8][64][8]
    const void *a_coarse, *wup;    // mode 1: fp16 input of m_up2's transposed conv [16][g2.Ptot][8], its weights [16][256][8] (plain column order)
    void* scratch;                 // mode 1: L2_SCRATCH_BYTES per CTA
    const void* wdown;             // mode 0: strided-conv weights [32][128][8]
    float* skip3;                  // mode 0: x3 fp32 [32][g2.Ptot][4] (U-Net skip + residual of level 2)
    void* x3_16;                   // mode 0: x3 fp16 [16][g2.Ptot][8] (operand of level 2's first conv)
    const float* skip32;           // mode 1: U-Net skip x2, fp32 [16][g1.Ptot][4]
    void* out16;                   // mode 1: fp16 (x + x2) [8][g1.Ptot][8] (input of m_up1's transposed conv)
};

template <int MODE>
__global__ void __launch_bounds__(L2_THREADS, 1) k_l2_chain(const L2ChainParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[L2B_COUNT];
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    unsigned char* w_smem = smem + L2_ACT_BYTES;
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    if (threadIdx.x == 0) {
        mbar_init(bar(L2B_X_FULL), 1); mbar_init(bar(L2B_X_FULL + 1), 1); mbar_init(bar(L2B_DOWN_FULL), 3); mbar_init(bar(L2B_DOWN_EMPTY), 12); mbar_init(bar(L2B_MMA_DONE), 3); mbar_init(bar(L2B_LO_DONE), 4);
        mbar_init(bar(L2B_ACC_FULL), 3); mbar_init(bar(L2B_ACC_FULL + 1), 3);
        for (int t = 0; t < L2_TILES; ++t) mbar_init(bar(L2B_TILE_DONE + t), 8);          // 2 channel halves x 4 quadrant warps
        for (int s = 0; s < L2_WSTAGES; ++s) { mbar_init(bar(L2B_W_FULL + s), 1); mbar_init(bar(L2B_W_EMPTY + s), 3); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // zero the activation planes once: gap rows, the zero row below the stamp and the pad pixels are never written afterwards
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < L2_ACT_BYTES / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int n_my = p.nb > (int)blockIdx.x ? (p.nb - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t Ptot1 = (uint32_t)p.g1.Ptot;
    auto row_off = [](int pl, int s) { return (uint32_t)((pl * L2_PSTRIDE + L2_GAP + s) * 16); };
    // every tile's barrier completes once per conv: completion n = 4 * item + conv has parity n & 1
    auto wait_tiles = [&](int t0, int t1, int n) {
        for (int t = t0; t <= t1; ++t) mbar_wait(bar(L2B_TILE_DONE + t), (uint32_t)(n & 1));
    };

    if (warp == 0) {
        // ===== producers: lane 0 streams the tap slices, lane 1 the activation planes (independent waits) =====
        if (lane == 0) {
            uint32_t wc = 0;                              // running weight-stage counter
            auto stage = [&](const void* src) {
                const uint32_t s = wc % L2_WSTAGES;
                mbar_wait(bar(L2B_W_EMPTY + s), ((wc / L2_WSTAGES) & 1) ^ 1);
                mbar_expect_tx(bar(L2B_W_FULL + s), (uint32_t)L2_WSTAGE);
                bulk_g2s(smem_u32(w_smem) + s * L2_WSTAGE, src, (uint32_t)L2_WSTAGE, bar(L2B_W_FULL + s));
                ++wc;
            };
            for (int k = 0; k < n_my; ++k) {
                if (MODE == 1)
                    for (int i = 0; i < L2_USTAGES; ++i) stage(reinterpret_cast<const unsigned char*>(p.wup) + (size_t)i * L2_WSTAGE);
                for (int c = 0; c < 4; ++c)
                    for (int tap = 0; tap < 9; ++tap) stage(reinterpret_cast<const unsigned char*>(p.w[c]) + (size_t)tap * L2_WSTAGE);
                if (MODE == 0)
                    for (int i = 0; i < L2_DSTAGES; ++i) stage(reinterpret_cast<const unsigned char*>(p.wdown) + (size_t)i * L2_WSTAGE);
            }
        } else if (lane == 1) {
            auto plane_src = [&](const void* src, int ch, int b) {
                return reinterpret_cast<const unsigned char*>(src) + ((size_t)ch * Ptot1 + (size_t)p.g1.base0 + (size_t)b * p.g1.S) * 16;
            };
            auto load_planes = [&](int b, const void* src, int pl0, int r0, int r1, uint32_t full) {
                mbar_expect_tx(full, 8u * (uint32_t)(r1 - r0) * 16u);
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    bulk_g2s(smem_u32(smem) + row_off(pl0 + ch, r0), plane_src(src, ch, b) + (size_t)r0 * 16, (uint32_t)(r1 - r0) * 16u, full);
            };
            for (int k = 0; k < n_my; ++k) {
                const int b = (int)blockIdx.x + k * (int)gridDim.x;
                // once per item: the previous item's last conv has been issued and completed, so every tile barrier has passed its
                // completion 4k - 2 and the parity waits below cannot alias an older phase (this lane is not tied to the weight ring)
                mbar_wait(bar(L2B_MMA_DONE), (uint32_t)((k & 1) ^ 1));
                if (MODE == 1) {
                    // the coarse map of the transposed conv -> T (conv 3 of the previous item has finished reading it: MMA_DONE)
                    auto coarse_src = [&](int ch, int bb) {
                        return reinterpret_cast<const unsigned char*>(p.a_coarse) + ((size_t)ch * p.g2.Ptot + (size_t)p.g2.base0 + (size_t)bb * L2_DS) * 16;
                    };
                    mbar_expect_tx(bar(L2B_X_FULL), 16u * L2_DS * 16u);
                    for (int ch = 0; ch < 16; ++ch)
                        bulk_g2s(smem_u32(smem) + (uint32_t)((8 * L2_PSTRIDE + ch * L2_DS) * 16), coarse_src(ch, b), (uint32_t)L2_DS * 16u, bar(L2B_X_FULL));
                    if (k + 1 < n_my)
                        for (int ch = 0; ch < 16; ++ch) bulk_prefetch_l2(coarse_src(ch, b + (int)gridDim.x), L2_DS * 16u);
                    for (int ch = 0; ch < 16; ++ch) bulk_prefetch_l2(plane_src(p.skip32, ch, b), L2_S * 16u);
                    continue;
                }
                // X in two parts, each as soon as the previous item's last epilogues (n = 4k - 1) have read their residual hi from it
                wait_tiles(0, L2_PH0_TILES, 4 * k - 1);
                load_planes(b, p.x_hi, 0, 0, L2_X0_ROWS, bar(L2B_X_FULL));
                wait_tiles(L2_PH0_TILES + 1, L2_TILES - 1, 4 * k - 1);
                load_planes(b, p.x_hi, 0, L2_X0_ROWS, L2_S, bar(L2B_X_FULL + 1));
                if (k + 1 < n_my) {                       // the next item's planes on their way into L2
                    const int bn = b + (int)gridDim.x;
                    for (int ch = 0; ch < 8; ++ch) { bulk_prefetch_l2(plane_src(p.x_hi, ch, bn), L2_S * 16u); bulk_prefetch_l2(plane_src(p.x_lo, ch, bn), L2_S * 16u); }
                }
            }
        }
    } else if (warp <= 3) {
        // ===== MMA issuers.  Tap-outer within a phase: phase 0 = tiles 0, 1 (warps 1, 2), phase 1 = tiles 2, 3, 4 (warps 3, 1, 2) =====
        const int jw = warp - 1;
        const uint32_t idesc = instr_desc_f16(MTILE, L2_C);
        const uint64_t a_desc0 = smem_desc(smem_u32(smem), L2_PSTRIDE * 16, 128);
        const uint64_t w_desc0 = smem_desc(smem_u32(w_smem), L2_C * 16, 128);
        constexpr uint32_t A_KK = 2 * L2_PSTRIDE, W_KK = 2 * L2_C;           // in 16-byte units
        uint32_t wc0 = 0;                                 // ring position of the conv's tap 0
        const int t0 = jw, t1 = L2_PH0_TILES + (jw + 1) % 3;                 // this warp's tile of phase 0 (jw < 2 only) and of phase 1
        for (int k = 0; k < n_my; ++k) {
            if (MODE == 1) {
                // ---- m_up2's transposed conv: two M128 N256 tiles (coarse rows 0..255), K = 128 in 8 steps, D over all 512 columns ----
                mbar_wait(bar(L2B_X_FULL), (uint32_t)(k & 1));               // the coarse map is in T
                wait_tiles(0, L2_TILES - 1, 4 * k - 1);                      // accumulators, lo stream and X of the previous item are dead
                tc_fence_after();
                const uint32_t idesc_u = instr_desc_f16(MTILE, L2_UN);
                const uint64_t au0 = smem_desc(smem_u32(smem) + 8 * L2_PSTRIDE * 16, L2_DS * 16, 128) + (uint64_t)(uint32_t)(jw * MTILE);
                const uint64_t wu0 = smem_desc(smem_u32(w_smem), L2_UN * 16, 128);
                for (int i = 0; i < L2_USTAGES; ++i) {
                    const uint32_t w = wc0 + (uint32_t)i, st = w % L2_WSTAGES;
                    mbar_wait(bar(L2B_W_FULL + st), (w / L2_WSTAGES) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        if (jw < 2) {
                            const uint64_t ad = au0 + (uint64_t)(uint32_t)(i * 2 * L2_DS), wd = wu0 + (uint64_t)(st * (L2_WSTAGE >> 4));
                            if (i == 0) tc_mma_f16(tmem + (uint32_t)(jw * L2_UN), ad, wd, idesc_u, 0u);
                            else tc_mma_f16_acc(tmem + (uint32_t)(jw * L2_UN), ad, wd, idesc_u);
                        }
                        tc_commit(bar(L2B_W_EMPTY + st));
                    }
                    __syncwarp();
                }
                if (elect_one()) tc_commit(bar(L2B_DOWN_FULL));              // (mode 1: "transposed conv issued and complete")
                __syncwarp();
                wc0 += L2_USTAGES;
            }
            for (int c = 0; c < 4; ++c, wc0 += 9) {
                const int n = 4 * k + c;
                const int src_pl = (c & 1) ? 8 : 0;
                auto issue = [&](int t, int tap, uint32_t s) {
                    const int off = (tap / 3 - 1) * L2_WP + (tap % 3 - 1);
                    const uint64_t wd = w_desc0 + (uint64_t)(s * (L2_WSTAGE >> 4));
                    const uint32_t d = tmem + (uint32_t)(t * L2_C);
                    const uint64_t ad = a_desc0 + (uint64_t)(uint32_t)(src_pl * L2_PSTRIDE + L2_GAP + t * MTILE + off);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (tap == 0 && kk == 0) tc_mma_f16(d, ad, wd, idesc, 0u);
                        else tc_mma_f16_acc(d, ad + (uint64_t)(kk * A_KK), wd + (uint64_t)(kk * W_KK), idesc);
                    }
                };
                // input rows written back (tile t reads the rows of tiles t-1 .. t+1) and accumulators drained: tiles 0..2 for phase 0
                wait_tiles(0, L2_PH0_TILES, n - 1);
                if (c == 0) {
                    if (MODE == 0) {
                        mbar_wait(bar(L2B_X_FULL), (uint32_t)(k & 1));
                        mbar_wait(bar(L2B_DOWN_EMPTY), (uint32_t)((k & 1) ^ 1));    // the previous item's strided-conv accumulators are drained
                    } else {
                        mbar_wait(bar(L2B_DOWN_EMPTY), (uint32_t)(k & 1));          // this item's transposed conv: hi scattered into X, D drained
                    }
                }
                tc_fence_after();
                constexpr int LAG = L2_LAG;
                for (int i = 0; i < 9 + LAG; ++i) {
                    if (i == LAG) {                    // ... tiles 3, 4 for phase 1
                        wait_tiles(L2_PH0_TILES + 1, L2_TILES - 1, n - 1);
                        if (MODE == 0 && c == 0) mbar_wait(bar(L2B_X_FULL + 1), (uint32_t)(k & 1));
                        tc_fence_after();
                    }
                    if (i < 9) {
                        const uint32_t w = wc0 + (uint32_t)i;
                        mbar_wait(bar(L2B_W_FULL + w % L2_WSTAGES), (w / L2_WSTAGES) & 1);
                        tc_fence_after();
                    }
                    if (elect_one()) {
                        if (i < 9 && jw < L2_PH0_TILES) issue(t0, i, (wc0 + (uint32_t)i) % L2_WSTAGES);
                        if (i == 8) tc_commit(bar(L2B_ACC_FULL));
                        if (i >= LAG) {
                            const uint32_t s = (wc0 + (uint32_t)(i - LAG)) % L2_WSTAGES;
                            issue(t1, i - LAG, s);
                            tc_commit(bar(L2B_W_EMPTY + s));
                        }
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    tc_commit(bar(L2B_ACC_FULL + 1));
                    if (c == 3) tc_commit(bar(L2B_MMA_DONE));
                }
                __syncwarp();
            }
            if (MODE == 0) {
                // ---- m_down2's strided conv: two M128 N128 tiles (coarse rows 0..255, 156 of them real), K = 256 in 16 steps ----
                wait_tiles(0, L2_TILES - 1, 4 * k + 3);           // the operand is complete and the conv accumulators are drained
                tc_fence_after();
                const uint32_t idesc_d = instr_desc_f16(MTILE, L2_DN);
                const uint64_t ad0 = smem_desc(smem_u32(smem) + 8 * L2_PSTRIDE * 16, L2_DS * 16, 128) + (uint64_t)(uint32_t)(jw * MTILE);
                const uint64_t wd0 = smem_desc(smem_u32(w_smem), L2_DN * 16, 128);
                for (int i = 0; i < L2_DSTAGES; ++i) {
                    const uint32_t w = wc0 + (uint32_t)i, st = w % L2_WSTAGES;
                    mbar_wait(bar(L2B_W_FULL + st), (w / L2_WSTAGES) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        if (jw < 2) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int kk = 2 * i + h;
                                const uint64_t ad = ad0 + (uint64_t)(uint32_t)(kk * 2 * L2_DS);
                                const uint64_t wd = wd0 + (uint64_t)(st * (L2_WSTAGE >> 4) + h * (L2_WSTAGE >> 5));
                                if (kk == 0) tc_mma_f16(tmem + (uint32_t)(jw * L2_DN), ad, wd, idesc_d, 0u);
                                else tc_mma_f16_acc(tmem + (uint32_t)(jw * L2_DN), ad, wd, idesc_d);
                            }
                        }
                        tc_commit(bar(L2B_W_EMPTY + st));
                    }
                    __syncwarp();
                }
                if (elect_one()) tc_commit(bar(L2B_DOWN_FULL));
                __syncwarp();
                wc0 += L2_DSTAGES;
            }
        }
    } else {
        // ===== epilogue: warps 4-15 = three groups of four quadrant warps; unit u = (tile u >> 1, channel half u & 1) goes to group
        // u mod 3.  Warps 4-7 (the group with three units) first move the lo halves of the item's input stream T -> tensor memory =====
        const int q = warp & 3, grp = (L2_EPI_GROUPS - 1) - ((warp - 4) >> 2);
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        for (int k = 0; k < n_my; ++k) {
            const int b = (int)blockIdx.x + k * (int)gridDim.x;
            uint4* scr = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(p.scratch) + (size_t)blockIdx.x * L2_SCRATCH_BYTES);
            if (MODE == 1) {
                // ---- transposed-conv epilogue: unit = (tile, sub-pixel): 64 channels of fine pixel (2cy+dy, 2cx+dx) per coarse row.  hi ->
                // X (the operand of conv 0), lo -> this CTA's scratch (L2), from where the helper warps below move it to tensor memory ----
                mbar_wait(bar(L2B_DOWN_FULL), (uint32_t)(k & 1));
                tc_fence_after();
                for (int u = grp; u < 8; u += L2_EPI_GROUPS) {
                    const int j = u >> 2, sp = u & 3;
                    const int cr = j * MTILE + q * 32 + lane, cy = (cr * 5042) >> 16, cx = cr - cy * 13;        // cr / 13 exactly for cr < 256
                    const bool ok = cr < 156 && cx < 12;
                    const int s = (2 * cy + (sp >> 1)) * L2_WP + 2 * cx + (sp & 1);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const uint32_t a = tmem + lane_base + (uint32_t)(j * L2_UN + sp * L2_C + hf * 32);
                        uint32_t d0[16], d1[16];
                        tc_ld16_nowait(a, d0); tc_ld16_nowait(a + 16, d1);
                        tc_ld_wait16(d0); tc_ld_wait16(d1);
                        if (ok) {
                            float v[32];
#pragma unroll
                            for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(d0[i]); v[16 + i] = __uint_as_float(d1[i]); }
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch) {
                                uint4 hi8, lo8;
                                split8_hilo(v + 8 * ch, hi8, lo8);
                                *reinterpret_cast<uint4*>(smem + row_off(4 * hf + ch, s)) = hi8;
                                __stcg(scr + (4 * hf + ch) * L2_S + s, lo8);
                            }
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(L2B_DOWN_EMPTY));
            }
            if (warp < 8) {
                if (MODE == 1) mbar_wait(bar(L2B_DOWN_EMPTY), (uint32_t)(k & 1));   // every warp has read D (the lo columns are part of it) and written its lo halves
                if (MODE == 1 || k > 0) {
                    // the strided conv's operand (DOWN) / the transposed conv's input (UP) overwrote T: restore the never-written zero rows (gap
                    // above each plane, pad column, zero rows below the stamp) before this item's conv 1 reads T; the valid rows are
                    // rewritten by conv 0's epilogue
                    for (int i = (warp - 4) * 32 + lane; i < 8 * 96 + L2_GAP; i += 128) {
                        int pl, s;
                        if (i < 8 * 96) {
                            const int j = i % 96;
                            pl = 8 + i / 96;
                            s = j < 32 ? j - 32 : j < 56 ? 25 * (j - 32) + 24 : 600 + (j - 56);
                        } else { pl = 15; s = 640 + (i - 8 * 96); }
                        *reinterpret_cast<uint4*>(smem + row_off(pl, s)) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                wait_tiles(0, L2_TILES - 1, 4 * k - 1);                      // the previous item's last epilogues have read the lo stream
                tc_fence_after();
                const uint4* lo_g = reinterpret_cast<const uint4*>(p.x_lo) + (size_t)p.g1.base0 + (size_t)b * p.g1.S;
                for (int t = 0; t < L2_TILES; ++t) {
                    const int s = t * MTILE + q * 32 + lane;
                    uint4 lo[8];
                    if (MODE == 1) {
                        const int yy = (s * 2622) >> 16;
                        const bool ok = s < L2_PIX && s - yy * L2_WP < 24;       // pad rows of the scratch are never written
#pragma unroll
                        for (int ch = 0; ch < 8; ++ch) lo[ch] = ok ? __ldcg(scr + ch * L2_S + s) : make_uint4(0u, 0u, 0u, 0u);
                    } else {
#pragma unroll
                        for (int ch = 0; ch < 8; ++ch)
                            lo[ch] = s < L2_S ? __ldg(lo_g + (size_t)ch * Ptot1 + s) : make_uint4(0u, 0u, 0u, 0u);
                    }
                    const uint32_t taddr = tmem + lane_base + (uint32_t)(L2_LO_COL + t * 32);
                    tc_st16(taddr, reinterpret_cast<const uint32_t*>(lo));
                    tc_st16(taddr + 16, reinterpret_cast<const uint32_t*>(lo) + 16);
                }
                tc_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(L2B_LO_DONE));
            }
            for (int c = 0; c < 4; ++c) {
                const int n = 4 * k + c;
                int ph_waited = -1;
                for (int u = grp; u < 2 * L2_TILES; u += L2_EPI_GROUPS) {
                    const int t = u >> 1, hf = u & 1, ph = t < L2_PH0_TILES ? 0 : 1;
                    if (ph != ph_waited) {
                        if (ph_waited < 0 && ph == 1) mbar_wait(bar(L2B_ACC_FULL), (uint32_t)(n & 1));     // keep every barrier at most one phase ahead
                        mbar_wait(bar(L2B_ACC_FULL + ph), (uint32_t)(n & 1));
                        // DOWN, last conv: its epilogue writes the strided conv's operand over T, which the MMAs of phase 1 still read
                        if (MODE == 0 && c == 3 && ph == 0) mbar_wait(bar(L2B_ACC_FULL + 1), (uint32_t)(n & 1));
                        if (c == 1 && ph_waited < 0) mbar_wait(bar(L2B_LO_DONE), (uint32_t)(k & 1));   // the lo halves of the input stream are in tensor memory
                        tc_fence_after();
                        ph_waited = ph;
                    }
                    const int s = t * MTILE + q * 32 + lane;
                    const int y = (s * 2622) >> 16, x = s - y * L2_WP;       // s / 25 exactly for 0 <= s < 2000
                    const bool valid = s < L2_PIX && x < 24;
                    const uint32_t acc_addr = tmem + lane_base + (uint32_t)(t * L2_C + hf * 32);
                    const uint32_t lo_addr = tmem + lane_base + (uint32_t)(L2_LO_COL + t * 32 + hf * 16);
                    uint32_t r0[16], r1[16], lo[16];
                    tc_ld16_nowait(acc_addr, r0); tc_ld16_nowait(acc_addr + 16, r1);
                    if (c & 1) tc_ld16_nowait(lo_addr, lo);
                    uint4 hi[4];
                    if (c & 1) {
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) hi[ch] = *reinterpret_cast<const uint4*>(smem + row_off(4 * hf + ch, s));
                    }
                    tc_ld_wait16(r0); tc_ld_wait16(r1);
                    if (c & 1) tc_ld_wait16(lo);
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(r0[i]); v[16 + i] = __uint_as_float(r1[i]); }
                    if (!(c & 1)) {
                        if (valid) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(smem + row_off(8 + 4 * hf + ch, s)) = pack8_half(v + 8 * ch);
                        }
                    } else {
                        const uint32_t* hi32 = reinterpret_cast<const uint32_t*>(hi);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float2 f = hilo_pair(hi32[i], lo[i]);
                            v[2 * i] += f.x; v[2 * i + 1] += f.y;
                        }
                        if (c == 1) {
                            uint4 nh[4], nl[4];
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch) split8_hilo(v + 8 * ch, nh[ch], nl[ch]);
                            tc_st16(lo_addr, reinterpret_cast<const uint32_t*>(nl));
                            if (valid) {
#pragma unroll
                                for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(smem + row_off(4 * hf + ch, s)) = nh[ch];
                            }
                            tc_st_wait();
                        } else if (valid) {
                            if (MODE == 0) {
                                // space-to-depth operand of m_down2's strided conv, [32][169][8] over the T planes: K chunk =
                                // (dy*2+dx)*8 + channel chunk, row = coarse padded-linear row
                                const int crow = (y >> 1) * 13 + (x >> 1), ctap = ((y & 1) << 1) | (x & 1);
                                uint4* dst = reinterpret_cast<uint4*>(smem + 8 * L2_PSTRIDE * 16) + (ctap * 8 + 4 * hf) * L2_DS + crow;
#pragma unroll
                                for (int ch = 0; ch < 4; ++ch) dst[ch * L2_DS] = pack8_half(v + 8 * ch);
                            } else {
                                const size_t row = (size_t)p.g1.base0 + (size_t)b * p.g1.S + (size_t)s;
                                const float4* sk = reinterpret_cast<const float4*>(p.skip32) + (size_t)(8 * hf) * Ptot1 + row;
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 a = __ldg(sk + (size_t)i * Ptot1);
                                    v[4 * i] += a.x; v[4 * i + 1] += a.y; v[4 * i + 2] += a.z; v[4 * i + 3] += a.w;
                                }
                                uint4* dst = reinterpret_cast<uint4*>(p.out16) + (size_t)(4 * hf) * Ptot1 + row;
#pragma unroll
                                for (int ch = 0; ch < 4; ++ch) dst[(size_t)ch * Ptot1] = pack8_half(v + 8 * ch);
                            }
                        }
                    }
                    tc_fence_before();
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(L2B_TILE_DONE + t));
                }
                // a group whose units all lie in phase 0 or all in phase 1 still has to follow both accumulator barriers
                if (ph_waited == 0) mbar_wait(bar(L2B_ACC_FULL + 1), (uint32_t)(n & 1));
            }
            if (MODE == 0) {
                // ---- strided-conv epilogue: unit = (tile, 32-channel quarter), eight units round robin over the three groups ----
                mbar_wait(bar(L2B_DOWN_FULL), (uint32_t)(k & 1));
                tc_fence_after();
                const Geom& g2 = p.g2;
                for (int u = grp; u < 8; u += L2_EPI_GROUPS) {
                    const int j = u >> 2, qq = u & 3;
                    const uint32_t a = tmem + lane_base + (uint32_t)(j * L2_DN + qq * 32);
                    uint32_t d0[16], d1[16];
                    tc_ld16_nowait(a, d0); tc_ld16_nowait(a + 16, d1);
                    const int cr = j * MTILE + q * 32 + lane, cy = (cr * 5042) >> 16, cx = cr - cy * 13;        // cr / 13 exactly for cr < 256
                    tc_ld_wait16(d0); tc_ld_wait16(d1);
                    if (cr < 156 && cx < 12) {
                        const size_t row = (size_t)g2.base0 + (size_t)b * g2.S + (size_t)cr;
                        float4* o32 = reinterpret_cast<float4*>(p.skip3) + (size_t)(qq * 8) * g2.Ptot + row;
                        uint4* o16 = reinterpret_cast<uint4*>(p.x3_16) + (size_t)(qq * 4) * g2.Ptot + row;
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(d0[i]); v[16 + i] = __uint_as_float(d1[i]); }
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) o32[(size_t)c4 * g2.Ptot] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) o16[(size_t)c8 * g2.Ptot] = pack8_half(v + 8 * c8);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(L2B_DOWN_EMPTY));
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

static int g_l2_sms = 0;

int conv_l2chain_init() {
    int dev;
    GD_CUDA_CHECK(cudaGetDevice(&dev));
    GD_CUDA_CHECK(cudaDeviceGetAttribute(&g_l2_sms, cudaDevAttrMultiProcessorCount, dev));
    GD_CUDA_CHECK(cudaFuncSetAttribute(k_l2_chain<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, L2_SMEM));
    GD_CUDA_CHECK(cudaFuncSetAttribute(k_l2_chain<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, L2_SMEM));
    return GD_OK;
}

size_t l2chain_scratch_bytes() { return (size_t)160 * L2_SCRATCH_BYTES; }      // one region per CTA of the persistent grid (<= 160 SMs)

// mode 0 (m_down2): x_hi / x_lo = the level's input stream; wdown, skip3, x3_16 = the strided conv and its outputs.
// mode 1 (m_up2):   a_coarse = fp16 input of the transposed conv (level 2), wup its weights, scratch; skip32 = x2, out16 = fp16 (x + x2).
int launch_l2chain(int mode, const Geom& g1, const Geom& g2, int nb, const void* x_hi, const void* x_lo, const void* const* w4, const void* wdown,
                   float* skip3, void* x3_16, const void* a_coarse, const void* wup, void* scratch, const float* skip32, void* out16, cudaStream_t st) {
    if (nb <= 0) return GD_OK;
    if (!g_l2_sms) { set_error("conv_l2chain: library not initialised"); return GD_ECUDA; }
    if (g1.Wp != L2_WP || g1.S != L2_S || g2.Wp != 13 || g2.S != L2_DS) { set_error("conv_l2chain: needs the 24x24 / 12x12 geometries"); return GD_EUNSUPPORTED; }
    L2ChainParams p;
    memset(&p, 0, sizeof(p));
    p.nb = nb; p.mode = mode; p.g1 = g1; p.g2 = g2; p.x_hi = x_hi; p.x_lo = x_lo; p.wdown = wdown; p.skip3 = skip3; p.x3_16 = x3_16; p.a_coarse = a_coarse; p.wup = wup; p.scratch = scratch; p.skip32 = skip32; p.out16 = out16;
    for (int i = 0; i < 4; ++i) p.w[i] = w4[i];
    if (!w4[0] || !w4[1] || !w4[2] || !w4[3] ||
        (mode == 0 ? (!x_hi || !x_lo || !wdown || !skip3 || !x3_16) : (!a_coarse || !wup || !scratch || !skip32 || !out16)) || g_l2_sms > 160) {
        set_error("conv_l2chain: missing buffer"); return GD_EBADSHAPE;
    }
    const int grid = nb < g_l2_sms ? nb : g_l2_sms;
    cudaEvent_t e1 = nullptr;
    const double flops = 4.0 * 2.0 * (double)nb * 576 * (double)L2_C * L2_C * 9 + (mode == 0 ? 2.0 * (double)nb * 144 * 256 * L2_DN : 2.0 * (double)nb * 144 * 128 * L2_UN);
    { int rc = conv_profile_mark(flops, st, &e1); if (rc != GD_OK) return rc; }
    if (mode == 0) k_l2_chain<0><<<grid, L2_THREADS, L2_SMEM, st>>>(p);
    else k_l2_chain<1><<<grid, L2_THREADS, L2_SMEM, st>>>(p);
    GD_LAUNCHED();
    if (e1) GD_CUDA_CHECK(cudaEventRecord(e1, st));
    return GD_OK;
}

}  // namespace gd
