// XDenseUNet denoiser of Tikhonet / ShapeNet (reference: models/XDenseUNet.py:5-115, models/Tikhonet.py:34-47) --
// SURVEY.md section 8f #1, the one model of the reference whose TRAINED weights are in the checkout.
//
// 415,745 parameters, ~150 MMAC per stamp, all of it depthwise-3x3 + pointwise-1x1 with 12-channel growth: there is no
// GEMM wide enough for the tensor cores to matter (N = 12), so the layers run on CUDA cores in fp32 on plain
// [stamp][channel][H][W] planes.  torch.cat never happens: every resolution owns ONE buffer whose channel ranges are
// laid out so that each dense layer reads a suffix [c_in0, c_in0 + C_in) and writes its 12 new channels right in front
// of it, and skip connections / up-sampled features are written straight into the range the consumer will read.
//
//   @48: OUT  [220] = [ out3..out0 of `output` (48) | x1 (112) = [x0 | out3..out0 of `input` | x0] | x6 (60) ]
//   @24: U2   [352] = [ out4..out0 of `up2` (60)    | x2 (220) = [d | out4..out0 of `down1` | d]   | x5 (72) ]
//   @12: U1   [508] = [ out5..out0 of `up1` (72)    | x3 (352) = [d | out5..out0 of `down2` | d]   | x4 (84) ]
//   @6 : BD   [296] = [ out6..out0 of `body` (84)   | d (212) ]
//
// Kernels: k_xd_in (3x3 1->32), k_xd_dense (BN + ReLU + depthwise 3x3 + pointwise C->12, one launch per dense layer),
// k_xd_pw (BN? + ReLU? + 1x1 conv (+bias) with three store modes: plain / 2x2 max-pool (Down) / nearest 2x up-sample (Up)).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/gdeconv.h"
#include "kernels.cuh"
#include "launch.cuh"

namespace gd {

struct XdDense { const float *scale, *shift, *dw, *pw; int c_in0, C_in, c_out0; };
struct XdPw {
    const float *scale, *shift;     // folded BN (nullptr: no BN / ReLU in front of the 1x1 conv)
    const float *wt, *bias;         // wt [C_in][C_out] (transposed), bias [C_out] or nullptr
    int in_buf, c_in0, C_in;        // input: buffer id, channel range
    int out_buf, c_out0, C_out;     // output buffer id (or -1: the user's output tensor), first channel
    int mode;                       // 0 plain, 1 max-pool 2x2 after the conv (Down), 2 nearest up-sampling x2 (Up)
};

constexpr int XD_NBUF = 4;
constexpr int XD_H[XD_NBUF] = {48, 24, 12, 6};
constexpr int XD_C[XD_NBUF] = {220, 352, 508, 296};

// ---- 3x3 conv 1 -> 32, 'same' padding, no bias; writes x0 twice (both ends of x1) ----
__global__ void __launch_bounds__(128) k_xd_in(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ buf,
                                               int Ctot, int c_a, int c_b) {
    const int b = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
    if (p >= NPIX) return;
    const int y = p / STAMP, x = p - y * STAMP;
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        v[t] = (yy >= 0 && yy < STAMP && xx >= 0 && xx < STAMP) ? in[(size_t)b * NPIX + yy * STAMP + xx] : 0.f;
    }
    float* o = buf + (size_t)b * Ctot * NPIX + p;
    for (int c = 0; c < 32; ++c) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) s = fmaf(w[c * 9 + t], v[t], s);
        o[(size_t)(c_a + c) * NPIX] = s;
        o[(size_t)(c_b + c) * NPIX] = s;
    }
}

// ---- one dense layer: out[0..12) = pointwise(depthwise3x3(relu(bn(y)))) with y = channels [c_in0, c_in0 + C_in) ----
__global__ void __launch_bounds__(128) k_xd_dense(float* __restrict__ buf, int Ctot, int H, XdDense L) {
    const int HW = H * H;
    const int b = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
    if (p >= HW) return;
    const int y = p / H, x = p - y * H;
    int off[9];
    bool ok[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        ok[t] = yy >= 0 && yy < H && xx >= 0 && xx < H;
        off[t] = ok[t] ? yy * H + xx : p;
    }
    float acc[12];
#pragma unroll
    for (int o = 0; o < 12; ++o) acc[o] = 0.f;
    float* base = buf + (size_t)b * Ctot * HW;
    const float* in = base + (size_t)L.c_in0 * HW;
    for (int c = 0; c < L.C_in; ++c, in += HW) {
        const float s = L.scale[c], sh = L.shift[c];
        const float* dw = L.dw + c * 9;
        float d = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float a = fmaxf(fmaf(in[off[t]], s, sh), 0.f);      // BN(eval) + ReLU; 'same' padding pads AFTER the activation
            d = fmaf(dw[t], ok[t] ? a : 0.f, d);
        }
        const float* pw = L.pw + c * 12;
#pragma unroll
        for (int o = 0; o < 12; ++o) acc[o] = fmaf(pw[o], d, acc[o]);
    }
    float* out = base + (size_t)L.c_out0 * HW + p;
#pragma unroll
    for (int o = 0; o < 12; ++o) out[(size_t)o * HW] = acc[o];
}

// ---- the same dense layer for the two large resolutions (48x48, 24x24: 85 % of the dense-layer work), tiled through shared memory.
// k_xd_dense above reloads and re-activates the 9 neighbours of every pixel and channel from global memory and fetches 23 weights per
// channel with uniform global loads: ~70 instructions per pixel and channel, bound by load latency.  Here a CTA owns a band of TR
// rows of one stamp, 8 input channels at a time are activated ONCE into a zero-bordered tile, the chunk's weights sit in shared
// memory, and a thread produces PX horizontally adjacent pixels from a 3 x (PX + 2) window: ~35 instructions per pixel and channel
// (the tile <-> image index decode is hoisted out of the channel loop: with it inside, the fill cost as much as the layer itself).
// The arithmetic (and its order) is k_xd_dense's, so the two kernels agree bit for bit (tools/gpu/xd_ab.py).  GDECONV_XD_TILED=0
// restores k_xd_dense everywhere.
template <int H, int PX, int TR>
__global__ void __launch_bounds__(128) k_xd_dense_tiled(float* __restrict__ buf, int Ctot, XdDense L) {
    constexpr int HW = H * H, RS = H + 2, TROWS = TR + 2, TILE = TROWS * RS, TPR = H / PX, NACT = TR * TPR, CHK = 8;
    static_assert(H % PX == 0 && H % TR == 0 && NACT <= 128, "band geometry");
    __shared__ float tile[CHK * TILE];
    __shared__ __align__(16) float wsm[CHK * 24];          // per channel: dw[9], 3 pad, pw[12]
    const int b = blockIdx.y, y0 = blockIdx.x * TR, tid = threadIdx.x;
    const int r = tid / TPR, x0 = (tid - r * TPR) * PX;
    const bool active = tid < NACT;
    float acc[PX][12];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int o = 0; o < 12; ++o) acc[p][o] = 0.f;
    // this thread's elements of the activated tile (the same for every channel): tile index tid + 128 i <-> image pixel, decoded once
    constexpr int NSLOT = (TILE + 127) / 128;
    int slot_off[NSLOT];
    bool slot_in[NSLOT], slot_ok[NSLOT];
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) {
        const int e = tid + 128 * i, rr = e / RS, cc = e - rr * RS, y = y0 + rr - 1, x = cc - 1;
        slot_in[i] = e < TILE;
        slot_ok[i] = slot_in[i] && y >= 0 && y < H && x >= 0 && x < H;
        slot_off[i] = slot_ok[i] ? y * H + x : 0;
    }
    float* base = buf + (size_t)b * Ctot * HW;
    const float* in = base + (size_t)L.c_in0 * HW;
    for (int c0 = 0; c0 < L.C_in; c0 += CHK) {
        const int n = L.C_in - c0 < CHK ? L.C_in - c0 : CHK;
        __syncthreads();                                   // the previous chunk's windows have been read
        for (int i = tid; i < n * 24; i += 128) {
            const int ch = i / 24, j = i - ch * 24;
            wsm[i] = j < 9 ? L.dw[(c0 + ch) * 9 + j] : j < 12 ? 0.f : L.pw[(c0 + ch) * 12 + j - 12];
        }
        for (int ch = 0; ch < n; ++ch) {
            const float s = L.scale[c0 + ch], sh = L.shift[c0 + ch];
            const float* src = in + (size_t)(c0 + ch) * HW;
            float* dst = tile + ch * TILE + tid;
#pragma unroll
            for (int i = 0; i < NSLOT; ++i) {
                // BN(eval) + ReLU; 'same' padding pads AFTER the activation (border elements of the tile are stored as zeros)
                if (slot_in[i]) dst[128 * i] = slot_ok[i] ? fmaxf(fmaf(src[slot_off[i]], s, sh), 0.f) : 0.f;
            }
        }
        __syncthreads();
        if (active) {
            for (int ch = 0; ch < n; ++ch) {
                const float* tw = tile + ch * TILE + r * RS + x0;
                float a[3][PX + 2];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int j = 0; j < PX + 2; ++j) a[ky][j] = tw[ky * RS + j];
                const float4* w4 = reinterpret_cast<const float4*>(wsm + ch * 24);
                float w[24];
#pragma unroll
                for (int q = 0; q < 6; ++q) { const float4 t = w4[q]; w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w; }
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    float d = 0.f;
#pragma unroll
                    for (int t = 0; t < 9; ++t) d = fmaf(w[t], a[t / 3][p + t % 3], d);
#pragma unroll
                    for (int o = 0; o < 12; ++o) acc[p][o] = fmaf(w[12 + o], d, acc[p][o]);
                }
            }
        }
    }
    if (active) {
        float* out = base + (size_t)L.c_out0 * HW + (y0 + r) * H + x0;
#pragma unroll
        for (int o = 0; o < 12; ++o)
#pragma unroll
            for (int p = 0; p < PX; ++p) out[(size_t)o * HW + p] = acc[p][o];
    }
}

// ---- 1x1 conv with optional BN+ReLU in front and pool / up-sample behind; OT output channels per thread ----
constexpr int XD_OT = 16;
template <int MODE>
__global__ void __launch_bounds__(128) k_xd_pw(const float* __restrict__ ibuf, int in_Ctot, int Hin, float* __restrict__ obuf,
                                               int out_Ctot, int Hout, XdPw L, const float* __restrict__ out_scale) {
    // a thread owns one OUTPUT-side site: mode 0/2: one input pixel; mode 1: one pooled pixel = 2x2 input pixels
    const int Hs = MODE == 1 ? Hin / 2 : Hin, HWs = Hs * Hs, HWin = Hin * Hin, HWout = Hout * Hout;
    const int b = blockIdx.z, o0 = blockIdx.y * XD_OT, p = blockIdx.x * 128 + threadIdx.x;
    if (p >= HWs) return;
    const int y = p / Hs, x = p - y * Hs;
    constexpr int nsub = MODE == 1 ? 4 : 1;
    int pin[4];
    if (MODE == 1) { pin[0] = (2 * y) * Hin + 2 * x; pin[1] = pin[0] + 1; pin[2] = pin[0] + Hin; pin[3] = pin[2] + 1; }
    else { pin[0] = p; pin[1] = pin[2] = pin[3] = p; }
    float acc[nsub][XD_OT];
#pragma unroll
    for (int q = 0; q < nsub; ++q)
#pragma unroll
        for (int o = 0; o < XD_OT; ++o) acc[q][o] = 0.f;
    const float* in = ibuf + ((size_t)b * in_Ctot + L.c_in0) * HWin;
    const int no = L.C_out - o0 < XD_OT ? L.C_out - o0 : XD_OT;
    for (int c = 0; c < L.C_in; ++c, in += HWin) {
        float a[nsub];
#pragma unroll
        for (int q = 0; q < nsub; ++q) {
            a[q] = in[pin[q]];
            if (L.scale) a[q] = fmaxf(fmaf(a[q], L.scale[c], L.shift[c]), 0.f);
        }
        const float* w = L.wt + (size_t)c * L.C_out + o0;
#pragma unroll
        for (int o = 0; o < XD_OT; ++o) {
            const float wv = o < no ? w[o] : 0.f;
#pragma unroll
            for (int q = 0; q < nsub; ++q) acc[q][o] = fmaf(wv, a[q], acc[q][o]);
        }
    }
    const float osc = out_scale ? out_scale[b] : 1.f;
#pragma unroll
    for (int o = 0; o < XD_OT; ++o) {
        if (o >= no) break;
        float v = acc[0][o];
        if (MODE == 1) v = fmaxf(fmaxf(acc[0][o], acc[nsub > 1 ? 1 : 0][o]), fmaxf(acc[nsub > 2 ? 2 : 0][o], acc[nsub > 3 ? 3 : 0][o]));
        if (L.bias) v += L.bias[o0 + o];
        v *= osc;
        float* dst = obuf + ((size_t)b * out_Ctot + L.c_out0 + o0 + o) * HWout;
        if (MODE == 2) {
            float* d2 = dst + (2 * y) * Hout + 2 * x;
            d2[0] = v; d2[1] = v; d2[Hout] = v; d2[Hout + 1] = v;
        } else {
            dst[p] = v;
        }
    }
}

}  // namespace gd

using namespace gd;

struct GdXDense {
    int device;
    void* blob;
    const float* w_in;               // [32][9]
    std::vector<XdDense> dense[7];   // input, down1, down2, body, up1, up2, output
    XdPw down[3], up[3], outc;
};

namespace {

struct XFinder {
    std::map<std::string, const GdTensorDesc*> m;
    std::string prefix;
    const GdTensorDesc* get(const std::string& k) const {
        auto it = m.find(prefix + k);
        return it == m.end() ? nullptr : it->second;
    }
};

struct XBlob {
    std::vector<float> h;
    size_t put(const float* p, size_t n) { size_t o = (h.size() + 3) / 4 * 4; h.resize(o + n); memcpy(h.data() + o, p, n * sizeof(float)); return o; }
    size_t put(const std::vector<float>& v) { return put(v.data(), v.size()); }
};

}  // namespace

#define XD_FAIL(...) do { gd::set_error(__VA_ARGS__); return GD_EBADSHAPE; } while (0)

// BN(eval) -> (scale, shift): y = x*scale + shift
static int fold_bn(const XFinder& F, const std::string& key, int C, XBlob& B, size_t* sc, size_t* sh) {
    const GdTensorDesc *w = F.get(key + ".weight"), *b = F.get(key + ".bias"), *m = F.get(key + ".running_mean"), *v = F.get(key + ".running_var");
    if (!w || !b || !m || !v || w->shape[0] != C) XD_FAIL("XDenseUNet: BatchNorm %s missing or not %d channels", key.c_str(), C);
    std::vector<float> s(C), t(C);
    for (int c = 0; c < C; ++c) { s[c] = w->data[c] / sqrtf(v->data[c] + 1e-5f); t[c] = b->data[c] - m->data[c] * s[c]; }
    *sc = B.put(s); *sh = B.put(t);
    return GD_OK;
}

extern "C" GD_API int gd_pack_xdense(const GdTensorDesc* tensors, int n_tensors, const char* prefix, int device, GdXDense** out) {
    if (!out || !tensors) XD_FAIL("gd_pack_xdense: NULL argument");
    *out = nullptr;
    XFinder F;
    F.prefix = prefix ? prefix : "";
    for (int i = 0; i < n_tensors; ++i)
        if (tensors[i].name && tensors[i].data) F.m[tensors[i].name] = &tensors[i];
    XBlob B;
    GdXDense* X = new GdXDense();
    X->device = device;
    struct Fix { const float** slot; size_t off; };
    std::vector<Fix> fix;
    auto slot = [&](const float** p, size_t off) { fix.push_back({p, off}); };
    const GdTensorDesc* t = F.get("input.0.weight");
    if (!t || t->ndim != 4 || t->shape[0] != 32 || t->shape[1] != 1 || t->shape[2] != 3) { delete X; XD_FAIL("XDenseUNet: input.0.weight must be (32,1,3,3)"); }
    slot(&X->w_in, B.put(t->data, 32 * 9));
    // dense blocks: name, #layers, first input channel count, c_in0 of layer 0 inside its buffer
    struct Blk { const char* name; int layers, C0, c_in0; };
    const Blk blks[7] = {{"input.1", 4, 32, 48 + 80}, {"down1.1", 5, 80, 60 + 140}, {"down2.1", 6, 140, 72 + 212}, {"body.1", 7, 212, 84},
                         {"up1.0", 6, 436, 72}, {"up2.0", 5, 292, 60}, {"output.0", 4, 172, 48}};
    // reserve first so that pointers into the vectors stay valid
    for (int k = 0; k < 7; ++k) X->dense[k].resize(blks[k].layers);
    for (int k = 0; k < 7; ++k)
        for (int i = 0; i < blks[k].layers; ++i) {
            const int C = blks[k].C0 + 12 * i;
            char key[128];
            snprintf(key, sizeof key, "%s.net.%d", blks[k].name, i);
            size_t sc, sh;
            int rc = fold_bn(F, std::string(key) + ".0", C, B, &sc, &sh);
            if (rc) { delete X; return rc; }
            const GdTensorDesc* dw = F.get(std::string(key) + ".2.depthewise.weight");
            const GdTensorDesc* pw = F.get(std::string(key) + ".2.pointwise.weight");
            if (!dw || !pw || dw->shape[0] != C || pw->shape[0] != 12 || pw->shape[1] != C) { delete X; XD_FAIL("XDenseUNet: %s separable conv missing or mis-shaped", key); }
            std::vector<float> pwt((size_t)C * 12);
            for (int o = 0; o < 12; ++o)
                for (int c = 0; c < C; ++c) pwt[(size_t)c * 12 + o] = pw->data[(size_t)o * C + c];
            XdDense& L = X->dense[k][i];
            L.C_in = C; L.c_in0 = blks[k].c_in0 - 12 * i; L.c_out0 = L.c_in0 - 12;
            slot(&L.scale, sc); slot(&L.shift, sh); slot(&L.dw, B.put(dw->data, (size_t)C * 9)); slot(&L.pw, B.put(pwt));
        }
    auto pack_pw = [&](XdPw& L, const std::string& wkey, const std::string& bnkey, int Cin, int Cout, bool bias) -> int {
        const GdTensorDesc* w = F.get(wkey + ".weight");
        if (!w || w->shape[0] != Cout || w->shape[1] != Cin) XD_FAIL("XDenseUNet: %s.weight missing or not (%d,%d,1,1)", wkey.c_str(), Cout, Cin);
        std::vector<float> wt((size_t)Cin * Cout);
        for (int o = 0; o < Cout; ++o)
            for (int c = 0; c < Cin; ++c) wt[(size_t)c * Cout + o] = w->data[(size_t)o * Cin + c];
        slot(&L.wt, B.put(wt));
        L.bias = nullptr; L.scale = nullptr; L.shift = nullptr;
        if (bias) {
            const GdTensorDesc* bb = F.get(wkey + ".bias");
            if (!bb || bb->shape[0] != Cout) XD_FAIL("XDenseUNet: %s.bias missing", wkey.c_str());
            slot(&L.bias, B.put(bb->data, Cout));
        }
        if (!bnkey.empty()) {
            size_t sc, sh;
            int rc = fold_bn(F, bnkey, Cin, B, &sc, &sh);
            if (rc) return rc;
            slot(&L.scale, sc); slot(&L.shift, sh);
        }
        L.C_in = Cin; L.C_out = Cout;
        return GD_OK;
    };
    int rc = GD_OK;
    // Down: BN + ReLU + 1x1 + MaxPool2 -> first channels `d` of the next resolution's x block, copied to its far end by the dense block layout
    if (!rc) rc = pack_pw(X->down[0], "down1.0.net.2", "down1.0.net.0", 112, 80, false);
    if (!rc) rc = pack_pw(X->down[1], "down2.0.net.2", "down2.0.net.0", 220, 140, false);
    if (!rc) rc = pack_pw(X->down[2], "body.0.net.2", "body.0.net.0", 352, 212, false);
    if (!rc) rc = pack_pw(X->up[0], "body.2.net.0", "", 296, 84, true);
    if (!rc) rc = pack_pw(X->up[1], "up1.1.net.0", "", 508, 72, true);
    if (!rc) rc = pack_pw(X->up[2], "up2.1.net.0", "", 352, 60, true);
    if (!rc) rc = pack_pw(X->outc, "output.1", "", 220, 1, true);
    if (rc) { delete X; return rc; }
    gd::DeviceScope scope(device);
    cudaError_t e = scope.err;
    // Tikhonet's Tikhonov step (k_wiener48) needs its opt-in shared-memory size even when this is the first call of the process
    if (e == cudaSuccess) { int irc = gd::ensure_device_init(device); if (irc != GD_OK) { delete X; return irc; } }
    if (e == cudaSuccess) e = cudaMalloc(&X->blob, B.h.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(X->blob, B.h.data(), B.h.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { gd::set_error("gd_pack_xdense: %s", cudaGetErrorString(e)); delete X; return GD_ECUDA; }
    for (auto& f : fix) *f.slot = (const float*)X->blob + f.off;
    *out = X;
    return GD_OK;
}

extern "C" GD_API void gd_free_xdense(GdXDense* x) {
    if (!x) return;
    {
        gd::DeviceScope scope(x->device);
        cudaFree(x->blob);
    }
    delete x;
}

static size_t xd_buf_floats(int i) { return (size_t)XD_C[i] * XD_H[i] * XD_H[i]; }

extern "C" GD_API size_t gd_xdense_workspace_bytes(int chunk) {
    if (chunk < 1) return 0;
    size_t per = 0;
    for (int i = 0; i < XD_NBUF; ++i) per += xd_buf_floats(i);
    return ((size_t)chunk * per + (size_t)chunk * NPIX) * sizeof(float) + 1024;
}

static int xd_forward_chunk(const GdXDense* X, const float* in, float* out, const float* out_scale, int nb, float* const* buf, cudaStream_t st) {
    auto dense = [&](int blk, int b) -> int {
        const int H = XD_H[b], HW = H * H;
        static int tiled = -1;
        if (tiled < 0) { const char* e = getenv("GDECONV_XD_TILED"); tiled = e ? atoi(e) : 1; }
        for (const XdDense& L : X->dense[blk]) {
            if (tiled && H == 48) k_xd_dense_tiled<48, 3, 8><<<dim3(6, nb), 128, 0, st>>>(buf[b], XD_C[b], L);
            else if (tiled && H == 24) k_xd_dense_tiled<24, 3, 12><<<dim3(2, nb), 128, 0, st>>>(buf[b], XD_C[b], L);
            else k_xd_dense<<<dim3((HW + 127) / 128, nb), 128, 0, st>>>(buf[b], XD_C[b], H, L);
            GD_LAUNCHED();
        }
        return GD_OK;
    };
    auto pw = [&](XdPw L, int ib, int c_in0, int ob, int c_out0, int mode, float* user_out, const float* osc) -> int {
        L.c_in0 = c_in0; L.c_out0 = c_out0; L.mode = mode;
        const int Hin = XD_H[ib], Hout = ob >= 0 ? XD_H[ob] : STAMP, Hs = mode == 1 ? Hin / 2 : Hin;
        dim3 grid((Hs * Hs + 127) / 128, (L.C_out + XD_OT - 1) / XD_OT, nb);
        float* optr = ob >= 0 ? buf[ob] : user_out;
        const int octot = ob >= 0 ? XD_C[ob] : 1;
        if (mode == 0) k_xd_pw<0><<<grid, 128, 0, st>>>(buf[ib], XD_C[ib], Hin, optr, octot, Hout, L, osc);
        else if (mode == 1) k_xd_pw<1><<<grid, 128, 0, st>>>(buf[ib], XD_C[ib], Hin, optr, octot, Hout, L, osc);
        else k_xd_pw<2><<<grid, 128, 0, st>>>(buf[ib], XD_C[ib], Hin, optr, octot, Hout, L, osc);
        GD_LAUNCHED();
        return GD_OK;
    };
    auto copy_ch = [&](int b, int c_src, int c_dst, int C) -> int {          // duplicate `d` at the far end of an x block
        const int HW = XD_H[b] * XD_H[b];
        GD_CUDA_CHECK(cudaMemcpy2DAsync(buf[b] + (size_t)c_dst * HW, (size_t)XD_C[b] * HW * 4, buf[b] + (size_t)c_src * HW, (size_t)XD_C[b] * HW * 4,
                                        (size_t)C * HW * 4, nb, cudaMemcpyDeviceToDevice, st));
        return GD_OK;
    };
    int rc;
    // input: x0 -> OUT[48..80) and OUT[128..160); dense block writes OUT[80..128)
    k_xd_in<<<dim3((NPIX + 127) / 128, nb), 128, 0, st>>>(in, X->w_in, buf[0], XD_C[0], 48, 128);
    GD_LAUNCHED();
    if ((rc = dense(0, 0))) return rc;
    // down1: x1 = OUT[48..160) -> d (80) at U2[60..140); dense -> U2[140..200); copy d -> U2[200..280)
    if ((rc = pw(X->down[0], 0, 48, 1, 60, 1, nullptr, nullptr))) return rc;
    if ((rc = copy_ch(1, 60, 200, 80))) return rc;
    if ((rc = dense(1, 1))) return rc;
    // down2: x2 = U2[60..280) -> d (140) at U1[72..212); dense -> U1[212..284); copy d -> U1[284..424)
    if ((rc = pw(X->down[1], 1, 60, 2, 72, 1, nullptr, nullptr))) return rc;
    if ((rc = copy_ch(2, 72, 284, 140))) return rc;
    if ((rc = dense(2, 2))) return rc;
    // body: x3 = U1[72..424) -> d (212) at BD[84..296); dense -> BD[0..84); Up(296 -> 84) -> x4 = U1[424..508)
    if ((rc = pw(X->down[2], 2, 72, 3, 84, 1, nullptr, nullptr))) return rc;
    if ((rc = dense(3, 3))) return rc;
    if ((rc = pw(X->up[0], 3, 0, 2, 424, 2, nullptr, nullptr))) return rc;
    // up1: cat(x3, x4) = U1[72..508); dense -> U1[0..72); Up(508 -> 72) -> x5 = U2[280..352)
    if ((rc = dense(4, 2))) return rc;
    if ((rc = pw(X->up[1], 2, 0, 1, 280, 2, nullptr, nullptr))) return rc;
    // up2: cat(x2, x5) = U2[60..352); dense -> U2[0..60); Up(352 -> 60) -> x6 = OUT[160..220)
    if ((rc = dense(5, 1))) return rc;
    if ((rc = pw(X->up[2], 1, 0, 0, 160, 2, nullptr, nullptr))) return rc;
    // output: cat(x1, x6) = OUT[48..220); dense -> OUT[0..48); 1x1 conv 220 -> 1 (+bias), times out_scale
    if ((rc = dense(6, 0))) return rc;
    return pw(X->outc, 0, 0, -1, 0, 0, out, out_scale);
}

static int xd_run(const GdXDense* X, int filter, float lam, const float* y, const float* psf, const float* alpha, const float* in, float* out,
                  const float* out_scale, int batch, void* ws, size_t ws_bytes, int chunk, cudaStream_t st) {
    if (!X) XD_FAIL("XDenseUNet weights are NULL");
    if (batch < 0 || chunk < 1 || !ws || ws_bytes < gd_xdense_workspace_bytes(chunk)) { gd::set_error("XDenseUNet: workspace too small for chunk %d", chunk); return GD_EWORKSPACE; }
    GD_DEVICE_SCOPE(X->device);
    { int irc = gd::ensure_device_init(X->device); if (irc != GD_OK) return irc; }
    float* base = (float*)ws;
    float* buf[XD_NBUF];
    for (int i = 0; i < XD_NBUF; ++i) { buf[i] = base; base += (size_t)chunk * xd_buf_floats(i); }
    float* tik = base;                                  // [chunk][2304] Tikhonov output
    for (int c0 = 0; c0 < batch; c0 += chunk) {
        const int nb = batch - c0 < chunk ? batch - c0 : chunk;
        const size_t o = (size_t)c0 * NPIX;
        const float* src = in ? in + o : tik;
        if (!in) {   // Tikhonet.forward (models/Tikhonet.py:41-47): clamp, Tikhonov step, denoise, * alpha
            int rc = launch_solver(filter | 0x100, 0, lam, y + o, psf + o, alpha + c0, tik, nb, st);
            if (rc) return rc;
        }
        int rc = xd_forward_chunk(X, src, out + o, out_scale ? out_scale + c0 : nullptr, nb, buf, st);
        if (rc) return rc;
    }
    return GD_OK;
}

extern "C" GD_API int gd_xdense_forward(const GdXDense* X, const float* in, float* out, int batch, void* ws, size_t ws_bytes, int chunk, void* stream) {
    if (batch && (!in || !out)) XD_FAIL("gd_xdense_forward: NULL buffers");
    return xd_run(X, 0, 0.f, nullptr, nullptr, nullptr, in, out, nullptr, batch, ws, ws_bytes, chunk, (cudaStream_t)stream);
}

extern "C" GD_API int gd_tikhonet_forward(const GdXDense* X, int filter, float lam, const float* y, const float* psf, const float* alpha, float* out,
                                          int batch, void* ws, size_t ws_bytes, int chunk, void* stream) {
    if (filter != GD_SOLVER_TIKHONOV_ID && filter != GD_SOLVER_TIKHONOV_LAP) XD_FAIL("gd_tikhonet_forward: filter must be GD_SOLVER_TIKHONOV_ID or _LAP");
    if (batch && (!y || !psf || !alpha || !out)) XD_FAIL("gd_tikhonet_forward: NULL buffers");
    return xd_run(X, filter, lam, y, psf, alpha, nullptr, out, alpha, batch, ws, ws_bytes, chunk, (cudaStream_t)stream);
}
