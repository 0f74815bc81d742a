// Shared-memory batched FFT building blocks (N = 48, 96, 128) for the Galaxy-Deconv hot path.
//
// Replaces the reference's torch.fft calls on 48x48 / 96x96 / 128x128 grids
// (utils/utils_torch.py:21-50, models/unrolled_admm_gaussian.py:66,91-92,114,121-122).
//
// Everything here is __host__ __device__ and written as "phases": a phase is a function of
// (work item index) that touches a disjoint set of elements, so a kernel runs
//     for (w = tid; w < items; w += nthreads) phase(w);  __syncthreads();
// and the CPU unit test (tests/cpu/test_fft_core.cpp) runs the very same code with a plain loop.
//
// 1-D transform of length N = R1*R2, two passes, in place, no scratch:
//   forward ("F", natural order in -> digit-swapped order out):
//     pass F1, item n2 in [0,R2): R1-point DFT over n1 of x[R2*n1+n2], times W_N^(n2*k1), stored at R2*k1+n2
//     pass F2, item k1 in [0,R1): R2-point DFT over n2 of the contiguous block R2*k1+[0,R2) -> slot R2*k1+k2
//   so frequency k = k1 + R1*k2 ends in slot(k) = R2*(k % R1) + k / R1.
//   inverse ("I", digit-swapped in -> natural order out) runs the two passes in the opposite order and
//   uses the re/im swap identity  N*idft(X) = swap(dft(swap(X)))  so only forward codelets exist.
//   No 1/N scaling is applied here; callers fold it into their pointwise step.
#pragma once
#include <cuda_runtime.h>

#ifndef GD_HD
#define GD_HD __host__ __device__ __forceinline__
#endif

namespace gdfft {

GD_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
GD_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GD_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
GD_HD float2 cswap(float2 a) { return make_float2(a.y, a.x); }
GD_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by (-i)^m
GD_HD float2 mul_negi_pow(float2 a, int m) {
    switch (m & 3) {
        case 0: return a;
        case 1: return make_float2(a.y, -a.x);
        case 2: return make_float2(-a.x, -a.y);
        default: return make_float2(-a.y, a.x);
    }
}

// ---- compile-time twiddles for the in-register codelets: W_R^m = exp(-2 pi i m / R) ----
template <int R> struct TwC;
template <> struct TwC<6> {
    GD_HD static float2 w(int m) {
        constexpr float c[6] = {1.f, 0.5f, -0.5f, -1.f, -0.5f, 0.5f};
        constexpr float s[6] = {0.f, -0.86602540378443865f, -0.86602540378443865f, 0.f, 0.86602540378443865f, 0.86602540378443865f};
        return make_float2(c[m], s[m]);
    }
};
template <> struct TwC<8> {
    GD_HD static float2 w(int m) {
        constexpr float h = 0.70710678118654752f;
        constexpr float c[8] = {1.f, h, 0.f, -h, -1.f, -h, 0.f, h};
        constexpr float s[8] = {0.f, -h, -1.f, -h, 0.f, h, 1.f, h};
        return make_float2(c[m], s[m]);
    }
};
template <> struct TwC<12> {
    GD_HD static float2 w(int m) {
        constexpr float a = 0.86602540378443865f;
        constexpr float c[12] = {1.f, a, 0.5f, 0.f, -0.5f, -a, -1.f, -a, -0.5f, 0.f, 0.5f, a};
        constexpr float s[12] = {0.f, -0.5f, -a, -1.f, -a, -0.5f, 0.f, 0.5f, a, 1.f, a, 0.5f};
        return make_float2(c[m], s[m]);
    }
};
template <> struct TwC<16> {
    GD_HD static float2 w(int m) {
        constexpr float h = 0.70710678118654752f, p = 0.92387953251128674f, q = 0.38268343236508977f;
        constexpr float c[16] = {1.f, p, h, q, 0.f, -q, -h, -p, -1.f, -p, -h, -q, 0.f, q, h, p};
        constexpr float s[16] = {0.f, -q, -h, -p, -1.f, -p, -h, -q, 0.f, q, h, p, 1.f, p, h, q};
        return make_float2(c[m], s[m]);
    }
};

template <> struct TwC<48> {          // m in [0, 30]: the twiddles n2*k1 of the 3 x 16 split
    GD_HD static float2 w(int m) {
        constexpr float c[31] = {1.f, 0.99144486137381038f, 0.96592582628906831f, 0.92387953251128674f, 0.86602540378443871f, 0.79335334029123517f, 0.70710678118654757f, 0.60876142900872066f, 0.5f, 0.38268343236508984f, 0.25881904510252074f, 0.13052619222005171f, 0.f, -0.1305261922200516f, -0.25881904510252063f, -0.3826834323650895f, -0.5f, -0.60876142900872066f, -0.70710678118654746f, -0.79335334029123505f, -0.86602540378443871f, -0.92387953251128674f, -0.9659258262890682f, -0.99144486137381038f, -1.f, -0.99144486137381038f, -0.96592582628906831f, -0.92387953251128685f, -0.86602540378443882f, -0.79335334029123517f, -0.70710678118654791f};
        constexpr float s[31] = {0.f, -0.13052619222005157f, -0.25881904510252074f, -0.38268343236508978f, -0.5f, -0.60876142900872066f, -0.70710678118654746f, -0.79335334029123517f, -0.8660254037844386f, -0.92387953251128674f, -0.96592582628906831f, -0.99144486137381038f, -1.f, -0.99144486137381038f, -0.96592582628906831f, -0.92387953251128685f, -0.86602540378443871f, -0.79335334029123517f, -0.70710678118654757f, -0.60876142900872088f, -0.5f, -0.38268343236508989f, -0.25881904510252102f, -0.13052619222005199f, 0.f, 0.13052619222005177f, 0.25881904510252079f, 0.38268343236508967f, 0.5f, 0.60876142900872066f, 0.70710678118654713f};
        return make_float2(c[m], s[m]);
    }
};

// ---- in-register forward DFT codelets, natural order in and out ----
template <int R> struct Dft;

template <> struct Dft<2> {
    GD_HD static void run(float2* v) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    }
};
template <> struct Dft<3> {
    GD_HD static void run(float2* v) {
        constexpr float s = 0.86602540378443865f;
        float2 a = v[0], t1 = cadd(v[1], v[2]), t2 = csub(v[1], v[2]);
        float2 m = make_float2(a.x - 0.5f * t1.x, a.y - 0.5f * t1.y);
        // -i * s * t2
        float2 r = make_float2(s * t2.y, -s * t2.x);
        v[0] = cadd(a, t1); v[1] = cadd(m, r); v[2] = csub(m, r);
    }
};
template <> struct Dft<4> {
    GD_HD static void run(float2* v) {
        float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        float2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        float2 nid = make_float2(d.y, -d.x);            // -i * d
        v[0] = cadd(a, c); v[2] = csub(a, c);
        v[1] = cadd(b, nid); v[3] = csub(b, nid);
    }
};
// R = RA*RB by one Cooley-Tukey step in registers
template <int RA, int RB> struct DftComposite {
    GD_HD static void run(float2* v) {
        constexpr int R = RA * RB;
        float2 t[R];
#pragma unroll
        for (int nb = 0; nb < RB; ++nb) {
            float2 a[RA];
#pragma unroll
            for (int na = 0; na < RA; ++na) a[na] = v[RB * na + nb];
            Dft<RA>::run(a);
#pragma unroll
            for (int ka = 0; ka < RA; ++ka)
                t[ka * RB + nb] = (nb * ka == 0) ? a[ka] : cmul(a[ka], TwC<R>::w((nb * ka) % R));
        }
#pragma unroll
        for (int ka = 0; ka < RA; ++ka) {
            float2 b[RB];
#pragma unroll
            for (int nb = 0; nb < RB; ++nb) b[nb] = t[ka * RB + nb];
            Dft<RB>::run(b);
#pragma unroll
            for (int kb = 0; kb < RB; ++kb) v[ka + RA * kb] = b[kb];
        }
    }
};
template <> struct Dft<6> : DftComposite<2, 3> {};
template <> struct Dft<8> : DftComposite<2, 4> {};
template <> struct Dft<12> : DftComposite<3, 4> {};
template <> struct Dft<16> : DftComposite<4, 4> {};

// ---- whole 48-point forward DFT in the registers of ONE thread (3 x 16 Cooley-Tukey, every index a compile-time constant) ----
// in: v[n] natural order;  out: frequency k is in v[Fft48::reg(k)]  (v[16*k1 + k2] = X[k1 + 3*k2]).
// The inverse uses idft(X) = conj(dft(conj(X))) (unscaled), i.e. conjugate on the way in and out.
struct Fft48 {
    GD_HD static constexpr int reg(int k) { return 16 * (k % 3) + k / 3; }
    GD_HD static void run(float2* v) {
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
            float2 a[3] = {v[n2], v[16 + n2], v[32 + n2]};
            Dft<3>::run(a);
            v[n2] = a[0];
            v[16 + n2] = n2 ? cmul(a[1], TwC<48>::w(n2)) : a[1];
            v[32 + n2] = n2 ? cmul(a[2], TwC<48>::w(2 * n2)) : a[2];
        }
        Dft<16>::run(v);
        Dft<16>::run(v + 16);
        Dft<16>::run(v + 32);
    }
};

// ---- two-pass length-N transform on one line held in (shared) memory ----
// tw[m] = exp(-2 pi i m / N), m in [0, N)
template <int N, int R1, int R2> struct Line {
    static_assert(R1 * R2 == N, "N = R1*R2");
    GD_HD static int slot(int k) { return R2 * (k % R1) + k / R1; }       // where frequency k lives
    GD_HD static int freq(int s) { return s / R2 + R1 * (s % R2); }       // which frequency slot s holds

    // forward pass 1, item n2 in [0,R2).  NZ1 = number of leading non-zero n1 groups (inputs with
    // n1 >= NZ1 are taken as zero and never read) -- the zero-padding prune.
    template <int NZ1 = R1>
    GD_HD static void f1(float2* x, int stride, int n2, const float2* tw) {
        float2 v[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) v[n1] = (n1 < NZ1) ? x[(R2 * n1 + n2) * stride] : make_float2(0.f, 0.f);
        Dft<R1>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) x[(R2 * k1 + n2) * stride] = (k1 == 0) ? v[0] : cmul(v[k1], tw[n2 * k1]);
    }
    // forward pass 2, item k1 in [0,R1)
    GD_HD static void f2(float2* x, int stride, int k1) {
        float2 v[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) v[n2] = x[(R2 * k1 + n2) * stride];
        Dft<R2>::run(v);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) x[(R2 * k1 + k2) * stride] = v[k2];
    }
    // inverse step A, item k1 in [0,R1): contiguous R2-DFT on swapped data, twiddle, stays swapped
    GD_HD static void i1(float2* x, int stride, int k1, const float2* tw) {
        float2 v[R2];
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) v[k2] = cswap(x[(R2 * k1 + k2) * stride]);
        Dft<R2>::run(v);
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) x[(R2 * k1 + n2) * stride] = (k1 == 0) ? v[n2] : cmul(v[n2], tw[n2 * k1]);
    }
    // inverse step B, item n2 in [0,R2): strided R1-DFT, un-swap, natural order out.
    // Only outputs n1 < NO1 are stored (crop prune).
    template <int NO1 = R1>
    GD_HD static void i2(float2* x, int stride, int n2) {
        float2 v[R1];
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) v[k1] = x[(R2 * k1 + n2) * stride];
        Dft<R1>::run(v);
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1)
            if (n1 < NO1) x[(R2 * n1 + n2) * stride] = cswap(v[n1]);
    }
};

// ---- 2-D real transform of an NIN x NIN real image zero-padded to N x N ----
// Buffers (shared memory):  Z[NIN/2][N]  packed row pairs,  S[N][SP]  half spectrum, SP >= N/2+1.
// S[s1][k2] holds the spectrum at (k1 = freq(s1), k2), k2 natural in [0, N/2].
template <int N, int R1, int R2, int NIN> struct Real2D {
    using L = Line<N, R1, R2>;
    static constexpr int NH = N / 2 + 1;
    static constexpr int ZL = NIN / 2;
    static constexpr int NZ1 = (NIN + R2 - 1) / R2;      // non-zero n1 groups when only n < NIN is non-zero
    static_assert(NIN % 2 == 0 && NIN <= N && (NIN % R2 == 0 || NIN == N), "prune needs NIN multiple of R2");

    // forward phases -------------------------------------------------------------------------
    // (caller) F0: Z[j][c] = (img[2j][c], img[2j+1][c]) for c < NIN; c >= NIN is never read.
    static constexpr int F1_ITEMS = ZL * R2;
    GD_HD static void F1(int w, float2* Z, const float2* tw) { L::template f1<NZ1>(Z + (w / R2) * N, 1, w % R2, tw); }
    static constexpr int F2_ITEMS = ZL * R1;
    GD_HD static void F2(int w, float2* Z) { L::f2(Z + (w / R1) * N, 1, w % R1); }
    // unpack: item (j, k2): A = row 2j, B = row 2j+1
    static constexpr int F3_ITEMS = ZL * NH;
    GD_HD static void F3(int w, const float2* Z, float2* S, int SP) {
        int j = w / NH, k2 = w % NH;
        float2 za = Z[j * N + L::slot(k2)], zb = Z[j * N + L::slot((N - k2) % N)];
        float2 a = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
        float2 d = make_float2(0.5f * (za.x - zb.x), 0.5f * (za.y + zb.y));   // (za - conj zb)/2
        S[(2 * j) * SP + k2] = a;
        S[(2 * j + 1) * SP + k2] = make_float2(d.y, -d.x);                    // -i * d
    }
    static constexpr int F4_ITEMS = NH * R2;
    GD_HD static void F4(int w, float2* S, int SP, const float2* tw) { L::template f1<NZ1>(S + (w % NH), SP, w / NH, tw); }
    static constexpr int F5_ITEMS = NH * R1;
    GD_HD static void F5(int w, float2* S, int SP) { L::f2(S + (w % NH), SP, w / NH); }

    // inverse phases (only the NIN x NIN corner of the output is produced) ---------------------
    static constexpr int I1_ITEMS = NH * R1;
    GD_HD static void I1(int w, float2* S, int SP, const float2* tw) { L::i1(S + (w % NH), SP, w / NH, tw); }
    static constexpr int I2_ITEMS = NH * R2;
    GD_HD static void I2(int w, float2* S, int SP) { L::template i2<NZ1>(S + (w % NH), SP, w / NH); }
    // pack rows (2j, 2j+1) of the row spectra into one Hermitian-extended complex line, digit-swapped
    static constexpr int I3_ITEMS = ZL * N;
    GD_HD static void I3(int w, const float2* S, int SP, float2* Z) {
        int j = w / N, k2 = w % N;
        float2 a, b;
        if (k2 < NH) { a = S[(2 * j) * SP + k2]; b = S[(2 * j + 1) * SP + k2]; }
        else { a = cconj(S[(2 * j) * SP + (N - k2)]); b = cconj(S[(2 * j + 1) * SP + (N - k2)]); }
        Z[j * N + L::slot(k2)] = make_float2(a.x - b.y, a.y + b.x);           // a + i b
    }
    static constexpr int I4_ITEMS = ZL * R1;
    GD_HD static void I4(int w, float2* Z, const float2* tw) { L::i1(Z + (w / R1) * N, 1, w % R1, tw); }
    static constexpr int I5_ITEMS = ZL * R2;
    GD_HD static void I5(int w, float2* Z) { L::template i2<NZ1>(Z + (w / R2) * N, 1, w % R2); }
    // (caller) I6: img[2j][c] = Z[j][c].x, img[2j+1][c] = Z[j][c].y for c < NIN (unscaled by 1/N^2).
};

}  // namespace gdfft
