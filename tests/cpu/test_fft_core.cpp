// CPU unit test of galaxy-deconv_b200/csrc/fft_core.cuh: the kernels' FFT phases run here with a plain
// loop over work items and are compared with a naive O(N^4)-free separable DFT in double precision.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include "fft_core.cuh"
using namespace gdfft;
typedef std::complex<double> cd;

template <int N, int R1, int R2, int NIN> static double run_case() {
    using F = Real2D<N, R1, R2, NIN>;
    const int SP = F::NH + 1;
    std::vector<float2> tw(N), Z(F::ZL * N), S(N * SP);
    for (int m = 0; m < N; ++m) tw[m] = make_float2((float)cos(2 * M_PI * m / N), (float)-sin(2 * M_PI * m / N));
    std::vector<float> img(NIN * NIN);
    srand(123 + N);
    for (auto& v : img) v = (float)rand() / RAND_MAX - 0.3f;
    // garbage everywhere first: phases must never read what they do not own
    for (auto& v : Z) v = make_float2(NAN, NAN);
    for (auto& v : S) v = make_float2(NAN, NAN);
    for (int j = 0; j < F::ZL; ++j) for (int c = 0; c < NIN; ++c) Z[j * N + c] = make_float2(img[2 * j * NIN + c], img[(2 * j + 1) * NIN + c]);
    for (int w = 0; w < F::F1_ITEMS; ++w) F::F1(w, Z.data(), tw.data());
    for (int w = 0; w < F::F2_ITEMS; ++w) F::F2(w, Z.data());
    for (int w = 0; w < F::F3_ITEMS; ++w) F::F3(w, Z.data(), S.data(), SP);
    for (int w = 0; w < F::F4_ITEMS; ++w) F::F4(w, S.data(), SP, tw.data());
    for (int w = 0; w < F::F5_ITEMS; ++w) F::F5(w, S.data(), SP);
    // reference spectrum (separable, double)
    std::vector<cd> rowf(NIN * N), ref(N * F::NH);
    for (int r = 0; r < NIN; ++r) for (int k = 0; k < N; ++k) { cd s = 0; for (int c = 0; c < NIN; ++c) s += (double)img[r * NIN + c] * std::polar(1.0, -2 * M_PI * k * c / N); rowf[r * N + k] = s; }
    double maxref = 0, err = 0;
    for (int k1 = 0; k1 < N; ++k1) for (int k2 = 0; k2 < F::NH; ++k2) { cd s = 0; for (int r = 0; r < NIN; ++r) s += rowf[r * N + k2] * std::polar(1.0, -2 * M_PI * k1 * r / N); ref[k1 * F::NH + k2] = s; maxref = std::max(maxref, std::abs(s)); }
    for (int s1 = 0; s1 < N; ++s1) for (int k2 = 0; k2 < F::NH; ++k2) {
        float2 g = S[s1 * SP + k2]; cd e = ref[F::L::freq(s1) * F::NH + k2];
        err = std::max(err, std::abs(cd(g.x, g.y) - e));
    }
    double fwd = err / maxref;
    // inverse of the forward must give back the image (times N*N)
    for (int w = 0; w < F::I1_ITEMS; ++w) F::I1(w, S.data(), SP, tw.data());
    for (int w = 0; w < F::I2_ITEMS; ++w) F::I2(w, S.data(), SP);
    for (auto& v : Z) v = make_float2(NAN, NAN);
    for (int w = 0; w < F::I3_ITEMS; ++w) F::I3(w, S.data(), SP, Z.data());
    for (int w = 0; w < F::I4_ITEMS; ++w) F::I4(w, Z.data(), tw.data());
    for (int w = 0; w < F::I5_ITEMS; ++w) F::I5(w, Z.data());
    double inv = 0;
    for (int j = 0; j < F::ZL; ++j) for (int c = 0; c < NIN; ++c) {
        inv = std::max(inv, (double)fabs(Z[j * N + c].x / (N * N) - img[2 * j * NIN + c]));
        inv = std::max(inv, (double)fabs(Z[j * N + c].y / (N * N) - img[(2 * j + 1) * NIN + c]));
    }
    printf("N=%d fwd_rel_err=%.3e inv_abs_err=%.3e\n", N, fwd, inv);
    return (std::isfinite(fwd) && std::isfinite(inv)) ? std::max(fwd, inv) : 1.0;
}

int main() {
    double e = 0;
    e = std::max(e, run_case<96, 8, 12, 48>());
    e = std::max(e, run_case<48, 4, 12, 48>());
    e = std::max(e, run_case<128, 8, 16, 48>());
    // codelets against the definition
    {
        float2 v[16]; cd x[16];
        for (int R : {2, 3, 4, 6, 8, 12, 16}) {
            for (int i = 0; i < R; ++i) { x[i] = cd(sin(i * 1.3 + R), cos(i * 0.7)); v[i] = make_float2((float)x[i].real(), (float)x[i].imag()); }
            switch (R) { case 2: Dft<2>::run(v); break; case 3: Dft<3>::run(v); break; case 4: Dft<4>::run(v); break; case 6: Dft<6>::run(v); break;
                         case 8: Dft<8>::run(v); break; case 12: Dft<12>::run(v); break; default: Dft<16>::run(v); }
            double m = 0;
            for (int k = 0; k < R; ++k) { cd s = 0; for (int n = 0; n < R; ++n) s += x[n] * std::polar(1.0, -2 * M_PI * n * k / R); m = std::max(m, std::abs(s - cd(v[k].x, v[k].y))); }
            printf("codelet R=%d err=%.3e\n", R, m);
            e = std::max(e, m / 4);
        }
    }
    {   // the whole 48-point transform of one thread (k_wiener48), forward and through the conjugation identity backward
        float2 v[48]; cd x[48];
        for (int n = 0; n < 48; ++n) { x[n] = cd(sin(n * 0.9 + 0.2), cos(n * 1.7) - 0.1); v[n] = make_float2((float)x[n].real(), (float)x[n].imag()); }
        Fft48::run(v);
        double m = 0, mx = 0;
        cd X[48];
        for (int k = 0; k < 48; ++k) {
            cd s = 0;
            for (int n = 0; n < 48; ++n) s += x[n] * std::polar(1.0, -2 * M_PI * n * k / 48);
            X[k] = s; mx = std::max(mx, std::abs(s));
            m = std::max(m, std::abs(s - cd(v[Fft48::reg(k)].x, v[Fft48::reg(k)].y)));
        }
        float2 w[48];
        for (int k = 0; k < 48; ++k) w[k] = make_float2(v[Fft48::reg(k)].x, -v[Fft48::reg(k)].y);
        Fft48::run(w);
        double mi = 0;
        for (int n = 0; n < 48; ++n) mi = std::max(mi, std::abs(cd(w[Fft48::reg(n)].x, -w[Fft48::reg(n)].y) / 48.0 - x[n]));
        printf("Fft48 fwd_rel_err=%.3e inv_abs_err=%.3e\n", m / mx, mi);
        e = std::max(e, std::max(m / mx, mi));
    }
    if (e > 5e-6) { printf("FAIL %.3e\n", e); return 1; }
    printf("OK\n");
    return 0;
}
