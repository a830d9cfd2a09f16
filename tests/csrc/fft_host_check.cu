// Host-side check of the warp-autonomous real FFT used by the log-mel kernel (csrc/fft1024.cuh): the very same
// __host__ __device__ code (32-point DFTs, pass-1 twiddle tree, 32 x 32 transpose indexing, Hermitian split with the
// partner taken from lane (32 - lane) % 32) is run lane by lane on the CPU against a naive float64 DFT of one windowed
// real 2048-sample frame, for an even and an odd frame offset of the de-interleaved span (hop 511 is odd).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../dcase2019_task4_b200/csrc/fft1024.cuh"

static double check_frame(const std::vector<float>& span, int o, const std::vector<float>& win, double* peak) {
    const int N = 2048;
    const int half = (int)span.size() / 2;
    std::vector<float> even(half + 1), odd(half + 1);
    for (size_t i = 0; i < span.size(); ++i) (i & 1 ? odd : even)[i >> 1] = span[i];
    const float* pre = (o & 1) ? odd.data() + (o >> 1) : even.data() + (o >> 1);
    const float* pim = (o & 1) ? even.data() + ((o + 1) >> 1) : odd.data() + (o >> 1);
    std::vector<cpx> buf(kXchgSize);
    std::vector<std::vector<cpx>> Z(32, std::vector<cpx>(32));
    for (int lane = 0; lane < 32; ++lane) {
        Pass1Twiddles tw;
        auto W = [&](int j) { double a = -2.0 * M_PI * lane * j / 1024.0; return cmake((float)cos(a), (float)sin(a)); };
        tw.w1 = W(1); tw.w2 = W(2); tw.w4 = W(4); tw.w8 = W(8); tw.w16 = W(16);
        cpx v[32];
        for (int n1 = 0; n1 < 32; ++n1) {
            const int n = 32 * n1 + lane;
            v[n1] = cmul_elem(cmake(pre[n], pim[n]), cmake(win[2 * n], win[2 * n + 1]));
        }
        stft_pass1(v, tw, lane, buf.data());
    }
    for (int lane = 0; lane < 32; ++lane) stft_pass2(Z[lane].data(), lane, buf.data());
    std::vector<double> mag(1025, -1.0);
    const float c64[16] = DCASE_W64_COS, s64[16] = DCASE_W64_SIN;
    for (int lane = 0; lane < 32; ++lane) {
        const double a = -2.0 * M_PI * lane / 2048.0;
        const cpx wl = cmake((float)cos(a), (float)sin(a));
        const int src = (32 - lane) & 31;
        for (int p = 0; p < 16; ++p) {
            const cpx zm = lane == 0 ? Z[0][(32 - p) & 31] : Z[src][31 - p];
            float lo, hi;
            stft_post_pair(Z[lane][p], zm, wl, c64[p], s64[p], lo, hi);
            mag[lane + 32 * p] = 0.5 * sqrt((double)lo);
            mag[1024 - lane - 32 * p] = 0.5 * sqrt((double)hi);
        }
        if (lane == 0) mag[512] = hypot(cre(Z[0][16]), cim(Z[0][16]));
    }
    double maxerr = 0;
    for (int k = 0; k <= 1024; ++k) {
        double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            const double x = (double)span[o + n] * (double)win[n];
            const double ang = -2.0 * M_PI * (double)((long)k * n % N) / N;
            sr += x * cos(ang);
            si += x * sin(ang);
        }
        const double ref = hypot(sr, si);
        if (mag[k] < 0) return 1e30;                      // a bin nobody wrote
        maxerr = fmax(maxerr, fabs(ref - mag[k]));
        *peak = fmax(*peak, ref);
    }
    return maxerr;
}

int main() {
    std::vector<float> win(2048);
    for (int n = 0; n < 2048; ++n) win[n] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * n / 2047.0));
    std::vector<float> span(7 * 511 + 2048 + 1);
    srand(1);
    for (auto& x : span) x = (float)(rand() / (double)RAND_MAX - 0.5);
    for (int i = 0; i < 400; ++i) span[100 + i] += (float)(0.8 * sin(0.3 * i));   // a tone burst: a dominant bin
    double peak = 0, err = 0;
    const int offsets[4] = {0, 511, 1022, 7 * 511};
    for (int o : offsets) err = fmax(err, check_frame(span, o, win, &peak));
    // a unit-test of the 32-point DFT alone
    cpx v[32];
    double xr[32], xi[32];
    for (int i = 0; i < 32; ++i) { xr[i] = rand() / (double)RAND_MAX - 0.5; xi[i] = rand() / (double)RAND_MAX - 0.5; v[i] = cmake((float)xr[i], (float)xi[i]); }
    dft32(v);
    double e32 = 0;
    for (int k = 0; k < 32; ++k) {
        double sr = 0, si = 0;
        for (int n = 0; n < 32; ++n) {
            const double ang = -2.0 * M_PI * (k * n % 32) / 32.0;
            sr += xr[n] * cos(ang) - xi[n] * sin(ang);
            si += xr[n] * sin(ang) + xi[n] * cos(ang);
        }
        e32 = fmax(e32, hypot(sr - cre(v[k]), si - cim(v[k])));
    }
    printf("max_abs_err %.3e max_mag %.3e dft32_err %.3e\n", err, peak, e32);
    return (err < 2e-6 * peak && e32 < 1e-5) ? 0 : 1;
}
