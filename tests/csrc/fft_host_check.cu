// Host-side check of the Stockham passes used by the log-mel kernel: runs the same
// __host__ __device__ code on the CPU against a naive float64 DFT.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../dcase2019_task4_b200/csrc/fft2048.cuh"

int main() {
    const int N = 2048;
    std::vector<cf32> tw(N), a(kFftPaddedSize), b(kFftPaddedSize);
    for (int m = 0; m < N; ++m) {
        double ang = -2.0 * M_PI * m / N;
        tw[m] = cf32{(float)cos(ang), (float)sin(ang)};
    }
    std::vector<double> xr(N), xi(N);
    srand(1);
    for (int i = 0; i < N; ++i) {
        xr[i] = rand() / (double)RAND_MAX - 0.5;
        xi[i] = rand() / (double)RAND_MAX - 0.5;
        a[fft_pad(i)] = cf32{(float)xr[i], (float)xi[i]};
    }
    for (int j = 0; j < 256; ++j) stockham_pass<8>(j, 1, a.data(), b.data(), tw.data());
    for (int j = 0; j < 256; ++j) stockham_pass<8>(j, 8, b.data(), a.data(), tw.data());
    for (int j = 0; j < 256; ++j) stockham_pass<8>(j, 64, a.data(), b.data(), tw.data());
    for (int j = 0; j < 512; ++j) stockham_pass<4>(j, 512, b.data(), a.data(), tw.data());
    double maxerr = 0, maxmag = 0;
    for (int k = 0; k < N; k += 7) {
        double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            double ang = -2.0 * M_PI * (double)((long)k * n % N) / N;
            sr += xr[n] * cos(ang) - xi[n] * sin(ang);
            si += xr[n] * sin(ang) + xi[n] * cos(ang);
        }
        maxerr = fmax(maxerr, hypot(sr - a[fft_pad(k)].x, si - a[fft_pad(k)].y));
        maxmag = fmax(maxmag, hypot(sr, si));
    }
    printf("max_abs_err %.3e max_mag %.3e\n", maxerr, maxmag);
    return maxerr < 1e-3 * maxmag / 10 ? 0 : 1;
}
