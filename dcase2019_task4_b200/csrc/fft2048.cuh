// 2048-point complex Stockham FFT building blocks (radix 8,8,8,4), shared between the
// log-mel kernel and the host-side self check (tests/csrc/fft_host_check.cu).
#pragma once
#include <cuda_runtime.h>

struct cf32 {
    float x, y;
};

__host__ __device__ __forceinline__ cf32 cmul(cf32 a, cf32 b) {
    return cf32{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
__host__ __device__ __forceinline__ cf32 cadd(cf32 a, cf32 b) { return cf32{a.x + b.x, a.y + b.y}; }
__host__ __device__ __forceinline__ cf32 csub(cf32 a, cf32 b) { return cf32{a.x - b.x, a.y - b.y}; }
// multiply by -i :  (x + iy)(-i) = y - ix
__host__ __device__ __forceinline__ cf32 cmul_mi(cf32 a) { return cf32{a.y, -a.x}; }

// 4-point forward DFT, natural order in and out.
__host__ __device__ __forceinline__ void fft4(cf32* v) {
    const cf32 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
    const cf32 a2 = cadd(v[1], v[3]), a3 = cmul_mi(csub(v[1], v[3]));
    v[0] = cadd(a0, a2);
    v[2] = csub(a0, a2);
    v[1] = cadd(a1, a3);
    v[3] = csub(a1, a3);
}

// 8-point forward DFT (decimation in frequency), natural order in and out.
__host__ __device__ __forceinline__ void fft8(cf32* v) {
    const float h = 0.70710678118654752440f;
    cf32 b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        b[i] = cadd(v[i], v[i + 4]);
        b[i + 4] = csub(v[i], v[i + 4]);
    }
    // twiddles W8^i on the lower half: W8^1 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
    b[5] = cf32{(b[5].x + b[5].y) * h, (b[5].y - b[5].x) * h};
    b[6] = cmul_mi(b[6]);
    b[7] = cf32{(b[7].y - b[7].x) * h, (-b[7].x - b[7].y) * h};
    // two 4-point DFTs: even outputs from b[0..3], odd outputs from b[4..7]
    fft4(b);
    fft4(b + 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = b[i];
        v[2 * i + 1] = b[4 + i];
    }
}

// Shared-memory index padding: one extra slot per 8 complex values makes the stride-8 writes of the first two
// passes (index 8 j + r across the lanes j) conflict-free while stride-1 accesses stay conflict-free.
__host__ __device__ __forceinline__ int fft_pad(int i) { return i + (i >> 3); }
constexpr int kFftPaddedSize = 2048 + 256;

// One Stockham pass of radix R for "thread" j (0 <= j < 2048 / R). Ns = product of the radices
// of the passes already done. tw[m] = exp(-2 pi i m / 2048).
template <int R>
__host__ __device__ __forceinline__ void stockham_pass(int j, int Ns, const cf32* __restrict__ src,
                                                       cf32* __restrict__ dst, const cf32* __restrict__ tw) {
    constexpr int N = 2048;
    cf32 v[R];
    const int k = j & (Ns - 1);
    const int tw_stride = N / (Ns * R);
    // The R - 1 twiddles of this butterfly are powers of w1 = tw[k * tw_stride]: one table load plus complex
    // multiplications on the (idle) FMA pipe instead of R - 1 loads on the load / store path that bounds the kernel.
    // Each power is at most 3 multiplications away from the table value (error ~3 ulp).
    cf32 w[R];
    w[1] = tw[k * tw_stride];
    w[2] = cmul(w[1], w[1]);
    w[3] = cmul(w[2], w[1]);
    if (R == 8) {
        w[4] = cmul(w[2], w[2]);
        w[5] = cmul(w[4], w[1]);
        w[6] = cmul(w[3], w[3]);
        w[7] = cmul(w[4], w[3]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        v[r] = src[fft_pad(j + r * (N / R))];
        if (r > 0) v[r] = cmul(v[r], w[r]);
    }
    if (R == 8) fft8(v); else fft4(v);
    const int d = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) dst[fft_pad(d + r * Ns)] = v[r];
}
