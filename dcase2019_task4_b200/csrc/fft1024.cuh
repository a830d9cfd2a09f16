// Warp-autonomous real FFT of one 2048-sample frame (baseline/DatasetDcase2019Task4.py:209-218: |librosa.stft| with
// n_fft = 2048) as ONE 1024-point complex FFT per warp, shared between the log-mel kernel (csrc/logmel.cu) and the
// host-side self check (tests/csrc/fft_host_check.cu compiles this very code for the CPU, lane by lane).
//
//   z[n] = xw[2n] + i xw[2n+1]              (xw = frame x Hamming window), n = 32 n1 + n2
//   pass 1 (lane = n2):  Y[k1] = DFT32 over n1 of z[32 n1 + lane];  times W_1024^(lane k1);  buf[k1][lane]   (smem)
//   pass 2 (lane = k1):  Z[lane + 32 k2] = DFT32 over n2 of buf[lane][n2]                                     (registers)
//   post   (k = lane + 32 p, p < 16):  A = Z[k] + conj Z[1024-k],  D = Z[k] - conj Z[1024-k],  T = W_2048^k (-i D)
//                        |X[k]| = |A + T| / 2,   |X[1024-k]| = |A - T| / 2     (Z[1024-k] comes from lane (32-lane)%32)
//
// Both 32-point DFTs live entirely in registers (4 x radix-8, constant twiddles, 8 x radix-4); the only traffic through
// shared memory is the 32 x 32 transpose between the passes.  No block barrier anywhere: a warp owns its frame.
//
// Complex values are `cpx`.  On the device a cpx is one 64-bit register pair and additions / subtractions / the window
// product are single packed instructions (add / sub / mul .f32x2 -> FADD2 / FMUL2 on sm_100); multiplications by
// twiddles and the -i rotations work on the two halves.  On the host it is a plain struct.
#pragma once
#include <cuda_runtime.h>

#if defined(__CUDA_ARCH__) && !defined(DCASE_NO_F32X2)
#define DCASE_PACKED_CPX 1
#else
#define DCASE_PACKED_CPX 0
#endif

#if DCASE_PACKED_CPX
struct cpx {
    unsigned long long v;
};
__device__ __forceinline__ cpx cmake(float x, float y) {
    cpx r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float cre(cpx a) {
    float x;
    asm("{ .reg .b32 hi; mov.b64 {%0, hi}, %1; }" : "=f"(x) : "l"(a.v));
    return x;
}
__device__ __forceinline__ float cim(cpx a) {
    float y;
    asm("{ .reg .b32 lo; mov.b64 {lo, %0}, %1; }" : "=f"(y) : "l"(a.v));
    return y;
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) {
    cpx r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ cpx csub(cpx a, cpx b) {
    cpx r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
// element-wise product (re * re, im * im): the window
__device__ __forceinline__ cpx cmul_elem(cpx a, cpx b) {
    cpx r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
// a + s * b element-wise with s = (sx, sy): conjugations folded into one FFMA2
__device__ __forceinline__ cpx cfma_elem(cpx b, cpx s, cpx a) {
    cpx r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(b.v), "l"(s.v), "l"(a.v));
    return r;
}
#else
struct cpx {
    float x, y;
};
__host__ __device__ __forceinline__ cpx cmake(float x, float y) { return cpx{x, y}; }
__host__ __device__ __forceinline__ float cre(cpx a) { return a.x; }
__host__ __device__ __forceinline__ float cim(cpx a) { return a.y; }
__host__ __device__ __forceinline__ cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
__host__ __device__ __forceinline__ cpx csub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
__host__ __device__ __forceinline__ cpx cmul_elem(cpx a, cpx b) { return cpx{a.x * b.x, a.y * b.y}; }
__host__ __device__ __forceinline__ cpx cfma_elem(cpx b, cpx s, cpx a) { return cpx{a.x + s.x * b.x, a.y + s.y * b.y}; }
#endif

// (a.x + i a.y)(c + i s)
__host__ __device__ __forceinline__ cpx cmul_cs(cpx a, float c, float s) {
    const float x = cre(a), y = cim(a);
    return cmake(x * c - y * s, x * s + y * c);
}
__host__ __device__ __forceinline__ cpx cmulc(cpx a, cpx b) { return cmul_cs(a, cre(b), cim(b)); }
// a + (-i) b  and  a - (-i) b  :  (-i)(x + iy) = y - ix
__host__ __device__ __forceinline__ cpx cadd_mi(cpx a, cpx b) { return cmake(cre(a) + cim(b), cim(a) - cre(b)); }
__host__ __device__ __forceinline__ cpx csub_mi(cpx a, cpx b) { return cmake(cre(a) - cim(b), cim(a) + cre(b)); }

// 4-point forward DFT, natural order in and out
__host__ __device__ __forceinline__ void dft4(cpx& v0, cpx& v1, cpx& v2, cpx& v3) {
    const cpx a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), d = csub(v1, v3);
    v0 = cadd(a0, a2);
    v2 = csub(a0, a2);
    v1 = cadd_mi(a1, d);
    v3 = csub_mi(a1, d);
}

// 8-point forward DFT (decimation in frequency), natural order in and out
__host__ __device__ __forceinline__ void dft8(cpx* v) {
    const float h = 0.70710678118654752440f;
    cpx b0 = cadd(v[0], v[4]), b1 = cadd(v[1], v[5]), b2 = cadd(v[2], v[6]), b3 = cadd(v[3], v[7]);
    cpx b4 = csub(v[0], v[4]), b5 = csub(v[1], v[5]), b6 = csub(v[2], v[6]), b7 = csub(v[3], v[7]);
    // twiddles W8^1 = (1 - i) / sqrt2, W8^2 = -i, W8^3 = (-1 - i) / sqrt2 on the lower half
    b5 = cmake((cre(b5) + cim(b5)) * h, (cim(b5) - cre(b5)) * h);
    b6 = cmake(cim(b6), -cre(b6));
    b7 = cmake((cim(b7) - cre(b7)) * h, -(cre(b7) + cim(b7)) * h);
    dft4(b0, b1, b2, b3);      // even outputs 0, 2, 4, 6
    dft4(b4, b5, b6, b7);      // odd outputs 1, 3, 5, 7
    v[0] = b0; v[2] = b1; v[4] = b2; v[6] = b3;
    v[1] = b4; v[3] = b5; v[5] = b6; v[7] = b7;
}

// cos / sin of 2 pi m / 32, m = 0 .. 21 (the products n2 * k1 of the 4 x 8 decomposition, n2 <= 3, k1 <= 7)
#define DCASE_W32_COS {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,                \
                       0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,                      \
                       0.19509032201612826785f, 0.0f, -0.19509032201612826785f, -0.38268343236508977173f,              \
                       -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,                   \
                       -0.92387953251128675613f, -0.98078528040323044913f, -1.0f, -0.98078528040323044913f,            \
                       -0.92387953251128675613f, -0.83146961230254523708f, -0.70710678118654752440f,                   \
                       -0.55557023301960222474f}
#define DCASE_W32_SIN {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,                 \
                       0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,                      \
                       0.98078528040323044913f, 1.0f, 0.98078528040323044913f, 0.92387953251128675613f,                \
                       0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,                      \
                       0.38268343236508977173f, 0.19509032201612826785f, 0.0f, -0.19509032201612826785f,               \
                       -0.38268343236508977173f, -0.55557023301960222474f, -0.70710678118654752440f,                   \
                       -0.83146961230254523708f}

// 32-point forward DFT in registers, natural order in and out (fully unrolled: every index is a compile-time constant).
//   n = 4 a + b:  U_b[k1] = DFT8 over a of v[4 a + b];  U_b[k1] *= W_32^(b k1);  X[k1 + 8 k2] = DFT4 over b of U_b[k1]
__host__ __device__ __forceinline__ void dft32(cpx* v) {
    constexpr float kc[22] = DCASE_W32_COS;
    constexpr float ks[22] = DCASE_W32_SIN;
    cpx u[4][8];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
        for (int a = 0; a < 8; ++a) u[b][a] = v[4 * a + b];
        dft8(u[b]);
    }
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
#pragma unroll
        for (int b = 1; b < 4; ++b) {
            const int m = b * k1;                                 // W_32^m = cos - i sin
            if (m == 0) continue;
            if (m == 8) u[b][k1] = cmake(cim(u[b][k1]), -cre(u[b][k1]));
            else u[b][k1] = cmul_cs(u[b][k1], kc[m], -ks[m]);
        }
        dft4(u[0][k1], u[1][k1], u[2][k1], u[3][k1]);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 8 * k2] = u[k2][k1];
    }
}

constexpr int kXchgPitch = 33;                        // complex values per row of the 32 x 32 transpose buffer
constexpr int kXchgSize = 32 * kXchgPitch;            // complex values per warp (8448 bytes)

// W_1024^(lane * j) for j = 1, 2, 4, 8, 16 (per lane, constant across frames), from which the 31 pass-1 twiddles of a
// lane are at most 4 complex products away.
struct Pass1Twiddles {
    cpx w1, w2, w4, w8, w16;
};

// pass 1 for one lane: v[n1] = window x frame at n = 32 n1 + lane already loaded by the caller
__host__ __device__ __forceinline__ void stft_pass1(cpx* v, const Pass1Twiddles& tw, int lane, cpx* buf) {
    dft32(v);
    // twiddle W_1024^(lane k1): powers built by a product tree over the bits of k1 (depth <= 4)
    cpx t[32];
    t[1] = tw.w1; t[2] = tw.w2; t[4] = tw.w4; t[8] = tw.w8; t[16] = tw.w16;
#pragma unroll
    for (int k1 = 3; k1 < 32; ++k1) {
        const int low = k1 & (-k1);                              // lowest set bit
        if (low == k1) continue;                                  // a pure power of two: given
        t[k1] = cmulc(t[k1 - low], t[low]);
    }
    buf[0 * kXchgPitch + lane] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) buf[k1 * kXchgPitch + lane] = cmulc(v[k1], t[k1]);
}

// pass 2 for one lane (= k1): v[k2] = Z[lane + 32 k2]
__host__ __device__ __forceinline__ void stft_pass2(cpx* v, int lane, const cpx* buf) {
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[n2] = buf[lane * kXchgPitch + n2];
    dft32(v);
}

// cos / sin of 2 pi p / 64, p = 0 .. 15:  W_2048^(lane + 32 p) = W_2048^lane * W_64^p
#define DCASE_W64_COS {1.0f, 0.99518472667219688624f, 0.98078528040323044913f, 0.95694033573220886494f,                \
                       0.92387953251128675613f, 0.88192126434835502971f, 0.83146961230254523708f,                      \
                       0.77301045336273696081f, 0.70710678118654752440f, 0.63439328416364549822f,                      \
                       0.55557023301960222474f, 0.47139673682599764856f, 0.38268343236508977173f,                      \
                       0.29028467725446236764f, 0.19509032201612826785f, 0.09801714032956060199f}
#define DCASE_W64_SIN {0.0f, 0.09801714032956060199f, 0.19509032201612826785f, 0.29028467725446236764f,                 \
                       0.38268343236508977173f, 0.47139673682599764856f, 0.55557023301960222474f,                      \
                       0.63439328416364549822f, 0.70710678118654752440f, 0.77301045336273696081f,                      \
                       0.83146961230254523708f, 0.88192126434835502971f, 0.92387953251128675613f,                      \
                       0.95694033573220886494f, 0.98078528040323044913f, 0.99518472667219688624f}

// squared magnitudes (x 4) of X[k] and X[1024 - k] for k = lane + 32 p from Zk = Z[k] and Zm = Z[1024 - k]:
// returns |A + T|^2 = |2 X[k]|^2 in `lo` and |A - T|^2 = |2 X[1024 - k]|^2 in `hi` (the kernel keeps the factor 2 and
// halves the mel weights instead, which is exact).  wl = W_2048^lane.
__host__ __device__ __forceinline__ void stft_post_pair(cpx zk, cpx zm, cpx wl, float c64, float s64, float& lo,
                                                        float& hi) {
    const cpx a = cfma_elem(zm, cmake(1.f, -1.f), zk);           // Zk + conj(Zm)
    const cpx d = cfma_elem(zm, cmake(-1.f, 1.f), zk);           // Zk - conj(Zm)
    const cpx o = cmake(cim(d), -cre(d));                        // -i D
    const cpx t = cmulc(cmul_cs(o, c64, -s64), wl);              // W_2048^k (-i D)
    const cpx p = cadd(a, t), q = csub(a, t);
    lo = cre(p) * cre(p) + cim(p) * cim(p);
    hi = cre(q) * cre(q) + cim(q) * cim(q);
}
