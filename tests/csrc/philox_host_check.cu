// Host-side check of the device RNG: philox4x32_10 in csrc/common.cuh is __host__ __device__, so the very code the
// dropout / teacher-noise kernels call is run on the CPU against the Random123 known-answer vectors and printed for a few
// counters that tests/test_cabi.py compares with the numpy restatement (oracle/philox.py).
#include <cstdio>
#include "../../dcase2019_task4_b200/csrc/common.cuh"

static int check(uint64_t row, uint32_t stream, uint32_t step, uint64_t seed, uint32_t e0, uint32_t e1, uint32_t e2,
                 uint32_t e3) {
    const uint4 r = philox4x32_10(row, stream, step, seed);
    return r.x == e0 && r.y == e1 && r.z == e2 && r.w == e3 ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += check(0ull, 0u, 0u, 0ull, 0x6627e8d5u, 0xe169c58du, 0xbc57ac4cu, 0x9b00dbd8u);
    bad += check(0xffffffffffffffffull, 0xffffffffu, 0xffffffffu, 0xffffffffffffffffull, 0x408f276du, 0x41c83b0eu,
                 0xa20bc7c6u, 0x6d5451fdu);
    bad += check(0x85a308d3243f6a88ull, 0x13198a2eu, 0x03707344u, 0x299f31d0a4093822ull, 0xd16cfe09u, 0x94fdccebu,
                 0x5001e420u, 0x24126ea1u);
    printf("kat_failures %d\n", bad);
    const uint64_t rows[3] = {0ull, 123456789ull, (1ull << 33) + 5ull};
    for (int i = 0; i < 3; ++i) {
        const uint4 r = philox4x32_10(rows[i], 8u * 1u + 2u, 41u + i, 0x0123456789abcdefull);
        printf("%llu %u %u %u %u\n", (unsigned long long)rows[i], r.x, r.y, r.z, r.w);
    }
    return bad;
}
