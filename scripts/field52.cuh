// field52.cuh -- BN254 Fq arithmetic on 5 x 52-bit limbs, multiplied on the FP64 pipe.
//
// Why: on B200 the FP64 pipe issues ~64 DFMA/clk/SM (measured: scripts/pipe_probe.cu), the same rate
// as 32-bit IMAD, but a DFMA pair yields a 52x52 -> 104-bit product where IMAD.WIDE (4 pipe cycles)
// yields 32x32 -> 64.  A 254-bit Montgomery product is 50 limb products here against 128 there.
//
// Product of two 52-bit integers held in doubles (a*b < 2^104):
//     hi = fma_rz(a, b, 2^104)                  mantissa field of hi = floor(a*b / 2^52)
//     lo = fma_rz(a, b, (2^104 + 2^52) - hi)    mantissa field of lo = a*b mod 2^52     (both exact)
// The bit patterns are accumulated as 64-bit integers per column; the exponent fields are multiples
// of 2^52 known at compile time per column and are cancelled by the accumulator's initial value.
//
// Representation: N52 = five 64-bit limbs, each < 2^52 ("normalised"), value < 2^260, read modulo q.
// The Montgomery radix here is 2^260.  Values are NOT kept canonical: a product of inputs below
// Ba*2^254 and Bb*2^254 is below (Ba*Bb/64 + 0.7562)*2^254, so sums and differences of a few
// products feed the next product unreduced (the bounds are stated where they are used, msm52.cuh).
//
// The same code compiles for the host (std::fma under FE_TOWARDZERO) so that tests/test_field52.py
// checks every routine against big-integer arithmetic without a GPU.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDA_ARCH__)
#define F52_DI __device__ __forceinline__
#define F52_FMA_RZ(a, b, c) __fma_rz((a), (b), (c))
#define F52_D2L(x) __double_as_longlong(x)
#define F52_L2D(x) __longlong_as_double(x)
#else
#include <cmath>
#if defined(__CUDACC__)
#define F52_DI __host__ __device__ inline
#else
#define F52_DI inline
#endif
// host: the caller runs under fesetround(FE_TOWARDZERO); every other sum formed here is exact
#define F52_FMA_RZ(a, b, c) std::fma((a), (b), (c))
static inline long long f52_d2l(double x) {
    long long r;
    std::memcpy(&r, &x, 8);
    return r;
}
static inline double f52_l2d(long long x) {
    double r;
    std::memcpy(&r, &x, 8);
    return r;
}
#define F52_D2L(x) f52_d2l(x)
#define F52_L2D(x) f52_l2d(x)
#endif

namespace h2b {

struct N52 {
    uint64_t l[5];  // normalised: every limb < 2^52
};
struct D52 {
    double l[5];  // the same limbs as exact doubles (a multiplication operand)
};

constexpr uint64_t kMask52 = (1ull << 52) - 1;
constexpr uint64_t kBitsC1 = 0x4670000000000000ull;  // bit pattern of 2^104
constexpr uint64_t kBitsC2 = 0x4330000000000000ull;  // bit pattern of 2^52

// q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47 in 52-bit limbs
F52_DI constexpr uint64_t fq52_q(int i) {
    return i == 0 ? 0x8c16d87cfd47ull
         : i == 1 ? 0x916871ca8d3c2ull
         : i == 2 ? 0x181585d97816aull
         : i == 3 ? 0xa029b85045b68ull
                  : 0x30644e72e131ull;
}
constexpr uint64_t kFq52Inv = 0x20782e4866389ull;  // -q^-1 mod 2^52

// How many hi / lo bit patterns land in column k of a Montgomery product (k = 0..9):
//   a*b  : lo of (i,j) -> column i+j, hi -> i+j+1, all 25 pairs (a squaring adds the i<j pairs twice)
//   m*q  : the same for rows i = 0..4, except that lo of (i, 0) is never formed (it only produces a carry)
F52_DI constexpr uint64_t f52_bias(int k) {
    uint64_t hi = 0, lo = 0;
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) {
            if (i + j == k) lo += 1;          // a*b lo
            if (i + j + 1 == k) hi += 1;      // a*b hi
            if (i + j == k && j >= 1) lo += 1;  // m*q lo
            if (i + j + 1 == k) hi += 1;      // m*q hi
        }
    return 0ull - (hi * kBitsC1 + lo * kBitsC2);
}

F52_DI D52 f52_to_d(const N52 &a) {
    D52 r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.l[i] = F52_L2D((long long)(a.l[i] | kBitsC2)) - 0x1p52;
    return r;
}

// hi and lo halves of a*b as raw bit patterns
F52_DI void f52_prod(double a, double b, uint64_t &hi, uint64_t &lo) {
    const double h = F52_FMA_RZ(a, b, 0x1p104);
    const double s = (0x1p104 + 0x1p52) - h;  // exact
    const double l = F52_FMA_RZ(a, b, s);
    hi = (uint64_t)F52_D2L(h);
    lo = (uint64_t)F52_D2L(l);
}
F52_DI uint64_t f52_prod_hi(double a, double b) { return (uint64_t)F52_D2L(F52_FMA_RZ(a, b, 0x1p104)); }

// Montgomery reduction of the ten columns T (a*b already accumulated) and normalisation.
F52_DI N52 f52_redc(uint64_t (&T)[10]) {
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const uint64_t t = T[i] & kMask52;
        const uint64_t m = (t * kFq52Inv) & kMask52;
        const double md = F52_L2D((long long)(m | kBitsC2)) - 0x1p52;
        // column i: t + lo(m*q0) is 0 or 2^52, so its only effect is a carry of (t != 0)
        T[i + 1] += f52_prod_hi(md, (double)fq52_q(0)) + ((T[i] >> 52) + (t != 0 ? 1u : 0u));
#pragma unroll
        for (int j = 1; j < 5; j++) {
            uint64_t hi, lo;
            f52_prod(md, (double)fq52_q(j), hi, lo);
            T[i + j] += lo;
            T[i + j + 1] += hi;
        }
    }
    N52 r;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        r.l[j] = T[5 + j] & kMask52;
        T[6 + j] += T[5 + j] >> 52;
    }
    r.l[4] = T[9];
    return r;
}

// a * b * 2^-260 mod q (not canonical): for a < Ba*2^254, b < Bb*2^254 the result is below
// (Ba*Bb/64 + 0.7562) * 2^254.  Both operands must be below 2^260.
F52_DI N52 f52_mul(const D52 &a, const D52 &b) {
    uint64_t T[10];
#pragma unroll
    for (int k = 0; k < 10; k++) T[k] = f52_bias(k);
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            uint64_t hi, lo;
            f52_prod(a.l[i], b.l[j], hi, lo);
            T[i + j] += lo;
            T[i + j + 1] += hi;
        }
    return f52_redc(T);
}
F52_DI N52 f52_sqr(const D52 &a) {
    uint64_t T[10];
#pragma unroll
    for (int k = 0; k < 10; k++) T[k] = f52_bias(k);
#pragma unroll
    for (int i = 0; i < 5; i++) {
        uint64_t hi, lo;
        f52_prod(a.l[i], a.l[i], hi, lo);
        T[2 * i] += lo;
        T[2 * i + 1] += hi;
#pragma unroll
        for (int j = i + 1; j < 5; j++) {
            f52_prod(a.l[i], a.l[j], hi, lo);
            T[i + j] += lo + lo;
            T[i + j + 1] += hi + hi;
        }
    }
    return f52_redc(T);
}

// Carry propagation of signed, un-normalised limbs whose value lies in [0, 2^260).
F52_DI N52 f52_norm(const int64_t (&s)[5]) {
    N52 r;
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t v = s[i] + c;
        r.l[i] = (uint64_t)v & kMask52;
        c = v >> 52;  // arithmetic
    }
    r.l[4] = (uint64_t)(s[4] + c);
    return r;
}
// a - b + K*q        (needs b <= K*q and a + K*q < 2^260)
template <int K>
F52_DI N52 f52_sub(const N52 &a, const N52 &b) {
    int64_t s[5];
#pragma unroll
    for (int i = 0; i < 5; i++) s[i] = (int64_t)a.l[i] - (int64_t)b.l[i] + (int64_t)(K * fq52_q(i));
    return f52_norm(s);
}
// a - b - 2c + K*q   (needs b + 2c <= K*q)
template <int K>
F52_DI N52 f52_sub_b_2c(const N52 &a, const N52 &b, const N52 &c) {
    int64_t s[5];
#pragma unroll
    for (int i = 0; i < 5; i++)
        s[i] = (int64_t)a.l[i] - (int64_t)b.l[i] - 2 * (int64_t)c.l[i] + (int64_t)(K * fq52_q(i));
    return f52_norm(s);
}
F52_DI N52 f52_add(const N52 &a, const N52 &b) {
    int64_t s[5];
#pragma unroll
    for (int i = 0; i < 5; i++) s[i] = (int64_t)(a.l[i] + b.l[i]);
    return f52_norm(s);
}
// K*q - a   (needs a <= K*q)
template <int K>
F52_DI N52 f52_neg(const N52 &a) {
    int64_t s[5];
#pragma unroll
    for (int i = 0; i < 5; i++) s[i] = (int64_t)(K * fq52_q(i)) - (int64_t)a.l[i];
    return f52_norm(s);
}
// a == 0 (mod q) for a < 2q: a is 0 or q
F52_DI bool f52_is_zero_mod_q_lt2q(const N52 &a) {
    uint64_t z = 0, e = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        z |= a.l[i];
        e |= a.l[i] ^ fq52_q(i);
    }
    return z == 0 || e == 0;
}

// 256-bit integer (8 x u32 little-endian words) <-> 52-bit limbs.  pack needs value < 2^256.
F52_DI N52 f52_unpack(const uint32_t (&w)[8]) {
    const uint64_t W0 = ((uint64_t)w[1] << 32) | w[0], W1 = ((uint64_t)w[3] << 32) | w[2];
    const uint64_t W2 = ((uint64_t)w[5] << 32) | w[4], W3 = ((uint64_t)w[7] << 32) | w[6];
    N52 r;
    r.l[0] = W0 & kMask52;
    r.l[1] = ((W0 >> 52) | (W1 << 12)) & kMask52;
    r.l[2] = ((W1 >> 40) | (W2 << 24)) & kMask52;
    r.l[3] = ((W2 >> 28) | (W3 << 36)) & kMask52;
    r.l[4] = W3 >> 16;
    return r;
}
F52_DI void f52_pack(const N52 &a, uint32_t (&w)[8]) {
    const uint64_t W0 = a.l[0] | (a.l[1] << 52), W1 = (a.l[1] >> 12) | (a.l[2] << 40);
    const uint64_t W2 = (a.l[2] >> 24) | (a.l[3] << 28), W3 = (a.l[3] >> 36) | (a.l[4] << 16);
    w[0] = (uint32_t)W0; w[1] = (uint32_t)(W0 >> 32);
    w[2] = (uint32_t)W1; w[3] = (uint32_t)(W1 >> 32);
    w[4] = (uint32_t)W2; w[5] = (uint32_t)(W2 >> 32);
    w[6] = (uint32_t)W3; w[7] = (uint32_t)(W3 >> 32);
}

}  // namespace h2b
