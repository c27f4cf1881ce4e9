// field.cuh -- BN254 Fr / Fq Montgomery arithmetic on 8 x 32-bit limbs for sm_100a.
//
// Replaces, on the device, the field layer the reference reaches through
// halo2curves 0.3.2 @9f5c508 src/bn256/{fr,fq}.rs (pinned by
// /root/reference/circuits/Cargo.lock:854-856): same modulus, same Montgomery
// radix R = 2^256, same canonical (< modulus) representatives, so a 32-byte
// element written by either side is read unchanged by the other (4 x u64 LE
// limbs == 8 x u32 LE limbs).
//
// Multiplication keeps two interleaved accumulators ("even" and "odd" columns) so
// that every 32x32->64 product lands on a register pair and a whole row of
// products is one mad.lo.cc / madc.hi.cc carry chain; ptxas lowers each lo/hi
// pair to IMAD.WIDE.U32 with carry.  The accumulators swap roles on every limb of
// b, which performs the Montgomery shift for free.  (The chain structure was
// validated against big-integer arithmetic instruction by instruction before it
// was written down here.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace h2b {

struct __align__(16) Fe {
    uint32_t l[8];
};

#define H2B_DI __device__ __forceinline__

// ------------------------------------------------------------------ field parameters
struct FrP {
    static constexpr uint32_t M0 = 0xefffffffu;
    H2B_DI static constexpr uint32_t n(int i) {
        constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    H2B_DI static constexpr uint32_t one(int i) {  // R mod r
        constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    H2B_DI static constexpr uint32_t r2(int i) {  // R^2 mod r
        constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return v[i];
    }
    H2B_DI static constexpr uint32_t n2(int i) {  // 2 r
        constexpr uint32_t v[8] = {0xe0000002u, 0x87c3eb27u, 0xf372e122u, 0x5067d090u,
                                   0x0302b0bau, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u};
        return v[i];
    }
};
struct FqP {
    static constexpr uint32_t M0 = 0xe4866389u;
    H2B_DI static constexpr uint32_t n(int i) {
        constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    H2B_DI static constexpr uint32_t one(int i) {  // R mod q
        constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    H2B_DI static constexpr uint32_t r2(int i) {  // R^2 mod q
        constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return v[i];
    }
    H2B_DI static constexpr uint32_t n2(int i) {  // 2 q
        constexpr uint32_t v[8] = {0xb0f9fa8eu, 0x7841182du, 0xd0e3951au, 0x2f02d522u,
                                   0x0302b0bbu, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u};
        return v[i];
    }
};

// ------------------------------------------------------------------ carry-chain rows
// acc[0..7] += x{0,2,4,6} * m laid out as four 64-bit products; carry out is added to `top`.
H2B_DI void cmad_row_fold(uint32_t (&e)[8], uint32_t &top, uint32_t x0, uint32_t x2, uint32_t x4,
                          uint32_t x6, uint32_t m) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]),
          "+r"(e[7]), "+r"(top)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(m));
}
// Same without a carry out (the caller knows the running total fits).
H2B_DI void cmad_row(uint32_t (&e)[8], uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6,
                     uint32_t m) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]),
          "+r"(e[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(m));
}
// e[0] += o[1]; then o <- (o >> 64) + x{1,3,5,7} * m, continuing the same carry chain.
H2B_DI void shift_mad_row(uint32_t &e0, uint32_t (&o)[8], uint32_t x1, uint32_t x3, uint32_t x5,
                          uint32_t x7, uint32_t m) {
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"
        "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
        "madc.hi.u32 %8, %12, %13, 0;"
        : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]),
          "+r"(o[6]), "+r"(o[7])
        : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(m));
}

// Row variants for the squaring: the first Z multiplicands are zero.  With no carry yet, the chain of cmad_row_fold
// simply starts at the first non-zero product; in shift_mad_row the skipped products become plain carry adds.
template <int Z>
H2B_DI void cmad_row_fold_z(uint32_t (&e)[8], uint32_t &top, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6,
                            uint32_t m) {
    if constexpr (Z == 0) {
        cmad_row_fold(e, top, x0, x2, x4, x6, m);
    } else if constexpr (Z == 1) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
            "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
            "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
            "addc.u32 %6, %6, 0;"
            : "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(top)
            : "r"(x2), "r"(x4), "r"(x6), "r"(m));
    } else if constexpr (Z == 2) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
            "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, %4, 0;"
            : "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(top)
            : "r"(x4), "r"(x6), "r"(m));
    } else if constexpr (Z == 3) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
            "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32 %2, %2, 0;"
            : "+r"(e[6]), "+r"(e[7]), "+r"(top)
            : "r"(x6), "r"(m));
    }
}
template <int Z>
H2B_DI void shift_mad_row_z(uint32_t &e0, uint32_t (&o)[8], uint32_t x1, uint32_t x3, uint32_t x5, uint32_t x7,
                            uint32_t m) {
    if constexpr (Z == 0) {
        shift_mad_row(e0, o, x1, x3, x5, x7, m);
    } else if constexpr (Z == 1) {
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.cc.u32 %1, %3, 0;\n\t"
            "addc.cc.u32 %2, %4, 0;\n\t"
            "madc.lo.cc.u32 %3, %9, %12, %5;\n\t"
            "madc.hi.cc.u32 %4, %9, %12, %6;\n\t"
            "madc.lo.cc.u32 %5, %10, %12, %7;\n\t"
            "madc.hi.cc.u32 %6, %10, %12, %8;\n\t"
            "madc.lo.cc.u32 %7, %11, %12, 0;\n\t"
            "madc.hi.u32 %8, %11, %12, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(x3), "r"(x5), "r"(x7), "r"(m));
    } else if constexpr (Z == 2) {
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.cc.u32 %1, %3, 0;\n\t"
            "addc.cc.u32 %2, %4, 0;\n\t"
            "addc.cc.u32 %3, %5, 0;\n\t"
            "addc.cc.u32 %4, %6, 0;\n\t"
            "madc.lo.cc.u32 %5, %9, %11, %7;\n\t"
            "madc.hi.cc.u32 %6, %9, %11, %8;\n\t"
            "madc.lo.cc.u32 %7, %10, %11, 0;\n\t"
            "madc.hi.u32 %8, %10, %11, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(x5), "r"(x7), "r"(m));
    } else if constexpr (Z == 3) {
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.cc.u32 %1, %3, 0;\n\t"
            "addc.cc.u32 %2, %4, 0;\n\t"
            "addc.cc.u32 %3, %5, 0;\n\t"
            "addc.cc.u32 %4, %6, 0;\n\t"
            "addc.cc.u32 %5, %7, 0;\n\t"
            "addc.cc.u32 %6, %8, 0;\n\t"
            "madc.lo.cc.u32 %7, %9, %10, 0;\n\t"
            "madc.hi.u32 %8, %9, %10, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(x7), "r"(m));
    }
}

template <class P>
struct Field {
    H2B_DI static Fe zero() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = 0;
        return r;
    }
    H2B_DI static Fe one() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::one(i);
        return r;
    }
    H2B_DI static Fe r2() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::r2(i);
        return r;
    }
    H2B_DI static bool is_zero(const Fe &a) {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) v |= a.l[i];
        return v == 0;
    }
    H2B_DI static bool eq(const Fe &a, const Fe &b) {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) v |= a.l[i] ^ b.l[i];
        return v == 0;
    }

    // r = a - N if a >= N else a      (a < 2N)
    H2B_DI static Fe reduce_once(const Fe &a) {
        Fe t;
        uint32_t borrow;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]),
              "=r"(t.l[6]), "=r"(t.l[7]), "=r"(borrow)
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
              "r"(a.l[6]), "r"(a.l[7]), "r"(P::n(0)), "r"(P::n(1)), "r"(P::n(2)), "r"(P::n(3)),
              "r"(P::n(4)), "r"(P::n(5)), "r"(P::n(6)), "r"(P::n(7)));
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = borrow ? a.l[i] : t.l[i];
        return r;
    }

    H2B_DI static Fe add(const Fe &a, const Fe &b) {
        Fe s;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(s.l[0]), "=r"(s.l[1]), "=r"(s.l[2]), "=r"(s.l[3]), "=r"(s.l[4]), "=r"(s.l[5]),
              "=r"(s.l[6]), "=r"(s.l[7])
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
              "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]),
              "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
        return reduce_once(s);  // both moduli < 2^254: no carry out of 256 bits
    }
    H2B_DI static Fe dbl(const Fe &a) { return add(a, a); }

    H2B_DI static Fe sub(const Fe &a, const Fe &b) {
        Fe d;
        uint32_t borrow;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(d.l[0]), "=r"(d.l[1]), "=r"(d.l[2]), "=r"(d.l[3]), "=r"(d.l[4]), "=r"(d.l[5]),
              "=r"(d.l[6]), "=r"(d.l[7]), "=r"(borrow)
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
              "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]),
              "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
        // borrow is 0 or 0xffffffff: add back N & borrow
        Fe r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]),
              "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(d.l[0]), "r"(d.l[1]), "r"(d.l[2]), "r"(d.l[3]), "r"(d.l[4]), "r"(d.l[5]),
              "r"(d.l[6]), "r"(d.l[7]), "r"(P::n(0) & borrow), "r"(P::n(1) & borrow),
              "r"(P::n(2) & borrow), "r"(P::n(3) & borrow), "r"(P::n(4) & borrow),
              "r"(P::n(5) & borrow), "r"(P::n(6) & borrow), "r"(P::n(7) & borrow));
        return r;
    }
    H2B_DI static Fe neg(const Fe &a) {
        if (is_zero(a)) return a;
        Fe n;
#pragma unroll
        for (int i = 0; i < 8; i++) n.l[i] = P::n(i);
        return sub_nored(n, a);
    }
    // a - b for a >= b (no correction)
    H2B_DI static Fe sub_nored(const Fe &a, const Fe &b) {
        Fe d;
        asm("sub.cc.u32 %0, %8, %16;\n\t"
            "subc.cc.u32 %1, %9, %17;\n\t"
            "subc.cc.u32 %2, %10, %18;\n\t"
            "subc.cc.u32 %3, %11, %19;\n\t"
            "subc.cc.u32 %4, %12, %20;\n\t"
            "subc.cc.u32 %5, %13, %21;\n\t"
            "subc.cc.u32 %6, %14, %22;\n\t"
            "subc.u32 %7, %15, %23;"
            : "=r"(d.l[0]), "=r"(d.l[1]), "=r"(d.l[2]), "=r"(d.l[3]), "=r"(d.l[4]), "=r"(d.l[5]),
              "=r"(d.l[6]), "=r"(d.l[7])
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
              "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]),
              "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
        return d;
    }

    // ---- lazy range [0, 2N): the butterflies of the NTT keep their values there and reduce once at the very end
    // a - 2N if a >= 2N else a      (a < 4N)
    H2B_DI static Fe reduce_2n(const Fe &a) {
        Fe t;
        uint32_t borrow;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]),
              "=r"(t.l[6]), "=r"(t.l[7]), "=r"(borrow)
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
              "r"(a.l[6]), "r"(a.l[7]), "r"(P::n2(0)), "r"(P::n2(1)), "r"(P::n2(2)), "r"(P::n2(3)),
              "r"(P::n2(4)), "r"(P::n2(5)), "r"(P::n2(6)), "r"(P::n2(7)));
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = borrow ? a.l[i] : t.l[i];
        return r;
    }
    // a + b brought back to [0, 2N)      (a, b < 2N; 4N < 2^256: no carry out)
    H2B_DI static Fe add_2n(const Fe &a, const Fe &b) {
        Fe s;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(s.l[0]), "=r"(s.l[1]), "=r"(s.l[2]), "=r"(s.l[3]), "=r"(s.l[4]), "=r"(s.l[5]),
              "=r"(s.l[6]), "=r"(s.l[7])
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
              "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]),
              "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
        return reduce_2n(s);
    }
    // a - b + 2N, in (0, 4N): no condition at all      (a, b < 2N)
    H2B_DI static Fe sub_2n(const Fe &a, const Fe &b) {
        Fe d;
        asm("sub.cc.u32 %0, %8, %16;\n\t"
            "subc.cc.u32 %1, %9, %17;\n\t"
            "subc.cc.u32 %2, %10, %18;\n\t"
            "subc.cc.u32 %3, %11, %19;\n\t"
            "subc.cc.u32 %4, %12, %20;\n\t"
            "subc.cc.u32 %5, %13, %21;\n\t"
            "subc.cc.u32 %6, %14, %22;\n\t"
            "subc.u32 %7, %15, %23;"
            : "=r"(d.l[0]), "=r"(d.l[1]), "=r"(d.l[2]), "=r"(d.l[3]), "=r"(d.l[4]), "=r"(d.l[5]),
              "=r"(d.l[6]), "=r"(d.l[7])
            : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
              "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]),
              "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
        Fe r;  // the wrap-around of a negative difference is undone by the addition
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]),
              "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(d.l[0]), "r"(d.l[1]), "r"(d.l[2]), "r"(d.l[3]), "r"(d.l[4]), "r"(d.l[5]),
              "r"(d.l[6]), "r"(d.l[7]), "r"(P::n2(0)), "r"(P::n2(1)), "r"(P::n2(2)), "r"(P::n2(3)),
              "r"(P::n2(4)), "r"(P::n2(5)), "r"(P::n2(6)), "r"(P::n2(7)));
        return r;
    }

    // One Montgomery step on the (E, O) accumulator pair: add N * m so that E[0] becomes 0.
    H2B_DI static void redc_step(uint32_t (&E)[8], uint32_t (&O)[8]) {
        uint32_t m = E[0] * P::M0;
        cmad_row(O, P::n(1), P::n(3), P::n(5), P::n(7), m);
        cmad_row_fold(E, O[7], P::n(0), P::n(2), P::n(4), P::n(6), m);
    }

    // Montgomery product without the final conditional subtraction: a * b * R^-1 + (a multiple of N), below
    // a * b / R + N.  For a < 4N and b < N that is below 2N (both moduli are below 2^254, so 4N < R), and the running
    // value of the interleaved accumulators stays below 5 N 2^32 < 2^288, inside their capacity.
    H2B_DI static Fe mul_lazy(const Fe &a, const Fe &b) {
        uint32_t ev[8], od[8];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            uint64_t p = (uint64_t)a.l[k] * b.l[0];
            ev[k] = (uint32_t)p;
            ev[k + 1] = (uint32_t)(p >> 32);
            uint64_t q = (uint64_t)a.l[k + 1] * b.l[0];
            od[k] = (uint32_t)q;
            od[k + 1] = (uint32_t)(q >> 32);
        }
        redc_step(ev, od);
#pragma unroll
        for (int i = 1; i < 8; i += 2) {
            // roles: E = od, O = ev
            shift_mad_row(od[0], ev, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
            cmad_row_fold(od, ev[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
            redc_step(od, ev);
            if (i + 1 < 8) {
                // roles: E = ev, O = od
                shift_mad_row(ev[0], od, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i + 1]);
                cmad_row_fold(ev, od[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i + 1]);
                redc_step(ev, od);
            }
        }
        // last step used E = od, O = ev: result = ev + (od >> 32)
        Fe r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, 0;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]),
              "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]),
              "r"(ev[7]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]),
              "r"(od[7]));
        return r;
    }
    // Montgomery product a * b * R^-1 mod N, fully reduced (a * b < N R).
    H2B_DI static Fe mul(const Fe &a, const Fe &b) { return reduce_once(mul_lazy(a, b)); }
    // (a * b + c * d) * R^-1 mod N, fully reduced: two products under ONE Montgomery reduction (24 instead of 32 limb
    // products per row pair).  Every term is non-negative and the running value stays below 3 N B < 2^288, so neither
    // accumulator can overflow; the final value is below N (1 + 2 N / R) < 2 N.  Inputs up to N inclusive.
    H2B_DI static Fe mul2_add(const Fe &a, const Fe &b, const Fe &c, const Fe &d) {
        uint32_t ev[8], od[8];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            uint64_t p = (uint64_t)a.l[k] * b.l[0];
            ev[k] = (uint32_t)p;
            ev[k + 1] = (uint32_t)(p >> 32);
            uint64_t q = (uint64_t)a.l[k + 1] * b.l[0];
            od[k] = (uint32_t)q;
            od[k + 1] = (uint32_t)(q >> 32);
        }
        cmad_row(od, c.l[1], c.l[3], c.l[5], c.l[7], d.l[0]);
        cmad_row_fold(ev, od[7], c.l[0], c.l[2], c.l[4], c.l[6], d.l[0]);
        redc_step(ev, od);
#pragma unroll
        for (int i = 1; i < 8; i += 2) {
            shift_mad_row(od[0], ev, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
            cmad_row_fold(od, ev[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
            cmad_row(ev, c.l[1], c.l[3], c.l[5], c.l[7], d.l[i]);
            cmad_row_fold(od, ev[7], c.l[0], c.l[2], c.l[4], c.l[6], d.l[i]);
            redc_step(od, ev);
            if (i + 1 < 8) {
                shift_mad_row(ev[0], od, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i + 1]);
                cmad_row_fold(ev, od[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i + 1]);
                cmad_row(od, c.l[1], c.l[3], c.l[5], c.l[7], d.l[i + 1]);
                cmad_row_fold(ev, od[7], c.l[0], c.l[2], c.l[4], c.l[6], d.l[i + 1]);
                redc_step(ev, od);
            }
        }
        Fe r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, 0;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]),
              "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]),
              "r"(ev[7]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]),
              "r"(od[7]));
        return reduce_once(r);
    }
    // a * b - c * d
    H2B_DI static Fe mul2_sub(const Fe &a, const Fe &b, const Fe &c, const Fe &d) { return mul2_add(a, b, neg(c), d); }

    // Montgomery square: the same rows as mul(), but row i multiplies a_i into
    //   [0 .. 0, a_i, 2 a_(i+1) .. 2 a_7]  (as the limbs of a_i B^i + 2 (a >> 32(i+1)) B^(i+1); a < 2^254 so nothing
    // is shifted out), i.e. every cross product a_i a_j is formed once, doubled: 36 limb products instead of 64.
    // The zero entries are skipped by the *_z row variants above (ptxas does not fold a mad.cc with a zero operand).
    H2B_DI static Fe sqr(const Fe &a) {
        uint32_t t[8], u[8];  // t_j = limb j of 2a, u_j = a_j << 1 (limb j of 2 (a >> 32 j) B^j)
        t[0] = u[0] = a.l[0] << 1;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            t[j] = __funnelshift_l(a.l[j - 1], a.l[j], 1);
            u[j] = a.l[j] << 1;
        }
        // c(i, j): multiplicand limb j of row i
#define H2B_SQ(i, j) ((j) < (i) ? 0u : ((j) == (i) ? a.l[(j) & 7] : ((j) == (i) + 1 ? u[(j) & 7] : t[(j) & 7])))
        uint32_t ev[8], od[8];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            uint64_t p = (uint64_t)H2B_SQ(0, k) * a.l[0];
            ev[k] = (uint32_t)p;
            ev[k + 1] = (uint32_t)(p >> 32);
            uint64_t q = (uint64_t)H2B_SQ(0, k + 1) * a.l[0];
            od[k] = (uint32_t)q;
            od[k + 1] = (uint32_t)(q >> 32);
        }
        redc_step(ev, od);
        // rows 1..7: even-position multiplicands have (i + 1) / 2 leading zeros, odd-position ones i / 2
#define H2B_SQ_ROW(i, E, O)                                                                                        \
    shift_mad_row_z<(i) / 2>(E[0], O, H2B_SQ(i, 1), H2B_SQ(i, 3), H2B_SQ(i, 5), H2B_SQ(i, 7), a.l[i]);              \
    cmad_row_fold_z<((i) + 1) / 2>(E, O[7], H2B_SQ(i, 0), H2B_SQ(i, 2), H2B_SQ(i, 4), H2B_SQ(i, 6), a.l[i]);        \
    redc_step(E, O);
        H2B_SQ_ROW(1, od, ev)
        H2B_SQ_ROW(2, ev, od)
        H2B_SQ_ROW(3, od, ev)
        H2B_SQ_ROW(4, ev, od)
        H2B_SQ_ROW(5, od, ev)
        H2B_SQ_ROW(6, ev, od)
        H2B_SQ_ROW(7, od, ev)
#undef H2B_SQ_ROW
#undef H2B_SQ
        Fe r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, 0;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]),
              "=r"(r.l[6]), "=r"(r.l[7])
            : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]),
              "r"(ev[7]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]),
              "r"(od[7]));
        return reduce_once(r);
    }
    H2B_DI static Fe sqr_by_mul(const Fe &a) { return mul(a, a); }

    // Reference multiplication on 64-bit temporaries (no inline PTX); used by the self-test
    // kernel to cross-check mul() on the device.
    H2B_DI static Fe mul_portable(const Fe &a, const Fe &b) {
        uint32_t t[10];
#pragma unroll
        for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                c += (uint64_t)a.l[j] * b.l[i] + t[j];
                t[j] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[8] = (uint32_t)c;
            t[9] = (uint32_t)(c >> 32);
            uint32_t m = t[0] * P::M0;
            c = (uint64_t)m * P::n(0) + t[0];
            c >>= 32;
#pragma unroll
            for (int j = 1; j < 8; j++) {
                c += (uint64_t)m * P::n(j) + t[j];
                t[j - 1] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[7] = (uint32_t)c;
            t[8] = t[9] + (uint32_t)(c >> 32);
        }
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = t[i];
        return reduce_once(r);  // t[8] == 0 because the result is < 2N < 2^255
    }

    // Montgomery -> canonical (PrimeField::to_repr): multiply by the integer 1.
    H2B_DI static Fe from_mont(const Fe &a) {
        Fe o = zero();
        o.l[0] = 1;
        return mul(a, o);
    }
    H2B_DI static Fe to_mont(const Fe &a) { return mul(a, r2()); }

    // a^(N-2); a == 0 -> 0.  Not inlined: only used outside hot loops.
    __device__ __noinline__ static Fe inv(Fe a) {
        Fe acc = one();
#pragma unroll 1
        for (int i = 7; i >= 0; i--) {
            uint32_t e = P::n(i) - (i == 0 ? 2u : 0u);  // low limb of both moduli is >= 2
#pragma unroll 1
            for (int bit = 31; bit >= 0; bit--) {
                acc = sqr(acc);
                if ((e >> bit) & 1) acc = mul(acc, a);
            }
        }
        return acc;
    }
    __device__ __noinline__ static Fe pow_u64(Fe a, uint64_t e) {
        Fe acc = one();
#pragma unroll 1
        for (int bit = 63; bit >= 0; bit--) {
            acc = sqr(acc);
            if ((e >> bit) & 1) acc = mul(acc, a);
        }
        return acc;
    }
};

using Fr = Field<FrP>;
using Fq = Field<FqP>;

// 32-byte element <-> two 128-bit memory transactions
H2B_DI Fe load_fe(const Fe *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    Fe r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
H2B_DI Fe load_fe_ro(const Fe *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fe r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
H2B_DI void store_fe(Fe *p, const Fe &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

}  // namespace h2b
