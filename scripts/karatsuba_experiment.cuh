// karatsuba_experiment.cuh -- NOT part of the product.  One level of Karatsuba on the 256 x 256-bit product of the
// Montgomery multiplication (3 x 16 instead of 64 limb products, reduction of the low half with the limb shift fused
// into its carry chains).  Bit-identical to Field<P>::mul, but measured SLOWER on B200 (scripts/f52_bench.cu: 58.5
// against 65.6 G modmul/s at 4 blocks/SM): ptxas places the ~100 extra carry additions on the FMA pipe as IMAD.X /
// IMAD.IADD, which cancels the 25 IMAD.WIDE saved.  Kept with the other pipe probes as evidence.
#pragma once
#include "../halo2-prover_b200/csrc/field.cuh"

namespace h2b {

// ------------------------------------------------------------------ 4 x 4 limb product (Karatsuba leaf)
// r[0..7] = x[0..3] * y[0..3].  Products whose limb offset is even accumulate in r, the others in a second
// register set one limb higher, so every row is two short carry chains; the two sets are summed at the end.
H2B_DI void mul4x4(uint32_t (&r)[8], const uint32_t (&x)[4], const uint32_t (&y)[4]) {
    uint32_t o0, o1, o2, o3, o4 = 0, o5 = 0, o6 = 0;
    {
        const uint64_t p0 = (uint64_t)x[0] * y[0], p2 = (uint64_t)x[2] * y[0];
        const uint64_t p1 = (uint64_t)x[1] * y[0], p3 = (uint64_t)x[3] * y[0];
        r[0] = (uint32_t)p0; r[1] = (uint32_t)(p0 >> 32); r[2] = (uint32_t)p2; r[3] = (uint32_t)(p2 >> 32);
        o0 = (uint32_t)p1; o1 = (uint32_t)(p1 >> 32); o2 = (uint32_t)p3; o3 = (uint32_t)(p3 >> 32);
        r[4] = r[5] = r[6] = r[7] = 0;
    }
    // y1: x0,x2 at limbs 1,3 (odd set), x1,x3 at limbs 2,4 (even set)
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(o0), "+r"(o1), "+r"(o2), "+r"(o3), "+r"(o4)
        : "r"(x[0]), "r"(x[2]), "r"(y[1]));
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6])
        : "r"(x[1]), "r"(x[3]), "r"(y[1]));
    // y2: x0,x2 at limbs 2,4 (even set), x1,x3 at limbs 3,5 (odd set)
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6])
        : "r"(x[0]), "r"(x[2]), "r"(y[2]));
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6)
        : "r"(x[1]), "r"(x[3]), "r"(y[2]));
    // y3: x0,x2 at limbs 3,5 (odd set), x1,x3 at limbs 4,6 (even set; the total is < 2^256: no carry out)
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6)
        : "r"(x[0]), "r"(x[2]), "r"(y[3]));
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"
        "madc.hi.u32 %3, %5, %6, %3;"
        : "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
        : "r"(x[1]), "r"(x[3]), "r"(y[3]));
    asm("add.cc.u32 %0, %0, %7;\n\t"
        "addc.cc.u32 %1, %1, %8;\n\t"
        "addc.cc.u32 %2, %2, %9;\n\t"
        "addc.cc.u32 %3, %3, %10;\n\t"
        "addc.cc.u32 %4, %4, %11;\n\t"
        "addc.cc.u32 %5, %5, %12;\n\t"
        "addc.u32 %6, %6, %13;"
        : "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
        : "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(o6));
}
// d = |hi - lo| over 4 limbs; returns 0xffffffff when hi < lo, else 0
H2B_DI uint32_t absdiff4(uint32_t (&d)[4], const uint32_t *hi, const uint32_t *lo) {
    uint32_t neg;
    asm("sub.cc.u32 %0, %5, %9;\n\t"
        "subc.cc.u32 %1, %6, %10;\n\t"
        "subc.cc.u32 %2, %7, %11;\n\t"
        "subc.cc.u32 %3, %8, %12;\n\t"
        "subc.u32 %4, 0, 0;"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(neg)
        : "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]));
    // two's complement negate when neg: (d ^ neg) + (neg & 1)
    d[0] ^= neg; d[1] ^= neg; d[2] ^= neg; d[3] ^= neg;
    asm("add.cc.u32 %0, %0, %4;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.u32 %3, %3, 0;"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(neg & 1u));
    return neg;
}

    // The same product with one level of Karatsuba on the 256 x 256-bit multiplication (3 x 16 instead of
    // 64 limb products) followed by the Montgomery reduction of the low half alone (64 limb products, the
    // limb shift fused into the carry chains as in mul()):  112 IMAD.WIDE instead of 128, ~100 more adds.
template <class P>
H2B_DI Fe mul_karatsuba(const Fe &a, const Fe &b) {
        uint32_t z0[8], z2[8], pp[8], da[4], db[4];
        const uint32_t alo[4] = {a.l[0], a.l[1], a.l[2], a.l[3]}, ahi[4] = {a.l[4], a.l[5], a.l[6], a.l[7]};
        const uint32_t blo[4] = {b.l[0], b.l[1], b.l[2], b.l[3]}, bhi[4] = {b.l[4], b.l[5], b.l[6], b.l[7]};
        mul4x4(z0, alo, blo);
        mul4x4(z2, ahi, bhi);
        const uint32_t na = absdiff4(da, ahi, alo), nb = absdiff4(db, bhi, blo);
        mul4x4(pp, da, db);
        // mid = alo*bhi + ahi*blo = z0 + z2 - sign * pp  (sign = +1 when the two differences have equal signs)
        const uint32_t m = ~(na ^ nb);  // all ones: subtract pp
        uint32_t t[9];
        asm("add.cc.u32 %0, %9, %17;\n\t"
            "addc.cc.u32 %1, %10, %18;\n\t"
            "addc.cc.u32 %2, %11, %19;\n\t"
            "addc.cc.u32 %3, %12, %20;\n\t"
            "addc.cc.u32 %4, %13, %21;\n\t"
            "addc.cc.u32 %5, %14, %22;\n\t"
            "addc.cc.u32 %6, %15, %23;\n\t"
            "addc.cc.u32 %7, %16, %24;\n\t"
            "addc.u32 %8, 0, 0;"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8])
            : "r"(z0[0]), "r"(z0[1]), "r"(z0[2]), "r"(z0[3]), "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]),
              "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
        {
            // t += (pp ^ m) + (m & 1) over 9 limbs (limb 8 of the complement is m); "add.cc m, m" sets the carry to m & 1
            uint32_t dummy;
            asm("add.cc.u32 %9, %10, %10;\n\t"
                "addc.cc.u32 %0, %0, %11;\n\t"
                "addc.cc.u32 %1, %1, %12;\n\t"
                "addc.cc.u32 %2, %2, %13;\n\t"
                "addc.cc.u32 %3, %3, %14;\n\t"
                "addc.cc.u32 %4, %4, %15;\n\t"
                "addc.cc.u32 %5, %5, %16;\n\t"
                "addc.cc.u32 %6, %6, %17;\n\t"
                "addc.cc.u32 %7, %7, %18;\n\t"
                "addc.u32 %8, %8, %10;"
                : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]),
                  "=r"(dummy)
                : "r"(m), "r"(pp[0] ^ m), "r"(pp[1] ^ m), "r"(pp[2] ^ m), "r"(pp[3] ^ m), "r"(pp[4] ^ m), "r"(pp[5] ^ m),
                  "r"(pp[6] ^ m), "r"(pp[7] ^ m));
        }
        // T = z0 + mid * 2^128 + z2 * 2^256: limbs 4..15 (limbs 0..3 are z0[0..3])
        uint32_t hi[8];  // T[8..15]
        asm("add.cc.u32 %0, %0, %12;\n\t"
            "addc.cc.u32 %1, %1, %13;\n\t"
            "addc.cc.u32 %2, %2, %14;\n\t"
            "addc.cc.u32 %3, %3, %15;\n\t"
            "addc.cc.u32 %4, %20, %16;\n\t"
            "addc.cc.u32 %5, %21, %17;\n\t"
            "addc.cc.u32 %6, %22, %18;\n\t"
            "addc.cc.u32 %7, %23, %19;\n\t"
            "addc.cc.u32 %8, %24, %28;\n\t"
            "addc.cc.u32 %9, %25, 0;\n\t"
            "addc.cc.u32 %10, %26, 0;\n\t"
            "addc.u32 %11, %27, 0;"
            : "+r"(z0[4]), "+r"(z0[5]), "+r"(z0[6]), "+r"(z0[7]), "=r"(hi[0]), "=r"(hi[1]), "=r"(hi[2]), "=r"(hi[3]),
              "=r"(hi[4]), "=r"(hi[5]), "=r"(hi[6]), "=r"(hi[7])
            : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
              "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]), "r"(t[8]));
        // Montgomery reduction of the low half: U = (T_lo + M * N) / 2^256, eight steps with the limb shift
        // folded into the odd-limb chain (the accumulator pair swaps roles every step, as in mul())
        uint32_t ev[8], od[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { ev[i] = z0[i]; od[i] = 0; }
        Field<P>::redc_step(ev, od);
#pragma unroll
        for (int i = 1; i < 8; i += 2) {
            {   // roles after the shift: E = od, O = ev
                const uint32_t mq = (od[0] + ev[1]) * P::M0;
                shift_mad_row(od[0], ev, P::n(1), P::n(3), P::n(5), P::n(7), mq);
                cmad_row_fold(od, ev[7], P::n(0), P::n(2), P::n(4), P::n(6), mq);
            }
            if (i + 1 < 8) {  // roles: E = ev, O = od
                const uint32_t mq = (ev[0] + od[1]) * P::M0;
                shift_mad_row(ev[0], od, P::n(1), P::n(3), P::n(5), P::n(7), mq);
                cmad_row_fold(ev, od[7], P::n(0), P::n(2), P::n(4), P::n(6), mq);
            }
        }
        // last step used E = od, O = ev: U = ev + (od >> 32); result = U + T_hi
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
        asm("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\t"
            "addc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\t"
            "addc.u32 %7, %7, %15;"
            : "+r"(r.l[0]), "+r"(r.l[1]), "+r"(r.l[2]), "+r"(r.l[3]), "+r"(r.l[4]), "+r"(r.l[5]), "+r"(r.l[6]), "+r"(r.l[7])
            : "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]));
        return Field<P>::reduce_once(r);
    }


}  // namespace h2b
