// curve.cuh -- BN254 G1 (y^2 = x^3 + 3 over Fq) group law for the MSM kernels.
//
// Memory formats are the reference's (halo2curves 0.3.2 @9f5c508 src/bn256/curve.rs,
// pinned by /root/reference/circuits/Cargo.lock:854-856):
//   G1Affine {x, y}     64 B, Montgomery, identity = (0, 0)
//   G1       {x, y, z}  96 B, HOMOGENEOUS projective (x = X/Z, y = Y/Z), Montgomery, identity z = 0
//                       ((0, R, 0) from G1::identity()).  Established by executing the reference's
//                       own compiled prover: its best_multiexp results are on the
//                       curve only under the homogeneous reading, not the Jacobian one.
// Internally buckets are kept in extended Jacobian "XYZZ" coordinates (x = X/ZZ,
// y = Y/ZZZ, ZZ^3 = ZZZ^2; identity ZZ = 0): a mixed addition is 8M + 2S, the figure
// SURVEY.md section 8d uses for the MSM work model.
#pragma once
#include "field.cuh"

namespace h2b {

struct __align__(16) Affine {
    Fe x, y;
};
struct __align__(16) XYZZ {
    Fe x, y, zz, zzz;
};
struct __align__(16) Projective {
    Fe x, y, z;
};

H2B_DI bool affine_is_identity(const Affine &p) { return Fq::is_zero(p.x) && Fq::is_zero(p.y); }
H2B_DI bool xyzz_is_identity(const XYZZ &p) { return Fq::is_zero(p.zz); }

H2B_DI XYZZ xyzz_identity() {
    XYZZ r;
    r.x = Fq::zero();
    r.y = Fq::zero();
    r.zz = Fq::zero();
    r.zzz = Fq::zero();
    return r;
}
H2B_DI XYZZ xyzz_from_affine(const Affine &p) {
    XYZZ r;
    if (affine_is_identity(p)) return xyzz_identity();
    r.x = p.x;
    r.y = p.y;
    r.zz = Fq::one();
    r.zzz = Fq::one();
    return r;
}

H2B_DI Affine load_affine(const Affine *p) {
    Affine r;
    r.x = load_fe_ro(&p->x);
    r.y = load_fe_ro(&p->y);
    return r;
}
H2B_DI XYZZ load_xyzz(const XYZZ *p) {
    XYZZ r;
    r.x = load_fe(&p->x);
    r.y = load_fe(&p->y);
    r.zz = load_fe(&p->zz);
    r.zzz = load_fe(&p->zzz);
    return r;
}
H2B_DI void store_xyzz(XYZZ *p, const XYZZ &v) {
    store_fe(&p->x, v.x);
    store_fe(&p->y, v.y);
    store_fe(&p->zz, v.zz);
    store_fe(&p->zzz, v.zzz);
}

// 2 * P  (dbl-2008-s-1 with a = 0)
H2B_DI XYZZ xyzz_dbl(const XYZZ &p) {
    if (xyzz_is_identity(p)) return p;
    Fe u = Fq::dbl(p.y);
    Fe v = Fq::sqr(u);
    Fe w = Fq::mul(u, v);
    Fe s = Fq::mul(p.x, v);
    Fe xx = Fq::sqr(p.x);
    Fe m = Fq::add(Fq::dbl(xx), xx);
    XYZZ r;
    r.x = Fq::sub(Fq::sub(Fq::sqr(m), s), s);
    r.y = Fq::mul2_sub(m, Fq::sub(s, r.x), w, p.y);  // two products, one reduction
    r.zz = Fq::mul(v, p.zz);
    r.zzz = Fq::mul(w, p.zzz);
    return r;
}

// acc += P (affine, not the identity)   (madd-2008-s; 8M + 2S)
H2B_DI void xyzz_madd(XYZZ &acc, const Affine &p) {
    if (xyzz_is_identity(acc)) {
        acc.x = p.x;
        acc.y = p.y;
        acc.zz = Fq::one();
        acc.zzz = Fq::one();
        return;
    }
    Fe u2 = Fq::mul(p.x, acc.zz);
    Fe s2 = Fq::mul(p.y, acc.zzz);
    Fe pp_ = Fq::sub(u2, acc.x);
    Fe rr = Fq::sub(s2, acc.y);
    if (Fq::is_zero(pp_)) {
        if (Fq::is_zero(rr)) {  // same point: double it
            XYZZ t;
            t.x = p.x;
            t.y = p.y;
            t.zz = Fq::one();
            t.zzz = Fq::one();
            acc = xyzz_dbl(t);
        } else {  // P + (-P)
            acc = xyzz_identity();
        }
        return;
    }
    Fe pp = Fq::sqr(pp_);
    Fe ppp = Fq::mul(pp_, pp);
    Fe q = Fq::mul(acc.x, pp);
    Fe x3 = Fq::sub(Fq::sub(Fq::sub(Fq::sqr(rr), ppp), q), q);
    Fe y3 = Fq::mul2_sub(rr, Fq::sub(q, x3), acc.y, ppp);  // two products, one reduction
    acc.x = x3;
    acc.y = y3;
    acc.zz = Fq::mul(acc.zz, pp);
    acc.zzz = Fq::mul(acc.zzz, ppp);
}

// acc += Q (XYZZ)   (add-2008-s; 12M + 2S).  Not inlined: used by the reduction kernels.
static __device__ __noinline__ void xyzz_add(XYZZ &acc, const XYZZ &q) {
    if (xyzz_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = q;
        return;
    }
    Fe u1 = Fq::mul(acc.x, q.zz);
    Fe u2 = Fq::mul(q.x, acc.zz);
    Fe s1 = Fq::mul(acc.y, q.zzz);
    Fe s2 = Fq::mul(q.y, acc.zzz);
    Fe pp_ = Fq::sub(u2, u1);
    Fe rr = Fq::sub(s2, s1);
    if (Fq::is_zero(pp_)) {
        if (Fq::is_zero(rr)) acc = xyzz_dbl(acc);
        else acc = xyzz_identity();
        return;
    }
    Fe pp = Fq::sqr(pp_);
    Fe ppp = Fq::mul(pp_, pp);
    Fe qq = Fq::mul(u1, pp);
    Fe x3 = Fq::sub(Fq::sub(Fq::sub(Fq::sqr(rr), ppp), qq), qq);
    Fe y3 = Fq::mul2_sub(rr, Fq::sub(qq, x3), s1, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = Fq::mul(Fq::mul(acc.zz, q.zz), pp);
    acc.zzz = Fq::mul(Fq::mul(acc.zzz, q.zzz), ppp);
}

static __device__ __noinline__ XYZZ xyzz_dbl_ni(const XYZZ &p) { return xyzz_dbl(p); }

// Jacobian (x = X/Z^2, y = Y/Z^3) doubling for a = 0 (dbl-2009-l: 2M + 5S), used by the long doubling chains of the
// SRS precomputation, where it is a quarter cheaper than the XYZZ form and where Z_(j+1) = 2 Y_j Z_j keeps all the
// denominators of a chain in one running product (one shared inversion per chain).  Z = 0 stays the identity.
struct Jac {
    Fe x, y, z;
};
static __device__ __noinline__ void jac_dbl_ni(Jac &p) {
    const Fe a = Fq::sqr(p.x), b = Fq::sqr(p.y), c = Fq::sqr(b);
    Fe d = Fq::sub(Fq::sub(Fq::sqr(Fq::add(p.x, b)), a), c);
    d = Fq::dbl(d);
    const Fe e = Fq::add(Fq::dbl(a), a), f = Fq::sqr(e);
    const Fe z3 = Fq::dbl(Fq::mul(p.y, p.z));
    p.x = Fq::sub(Fq::sub(f, d), d);
    Fe c8 = Fq::dbl(Fq::dbl(Fq::dbl(c)));
    p.y = Fq::sub(Fq::mul(e, Fq::sub(d, p.x)), c8);
    p.z = z3;
}

// XYZZ -> a homogeneous projective representative with Z = ZZ * ZZZ (no inversion);
// identity -> (0, R, 0).
H2B_DI Projective xyzz_to_projective(const XYZZ &p) {
    Projective r;
    if (xyzz_is_identity(p)) {
        r.x = Fq::zero();
        r.y = Fq::one();
        r.z = Fq::zero();
        return r;
    }
    r.x = Fq::mul(p.x, p.zzz);  // X/ZZ = X*ZZZ / (ZZ*ZZZ)
    r.y = Fq::mul(p.y, p.zz);   // Y/ZZZ = Y*ZZ / (ZZ*ZZZ)
    r.z = Fq::mul(p.zz, p.zzz);
    return r;
}
// homogeneous (X:Y:Z) -> XYZZ (X*Z, Y*Z^2, Z^2, Z^3)
H2B_DI XYZZ projective_to_xyzz(const Projective &p) {
    XYZZ r;
    if (Fq::is_zero(p.z)) return xyzz_identity();
    r.zz = Fq::sqr(p.z);
    r.zzz = Fq::mul(r.zz, p.z);
    r.x = Fq::mul(p.x, p.z);
    r.y = Fq::mul(p.y, r.zz);
    return r;
}

}  // namespace h2b
