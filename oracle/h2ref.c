/*
 * h2ref.c -- CPU restatement (plain C, pthreads) of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Built into oracle/libh2ref.so by oracle/Makefile and
 * loaded (ctypes) by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product library (libh2b200.so) never links, loads
 * or calls it.
 *
 * The algorithm lives in un-vendored dependencies of /root/reference
 * (circuits/Cargo.lock:836-838 halo2_proofs 0.2.0 @6b43b6b, :854-856 halo2curves
 * 0.3.2 @9f5c508); this file restates their published algorithm:
 *   - Fr/Fq: 4 x u64 Montgomery, R = 2^256    (halo2curves src/bn256/{fr,fq}.rs)
 *   - multiexp_serial / best_multiexp           (halo2_proofs src/arithmetic.rs:28-140, :147-180)
 *   - best_fft / recursive_butterfly_arithmetic (src/arithmetic.rs:185-250, :253-290)
 *   - parallelize                               (src/arithmetic.rs:~395-425)
 *   - EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended,
 *     extended_to_coeff, divide_by_vanishing_poly} (src/poly/domain.rs:~40-140, :227, :244, :311)
 * rayon is replaced by pthreads with the same work split (contiguous chunks of
 * len/num_threads for the MSM; 2^log_threads independent sub-transforms, then the
 * joined butterfly levels, for the FFT).
 *
 * Parity: cross-checked bit-for-bit against oracle/bn254.py (big-integer spec) and
 * against the vectors captured from the reference's own compiled prover
 * (tests/golden/, see oracle/wasm/).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

typedef struct {
    fe p;        /* modulus */
    uint64_t inv; /* -p^{-1} mod 2^64 */
    fe one;      /* R mod p */
    fe r2;       /* R^2 mod p */
} field;

static const field FR = {
    {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}},
    0xc2e1f593efffffffULL,
    {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}},
    {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}},
};
static const field FQ = {
    {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}},
    0x87d20782e4866389ULL,
    {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
    {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}},
};

static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe *a, const fe *b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int fe_geq(const fe *a, const fe *b) {
    for (int i = 3; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static inline uint64_t fe_sub_raw(fe *r, const fe *a, const fe *b) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a->l[i] - b->l[i] - (uint64_t)br;
        r->l[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
    return (uint64_t)br;
}
static inline void f_add(const field *F, fe *r, const fe *a, const fe *b) {
    u128 c = 0;
    fe t;
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        t.l[i] = (uint64_t)c;
        c >>= 64;
    }
    /* both moduli are < 2^254 so no carry out of 256 bits */
    if (fe_geq(&t, &F->p)) fe_sub_raw(r, &t, &F->p); else *r = t;
}
static inline void f_sub(const field *F, fe *r, const fe *a, const fe *b) {
    fe t;
    if (fe_sub_raw(&t, a, b)) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t.l[i] + F->p.l[i];
            t.l[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    *r = t;
}
static inline void f_dbl(const field *F, fe *r, const fe *a) { f_add(F, r, a, a); }
static inline void f_neg(const field *F, fe *r, const fe *a) {
    if (fe_is_zero(a)) { *r = *a; return; }
    fe_sub_raw(r, &F->p, a);
}
/* Montgomery multiplication, CIOS, 4 x 64-bit limbs. */
static inline void f_mul(const field *F, fe *r, const fe *a, const fe *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->p.l[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * F->p.l[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe o = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fe_geq(&o, &F->p)) fe_sub_raw(&o, &o, &F->p);
    *r = o;
}
static inline void f_sqr(const field *F, fe *r, const fe *a) { f_mul(F, r, a, a); }
static void f_pow(const field *F, fe *r, const fe *a, const uint64_t e[4]) {
    fe acc = F->one;
    for (int i = 255; i >= 0; i--) {
        f_sqr(F, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) f_mul(F, &acc, &acc, a);
    }
    *r = acc;
}
static void f_inv(const field *F, fe *r, const fe *a) {
    fe e = F->p;
    e.l[0] -= 2; /* p is odd and p.l[0] >= 2 */
    f_pow(F, r, a, e.l);
}
static void f_from_u64(const field *F, fe *r, uint64_t v) {
    fe t = {{v, 0, 0, 0}};
    f_mul(F, r, &t, &F->r2);
}
/* Montgomery -> canonical (PrimeField::to_repr): multiply by 1. */
static void f_to_canonical(const field *F, fe *r, const fe *a) {
    fe one = {{1, 0, 0, 0}};
    f_mul(F, r, a, &one);
}

/* ---------------------------------------------------------------- G1, Jacobian, a = 0, b = 3 */
typedef struct { fe x, y; } g1a;    /* identity = (0,0) */
typedef struct { fe x, y, z; } g1j; /* identity: z = 0 */

static inline int g1a_is_identity(const g1a *p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void g1j_identity(g1j *p) {
    memset(p, 0, sizeof *p);
    p->y = FQ.one; /* (0, R, 0), as G1::identity() */
}
static void g1j_double(g1j *r, const g1j *p) {
    if (fe_is_zero(&p->z)) { *r = *p; return; }
    fe a, b, c, d, e, f, t, x3, y3, z3;
    f_sqr(&FQ, &a, &p->x);
    f_sqr(&FQ, &b, &p->y);
    f_sqr(&FQ, &c, &b);
    f_add(&FQ, &d, &p->x, &b);
    f_sqr(&FQ, &d, &d);
    f_sub(&FQ, &d, &d, &a);
    f_sub(&FQ, &d, &d, &c);
    f_dbl(&FQ, &d, &d);
    f_dbl(&FQ, &e, &a);
    f_add(&FQ, &e, &e, &a);
    f_sqr(&FQ, &f, &e);
    f_mul(&FQ, &z3, &p->z, &p->y);
    f_dbl(&FQ, &z3, &z3);
    f_dbl(&FQ, &t, &d);
    f_sub(&FQ, &x3, &f, &t);
    f_dbl(&FQ, &c, &c);
    f_dbl(&FQ, &c, &c);
    f_dbl(&FQ, &c, &c);
    f_sub(&FQ, &t, &d, &x3);
    f_mul(&FQ, &y3, &e, &t);
    f_sub(&FQ, &y3, &y3, &c);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_add_mixed(g1j *r, const g1j *p, const g1a *q) {
    if (g1a_is_identity(q)) { *r = *p; return; }
    if (fe_is_zero(&p->z)) { r->x = q->x; r->y = q->y; r->z = FQ.one; return; }
    fe z1z1, u2, s2, h, hh, i, j, rr, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z);
    f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s2, &q->y, &p->z);
    f_mul(&FQ, &s2, &s2, &z1z1);
    if (fe_eq(&u2, &p->x)) {
        if (fe_eq(&s2, &p->y)) { g1j_double(r, p); return; }
        g1j_identity(r);
        return;
    }
    f_sub(&FQ, &h, &u2, &p->x);
    f_sqr(&FQ, &hh, &h);
    f_dbl(&FQ, &i, &hh);
    f_dbl(&FQ, &i, &i);
    f_mul(&FQ, &j, &h, &i);
    f_sub(&FQ, &rr, &s2, &p->y);
    f_dbl(&FQ, &rr, &rr);
    f_mul(&FQ, &v, &p->x, &i);
    f_sqr(&FQ, &x3, &rr);
    f_sub(&FQ, &x3, &x3, &j);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &x3, &x3, &v);
    f_mul(&FQ, &j, &p->y, &j);
    f_dbl(&FQ, &j, &j);
    f_sub(&FQ, &t, &v, &x3);
    f_mul(&FQ, &y3, &rr, &t);
    f_sub(&FQ, &y3, &y3, &j);
    f_add(&FQ, &z3, &p->z, &h);
    f_sqr(&FQ, &z3, &z3);
    f_sub(&FQ, &z3, &z3, &z1z1);
    f_sub(&FQ, &z3, &z3, &hh);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_add(g1j *r, const g1j *p, const g1j *q) {
    if (fe_is_zero(&p->z)) { *r = *q; return; }
    if (fe_is_zero(&q->z)) { *r = *p; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z);
    f_sqr(&FQ, &z2z2, &q->z);
    f_mul(&FQ, &u1, &p->x, &z2z2);
    f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s1, &p->y, &q->z);
    f_mul(&FQ, &s1, &s1, &z2z2);
    f_mul(&FQ, &s2, &q->y, &p->z);
    f_mul(&FQ, &s2, &s2, &z1z1);
    if (fe_eq(&u1, &u2)) {
        if (fe_eq(&s1, &s2)) { g1j_double(r, p); return; }
        g1j_identity(r);
        return;
    }
    f_sub(&FQ, &h, &u2, &u1);
    f_dbl(&FQ, &i, &h);
    f_sqr(&FQ, &i, &i);
    f_mul(&FQ, &j, &h, &i);
    f_sub(&FQ, &rr, &s2, &s1);
    f_dbl(&FQ, &rr, &rr);
    f_mul(&FQ, &v, &u1, &i);
    f_sqr(&FQ, &x3, &rr);
    f_sub(&FQ, &x3, &x3, &j);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &x3, &x3, &v);
    f_mul(&FQ, &s1, &s1, &j);
    f_dbl(&FQ, &s1, &s1);
    f_sub(&FQ, &t, &v, &x3);
    f_mul(&FQ, &y3, &rr, &t);
    f_sub(&FQ, &y3, &y3, &s1);
    f_add(&FQ, &z3, &p->z, &q->z);
    f_sqr(&FQ, &z3, &z3);
    f_sub(&FQ, &z3, &z3, &z1z1);
    f_sub(&FQ, &z3, &z3, &z2z2);
    f_mul(&FQ, &z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}

/* The reference's G1 {x,y,z} is HOMOGENEOUS projective (x = X/Z, y = Y/Z): established by executing
 * its compiled prover (oracle/wasm).  This file computes in Jacobian coordinates internally (the
 * group element is what matters) and converts at the API boundary so that buffers have the
 * reference's meaning. */
static void jac_to_hom(uint64_t out[12], const g1j *p) {
    g1j r;
    if (fe_is_zero(&p->z)) { g1j_identity(&r); }
    else {
        fe z2;
        f_sqr(&FQ, &z2, &p->z);
        f_mul(&FQ, &r.x, &p->x, &p->z);  /* X/Z^2 = X*Z / Z^3 */
        r.y = p->y;                      /* Y/Z^3 */
        f_mul(&FQ, &r.z, &z2, &p->z);
    }
    memcpy(out, &r, sizeof r);
}
static void hom_to_jac(g1j *r, const uint64_t in[12]) {
    g1j p;
    memcpy(&p, in, sizeof p);
    if (fe_is_zero(&p.z)) { g1j_identity(r); return; }
    /* (X:Y:Z) -> Jacobian (X*Z, Y*Z^2, Z) */
    fe z2;
    f_sqr(&FQ, &z2, &p.z);
    f_mul(&FQ, &r->x, &p.x, &p.z);
    f_mul(&FQ, &r->y, &p.y, &z2);
    r->z = p.z;
}

/* ---------------------------------------------------------------- multiexp_serial (arithmetic.rs:28-140) */
enum { B_NONE = 0, B_AFFINE = 1, B_PROJ = 2 };
typedef struct { int tag; g1j p; } bucket; /* Affine uses p.x,p.y */

static inline unsigned get_at(unsigned segment, unsigned c, const uint8_t bytes[32]) {
    unsigned skip_bits = segment * c;
    unsigned skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0};
    unsigned avail = 32 - skip_bytes;
    memcpy(v, bytes + skip_bytes, avail < 8 ? avail : 8);
    uint64_t tmp;
    memcpy(&tmp, v, 8); /* little-endian host */
    tmp >>= skip_bits - skip_bytes * 8;
    tmp %= (1ull << c);
    return (unsigned)tmp;
}

static unsigned window_rule(size_t n) {
    if (n < 4) return 1;
    if (n < 32) return 3;
    return (unsigned)ceil(log((double)(uint32_t)n));
}

static void multiexp_serial(const fe *coeffs, const g1a *bases, size_t n, g1j *acc) {
    fe *repr = (fe *)malloc(n * sizeof(fe) + 8);
    for (size_t i = 0; i < n; i++) f_to_canonical(&FR, &repr[i], &coeffs[i]);
    unsigned c = window_rule(n);
    unsigned segments = 256 / c + 1;
    size_t nb = ((size_t)1 << c) - 1;
    bucket *buckets = (bucket *)malloc(nb * sizeof(bucket));
    for (int seg = (int)segments - 1; seg >= 0; seg--) {
        for (unsigned d = 0; d < c; d++) g1j_double(acc, acc);
        for (size_t b = 0; b < nb; b++) buckets[b].tag = B_NONE;
        for (size_t i = 0; i < n; i++) {
            unsigned w = get_at((unsigned)seg, c, (const uint8_t *)&repr[i]);
            if (w == 0) continue;
            bucket *bk = &buckets[w - 1]; /* arithmetic.rs:102 */
            if (bk->tag == B_NONE) {
                bk->tag = B_AFFINE;
                bk->p.x = bases[i].x;
                bk->p.y = bases[i].y;
            } else if (bk->tag == B_AFFINE) {
                g1a a = {bk->p.x, bk->p.y};
                g1j t;
                if (g1a_is_identity(&a)) g1j_identity(&t);
                else { t.x = a.x; t.y = a.y; t.z = FQ.one; }
                g1j_add_mixed(&bk->p, &t, &bases[i]);
                bk->tag = B_PROJ;
            } else {
                g1j_add_mixed(&bk->p, &bk->p, &bases[i]);
            }
        }
        g1j running;
        g1j_identity(&running);
        for (size_t b = nb; b-- > 0;) {
            bucket *bk = &buckets[b];
            if (bk->tag == B_AFFINE) {
                g1a a = {bk->p.x, bk->p.y};
                g1j_add_mixed(&running, &running, &a);
            } else if (bk->tag == B_PROJ) {
                g1j_add(&running, &running, &bk->p);
            }
            g1j_add(acc, acc, &running);
        }
    }
    free(buckets);
    free(repr);
}

typedef struct { const fe *c; const g1a *b; size_t n; g1j acc; } msm_job;
static void *msm_worker(void *arg) {
    msm_job *j = (msm_job *)arg;
    g1j_identity(&j->acc);
    multiexp_serial(j->c, j->b, j->n, &j->acc);
    return NULL;
}

/* best_multiexp (arithmetic.rs:147-180). out = homogeneous projective, Montgomery, 12 x u64. */
void h2ref_best_multiexp(const uint64_t *coeffs, const uint64_t *bases, size_t n, int num_threads, uint64_t out[12]) {
    g1j acc;
    g1j_identity(&acc);
    if (num_threads < 1) num_threads = 1;
    if (n > (size_t)num_threads) {
        size_t chunk = n / (size_t)num_threads; /* arithmetic.rs:152 */
        size_t num_chunks = (n + chunk - 1) / chunk;
        msm_job *jobs = (msm_job *)malloc(num_chunks * sizeof(msm_job));
        pthread_t *th = (pthread_t *)malloc(num_chunks * sizeof(pthread_t));
        for (size_t i = 0; i < num_chunks; i++) {
            size_t s = i * chunk, e = s + chunk > n ? n : s + chunk;
            jobs[i].c = (const fe *)coeffs + s;
            jobs[i].b = (const g1a *)bases + s;
            jobs[i].n = e - s;
            pthread_create(&th[i], NULL, msm_worker, &jobs[i]);
        }
        for (size_t i = 0; i < num_chunks; i++) {
            pthread_join(th[i], NULL);
            g1j_add(&acc, &acc, &jobs[i].acc);
        }
        free(th);
        free(jobs);
    } else {
        multiexp_serial((const fe *)coeffs, (const g1a *)bases, n, &acc);
    }
    jac_to_hom(out, &acc);
}

/* projective (12 x u64) -> affine (8 x u64), identity -> (0,0); what callers do before the transcript. */
void h2ref_g1_to_affine(const uint64_t in[12], uint64_t out[8]) {
    const g1j *p = (const g1j *)in;
    g1a r;
    if (fe_is_zero(&p->z)) { memset(&r, 0, sizeof r); }
    else {
        fe zi;
        f_inv(&FQ, &zi, &p->z);
        f_mul(&FQ, &r.x, &p->x, &zi);
        f_mul(&FQ, &r.y, &p->y, &zi);
    }
    memcpy(out, &r, sizeof r);
}

/* ---------------------------------------------------------------- parallelize (arithmetic.rs:~395-425) */
typedef void (*par_fn)(fe *chunk, size_t len, size_t start, void *ctx);
typedef struct { par_fn f; fe *v; size_t len, start; void *ctx; } par_job;
static void *par_worker(void *arg) {
    par_job *j = (par_job *)arg;
    j->f(j->v, j->len, j->start, j->ctx);
    return NULL;
}
static void parallelize(fe *v, size_t n, int num_threads, par_fn f, void *ctx) {
    if (num_threads < 1) num_threads = 1;
    size_t chunk = n / (size_t)num_threads;
    if (chunk < (size_t)num_threads || num_threads == 1) { f(v, n, 0, ctx); return; }
    size_t njobs = (n + chunk - 1) / chunk;
    par_job *jobs = (par_job *)malloc(njobs * sizeof(par_job));
    pthread_t *th = (pthread_t *)malloc(njobs * sizeof(pthread_t));
    for (size_t i = 0; i < njobs; i++) {
        size_t s = i * chunk, e = s + chunk > n ? n : s + chunk;
        jobs[i] = (par_job){f, v + s, e - s, s, ctx};
        pthread_create(&th[i], NULL, par_worker, &jobs[i]);
    }
    for (size_t i = 0; i < njobs; i++) pthread_join(th[i], NULL);
    free(th);
    free(jobs);
}

/* ---------------------------------------------------------------- best_fft (arithmetic.rs:185-290) */
static inline size_t bitreverse(size_t n, unsigned l) {
    size_t r = 0;
    for (unsigned i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
static void butterfly_level(fe *a, size_t n, size_t chunk, size_t twiddle_chunk, const fe *tw) {
    size_t half = chunk / 2;
    for (size_t s = 0; s < n; s += chunk) {
        fe *left = a + s, *right = a + s + half;
        fe t = right[0];
        f_sub(&FR, &right[0], &left[0], &t);
        f_add(&FR, &left[0], &left[0], &t);
        for (size_t i = 1; i < half; i++) {
            f_mul(&FR, &t, &right[i], &tw[i * twiddle_chunk]);
            f_sub(&FR, &right[i], &left[i], &t);
            f_add(&FR, &left[i], &left[i], &t);
        }
    }
}
static void fft_serial_levels(fe *a, size_t n, unsigned log_n, size_t top_twiddle_chunk, const fe *tw) {
    /* levels chunk = 2 .. n of a length-n block whose final level uses top_twiddle_chunk */
    if (log_n == 0) return;
    size_t chunk = 2, tc = top_twiddle_chunk << (log_n - 1);
    for (unsigned l = 0; l < log_n; l++) {
        butterfly_level(a, n, chunk, tc, tw);
        chunk *= 2;
        tc /= 2;
    }
}
typedef struct { fe *a; size_t n; unsigned log_n; size_t tc; const fe *tw; } sub_job;
static void *sub_worker(void *arg) {
    sub_job *j = (sub_job *)arg;
    fft_serial_levels(j->a, j->n, j->log_n, j->tc, j->tw);
    return NULL;
}
typedef struct { fe *a; size_t half, tc, i0, i1, blocks, chunk; const fe *tw; } lvl_job;
static void *lvl_worker(void *arg) {
    lvl_job *j = (lvl_job *)arg;
    for (size_t b = 0; b < j->blocks; b++) {
        fe *left = j->a + b * j->chunk, *right = left + j->half;
        for (size_t i = j->i0; i < j->i1; i++) {
            fe t;
            if (i == 0) t = right[0]; else f_mul(&FR, &t, &right[i], &j->tw[i * j->tc]);
            f_sub(&FR, &right[i], &left[i], &t);
            f_add(&FR, &left[i], &left[i], &t);
        }
    }
    return NULL;
}

void h2ref_best_fft(uint64_t *a_, const uint64_t omega_[4], uint32_t log_n, int num_threads) {
    fe *a = (fe *)a_;
    fe omega;
    memcpy(&omega, omega_, sizeof omega);
    size_t n = (size_t)1 << log_n;
    if (num_threads < 1) num_threads = 1;
    unsigned log_threads = 0;
    while ((2u << log_threads) <= (unsigned)num_threads) log_threads++;
    for (size_t k = 0; k < n; k++) { /* arithmetic.rs:204 */
        size_t rk = bitreverse(k, log_n);
        if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    size_t nt = n / 2 ? n / 2 : 1;
    fe *tw = (fe *)malloc(nt * sizeof(fe));
    fe w = FR.one;
    for (size_t i = 0; i < n / 2; i++) { tw[i] = w; f_mul(&FR, &w, &w, &omega); }
    if (log_n <= log_threads || log_threads == 0) {
        fft_serial_levels(a, n, log_n, 1, tw);
    } else {
        /* recursive_butterfly_arithmetic with rayon::join: 2^log_threads independent
         * halves-of-halves, then the joined top levels. */
        unsigned sub_log = log_n - log_threads;
        size_t subs = (size_t)1 << log_threads, sub_n = (size_t)1 << sub_log;
        sub_job *jobs = (sub_job *)malloc(subs * sizeof(sub_job));
        pthread_t *th = (pthread_t *)malloc(subs * sizeof(pthread_t));
        for (size_t s = 0; s < subs; s++) {
            jobs[s] = (sub_job){a + s * sub_n, sub_n, sub_log, subs, tw};
            pthread_create(&th[s], NULL, sub_worker, &jobs[s]);
        }
        for (size_t s = 0; s < subs; s++) pthread_join(th[s], NULL);
        lvl_job *lj = (lvl_job *)malloc(subs * sizeof(lvl_job));
        size_t chunk = sub_n * 2, tc = subs / 2;
        for (unsigned l = 0; l < log_threads; l++) {
            size_t half = chunk / 2, per = half / subs ? half / subs : half;
            size_t njobs = half / per;
            for (size_t s = 0; s < njobs; s++) {
                lj[s] = (lvl_job){a, half, tc, s * per, (s + 1) * per, n / chunk, chunk, tw};
                pthread_create(&th[s], NULL, lvl_worker, &lj[s]);
            }
            for (size_t s = 0; s < njobs; s++) pthread_join(th[s], NULL);
            chunk *= 2;
            tc /= 2;
        }
        free(lj);
        free(th);
        free(jobs);
    }
    free(tw);
}

/* ---------------------------------------------------------------- EvaluationDomain (poly/domain.rs) */
typedef struct {
    uint32_t k, extended_k, j, n_t; /* n_t = 2^(extended_k-k) */
    uint64_t omega[4], omega_inv[4], extended_omega[4], extended_omega_inv[4];
    uint64_t g_coset[4], g_coset_inv[4], ifft_divisor[4], extended_ifft_divisor[4];
    uint64_t t_evaluations[8 * 4]; /* inverted (X^n - 1) on the coset; up to ext_k - k = 3 */
} h2ref_domain;

static const fe FR_ROOT_OF_UNITY_CANON = {{0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL}};
/* Fr::ZETA (halo2curves 0.3.2); SURVEY.md quotes its square, the reference's recorded coeff_to_extended calls use this one */
static const fe FR_ZETA_CANON = {{0x8b17ea66b99c90ddULL, 0x5bfc41088d8daaa7ULL, 0xb3c4d79d41a91758ULL, 0x0ULL}};
#define FR_S 28

int h2ref_domain_new(uint32_t j, uint32_t k, h2ref_domain *d) {
    memset(d, 0, sizeof *d);
    uint64_t qdeg = j - 1, n = 1ull << k;
    uint32_t ek = k;
    while ((1ull << ek) < n * qdeg) ek++;
    if (ek > FR_S || ek - k > 3) return -1;
    fe w, zeta, t;
    f_mul(&FR, &w, &FR_ROOT_OF_UNITY_CANON, &FR.r2);
    f_mul(&FR, &zeta, &FR_ZETA_CANON, &FR.r2);
    for (uint32_t i = ek; i < FR_S; i++) f_sqr(&FR, &w, &w);
    fe ext_w = w;
    for (uint32_t i = k; i < ek; i++) f_sqr(&FR, &w, &w);
    fe omega = w, zeta2;
    f_sqr(&FR, &zeta2, &zeta);
    d->k = k; d->extended_k = ek; d->j = j; d->n_t = 1u << (ek - k);
    memcpy(d->omega, &omega, 32);
    memcpy(d->extended_omega, &ext_w, 32);
    f_inv(&FR, &t, &omega); memcpy(d->omega_inv, &t, 32);
    f_inv(&FR, &t, &ext_w); memcpy(d->extended_omega_inv, &t, 32);
    memcpy(d->g_coset, &zeta, 32);
    memcpy(d->g_coset_inv, &zeta2, 32);
    uint64_t e[4] = {n, 0, 0, 0};
    fe orig, step, cur;
    f_pow(&FR, &orig, &zeta, e);
    f_pow(&FR, &step, &ext_w, e);
    cur = orig;
    uint32_t cnt = 0;
    do {
        if (cnt >= 8) return -2;
        fe v;
        f_sub(&FR, &v, &cur, &FR.one);
        f_inv(&FR, &v, &v);
        memcpy(&d->t_evaluations[4 * cnt], &v, 32);
        cnt++;
        f_mul(&FR, &cur, &cur, &step);
    } while (!fe_eq(&cur, &orig));
    if (cnt != d->n_t) return -3; /* domain.rs:101 */
    f_from_u64(&FR, &t, 1ull << k); f_inv(&FR, &t, &t); memcpy(d->ifft_divisor, &t, 32);
    f_from_u64(&FR, &t, 1ull << ek); f_inv(&FR, &t, &t); memcpy(d->extended_ifft_divisor, &t, 32);
    return 0;
}

static void scale_fn(fe *v, size_t len, size_t start, void *ctx) {
    (void)start;
    const fe *s = (const fe *)ctx;
    for (size_t i = 0; i < len; i++) f_mul(&FR, &v[i], &v[i], s);
}
static void zeta_fn(fe *v, size_t len, size_t start, void *ctx) {
    const fe *cp = (const fe *)ctx; /* [2] */
    size_t index = start;
    for (size_t i = 0; i < len; i++, index++) {
        size_t m = index % 3;
        if (m != 0) f_mul(&FR, &v[i], &v[i], &cp[m - 1]);
    }
}
static void ifft(fe *a, const uint64_t omega_inv[4], uint32_t log_n, const uint64_t divisor[4], int threads) {
    h2ref_best_fft((uint64_t *)a, omega_inv, log_n, threads);
    parallelize(a, (size_t)1 << log_n, threads, scale_fn, (void *)divisor);
}
static void distribute_powers_zeta(const h2ref_domain *d, fe *a, size_t n, int into_coset, int threads) {
    fe cp[2];
    memcpy(&cp[0], into_coset ? d->g_coset : d->g_coset_inv, 32);
    memcpy(&cp[1], into_coset ? d->g_coset_inv : d->g_coset, 32);
    parallelize(a, n, threads, zeta_fn, cp);
}

/* domain.rs:227 */
void h2ref_lagrange_to_coeff(const h2ref_domain *d, uint64_t *a, int threads) {
    ifft((fe *)a, d->omega_inv, d->k, d->ifft_divisor, threads);
}
/* domain.rs:244. in: 2^k, out: 2^extended_k */
void h2ref_coeff_to_extended(const h2ref_domain *d, const uint64_t *in, uint64_t *out, int threads) {
    size_t n = (size_t)1 << d->k, en = (size_t)1 << d->extended_k;
    memcpy(out, in, n * 32);
    distribute_powers_zeta(d, (fe *)out, n, 1, threads);
    memset(out + 4 * n, 0, (en - n) * 32);
    h2ref_best_fft(out, d->extended_omega, d->extended_k, threads);
}
/* domain.rs:311. a: 2^extended_k (clobbered), out: n*(j-1) */
void h2ref_extended_to_coeff(const h2ref_domain *d, uint64_t *a, uint64_t *out, int threads) {
    size_t en = (size_t)1 << d->extended_k;
    ifft((fe *)a, d->extended_omega_inv, d->extended_k, d->extended_ifft_divisor, threads);
    distribute_powers_zeta(d, (fe *)a, en, 0, threads);
    memcpy(out, a, ((size_t)(d->j - 1) << d->k) * 32);
}
typedef struct { const fe *t; size_t m; } dv_ctx;
static void dv_fn(fe *v, size_t len, size_t start, void *ctx) {
    const dv_ctx *c = (const dv_ctx *)ctx;
    for (size_t i = 0; i < len; i++) f_mul(&FR, &v[i], &v[i], &c->t[(start + i) % c->m]);
}
void h2ref_divide_by_vanishing_poly(const h2ref_domain *d, uint64_t *a, int threads) {
    dv_ctx c = {(const fe *)d->t_evaluations, d->n_t};
    parallelize((fe *)a, (size_t)1 << d->extended_k, threads, dv_fn, &c);
}

/* ---------------------------------------------------------------- small helpers for tests / input generation */
void h2ref_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) f_mul(&FR, (fe *)r + i, (const fe *)a + i, (const fe *)b + i);
}
void h2ref_fq_mul(const uint64_t *a, const uint64_t *b, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) f_mul(&FQ, (fe *)r + i, (const fe *)a + i, (const fe *)b + i);
}
/* canonical -> Montgomery, in place; which = 0 Fr, 1 Fq */
void h2ref_to_mont(uint64_t *a, size_t n, int which) {
    const field *F = which ? &FQ : &FR;
    for (size_t i = 0; i < n; i++) f_mul(F, (fe *)a + i, (fe *)a + i, &F->r2);
}
void h2ref_from_mont(uint64_t *a, size_t n, int which) {
    const field *F = which ? &FQ : &FR;
    for (size_t i = 0; i < n; i++) f_to_canonical(F, (fe *)a + i, (fe *)a + i);
}

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* Same stream as oracle/bn254.py random_fr: Montgomery-form output. */
void h2ref_random_fr(uint64_t *out, size_t n, uint64_t seed) {
    uint64_t s = seed;
    size_t i = 0;
    while (i < n) {
        fe v;
        for (int k = 0; k < 4; k++) v.l[k] = splitmix64(&s);
        v.l[3] &= (1ull << 62) - 1;
        if (fe_geq(&v, &FR.p)) continue;
        f_mul(&FR, (fe *)out + i, &v, &FR.r2);
        i++;
    }
}
/* Same stream as oracle/bn254.py random_g1 (try-and-increment, sign from a PRNG bit). */
void h2ref_random_g1(uint64_t *out, size_t n, uint64_t seed) {
    uint64_t s = seed;
    /* (q+1)/4 */
    fe e = FQ.p;
    e.l[0] += 1; /* no carry: low limb of q is ...fd47 */
    for (int k = 0; k < 4; k++) e.l[k] = (e.l[k] >> 2) | (k < 3 ? e.l[k + 1] << 62 : 0);
    fe three;
    f_from_u64(&FQ, &three, 3);
    for (size_t i = 0; i < n; i++) {
        fe x;
        for (int k = 0; k < 4; k++) x.l[k] = splitmix64(&s);
        x.l[3] &= (1ull << 62) - 1;
        while (fe_geq(&x, &FQ.p)) fe_sub_raw(&x, &x, &FQ.p);
        unsigned sign = (unsigned)(splitmix64(&s) & 1);
        fe xm, y, rhs, y2;
        f_mul(&FQ, &xm, &x, &FQ.r2);
        for (;;) {
            f_sqr(&FQ, &rhs, &xm);
            f_mul(&FQ, &rhs, &rhs, &xm);
            f_add(&FQ, &rhs, &rhs, &three);
            f_pow(&FQ, &y, &rhs, e.l);
            f_sqr(&FQ, &y2, &y);
            if (fe_eq(&y2, &rhs)) break;
            f_add(&FQ, &xm, &xm, &FQ.one);
        }
        fe yc;
        f_to_canonical(&FQ, &yc, &y);
        if ((yc.l[0] & 1) != sign) f_neg(&FQ, &y, &y);
        memcpy(out + 8 * i, &xm, 32);
        memcpy(out + 8 * i + 4, &y, 32);
    }
}
/* out[i] = k[i] * P (double-and-add) -- slow, for building structured test bases. */
void h2ref_g1_scalar_mul(const uint64_t p_[8], const uint64_t k_mont[4], uint64_t out[12]) {
    fe k;
    f_to_canonical(&FR, &k, (const fe *)k_mont);
    g1j acc;
    g1j_identity(&acc);
    const g1a *p = (const g1a *)p_;
    for (int i = 255; i >= 0; i--) {
        g1j_double(&acc, &acc);
        if ((k.l[i / 64] >> (i % 64)) & 1) g1j_add_mixed(&acc, &acc, p);
    }
    jac_to_hom(out, &acc);
}
void h2ref_g1_add(const uint64_t a[12], const uint64_t b[12], uint64_t out[12]) {
    g1j r, x, y;
    hom_to_jac(&x, a);
    hom_to_jac(&y, b);
    g1j_add(&r, &x, &y);
    jac_to_hom(out, &r);
}

/* n DISTINCT points B + i*G, i in [0, n), B = [seed-derived scalar] G: synthetic MSM bases at benchmark sizes, where
 * try-and-increment (h2ref_random_g1, ~30 us per point) is too slow.  Each thread owns a contiguous range and walks
 * K interleaved chains by adding D = K*G to all K chain heads with one shared inversion per step (Montgomery's
 * trick), ~7 products per point.  Test / benchmark infrastructure only: the cost of an MSM does not depend on
 * which distinct points it is given. */
typedef struct { uint64_t *out; size_t lo, hi; fe start_scalar; } prog_job;
static void g1_affine_from_jac(g1a *r, const g1j *p) {
    uint64_t hom[12], aff[8];
    jac_to_hom(hom, p);
    h2ref_g1_to_affine(hom, aff);
    memcpy(r, aff, 64);
}
static void g1a_scalar_mul_canonical(g1a *r, const g1a *p, const fe *k) {
    g1j acc;
    g1j_identity(&acc);
    for (int i = 255; i >= 0; i--) {
        g1j_double(&acc, &acc);
        if ((k->l[i / 64] >> (i % 64)) & 1) g1j_add_mixed(&acc, &acc, p);
    }
    g1_affine_from_jac(r, &acc);
}
static void *prog_worker(void *arg) {
    prog_job *j = (prog_job *)arg;
    enum { K = 256 };
    const size_t len = j->hi - j->lo;
    if (len == 0) return NULL;
    g1a G;
    f_from_u64(&FQ, &G.x, 1);
    f_from_u64(&FQ, &G.y, 2);
    /* chain heads Q_t = [start + lo + t] G, t < K; D = [K] G */
    g1a Q[K], D;
    fe s = j->start_scalar, lo_fe = {{(uint64_t)j->lo, 0, 0, 0}}, one = {{1, 0, 0, 0}}, kk = {{K, 0, 0, 0}};
    /* canonical integer arithmetic on small values: start < 2^200, so no reduction is ever needed */
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) { c += (unsigned __int128)s.l[i] + lo_fe.l[i]; s.l[i] = (uint64_t)c; c >>= 64; }
    for (size_t t = 0; t < K && t < len; t++) {
        g1a_scalar_mul_canonical(&Q[t], &G, &s);
        c = 0;
        for (int i = 0; i < 4; i++) { c += (unsigned __int128)s.l[i] + one.l[i]; s.l[i] = (uint64_t)c; c >>= 64; }
    }
    g1a_scalar_mul_canonical(&D, &G, &kk);
    fe den[K], pre[K];
    for (size_t base = 0; base < len; base += K) {
        const size_t cnt = len - base < K ? len - base : K;
        memcpy(j->out + 8 * (j->lo + base), Q, cnt * 64);
        if (base + K >= len) break;
        /* Q_t += D for every chain that still has a successor */
        const size_t nxt = len - (base + K) < K ? len - (base + K) : K;
        for (size_t t = 0; t < nxt; t++) {
            f_sub(&FQ, &den[t], &D.x, &Q[t].x);
            if (t == 0) pre[0] = den[0]; else f_mul(&FQ, &pre[t], &pre[t - 1], &den[t]);
        }
        fe inv;
        f_inv(&FQ, &inv, &pre[nxt - 1]);
        for (size_t t = nxt; t-- > 0;) {
            fe di;
            if (t) { f_mul(&FQ, &di, &inv, &pre[t - 1]); f_mul(&FQ, &inv, &inv, &den[t]); } else di = inv;
            fe lam, x3, y3, tmp;
            f_sub(&FQ, &tmp, &D.y, &Q[t].y);
            f_mul(&FQ, &lam, &tmp, &di);
            f_sqr(&FQ, &x3, &lam);
            f_sub(&FQ, &x3, &x3, &Q[t].x);
            f_sub(&FQ, &x3, &x3, &D.x);
            f_sub(&FQ, &tmp, &Q[t].x, &x3);
            f_mul(&FQ, &y3, &lam, &tmp);
            f_sub(&FQ, &y3, &y3, &Q[t].y);
            Q[t].x = x3;
            Q[t].y = y3;
        }
    }
    return NULL;
}
void h2ref_progression_g1(uint64_t *out, size_t n, uint64_t seed, int num_threads) {
    if (num_threads < 1) num_threads = 1;
    if ((size_t)num_threads > n / 1024 + 1) num_threads = (int)(n / 1024 + 1);
    fe start = {{0, 0, 0, 0}};
    uint64_t st = seed;
    for (int i = 0; i < 3; i++) start.l[i] = splitmix64(&st);  /* < 2^192 */
    pthread_t th[256];
    prog_job jobs[256];
    if (num_threads > 256) num_threads = 256;
    const size_t per = (n + num_threads - 1) / num_threads;
    for (int t = 0; t < num_threads; t++) {
        jobs[t].out = out;
        jobs[t].lo = (size_t)t * per < n ? (size_t)t * per : n;
        jobs[t].hi = (size_t)(t + 1) * per < n ? (size_t)(t + 1) * per : n;
        jobs[t].start_scalar = start;
        pthread_create(&th[t], NULL, prog_worker, &jobs[t]);
    }
    for (int t = 0; t < num_threads; t++) pthread_join(th[t], NULL);
}
