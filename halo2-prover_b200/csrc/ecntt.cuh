// ecntt.cuh -- the G = G1 instantiation of best_fft, as g_to_lagrange uses it.
//
// halo2_proofs @6b43b6b src/arithmetic.rs (g_to_lagrange): best_fft over group elements with omega_inv, then
// every point times 1/2^k, then batch normalisation -- how ParamsKZG::{setup, from_parts} obtain
// g_lagrange from g (src/poly/kzg/commitment.rs:68-114).  Setup-time work: a plain
// radix-2 decimation-in-time transform over XYZZ points in HBM, one thread per butterfly, the twiddle
// applied as a 254-bit double-and-add.  Natural order in and out like best_fft.
#pragma once
#include "curve.cuh"

namespace h2b {

// [k] P for a canonical (non-Montgomery) scalar k < 2^254, most significant bit first
static __device__ __noinline__ XYZZ xyzz_scalar_mul(const XYZZ &p, const Fe &k) {
    XYZZ acc = xyzz_identity();
    if (xyzz_is_identity(p)) return acc;
    int top = -1;
#pragma unroll 1
    for (int i = 7; i >= 0 && top < 0; i--)
        if (k.l[i]) top = 32 * i + 31 - __clz(k.l[i]);
#pragma unroll 1
    for (int bit = top; bit >= 0; bit--) {
        acc = xyzz_dbl_ni(acc);
        if ((k.l[bit >> 5] >> (bit & 31)) & 1) xyzz_add(acc, p);
    }
    return acc;
}

__global__ void __launch_bounds__(128)
ec_ntt_load_kernel(const Affine *__restrict__ in, uint32_t log_n, XYZZ *__restrict__ work) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << log_n)) return;
    const uint32_t r = log_n ? __brev(i) >> (32 - log_n) : 0u;
    store_xyzz(&work[r], xyzz_from_affine(load_affine(&in[i])));
}

// stage s (butterfly span 2^s): (a, b) <- (a + w b, a - w b), w = omega^(j * n / 2^(s+1))
__global__ void __launch_bounds__(128)
ec_ntt_stage_kernel(XYZZ *__restrict__ a, uint32_t log_n, uint32_t s, const Fe *__restrict__ W) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (1u << log_n) / 2) return;
    const uint32_t half = 1u << s, j = t & (half - 1);
    const uint32_t i0 = ((t >> s) << (s + 1)) + j, i1 = i0 + half;
    XYZZ b = load_xyzz(&a[i1]);
    if (j != 0) b = xyzz_scalar_mul(b, Fr::from_mont(load_fe_ro(&W[(size_t)j << (log_n - 1 - s)])));
    const XYZZ x = load_xyzz(&a[i0]);
    XYZZ sum = x, dif = x;
    xyzz_add(sum, b);
    if (!xyzz_is_identity(b)) b.y = Fq::neg(b.y);
    xyzz_add(dif, b);
    store_xyzz(&a[i0], sum);
    store_xyzz(&a[i1], dif);
}

// every point times `scale`, normalised to affine ((0,0) for the identity)
__global__ void __launch_bounds__(128)
ec_ntt_finish_kernel(const XYZZ *__restrict__ a, uint32_t n, Fe scale, Affine *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const XYZZ p = xyzz_scalar_mul(load_xyzz(&a[i]), Fr::from_mont(scale));  // `scale` arrives in Montgomery form
    Affine r;
    if (xyzz_is_identity(p)) {
        r.x = Fq::zero();
        r.y = Fq::zero();
    } else {
        const Fe t = Fq::inv(Fq::mul(p.zz, p.zzz));
        r.x = Fq::mul(Fq::mul(p.x, t), p.zzz);  // X / ZZ
        r.y = Fq::mul(Fq::mul(p.y, t), p.zz);   // Y / ZZZ
    }
    store_fe(&out[i].x, r.x);
    store_fe(&out[i].y, r.y);
}

}  // namespace h2b
