// Host harness around scripts/field52.cuh (the device's FP64-pipe field arithmetic compiled
// for the CPU with std::fma under FE_TOWARDZERO) so tests/test_field52.py can check it against big integers.
#include <cfenv>
#include "../../scripts/field52.cuh"
using namespace h2b;

namespace {
struct RZ {
    int old;
    RZ() : old(fegetround()) { fesetround(FE_TOWARDZERO); }
    ~RZ() { fesetround(old); }
};
N52 ld(const uint64_t *p) { N52 r; for (int i = 0; i < 5; i++) r.l[i] = p[i]; return r; }
void st(uint64_t *p, const N52 &a) { for (int i = 0; i < 5; i++) p[i] = a.l[i]; }
}  // namespace

extern "C" {
void f52h_mul(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    RZ rz;
    for (size_t i = 0; i < n; i++) st(out + 5 * i, f52_mul(f52_to_d(ld(a + 5 * i)), f52_to_d(ld(b + 5 * i))));
}
void f52h_sqr(const uint64_t *a, uint64_t *out, size_t n) {
    RZ rz;
    for (size_t i = 0; i < n; i++) st(out + 5 * i, f52_sqr(f52_to_d(ld(a + 5 * i))));
}
void f52h_sub(int k, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        N52 x = ld(a + 5 * i), y = ld(b + 5 * i), r;
        switch (k) {
            case 2: r = f52_sub<2>(x, y); break;
            case 4: r = f52_sub<4>(x, y); break;
            case 6: r = f52_sub<6>(x, y); break;
            default: r = f52_sub<8>(x, y); break;
        }
        st(out + 5 * i, r);
    }
}
void f52h_sub_b_2c(const uint64_t *a, const uint64_t *b, const uint64_t *c, uint64_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) st(out + 5 * i, f52_sub_b_2c<4>(ld(a + 5 * i), ld(b + 5 * i), ld(c + 5 * i)));
}
void f52h_add(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) st(out + 5 * i, f52_add(ld(a + 5 * i), ld(b + 5 * i)));
}
void f52h_unpack(const uint32_t *w, uint64_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint32_t t[8];
        for (int j = 0; j < 8; j++) t[j] = w[8 * i + j];
        st(out + 5 * i, f52_unpack(t));
    }
}
void f52h_pack(const uint64_t *a, uint32_t *w, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint32_t t[8];
        f52_pack(ld(a + 5 * i), t);
        for (int j = 0; j < 8; j++) w[8 * i + j] = t[j];
    }
}
int f52h_is_zero(const uint64_t *a) { return f52_is_zero_mod_q_lt2q(ld(a)) ? 1 : 0; }
}
