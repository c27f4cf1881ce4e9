/* abi_c_test.c -- the C ABI of libh2b200.so driven from plain C with buffers laid out as the reference's Rust types
 * (halo2curves 0.3.2: Fr / Fq = [u64; 4] Montgomery limbs, G1Affine {x, y}, G1 {x, y, z}), the way
 * a `-sys` crate would pass them (tests/test_abi.py builds and runs this).
 *
 *   abi_c_test nogpu                 h2b_init must fail with H2B_ERR_CUDA and every compute entry point with H2B_ERR_STATE
 *   abi_c_test run <in.bin> <out.bin> read n, log_n, scalars, bases, omega, a[2^log_n]; write best_multiexp, commit,
 *                                    commit of a prefix, best_fft, lagrange_to_coeff, coeff_to_extended results
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/h2b200.h"

typedef struct { uint64_t l[4]; } Fr;
typedef struct { uint64_t l[4]; } Fq;
typedef struct { Fq x, y; } G1Affine;
typedef struct { Fq x, y, z; } G1;
_Static_assert(sizeof(Fr) == 32 && _Alignof(Fr) == 8, "Fr is [u64; 4]");
_Static_assert(sizeof(G1Affine) == 64, "G1Affine is two Fq");
_Static_assert(sizeof(G1) == 96, "G1 is three Fq");
_Static_assert(sizeof(h2b_domain) == 16 + 8 * 32 + 32 * 32 + 3 * 32, "h2b_domain layout");

/* a Rust reference to a struct / slice crosses the FFI as a pointer to its first limb */
#define P(x) ((uint64_t *)&(x))

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != H2B_OK) {                                                         \
            fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, h2b_last_error());        \
            return 1;                                                                \
        }                                                                            \
    } while (0)

static int expect(int got, int want, const char *what) {
    if (got != want) {
        fprintf(stderr, "%s: status %d, expected %d (%s)\n", what, got, want, h2b_last_error());
        return 1;
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "nogpu") == 0) {
        G1 out;
        Fr s = {{1, 0, 0, 0}};
        G1Affine b = {{{1, 0, 0, 0}}, {{2, 0, 0, 0}}};
        if (expect(h2b_best_multiexp(P(s), P(b), 1, P(out)), H2B_ERR_STATE, "best_multiexp before init")) return 1;
        if (expect(h2b_init(0), H2B_ERR_CUDA, "h2b_init without a GPU")) return 1;
        if (strlen(h2b_last_error()) == 0) return 1;
        if (expect(h2b_best_fft(P(s), P(s), 0), H2B_ERR_STATE, "best_fft without a context")) return 1;
        printf("nogpu ok: %s\n", h2b_last_error());
        return 0;
    }
    if (argc < 4 || strcmp(argv[1], "run") != 0) {
        fprintf(stderr, "usage: abi_c_test nogpu | run in.bin out.bin\n");
        return 2;
    }
    FILE *f = fopen(argv[2], "rb");
    if (!f) return 2;
    uint64_t n = 0, log_n = 0;
    if (fread(&n, 8, 1, f) != 1 || fread(&log_n, 8, 1, f) != 1) return 2;
    const size_t m = (size_t)1 << log_n;
    Fr *scalars = malloc(n * sizeof(Fr)), *a = malloc(m * sizeof(Fr)), omega;
    G1Affine *bases = malloc(n * sizeof(G1Affine));
    if (fread(scalars, sizeof(Fr), n, f) != n || fread(bases, sizeof(G1Affine), n, f) != n ||
        fread(&omega, sizeof(Fr), 1, f) != 1 || fread(a, sizeof(Fr), m, f) != m)
        return 2;
    fclose(f);

    CHECK(h2b_init(0));
    CHECK(h2b_init(0)); /* idempotent */
    if (h2b_abi_version() != 2) return 1;
    G1 msm, commit, commit_half;
    CHECK(h2b_best_multiexp(P(scalars[0]), P(bases[0]), n, P(msm)));
    uint64_t srs = 0;
    CHECK(h2b_srs_register(P(bases[0]), n, &srs));
    CHECK(h2b_commit(srs, P(scalars[0]), n, P(commit)));
    CHECK(h2b_commit(srs, P(scalars[0]), n / 2, P(commit_half)));
    /* upstream asserts bases.len() >= size (kzg/commitment.rs:319, :363): an error code here, never a crash */
    Fr *too_many = calloc(n + 1, sizeof(Fr));
    if (expect(h2b_commit(srs, P(too_many[0]), n + 1, P(commit)), H2B_ERR_ARG, "commit beyond the SRS")) return 1;
    free(too_many);
    CHECK(h2b_commit(srs, P(scalars[0]), n, P(commit)));
    if (expect(h2b_commit(srs + 12345, P(scalars[0]), n, P(commit_half)), H2B_ERR_STATE, "unknown handle")) return 1;
    CHECK(h2b_commit(srs, P(scalars[0]), n / 2, P(commit_half)));
    CHECK(h2b_srs_release(srs));
    if (expect(h2b_best_multiexp(NULL, P(bases[0]), n, P(msm)), H2B_ERR_ARG, "null coeffs")) return 1;
    CHECK(h2b_best_multiexp(P(scalars[0]), P(bases[0]), n, P(msm)));

    Fr *fft = malloc(m * sizeof(Fr)), *lag = malloc(m * sizeof(Fr));
    memcpy(fft, a, m * sizeof(Fr));
    memcpy(lag, a, m * sizeof(Fr));
    CHECK(h2b_best_fft(P(fft[0]), P(omega), (uint32_t)log_n));
    h2b_domain d;
    CHECK(h2b_domain_new(4, (uint32_t)log_n, &d));
    if (d.k != log_n || d.extended_k != log_n + 2) return 1;
    CHECK(h2b_lagrange_to_coeff(&d, P(lag[0])));
    Fr *ext = malloc(((size_t)1 << d.extended_k) * sizeof(Fr));
    CHECK(h2b_coeff_to_extended(&d, P(a[0]), P(ext[0])));
    if (expect(h2b_domain_new(1, 4, &d), H2B_ERR_ARG, "domain with j < 2")) return 1;

    f = fopen(argv[3], "wb");
    if (!f) return 2;
    fwrite(&msm, sizeof msm, 1, f);
    fwrite(&commit, sizeof commit, 1, f);
    fwrite(&commit_half, sizeof commit_half, 1, f);
    fwrite(fft, sizeof(Fr), m, f);
    fwrite(lag, sizeof(Fr), m, f);
    fwrite(ext, sizeof(Fr), (size_t)1 << (log_n + 2), f);
    fclose(f);
    h2b_shutdown();
    printf("abi_c_test ok\n");
    return 0;
}
