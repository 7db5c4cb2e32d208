// TEST INFRASTRUCTURE (CPU tier) -- a CholQR2 orthonormalisation (the set-up of the deflation projector,
// krypy/utils.py:680-707 / deflation.py:33-56 as utils._cholqr2 runs it) with the DEVICE code of its two kernels,
// kry_gram (gram_kernel: cp.async double-buffered staging, 4x4 register blocks, upper block triangle for X^H X) and
// kry_block_trsm (block_trsm_kernel), included unchanged (krypy_b200/csrc/kry_block_kernels.cuh) and run over the
// CUDA execution emulator of tests/csrc/cuda_emul; the Cholesky factor is formed on the host as in the product.
//     block_emul_host cholqr2 <dtype f64|f32> <n> <d> <grid>
//     block_emul_host gram <dtype f64|f32> <n> <kx> <ky> <grid>          (X^H Y for two different blocks)
#define KRY_EMUL 1
#include <random>

#include "emul_runtime.h"

// cp.async wrappers of kry_block_kernels.cuh (#ifndef KRY_EMUL): an immediate copy / zero fill, nothing to wait for
template <int BYTES> static inline void cp_async_zfill(void* smem_dst, const void* gsrc, bool valid) {
    if (valid) memcpy(smem_dst, gsrc, BYTES);
    else memset(smem_dst, 0, BYTES);
}
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}

#include "kry_block_kernels.cuh"

typedef long double LD;

template <typename T> static size_t gram_smem(int nvec) { return 2ull * nvec * GRAM_LD * sizeof(T); }

template <typename T>
static bool gram(int G, long long n, const T* X, long long ldx, int kx, const T* Y, long long ldy, int ky, int same,
                 double* partials, unsigned int* ticket, double* out) {
    return emul_launch(G, GRAM_THREADS, gram_smem<T>(same ? kx : kx + ky), [=]() {
        gram_kernel<T>(n, X, ldx, kx, Y, ldy, ky, same, partials, ticket, out);
    });
}

template <typename T>
static int run_gram(long long n, int kx, int ky, int G) {
    std::mt19937_64 rng(5 + n + kx);
    std::normal_distribution<double> nd;
    const long long ld = n + 5;
    T* X = dev_alloc<T>((size_t)kx * ld);
    T* Y = dev_alloc<T>((size_t)ky * ld);
    double* partials = dev_alloc<double>((size_t)G * kx * ky + 8);
    unsigned int* ticket = dev_alloc<unsigned int>(2);
    double* out = dev_alloc<double>(kx * ky);
    for (long long i = 0; i < (long long)kx * ld; ++i) X[i] = (T)nd(rng);
    for (long long i = 0; i < (long long)ky * ld; ++i) Y[i] = (T)nd(rng);
    if (!gram<T>(G, n, X, ld, kx, Y, ld, ky, 0, partials, ticket, out)) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    double err = 0;
    for (int i = 0; i < kx; ++i)
        for (int j = 0; j < ky; ++j) {
            LD s = 0;
            for (long long e = 0; e < n; ++e) s += (LD)X[(long long)i * ld + e] * (LD)Y[(long long)j * ld + e];
            err = fmax(err, fabs(out[i * ky + j] - (double)s) / std::sqrt((double)n));
        }
    const bool ok = err <= 1e-13 && ticket[0] == 0;
    printf("%s gram T=%s n=%lld kx=%d ky=%d G=%d: err %.2e\n", ok ? "ok" : "FAIL", sizeof(T) == 8 ? "f64" : "f32", n, kx, ky,
           G, err);
    return ok ? 0 : 1;
}

template <typename T>
static int run_cholqr2(long long n, int d, int G) {
    std::mt19937_64 rng(9 + n + d);
    std::normal_distribution<double> nd;
    const long long ld = (n + 7) / 8 * 8;
    T* X = dev_alloc<T>((size_t)d * ld);
    double* partials = dev_alloc<double>((size_t)G * d * d + 8);
    unsigned int* ticket = dev_alloc<unsigned int>(2);
    double* Gm = dev_alloc<double>(d * d);
    double* Rm = dev_alloc<double>(d * d);
    // a mildly ill-conditioned block: random columns plus a common component
    std::vector<double> common(n);
    for (long long e = 0; e < n; ++e) common[e] = nd(rng);
    for (int j = 0; j < d; ++j)
        for (long long e = 0; e < n; ++e) X[(long long)j * ld + e] = (T)(nd(rng) + 3.0 * common[e]);
    for (int round = 0; round < 2; ++round) {
        if (!gram<T>(G, n, X, ld, d, X, ld, d, 1, partials, ticket, Gm)) {
            printf("FAIL a CTA died\n");
            return 1;
        }
        // host Cholesky G = R^T R (upper triangular R, row-major), as utils._cholqr2 does with numpy
        for (int i = 0; i < d * d; ++i) Rm[i] = 0.0;
        for (int i = 0; i < d; ++i) {
            for (int j = i; j < d; ++j) {
                LD s = (LD)Gm[i * d + j];
                for (int l = 0; l < i; ++l) s -= (LD)Rm[l * d + i] * (LD)Rm[l * d + j];
                if (j == i) {
                    if (s <= 0) {
                        printf("FAIL Gram matrix not positive definite\n");
                        return 1;
                    }
                    Rm[i * d + i] = (double)sqrtl(s);
                } else {
                    Rm[i * d + j] = (double)(s / (LD)Rm[i * d + i]);
                }
            }
        }
        bool ran;
        if (d <= 8)
            ran = emul_launch(G, KRY_THREADS, 0, [=]() { block_trsm_kernel<T, 8>(n, X, ld, d, Rm, X, ld); });
        else if (d <= 16)
            ran = emul_launch(G, KRY_THREADS, 0, [=]() { block_trsm_kernel<T, 16>(n, X, ld, d, Rm, X, ld); });
        else if (d <= 24)
            ran = emul_launch(G, KRY_THREADS, 0, [=]() { block_trsm_kernel<T, 24>(n, X, ld, d, Rm, X, ld); });
        else
            ran = emul_launch(G, KRY_THREADS, 0, [=]() { block_trsm_kernel<T, 32>(n, X, ld, d, Rm, X, ld); });
        if (!ran) {
            printf("FAIL a CTA died\n");
            return 1;
        }
    }
    double eo = 0;
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            LD s = 0;
            for (long long e = 0; e < n; ++e) s += (LD)X[(long long)i * ld + e] * (LD)X[(long long)j * ld + e];
            eo = fmax(eo, fabs((double)s - (i == j ? 1.0 : 0.0)));
        }
    const double tol = sizeof(T) == 8 ? 1e-13 : 1e-5;
    const bool ok = eo <= tol && ticket[0] == 0;
    printf("%s cholqr2 T=%s n=%lld d=%d G=%d: |Q^T Q - I| %.2e\n", ok ? "ok" : "FAIL", sizeof(T) == 8 ? "f64" : "f32", n, d, G, eo);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 6 && !strcmp(argv[1], "cholqr2")) {
        const bool f64 = !strcmp(argv[2], "f64");
        return f64 ? run_cholqr2<double>(atoll(argv[3]), atoi(argv[4]), atoi(argv[5]))
                   : run_cholqr2<float>(atoll(argv[3]), atoi(argv[4]), atoi(argv[5]));
    }
    if (argc >= 7 && !strcmp(argv[1], "gram")) {
        const bool f64 = !strcmp(argv[2], "f64");
        return f64 ? run_gram<double>(atoll(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]))
                   : run_gram<float>(atoll(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
    }
    fprintf(stderr, "usage: see the header of this file\n");
    return 2;
}
