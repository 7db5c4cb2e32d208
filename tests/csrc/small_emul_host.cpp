// TEST INFRASTRUCTURE (CPU tier) -- the small device recurrences (krypy_b200/csrc/kry_small_kernels.cuh, included
// unchanged) over the CUDA execution emulator of tests/csrc/cuda_emul, in the chains the solvers run them in:
//   gmres : kry_givens_update column after column on a random Hessenberg matrix (R left on the "device", column k
//           in row k), then kry_tri_solve_t and kry_tri_solve: the residual norms of the least-squares problems
//           min || beta e_1 - H_k y || (krypy/linsys.py:982-993) and the final solution (linsys.py:946)
//   minres: a MINRES iteration chain for config C5 (krypy/linsys.py:791-853 with a diagonal ip_B): SpMV ->
//           kry_lanczos_diag -> kry_minres_recur -> kry_minres_update; the residual norm the recurrence reports
//           must be the TRUE ||b - A x_k||_B of the iterate the update kernel accumulates
//   qr    : kry_small_qr_apply (R^-1 Q^H c)
//     small_emul_host gmres <m> | minres <n> <iterations> <grid> | qr <d>
#define KRY_EMUL 1
#include <random>

#include "emul_runtime.h"

static inline void mbar_init(uint64_t* bar, uint32_t count) { z_mbar_init(bar, count); }
static inline void mbar_fence_init() {}
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { z_mbar_expect_tx(bar, bytes); }
static inline void mbar_arrive(uint64_t* bar) { z_mbar_arrive(bar); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { z_mbar_wait(bar, parity); }
static inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { z_bulk_g2s(dst, src, bytes, bar); }
static inline void consumer_bar_sync() { z_consumer_bar_sync(); }

#include "kry_small_kernels.cuh"
#include "kry_spmv_kernels.cuh"
#include "kry_lanczos_kernels.cuh"
#include "kry_update_kernels.cuh"

typedef long double LD;

static int run_gmres(int m) {
    std::mt19937_64 rng(3 + m);
    std::normal_distribution<double> nd;
    const int ld = m + 2;
    std::vector<LD> H((size_t)(m + 1) * m, 0.0L);
    for (int j = 0; j < m; ++j)
        for (int i = 0; i <= j + 1; ++i) H[(size_t)i * m + j] = (i == j + 1) ? fabs(nd(rng)) + 0.5 : nd(rng);
    const double beta = 2.5;
    double* hcol = dev_alloc<double>(m + 3);
    double* Rt = dev_alloc<double>((size_t)m * ld);
    double* Rrow = dev_alloc<double>((size_t)m * ld);
    double* cs = dev_alloc<double>(2 * m + 2);
    double* y = dev_alloc<double>(m + 2);
    double* mailbox = dev_alloc<double>((size_t)m * (2 * m + 8));
    double* out = dev_alloc<double>(m + 1);
    double* out2 = dev_alloc<double>(m + 1);
    y[0] = beta;
    double eres = 0;
    for (int k = 0; k < m; ++k) {
        for (int i = 0; i < k + 2; ++i) hcol[i] = (double)H[(size_t)i * m + k];
        double* mb = mailbox + (size_t)k * (2 * m + 8);
        double* rcol = Rt + (size_t)k * ld;
        if (!emul_launch(1, 128, sizeof(double) * (3 * k + 2) + 64, [=]() { givens_kernel(k, hcol, rcol, cs, y, mb); })) {
            printf("FAIL a CTA died\n");
            return 1;
        }
        for (int i = 0; i < k + 2; ++i) {
            if (hcol[i] != 0.0) eres = 1.0;                              // the accumulator is left zeroed
            if (mb[1 + i] != (double)H[(size_t)i * m + k]) eres = 1.0;   // the raw column goes to the host
            Rrow[(size_t)i * ld + k] = rcol[i];                          // the same R row-major, for kry_tri_solve
        }
        // least-squares residual of the leading (k+2) x (k+1) problem by normal equations in long double
        const int r = k + 2, c = k + 1;
        std::vector<LD> N((size_t)c * c, 0.0L), g(c, 0.0L), sol(c);
        for (int a = 0; a < c; ++a) {
            for (int b2 = 0; b2 < c; ++b2)
                for (int i = 0; i < r; ++i) N[(size_t)a * c + b2] += H[(size_t)i * m + a] * H[(size_t)i * m + b2];
            g[a] = H[a] * (LD)beta;                                      // H[0, a] * beta
        }
        for (int a = 0; a < c; ++a) {                                    // Gaussian elimination (SPD)
            for (int b2 = a + 1; b2 < c; ++b2) {
                const LD f = N[(size_t)b2 * c + a] / N[(size_t)a * c + a];
                for (int t = a; t < c; ++t) N[(size_t)b2 * c + t] -= f * N[(size_t)a * c + t];
                g[b2] -= f * g[a];
            }
        }
        for (int a = c - 1; a >= 0; --a) {
            LD s = g[a];
            for (int t = a + 1; t < c; ++t) s -= N[(size_t)a * c + t] * sol[t];
            sol[a] = s / N[(size_t)a * c + a];
        }
        LD rn = 0;
        for (int i = 0; i < r; ++i) {
            LD s = (i == 0 ? (LD)beta : 0.0L);
            for (int a = 0; a < c; ++a) s -= H[(size_t)i * m + a] * sol[a];
            rn += s * s;
        }
        eres = fmax(eres, fabs(mb[0] - (double)sqrtl(rn)) / beta);
        if (k == m - 1) {
            bool ok1 = emul_launch(1, 128, sizeof(double) * m + 64, [=]() { tri_solve_kernel<true>(m, Rt, ld, y, out); });
            bool ok2 = emul_launch(1, 128, sizeof(double) * m + 64, [=]() { tri_solve_kernel<false>(m, Rrow, ld, y, out2); });
            if (!ok1 || !ok2) {
                printf("FAIL a CTA died\n");
                return 1;
            }
            double esol = 0, smax = 0;
            for (int a = 0; a < c; ++a) smax = fmax(smax, fabs((double)sol[a]));
            for (int a = 0; a < c; ++a) {
                esol = fmax(esol, fabs(out[a] - (double)sol[a]) / smax);
                if (out[a] != out2[a]) esol = 1.0;                       // both layouts, the same arithmetic
            }
            const bool ok = eres <= 1e-11 && esol <= 1e-9;
            printf("%s gmres recurrences m=%d: residual norms %.2e solution %.2e\n", ok ? "ok" : "FAIL", m, eres, esol);
            return ok ? 0 : 1;
        }
    }
    return 1;
}

static int run_minres(long long n, int K, int G) {
    std::mt19937_64 rng(11 + n);
    std::normal_distribution<double> nd;
    std::uniform_real_distribution<double> ud(1.0, 2.0);
    // A = B^-1 L, L = tridiag(-1, 2.3, -1) - 0.3 I (symmetric), B = diag in [1, 2]: A is self-adjoint in <., .>_B
    std::vector<double> bd(n);
    for (long long i = 0; i < n; ++i) bd[i] = ud(rng);
    std::vector<int> rp(1, 0), ci;
    std::vector<double> va;
    for (long long i = 0; i < n; ++i) {
        if (i > 0) { ci.push_back((int)i - 1); va.push_back(-1.0 / bd[i]); }
        ci.push_back((int)i); va.push_back(2.0 / bd[i]);
        if (i < n - 1) { ci.push_back((int)i + 1); va.push_back(-1.0 / bd[i]); }
        rp.push_back((int)ci.size());
    }
    const long long nnz = (long long)ci.size();
    int* rowptr = dev_alloc<int>(n + 1);
    int* colidx = dev_alloc<int>(nnz + 4);
    double* vals = dev_alloc<double>(nnz + 4);
    memcpy(rowptr, rp.data(), sizeof(int) * (n + 1));
    memcpy(colidx, ci.data(), sizeof(int) * nnz);
    memcpy(vals, va.data(), sizeof(double) * nnz);
    const long long ld = (n + 7) / 8 * 8;
    double* V = dev_alloc<double>((size_t)(K + 2) * ld);
    double *q = dev_alloc<double>(n + 8), *bdev = dev_alloc<double>(n + 8), *w0 = dev_alloc<double>(n + 8),
           *w1 = dev_alloc<double>(n + 8), *yk = dev_alloc<double>(n + 8);
    double* h3 = dev_alloc<double>(4);
    double* st = dev_alloc<double>(16);
    double* mailbox = dev_alloc<double>((size_t)K * 8 + 8);
    double* partials = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
    unsigned int* ticket = dev_alloc<unsigned int>(4);
    std::vector<LD> rhs(n);
    LD nb = 0;
    for (long long i = 0; i < n; ++i) {
        rhs[i] = nd(rng);
        bdev[i] = bd[i];
        nb += rhs[i] * (LD)bd[i] * rhs[i];
    }
    const LD beta0 = sqrtl(nb);
    for (long long i = 0; i < n; ++i) V[i] = (double)(rhs[i] / beta0);       // v_0 = r_0 / ||r_0||_B, x_0 = 0
    st[6] = (double)beta0;
    PeerArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.world = 1;
    MDotArgs<double> md;
    memset(&md, 0, sizeof(md));
    double eres = 0, prev = (double)beta0;
    bool monotone = true;
    for (int k = 0; k < K; ++k) {
        double* vk = V + (size_t)k * ld;
        double* vprev = k > 0 ? V + (size_t)(k - 1) * ld : nullptr;
        double* vnext = V + (size_t)(k + 1) * ld;
        bool ok = emul_launch(G, SPMV_THREADS, SpmvCfg<double, 6, 2>::SMEM_BYTES, [=]() {
            spmv_staged_kernel<double, 6, 2, false, 0>(n, nnz, rowptr, colidx, vals, vk, q, nullptr, partials, ticket, nullptr, md);
        });
        LanczosArgs<double> la = {n, vprev, vk, bdev, q, k > 0 ? h3 : nullptr, h3, vnext, partials, pa};
        ok = ok && emul_launch(G, KRY_THREADS, 0, [=]() { lanczos_diag_kernel<double, 2, false>(la); });
        double* mb = mailbox + 8 * k;
        ok = ok && emul_launch(1, 32, 0, [=]() { minres_recur_kernel(k, h3, st, 1, mb); });
        double* wa = (k & 1) ? w1 : w0;                   // W = [older, newer]: z overwrites the older column
        double* wb = (k & 1) ? w0 : w1;
        ok = ok && emul_launch(G, KRY_THREADS, 0, [=]() { minres_update_kernel<double, 2>(n, vk, wa, wb, yk, st); });
        if (!ok) {
            printf("FAIL a CTA died\n");
            return 1;
        }
        // the true residual of the accumulated iterate in the B-norm
        LD rn = 0;
        for (long long i = 0; i < n; ++i) {
            LD s = rhs[i];
            for (int t = rp[i]; t < rp[i + 1]; ++t) s -= (LD)va[t] * (LD)yk[ci[t]];
            rn += s * (LD)bd[i] * s;
        }
        eres = fmax(eres, fabs(mb[0] - (double)sqrtl(rn)) / (double)beta0);
        if (mb[0] > prev * (1 + 1e-12)) monotone = false;
        prev = mb[0];
    }
    const bool ok = eres <= 1e-10 && monotone;
    printf("%s minres chain n=%lld iterations=%d G=%d: |reported - true B-norm residual| %.2e, monotone %d, final relative "
           "residual %.2e\n", ok ? "ok" : "FAIL", n, K, G, eres, (int)monotone, prev / (double)beta0);
    return ok ? 0 : 1;
}

static int run_qr(int d) {
    std::mt19937_64 rng(23 + d);
    std::normal_distribution<double> nd;
    double *Q = dev_alloc<double>(d * d), *R = dev_alloc<double>(d * d), *c = dev_alloc<double>(d), *o = dev_alloc<double>(d);
    for (int i = 0; i < d; ++i) {
        c[i] = nd(rng);
        for (int j = 0; j < d; ++j) {
            Q[i * d + j] = nd(rng) / std::sqrt((double)d);
            R[i * d + j] = j > i ? 0.3 * nd(rng) : (j == i ? 2.0 : 0.0);
        }
    }
    if (!emul_launch(1, 128, sizeof(double) * 2 * d + 64, [=]() { small_qr_apply_kernel(d, Q, R, c, o); })) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    std::vector<LD> t(d, 0.0L);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) t[i] += (LD)Q[j * d + i] * (LD)c[j];
    for (int j = d - 1; j >= 0; --j) {
        t[j] /= (LD)R[j * d + j];
        for (int i = 0; i < j; ++i) t[i] -= t[j] * (LD)R[i * d + j];
    }
    double e = 0;
    for (int i = 0; i < d; ++i) e = fmax(e, fabs(o[i] - (double)t[i]));
    const bool ok = e <= 1e-13;
    printf("%s small qr apply d=%d: %.2e\n", ok ? "ok" : "FAIL", d, e);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 3 && !strcmp(argv[1], "gmres")) return run_gmres(atoi(argv[2]));
    if (argc >= 5 && !strcmp(argv[1], "minres")) return run_minres(atoll(argv[2]), atoi(argv[3]), atoi(argv[4]));
    if (argc >= 3 && !strcmp(argv[1], "qr")) return run_qr(atoi(argv[2]));
    fprintf(stderr, "usage: see the header of this file\n");
    return 2;
}
