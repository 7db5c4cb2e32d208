// TEST INFRASTRUCTURE (CPU tier) -- runs the DEVICE code of kry_orth_fused_z / kry_spmv_csr_z
// (krypy_b200/csrc/kry_zorth.cuh, kry_zspmv.cuh, included unchanged) on the host over the CUDA execution
// emulator of tests/csrc/cuda_emul and compares with extended-precision references.  See cuda_runtime.h there
// for what the emulation is and what it proves.  Driven by tests/test_cplx_emul_cpu.py:
//     cplx_emul_host orth <algo 0|1> <passes> <nv> <j0> <n> <grid> <separate_P 0|1>
//     cplx_emul_host spmv <kind> <cplx 0|1> <grid>
// prints one line "ok <max error> ..." or "FAIL ..." and exits 0 / 1.
#define KRY_EMUL 1
#include <complex>
#include <random>

#include "emul_runtime.h"
#include "emul_matrices.h"

#include "kry_zorth.cuh"
#include "kry_zspmv.cuh"

typedef std::complex<long double> CL;

// ------------------------------------------------------------------ kry_orth_fused_z
static int run_orth(int algo, int passes, int nv, int j0, long long n, int G, int separate_P) {
    std::mt19937_64 rng(1234 + 7 * nv + n);
    std::normal_distribution<double> nd;
    const long long ldv = n + 3;                          // rows need not be contiguous
    const int nvs = nv > 0 ? nv : 1;
    Z* V = dev_alloc<Z>((size_t)nvs * ldv);
    Z* P = separate_P ? dev_alloc<Z>((size_t)nvs * ldv) : V;
    Z* q = dev_alloc<Z>(n);
    Z* vnext = dev_alloc<Z>(n);
    double* h = dev_alloc<double>(2 * nvs + 2);
    double* partials = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
    const double sc = 1.0 / std::sqrt(2.0 * (double)n);
    for (long long i = 0; i < (long long)nvs * ldv; ++i) {
        V[i] = make_double2(nd(rng) * sc, nd(rng) * sc);
        if (separate_P) P[i] = make_double2(nd(rng) * sc, nd(rng) * sc);
    }
    std::vector<CD> q0(n);
    for (long long i = 0; i < n; ++i) {
        q[i] = make_double2(nd(rng), nd(rng));
        q0[i] = CD(q[i].x, q[i].y);
        vnext[i] = make_double2(NAN, NAN);
    }
    for (int j = 0; j < 2 * nvs + 2; ++j) h[j] = (j & 1) ? -0.25 : 0.5;
    // reference (krypy/utils.py:1012-1029 for MGS; block classical Gram-Schmidt for CGS), long double
    std::vector<CL> qr(q0.begin(), q0.end()), hr(nvs, CL(0, 0));
    auto Vc = [&](const Z* B, int j, long long i) { return CL(B[(long long)j * ldv + i].x, B[(long long)j * ldv + i].y); };
    for (int p = 0; p < passes; ++p) {
        if (algo == KRY_ORTH_CGS) {
            std::vector<CL> c(nvs, CL(0, 0));
            for (int j = j0; j < nv; ++j)
                for (long long i = 0; i < n; ++i) c[j] += std::conj(Vc(V, j, i)) * qr[i];
            for (int j = j0; j < nv; ++j) {
                hr[j] += c[j];
                for (long long i = 0; i < n; ++i) qr[i] -= c[j] * Vc(P, j, i);
            }
        } else {
            for (int j = j0; j < nv; ++j) {
                CL c(0, 0);
                for (long long i = 0; i < n; ++i) c += std::conj(Vc(V, j, i)) * qr[i];
                hr[j] += c;
                for (long long i = 0; i < n; ++i) qr[i] -= c * Vc(P, j, i);
            }
        }
    }
    long double nr2 = 0;
    for (long long i = 0; i < n; ++i) nr2 += std::norm(qr[i]);
    const long double nr = sqrtl(nr2);

    ZOrthArgs a = {n, V, P, ldv, j0, nv, passes, algo, q, h, h + 2 * nvs, vnext, partials};
    if (!emul_launch(G, KRY_THREADS, 0, [a]() { zorth_kernel(a); })) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    double eh = 0, eq = 0, ev = 0, en;
    for (int j = 0; j < nvs; ++j) {
        const CD want = (j >= j0 && j < nv) ? CD(0.5, -0.25) + CD((double)hr[j].real(), (double)hr[j].imag()) : CD(0.5, -0.25);
        eh = fmax(eh, std::abs(CD(h[2 * j], h[2 * j + 1]) - want));
    }
    en = fabs(h[2 * nvs] - (double)nr) / (double)nr;
    if (h[2 * nvs + 1] != -0.25) en = 1.0;                // only one double is written for the norm
    double qmax = 0;
    for (long long i = 0; i < n; ++i) qmax = fmax(qmax, std::abs(q0[i]));
    for (long long i = 0; i < n; ++i) {
        const CD w((double)qr[i].real(), (double)qr[i].imag());
        eq = fmax(eq, std::abs(CD(q[i].x, q[i].y) - w) / qmax);
        ev = fmax(ev, std::abs(CD(vnext[i].x, vnext[i].y) - w / (double)nr));
    }
    const bool ok = eh <= 1e-13 && eq <= 1e-13 && ev <= 1e-13 && en <= 1e-13;
    printf("%s orth algo=%d passes=%d nv=%d j0=%d n=%lld G=%d: h %.2e q %.2e vnext %.2e nrm %.2e\n", ok ? "ok" : "FAIL",
           algo, passes, nv, j0, n, G, eh, eq, ev, en);
    return ok ? 0 : 1;
}

// ------------------------------------------------------------------ kry_spmv_csr_z
template <typename TV> static TV to_val(CD v);
template <> double2 to_val<double2>(CD v) { return make_double2(v.real(), v.imag()); }
template <> double to_val<double>(CD v) { return v.real(); }

template <typename TV>
static int run_spmv_t(const char* kind, int G) {
    std::mt19937_64 rng(99);
    std::normal_distribution<double> nd;
    Csr A = make_matrix(kind, rng);
    const long long nnz = (long long)A.vals.size();
    int* rowptr = dev_alloc<int>(A.nrows + 1);
    int* colidx = dev_alloc<int>(nnz + 4);
    TV* vals = dev_alloc<TV>(nnz + 4);
    Z* x = dev_alloc<Z>(A.ncols);
    Z* y = dev_alloc<Z>(A.nrows);
    memcpy(rowptr, A.rowptr.data(), sizeof(int) * (A.nrows + 1));
    memcpy(colidx, A.colidx.data(), sizeof(int) * nnz);
    for (long long k = 0; k < nnz; ++k) vals[k] = to_val<TV>(A.vals[k]);
    for (long long i = 0; i < A.ncols; ++i) x[i] = make_double2(nd(rng), nd(rng));
    for (long long i = 0; i < A.nrows; ++i) y[i] = make_double2(NAN, NAN);
    const double avg = (double)nnz / (double)A.nrows;
    const long long nrows = A.nrows;
    bool ran;
    const char* path;
    if (avg <= 5.5) {
        path = "staged6";
        ran = emul_launch(G, ZSPMV_THREADS, ZSpmvCfg<TV, 6, 2>::SMEM_BYTES,
                          [=]() { zspmv_staged_kernel<TV, 6, 2>(nrows, nnz, rowptr, colidx, vals, x, y); });
    } else if (avg <= 7.5) {
        path = "staged8";
        ran = emul_launch(G, ZSPMV_THREADS, ZSpmvCfg<TV, 8, 2>::SMEM_BYTES,
                          [=]() { zspmv_staged_kernel<TV, 8, 2>(nrows, nnz, rowptr, colidx, vals, x, y); });
    } else if (avg <= 15.0) {
        path = "staged16";
        ran = emul_launch(G, ZSPMV_THREADS, ZSpmvCfg<TV, 16, 2>::SMEM_BYTES,
                          [=]() { zspmv_staged_kernel<TV, 16, 2>(nrows, nnz, rowptr, colidx, vals, x, y); });
    } else {
        path = "warp";
        ran = emul_launch(G, KRY_THREADS, 0, [=]() { zspmv_warp_kernel<TV>(nrows, rowptr, colidx, vals, x, y); });
    }
    if (!ran) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    double err = 0;
    long long exact = 0;
    for (long long r = 0; r < A.nrows; ++r) {
        CL s(0, 0);
        long double scale = 1e-300L;
        CD sd(0, 0);                                      // the kernel's own order and roundings (staged path)
        for (int k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k) {
            const CD a = sizeof(TV) == 16 ? A.vals[k] : CD(A.vals[k].real(), 0.0);
            const CD xv(x[A.colidx[k]].x, x[A.colidx[k]].y);
            s += CL(a) * CL(xv);
            scale += std::abs(a) * std::abs(xv);
            const double pr = sizeof(TV) == 16 ? a.real() * xv.real() - a.imag() * xv.imag() : a.real() * xv.real();
            const double pi = sizeof(TV) == 16 ? a.real() * xv.imag() + a.imag() * xv.real() : a.real() * xv.imag();
            sd = CD(sd.real() + pr, sd.imag() + pi);
        }
        const CD got(y[r].x, y[r].y);
        if (got == sd) ++exact;
        const double e = (double)(std::abs(CL(got) - s) / scale);
        if (!(e <= err)) err = e;                         // (NaN propagates into err)
    }
    const int maxrow = [&]() { int m = 0; for (long long r = 0; r < A.nrows; ++r) m = std::max(m, A.rowptr[r + 1] - A.rowptr[r]); return m; }();
    const bool ok = err <= 3e-16 * std::max(1, maxrow) && (strcmp(path, "warp") == 0 || exact == A.nrows);
    printf("%s spmv %s vals=%s path=%s G=%d rows=%lld nnz=%lld: err %.2e, rows bit-identical to the ordered sum %lld\n",
           ok ? "ok" : "FAIL", kind, sizeof(TV) == 16 ? "complex" : "real", path, G, A.nrows, nnz, err, exact);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 9 && !strcmp(argv[1], "orth"))
        return run_orth(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoll(argv[6]), atoi(argv[7]),
                        atoi(argv[8]));
    if (argc >= 5 && !strcmp(argv[1], "spmv"))
        return atoi(argv[3]) ? run_spmv_t<double2>(argv[2], atoi(argv[4])) : run_spmv_t<double>(argv[2], atoi(argv[4]));
    fprintf(stderr, "usage: see the header of this file\n");
    return 2;
}
