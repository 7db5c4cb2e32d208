// TEST INFRASTRUCTURE (CPU tier) -- runs the DEVICE code of kry_spmv_csr / kry_spmv_csr_mdot
// (krypy_b200/csrc/kry_spmv_kernels.cuh, included unchanged: the TMA-staged warp-specialised kernel with its
// <w, y> and multi-vector dot epilogues, the warp-per-row kernel) on the host over the CUDA execution emulator of
// tests/csrc/cuda_emul.  The staged kernel sums every row in storage order with separately rounded products and
// sums -- scipy's csr_matvec -- so its rows must be BIT-IDENTICAL to that ordered sum (bit-exact for
// the SpMV in fp64).  Driven by tests/test_spmv_emul_cpu.py:
//     spmv_emul_host <kind> <dtype f64|f32> <mode plain|dot|mdot> <grid> [nb]
#define KRY_EMUL 1
#include "emul_runtime.h"
#include "emul_matrices.h"

// the PTX wrappers of kry_spmv_kernels.cuh (#ifndef KRY_EMUL) on the emulator's mbarrier / bulk-copy model
static inline void mbar_init(uint64_t* bar, uint32_t count) { z_mbar_init(bar, count); }
static inline void mbar_fence_init() {}
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { z_mbar_expect_tx(bar, bytes); }
static inline void mbar_arrive(uint64_t* bar) { z_mbar_arrive(bar); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { z_mbar_wait(bar, parity); }
static inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { z_bulk_g2s(dst, src, bytes, bar); }
static inline void consumer_bar_sync() { z_consumer_bar_sync(); }

#include "kry_spmv_kernels.cuh"

typedef long double LD;

template <typename T, int CPR, bool DOT, int NACC>
static bool launch(int G, long long nrows, long long nnz, const int* rowptr, const int* colidx, const T* vals, const T* x,
                   T* y, const T* w, double* partials, unsigned int* ticket, double* dot_out, MDotArgs<T> md) {
    return emul_launch(G, SPMV_THREADS, SpmvCfg<T, CPR, 2>::SMEM_BYTES, [=]() {
        spmv_staged_kernel<T, CPR, 2, DOT, NACC>(nrows, nnz, rowptr, colidx, vals, x, y, w, partials, ticket, dot_out, md);
    });
}

template <typename T>
static int run(const char* kind, const char* mode, int G, int nb) {
    std::mt19937_64 rng(77);
    std::normal_distribution<double> nd;
    Csr A = make_matrix(kind, rng);
    const long long nnz = (long long)A.vals.size(), nrows = A.nrows;
    int* rowptr = dev_alloc<int>(nrows + 1);
    int* colidx = dev_alloc<int>(nnz + 4);
    T* vals = dev_alloc<T>(nnz + 4);
    T* x = dev_alloc<T>(A.ncols);
    T* y = dev_alloc<T>(nrows);
    T* w = dev_alloc<T>(nrows);
    T* B = dev_alloc<T>((size_t)(nb > 0 ? nb : 1) * (nrows + 8));
    double* partials = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
    unsigned int* ticket = dev_alloc<unsigned int>(4);
    double* dot_out = dev_alloc<double>(PEER_SLOT + 2);
    memcpy(rowptr, A.rowptr.data(), sizeof(int) * (nrows + 1));
    memcpy(colidx, A.colidx.data(), sizeof(int) * nnz);
    for (long long k = 0; k < nnz; ++k) vals[k] = (T)A.vals[k].real();
    for (long long i = 0; i < A.ncols; ++i) x[i] = (T)nd(rng);
    for (long long i = 0; i < nrows; ++i) {
        y[i] = (T)NAN;
        w[i] = (T)nd(rng);
    }
    const long long ldb = nrows + 8;
    for (long long i = 0; i < (long long)(nb > 0 ? nb : 1) * ldb; ++i) B[i] = (T)nd(rng);
    for (int j = 0; j < PEER_SLOT + 2; ++j) dot_out[j] = -7.0;

    const bool dot = !strcmp(mode, "dot"), mdot = !strcmp(mode, "mdot");
    MDotArgs<T> md;
    memset(&md, 0, sizeof(md));
    if (mdot) {
        md.B = B;
        md.ldb = ldb;
        md.nb = nb;
        md.want_sq = 1;
        md.out = dot_out;
        md.pa.world = 1;
    }
    const double avg = (double)nnz / (double)nrows;
    const char* path;
    bool ran;
#define LAUNCH(CPR)                                                                                                       \
    (mdot ? launch<T, CPR, false, 8>(G, nrows, nnz, rowptr, colidx, vals, x, y, nullptr, partials, ticket, nullptr, md)     \
          : dot ? launch<T, CPR, true, 0>(G, nrows, nnz, rowptr, colidx, vals, x, y, w, partials, ticket, dot_out, md)      \
                : launch<T, CPR, false, 0>(G, nrows, nnz, rowptr, colidx, vals, x, y, nullptr, partials, ticket, nullptr, md))
    if (avg <= 5.5) {
        path = "staged6";
        ran = LAUNCH(6);
    } else if (avg <= 7.5) {
        path = "staged8";
        ran = LAUNCH(8);
    } else if (avg <= 15.0) {
        path = "staged16";
        ran = LAUNCH(16);
    } else {
        path = "warp";
        if (mdot) {
            printf("FAIL the warp-per-row kernel has no multi-dot epilogue\n");
            return 1;
        }
        if (dot)
            ran = emul_launch(G, KRY_THREADS, 0, [=]() {
                spmv_warp_kernel<T, true>(nrows, rowptr, colidx, vals, x, y, w, partials, ticket, dot_out);
            });
        else
            ran = emul_launch(G, KRY_THREADS, 0, [=]() {
                spmv_warp_kernel<T, false>(nrows, rowptr, colidx, vals, x, y, nullptr, partials, ticket, nullptr);
            });
    }
    if (!ran) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    double err = 0;
    long long exact = 0;
    LD dref = 0, sq = 0;
    std::vector<LD> bref(nb > 0 ? nb : 1, 0.0L);
    for (long long r = 0; r < nrows; ++r) {
        LD s = 0, scale = 1e-300L;
        double sd = 0.0;                                  // storage order, product and sum rounded separately
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
            s += (LD)vals[k] * (LD)x[colidx[k]];
            scale += fabsl((LD)vals[k] * (LD)x[colidx[k]]);
            sd = sd + (double)vals[k] * (double)x[colidx[k]];
        }
        if (y[r] == (T)sd) ++exact;
        const double e = (double)(fabsl((LD)y[r] - s) / scale);
        if (!(e <= err)) err = e;
        dref += (LD)w[r] * (LD)y[r];
        sq += (LD)y[r] * (LD)y[r];
        for (int j = 0; j < nb; ++j) bref[j] += (LD)B[(long long)j * ldb + r] * (LD)y[r];
    }
    double edot = 0;
    if (dot) edot = fabs(dot_out[0] - (double)dref) / fmax(1.0, fabs((double)dref));
    if (mdot) {
        for (int j = 0; j < nb; ++j) edot = fmax(edot, fabs(dot_out[j] - (double)bref[j]) / fmax(1.0, fabs((double)bref[j])));
        edot = fmax(edot, fabs(dot_out[nb] - (double)sq) / (double)sq);
        if (dot_out[nb + 1] != -7.0) edot = 1.0;          // nothing written beyond the nb + 1 sums
    }
    if ((dot || mdot) && ticket[0] + ticket[1] + ticket[2] + ticket[3] != 0) edot = 1.0;   // tickets are left reset
    int maxrow = 0;
    for (long long r = 0; r < nrows; ++r) maxrow = std::max(maxrow, rowptr[r + 1] - rowptr[r]);
    const double eps = sizeof(T) == 8 ? 2.3e-16 : 1.2e-7;
    const bool ok = err <= eps * std::max(1, maxrow) && edot <= 100 * eps && (!strcmp(path, "warp") || exact == nrows);
    printf("%s spmv %s T=%s mode=%s path=%s G=%d rows=%lld nnz=%lld: err %.2e dots %.2e, rows bit-identical to the ordered sum %lld\n",
           ok ? "ok" : "FAIL", kind, sizeof(T) == 8 ? "f64" : "f32", mode, path, G, nrows, nnz, err, edot, exact);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: see the header of this file\n");
        return 2;
    }
    const int nb = argc > 5 ? atoi(argv[5]) : 0;
    if (!strcmp(argv[2], "f64")) return run<double>(argv[1], argv[3], atoi(argv[4]), nb);
    return run<float>(argv[1], argv[3], atoi(argv[4]), nb);
}
