// TEST INFRASTRUCTURE (CPU tier) -- a preconditioned CG iteration chain (krypy/linsys.py:593-689, Jacobi M: BASELINE
// config C3) with the DEVICE code of its kernels, included unchanged and run over the CUDA execution emulator of
// tests/csrc/cuda_emul:  SpMV with the <p, Ap> epilogue (spmv_staged_kernel, kry_spmv_kernels.cuh)  ->
// kry_cg_update_dev (x, r, z and the local <r, z> in one sweep)  ->  kry_cg_scalars (the scalar recurrence on the
// device)  ->  kry_xpby_dev (p = z + beta p), the scalars never leaving "device" memory -- against a long-double
// CG.  Second mode: kry_cg_scalars over emulated ranks (the global sum of the new rho through the peer slots).
//     cg_emul_host chain <nx> <ny> <iterations> <grid>
//     cg_emul_host scalars <ranks>
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

#include "kry_spmv_kernels.cuh"
#include "kry_update_kernels.cuh"

typedef long double LD;

static int run_chain(int nx, int ny, int K, int G) {
    const long long n = (long long)nx * ny;
    std::vector<int> rp(1, 0), ci;
    std::vector<double> va;
    for (int r = 0; r < n; ++r) {                       // 5-point Laplacian with a varying diagonal (SPD)
        const int i = r / ny, j = r % ny;
        if (i > 0) { ci.push_back(r - ny); va.push_back(-1.0); }
        if (j > 0) { ci.push_back(r - 1); va.push_back(-1.0); }
        ci.push_back(r); va.push_back(4.0 + 0.5 * ((r * 7) % 5));
        if (j < ny - 1) { ci.push_back(r + 1); va.push_back(-1.0); }
        if (i < nx - 1) { ci.push_back(r + ny); va.push_back(-1.0); }
        rp.push_back((int)ci.size());
    }
    const long long nnz = (long long)ci.size();
    int* rowptr = dev_alloc<int>(n + 1);
    int* colidx = dev_alloc<int>(nnz + 4);
    double* vals = dev_alloc<double>(nnz + 4);
    memcpy(rowptr, rp.data(), sizeof(int) * (n + 1));
    memcpy(colidx, ci.data(), sizeof(int) * nnz);
    memcpy(vals, va.data(), sizeof(double) * nnz);
    double *p = dev_alloc<double>(n + 8), *Ap = dev_alloc<double>(n + 8), *y = dev_alloc<double>(n + 8),
           *r = dev_alloc<double>(n + 8), *z = dev_alloc<double>(n + 8), *dinv = dev_alloc<double>(n + 8);
    double* st = dev_alloc<double>(16);
    double* mailbox = dev_alloc<double>((size_t)K * 4 + 4);
    double* partials = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
    unsigned int* ticket = dev_alloc<unsigned int>(8);
    std::mt19937_64 rng(31);
    std::normal_distribution<double> nd;
    std::vector<LD> br(n), xr(n, 0.0L), rr(n), zr(n), pr(n), Apr(n);
    LD rho = 0;
    for (long long i = 0; i < n; ++i) {
        dinv[i] = 1.0 / va[rp[i] + (i >= ny ? 1 : 0) + (i % ny > 0 ? 1 : 0)];     // 1 / diagonal entry
        r[i] = nd(rng);
        z[i] = dinv[i] * r[i];
        p[i] = z[i];
        y[i] = 0.0;
        br[i] = rr[i] = (LD)r[i];
        zr[i] = (LD)z[i];
        pr[i] = zr[i];
        rho += rr[i] * zr[i];
    }
    st[1] = (double)rho;
    const LD rho0 = rho;
    PeerArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.world = 1;
    MDotArgs<double> md;
    memset(&md, 0, sizeof(md));
    double ehist = 0;
    for (int k = 0; k < K; ++k) {
        bool ok = emul_launch(G, SPMV_THREADS, SpmvCfg<double, 6, 2>::SMEM_BYTES, [=]() {
            spmv_staged_kernel<double, 6, 2, true, 0>(n, nnz, rowptr, colidx, vals, p, Ap, p, partials, ticket + 1, st + 2, md);
        });
        ok = ok && emul_launch(G, KRY_THREADS, 0, [=]() {
            cg_update_kernel<double, 2>(n, Ap, p, y, r, z, dinv, 0.0, nullptr, partials, ticket + 2, nullptr, st);
        });
        double* mb = mailbox + 4 * k;
        ok = ok && emul_launch(1, 64, 0, [=]() { cg_scalars_kernel(st, mb, pa); });
        ok = ok && emul_launch(G, KRY_THREADS, 0, [=]() { xpby_dev_kernel<double, 2>(n, z, st + 4, p, p); });
        if (!ok) {
            printf("FAIL a CTA died\n");
            return 1;
        }
        // reference iteration (linsys.py:627-665)
        LD pap = 0;
        for (long long i = 0; i < n; ++i) {
            LD s = 0;
            for (int t = rp[i]; t < rp[i + 1]; ++t) s += (LD)va[t] * pr[ci[t]];
            Apr[i] = s;
            pap += pr[i] * s;
        }
        const LD alpha = rho / pap;
        LD rho_new = 0;
        for (long long i = 0; i < n; ++i) {
            xr[i] += alpha * pr[i];
            rr[i] -= alpha * Apr[i];
            zr[i] = (LD)dinv[i] * rr[i];
            rho_new += rr[i] * zr[i];
        }
        const LD beta = rho_new / rho;
        for (long long i = 0; i < n; ++i) pr[i] = zr[i] + beta * pr[i];
        rho = rho_new;
        ehist = fmax(ehist, fabs(sqrt(mb[0]) - (double)sqrtl(rho)) / (double)sqrtl(rho0));
        if (fabs(mb[1] - (double)alpha) > 1e-10 * fabs((double)alpha)) ehist = 1.0;
    }
    double ex = 0, xmax = 0;
    for (long long i = 0; i < n; ++i) xmax = fmax(xmax, fabs((double)xr[i]));
    for (long long i = 0; i < n; ++i) ex = fmax(ex, fabs(y[i] - (double)xr[i]) / xmax);
    const bool tickets = ticket[1] == 0 && ticket[2] == 0;
    const bool ok = ehist <= 1e-11 && ex <= 1e-10 && tickets;
    printf("%s cg chain n=%lld nnz=%lld iterations=%d G=%d: residual history %.2e x %.2e (final relative residual %.2e)\n",
           ok ? "ok" : "FAIL", n, nnz, K, G, ehist, ex, (double)sqrtl(rho / rho0));
    return ok ? 0 : 1;
}

static int run_scalars(int R) {
    unsigned long long** flag_tab = dev_alloc<unsigned long long*>(R);
    double** slot_tab = dev_alloc<double*>(R);
    std::vector<double*> st(R), mb(R);
    std::vector<unsigned long long*> epoch(R);
    LD total = 0;
    for (int r = 0; r < R; ++r) {
        flag_tab[r] = dev_alloc<unsigned long long>(PEER_MAX_RANKS);
        slot_tab[r] = dev_alloc<double>(2ull * PEER_MAX_RANKS * PEER_SLOT);
        st[r] = dev_alloc<double>(16);
        mb[r] = dev_alloc<double>(4);
        epoch[r] = dev_alloc<unsigned long long>(1);
        epoch[r][0] = 10;
        for (int p = 0; p < R; ++p) flag_tab[r][p] = 10;
        st[r][1] = 2.0;                                   // rho_k
        st[r][2] = 0.7;
        st[r][3] = 0.3;
        st[r][5] = 0.1 + 0.37 * r;                        // local share of the new rho
        total += (LD)st[r][5];
    }
    if (!emul_launch_ranks(R, 1, 64, 0, [&](int r) {
            PeerArgs pa;
            pa.world = R;
            pa.rank = r;
            pa.epoch_dev = epoch[r];
            pa.slots = slot_tab;
            pa.flags = flag_tab;
            cg_scalars_kernel(st[r], mb[r], pa);
        })) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    bool ok = true;
    for (int r = 0; r < R; ++r) {
        const double nrm = sqrt(fabs((double)total));
        ok = ok && fabs(st[r][1] - nrm * nrm) <= 1e-15 && st[r][0] == 2.0 && fabs(st[r][4] - st[r][1] / 2.0) <= 1e-15 &&
             st[r][1] == st[0][1] && mb[r][0] == mb[0][0] && epoch[r][0] == 11;
    }
    printf("%s cg scalars over %d ranks: rho %.17g\n", ok ? "ok" : "FAIL", R, st[0][1]);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 6 && !strcmp(argv[1], "chain")) return run_chain(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
    if (argc >= 3 && !strcmp(argv[1], "scalars")) return run_scalars(atoi(argv[2]));
    fprintf(stderr, "usage: see the header of this file\n");
    return 2;
}
