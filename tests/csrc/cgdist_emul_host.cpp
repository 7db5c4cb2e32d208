// TEST INFRASTRUCTURE (CPU tier) -- a ROW-PARTITIONED Jacobi-preconditioned CG (BASELINE config C3 at 8 GPUs:
// krypy/linsys.py:593-689 on krypy_b200/dist.py's row blocks) over EMULATED ranks with the DEVICE code of every
// kernel of its iteration, included unchanged:
//     kry_xpby_dev (p_k into one of two peer-visible buffers)  ->  kry_dist_halo (flag handshake + P2P gather of the
//     remote entries of p_k)  ->  kry_spmv_csr with the <p, Ap> epilogue on [local | halo]  ->  kry_peer_allreduce
//     of <p, Ap>  ->  kry_cg_update_dev  ->  kry_cg_scalars (global sum of the new rho through the peer slots)
// Three peer operations per iteration share one epoch counter and the parity-double-buffered slot arrays.  Every
// rank (a group of processes, one per CTA) runs ALL iterations back to back, synchronised with the other ranks by
// nothing but the kernels' own flag protocol, so ranks run ahead of each other as GPUs do (EMUL_JITTER,
// EMUL_SLOW_RANK shake the schedule).  Checked against a long-double CG on the global system: rho and alpha of
// every iteration (bitwise identical on all ranks), the iterate, the epoch counters.
//     cgdist_emul_host <ranks> <local grid rows> <grid columns> <iterations> <CTAs per rank>
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
// PTX wrappers of kry_peer_kernels.cuh
static inline void st_release_sys(unsigned long long* p, unsigned long long v) { dst_release_sys(p, v); }
static inline unsigned long long ld_acquire_sys(const unsigned long long* p) { return dld_acquire_sys(p); }
static inline unsigned long long global_timer_ns() { return dglobal_timer_ns(); }
static inline double ld_volatile_f64(const double* p) { return dld_volatile_f64(p); }

#include "kry_spmv_kernels.cuh"
#include "kry_peer_kernels.cuh"
#include "kry_dist_kernels.cuh"
#define KRY_ROUND_AS_DEFINED          // (kry_sweeps.cuh, included by the dist kernels, has defined round_as)
#include "kry_update_kernels.cuh"

typedef long double LD;

struct RankMem {
    long long n, nhalo, next;
    int *rowptr, *colidx, *halo_peer, *halo_off;
    double *vals, *preg, *Ap, *y, *r, *z, *dinv, *st, *mailbox, *partials;
    unsigned int* ticket;
    unsigned long long* epoch;
    long long nnz;
};

int main(int argc, char** argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: see the header of this file\n");
        return 2;
    }
    const int R = atoi(argv[1]), nxl = atoi(argv[2]), ny = atoi(argv[3]), K = atoi(argv[4]), G = atoi(argv[5]);
    const long long n = (long long)nxl * ny, NG = n * R;
    const long long pld = (n + 2 * ny + 7) / 8 * 8;          // one extended p buffer: [local | left halo | right halo]
    std::mt19937_64 rng(4242 + R);
    std::normal_distribution<double> nd;
    auto diag = [](long long g) { return 4.0 + 0.5 * (double)((g * 7) % 5); };

    std::vector<RankMem> M(R);
    double** p_tab = dev_alloc<double*>(R);
    unsigned long long** flag_tab = dev_alloc<unsigned long long*>(R);
    double** slot_tab = dev_alloc<double*>(R);
    std::vector<LD> bglob(NG);
    for (long long g = 0; g < NG; ++g) bglob[g] = nd(rng);
    for (int r = 0; r < R; ++r) {
        RankMem& m = M[r];
        m.n = n;
        const long long hl = r > 0 ? ny : 0, hr = r < R - 1 ? ny : 0;
        m.nhalo = hl + hr;
        m.next = n + m.nhalo;
        std::vector<int> rp(1, 0), ci;
        std::vector<double> va;
        for (long long l = 0; l < n; ++l) {
            const long long g = r * n + l, gi = g / ny, gj = g % ny;
            auto col = [&](long long c) {                 // global column -> position in [local | halo]
                if (c >= r * n && c < (r + 1) * n) return (int)(c - r * n);
                if (c < r * n) return (int)(n + (c - (r * n - ny)));
                return (int)(n + hl + (c - (r + 1) * n));
            };
            if (gi > 0) { ci.push_back(col(g - ny)); va.push_back(-1.0); }
            if (gj > 0) { ci.push_back(col(g - 1)); va.push_back(-1.0); }
            ci.push_back(col(g)); va.push_back(diag(g));
            if (gj < ny - 1) { ci.push_back(col(g + 1)); va.push_back(-1.0); }
            if (gi < (long long)R * nxl - 1) { ci.push_back(col(g + ny)); va.push_back(-1.0); }
            rp.push_back((int)ci.size());
        }
        m.nnz = (long long)ci.size();
        m.rowptr = dev_alloc<int>(n + 1);
        m.colidx = dev_alloc<int>(m.nnz + 4);
        m.vals = dev_alloc<double>(m.nnz + 4);
        memcpy(m.rowptr, rp.data(), sizeof(int) * (n + 1));
        memcpy(m.colidx, ci.data(), sizeof(int) * m.nnz);
        memcpy(m.vals, va.data(), sizeof(double) * m.nnz);
        m.halo_peer = dev_alloc<int>(2 * ny + 1);
        m.halo_off = dev_alloc<int>(2 * ny + 1);
        int t = 0;
        for (long long i = 0; i < hl; ++i, ++t) {
            m.halo_peer[t] = r - 1;
            m.halo_off[t] = (int)(n - ny + i);
        }
        for (long long i = 0; i < hr; ++i, ++t) {
            m.halo_peer[t] = r + 1;
            m.halo_off[t] = (int)i;
        }
        m.preg = dev_alloc<double>(2 * pld);
        m.Ap = dev_alloc<double>(n + 8);
        m.y = dev_alloc<double>(n + 8);
        m.r = dev_alloc<double>(n + 8);
        m.z = dev_alloc<double>(n + 8);
        m.dinv = dev_alloc<double>(n + 8);
        m.st = dev_alloc<double>(16);
        m.mailbox = dev_alloc<double>((size_t)K * 4 + 4);
        m.partials = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
        m.ticket = dev_alloc<unsigned int>(8);
        m.epoch = dev_alloc<unsigned long long>(1);
        p_tab[r] = m.preg;
        flag_tab[r] = dev_alloc<unsigned long long>(PEER_MAX_RANKS);
        slot_tab[r] = dev_alloc<double>(2ull * PEER_MAX_RANKS * PEER_SLOT);
    }
    // x0 = 0: r = b, z = M r, p_0 = z, rho_0 = <r, z> (global, set-up is host work in the product too)
    LD rho = 0;
    for (long long g = 0; g < NG; ++g) rho += bglob[g] * (bglob[g] / (LD)diag(g));
    for (int r = 0; r < R; ++r) {
        RankMem& m = M[r];
        for (long long l = 0; l < n; ++l) {
            const long long g = r * n + l;
            m.dinv[l] = 1.0 / diag(g);
            m.r[l] = (double)bglob[g];
            m.z[l] = m.dinv[l] * m.r[l];
            m.preg[l] = m.z[l];                           // p_0 in buffer 0
            m.y[l] = 0.0;
        }
        m.st[1] = (double)rho;
    }
    const size_t smem = SpmvCfg<double, 6, 2>::SMEM_BYTES;

    auto body = [&](int r) {
        RankMem& m = M[r];
        PeerArgs pa;
        pa.world = R;
        pa.rank = r;
        pa.epoch_dev = m.epoch;
        pa.slots = slot_tab;
        pa.flags = flag_tab;
        MDotArgs<double> md;
        memset(&md, 0, sizeof(md));
        int slow_rank = -1, slow_us = 0;
        if (const char* e = getenv("EMUL_SLOW_RANK")) sscanf(e, "%d:%d", &slow_rank, &slow_us);
        auto boundary = [&]() {
            if (r == slow_rank && threadIdx.x == 0) usleep((useconds_t)slow_us);
            kry_emul_grid_sync();
        };
        for (int k = 0; k < K; ++k) {
            double* p = m.preg + (k & 1) * pld;
            if (k > 0) {
                xpby_dev_kernel<double, 2>(n, m.z, m.st + 4, m.preg + ((k - 1) & 1) * pld, p);      // linsys.py:627
                boundary();
            }
            dist_halo_kernel<double>(m.nhalo, p_tab, (long long)(k & 1) * pld, m.halo_peer, m.halo_off, p + n, m.ticket, pa);
            boundary();
            spmv_staged_kernel<double, 6, 2, true, 0>(n, m.nnz, m.rowptr, m.colidx, m.vals, p, m.Ap, p, m.partials,
                                                      m.ticket + 1, m.st + 2, md);                  // linsys.py:631-634
            boundary();
            if (blockIdx.x == 0)
                peer_allreduce_kernel(R, r, m.epoch, 1, m.st + 2, slot_tab, flag_tab, 0, nullptr);
            boundary();
            cg_update_kernel<double, 2>(n, m.Ap, p, m.y, m.r, m.z, m.dinv, 0.0, nullptr, m.partials, m.ticket + 2, nullptr, m.st);
            boundary();
            if (blockIdx.x == 0) cg_scalars_kernel(m.st, m.mailbox + 4 * k, pa);                    // linsys.py:655-665
            boundary();
        }
    };
    if (!emul_launch_ranks(R, G, SPMV_THREADS, smem, body)) {
        printf("FAIL a CTA died\n");
        return 1;
    }

    // reference: CG on the global system, long double
    std::vector<LD> xr(NG, 0.0L), rr(bglob), zr(NG), pr(NG), Apr(NG);
    for (long long g = 0; g < NG; ++g) pr[g] = zr[g] = rr[g] / (LD)diag(g);
    const LD rho0 = rho;
    double ehist = 0, ealpha = 0;
    bool same = true;
    for (int k = 0; k < K; ++k) {
        LD pap = 0;
        for (long long g = 0; g < NG; ++g) {
            const long long gi = g / ny, gj = g % ny;
            LD s = (LD)diag(g) * pr[g];
            if (gi > 0) s -= pr[g - ny];
            if (gj > 0) s -= pr[g - 1];
            if (gj < ny - 1) s -= pr[g + 1];
            if (gi < (long long)R * nxl - 1) s -= pr[g + ny];
            Apr[g] = s;
            pap += pr[g] * s;
        }
        const LD alpha = rho / pap;
        LD rho_new = 0;
        for (long long g = 0; g < NG; ++g) {
            xr[g] += alpha * pr[g];
            rr[g] -= alpha * Apr[g];
            zr[g] = rr[g] / (LD)diag(g);
            rho_new += rr[g] * zr[g];
        }
        const LD beta = rho_new / rho;
        for (long long g = 0; g < NG; ++g) pr[g] = zr[g] + beta * pr[g];
        rho = rho_new;
        for (int r = 0; r < R; ++r) {
            const double* mb = M[r].mailbox + 4 * k;
            ehist = fmax(ehist, fabs(sqrt(fabs(mb[0])) - (double)sqrtl(rho)) / (double)sqrtl(rho0));
            ealpha = fmax(ealpha, fabs(mb[1] - (double)alpha) / fabs((double)alpha));
            if (mb[0] != M[0].mailbox[4 * k] || mb[1] != M[0].mailbox[4 * k + 1] || mb[2] != M[0].mailbox[4 * k + 2]) same = false;
        }
    }
    double ex = 0, xmax = 0;
    for (long long g = 0; g < NG; ++g) xmax = fmax(xmax, fabs((double)xr[g]));
    for (long long g = 0; g < NG; ++g) ex = fmax(ex, fabs(M[g / n].y[g % n] - (double)xr[g]) / xmax);
    bool epochs = true;
    for (int r = 0; r < R; ++r)
        if (M[r].epoch[0] != 3ull * K) epochs = false;      // halo handshake, <p,Ap>, rho: three peer operations per iteration
    const bool ok = ehist <= 1e-11 && ealpha <= 1e-10 && ex <= 1e-10 && same && epochs;
    printf("%s cg over %d emulated ranks, %lld rows each (grid %d x %d per rank), %d iterations, %d CTAs per rank: residual history "
           "%.2e alpha %.2e x %.2e | scalars bitwise identical on all ranks %d, epochs %llu (want %d), final relative residual %.2e\n",
           ok ? "ok" : "FAIL", R, n, nxl, ny, K, G, ehist, ealpha, ex, (int)same, M[0].epoch[0], 3 * K, (double)sqrtl(rho / rho0));
    return ok ? 0 : 1;
}
