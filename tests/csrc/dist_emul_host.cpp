// TEST INFRASTRUCTURE (CPU tier) -- a row-partitioned Arnoldi process over EMULATED ranks, running the DEVICE code
// of the one-wait step (krypy_b200/csrc/kry_dist_kernels.cuh, included unchanged): kry_dist_dot with <w, w>
// (dist_dot_kernel) and kry_dist_update_scale (dist_update_scale_kernel: acquire, norm from <w, w> - sum c^2 with
// the exact-norm guard, update + normalised store in one sweep, halo of v_{k+1} from the peers' w, Givens update
// in an extra CTA).  Every rank is a group of processes (one per CTA) that runs ALL steps back to back --
//     w = "A v_k" into one of two buffers (step parity) | kernel boundary | dot | boundary | update_scale | boundary
// -- with no synchronisation between ranks except the kernels' own flag protocol over shared memory, so ranks run
// ahead of each other exactly as GPUs do.  Checked against a long-double reference of the same process: the
// basis, the Hessenberg columns (bitwise identical on all ranks), the halo copies (bitwise the owners' values),
// the Givens residuals, the epoch counters.  Driven by tests/test_dist_emul_cpu.py:
//     dist_emul_host <ranks> <sweep CTAs> <steps> <n per rank> <givens 0|1> <guard step or -1>
// Environment: EMUL_JITTER=<max us> (random schedule jitter, emul_runtime.h), EMUL_SLOW_RANK=<rank>:<us>.
#define KRY_EMUL 1
#include <random>

#include "emul_runtime.h"

#include "kry_dist_kernels.cuh"

typedef long double LD;

struct RankMem {
    double* V;          // (steps + 2) rows of ldv: [local n | halo]
    double* qreg;       // two w buffers (step parity), qld apart
    double* wgen;       // steps x n: the "A v_k" of every step
    double* hcol;       // accumulator (steps + 3)
    double* Hrec;       // steps x (steps + 3): the raw Hessenberg columns as the host would book them
    double *Rt, *cs, *y, *mailbox;
    double* partials;
    unsigned int* ticket;
    int *halo_peer, *halo_off;
    long long nhalo;
    unsigned long long* epoch;
};

int main(int argc, char** argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: see the header of this file\n");
        return 2;
    }
    const int R = atoi(argv[1]), Gs = atoi(argv[2]), K = atoi(argv[3]);
    const long long n = atoll(argv[4]);
    const int givens = atoi(argv[5]), guard_step = atoi(argv[6]);
    const int G = Gs + (givens ? 1 : 0);
    const int HL = 3, HR = 5;                                  // entries taken from the left / right neighbour
    const long long ldv = (n + HL + HR + 7) / 8 * 8, qld = (n + 7) / 8 * 8;
    const int mcol = K + 3;
    std::mt19937_64 rng(2024 + R + 10 * K);
    std::normal_distribution<double> nd;

    std::vector<RankMem> M(R);
    double** q_tab = dev_alloc<double*>(R);
    unsigned long long** flag_tab = dev_alloc<unsigned long long*>(R);
    double** slot_tab = dev_alloc<double*>(R);
    for (int r = 0; r < R; ++r) {
        RankMem& m = M[r];
        m.V = dev_alloc<double>((size_t)(K + 2) * ldv);
        m.qreg = dev_alloc<double>(2 * qld);
        m.wgen = dev_alloc<double>((size_t)K * n);
        m.hcol = dev_alloc<double>(mcol);
        m.Hrec = dev_alloc<double>((size_t)K * mcol);
        m.Rt = dev_alloc<double>((size_t)K * mcol);
        m.cs = dev_alloc<double>(2 * K + 2);
        m.y = dev_alloc<double>(K + 2);
        m.mailbox = dev_alloc<double>((size_t)K * 64);
        m.partials = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
        m.ticket = dev_alloc<unsigned int>(8);
        m.epoch = dev_alloc<unsigned long long>(1);
        m.nhalo = (r > 0 ? HL : 0) + (r < R - 1 ? HR : 0);
        m.halo_peer = dev_alloc<int>(HL + HR);
        m.halo_off = dev_alloc<int>(HL + HR);
        int t = 0;
        if (r > 0)
            for (int i = 0; i < HL; ++i, ++t) {
                m.halo_peer[t] = r - 1;
                m.halo_off[t] = (int)(n - HL + i);
            }
        if (r < R - 1)
            for (int i = 0; i < HR; ++i, ++t) {
                m.halo_peer[t] = r + 1;
                m.halo_off[t] = i;
            }
        q_tab[r] = m.qreg;
        flag_tab[r] = dev_alloc<unsigned long long>(PEER_MAX_RANKS);
        slot_tab[r] = dev_alloc<double>(2ull * PEER_MAX_RANKS * PEER_SLOT);
        for (long long i = 0; i < (long long)K * n; ++i) m.wgen[i] = nd(rng);
        m.y[0] = 2.5;
    }
    // v_0: a normalised global vector, its halo copies behind every rank's row 0
    const long long NG = n * R;
    std::vector<LD> v0(NG);
    LD s0 = 0;
    for (long long i = 0; i < NG; ++i) {
        v0[i] = nd(rng);
        s0 += v0[i] * v0[i];
    }
    for (long long i = 0; i < NG; ++i) M[i / n].V[i % n] = (double)(v0[i] / sqrtl(s0));
    for (int r = 0; r < R; ++r)
        for (long long t = 0; t < M[r].nhalo; ++t) M[r].V[n + t] = M[M[r].halo_peer[t]].V[M[r].halo_off[t]];

    // ---------------- the emulated run: every rank executes all steps ----------------
    auto body = [&](int r) {
        RankMem& m = M[r];
        PeerArgs pa;
        pa.world = R;
        pa.rank = r;
        pa.epoch_dev = m.epoch;
        pa.slots = slot_tab;
        pa.flags = flag_tab;
        const long long stride = (long long)gridDim.x * blockDim.x;
        // EMUL_SLOW_RANK=<rank>:<microseconds>: one rank is held back at every kernel boundary, the others run as far
        // ahead as the flag protocol lets them
        int slow_rank = -1, slow_us = 0;
        if (const char* e = getenv("EMUL_SLOW_RANK")) sscanf(e, "%d:%d", &slow_rank, &slow_us);
        auto boundary = [&]() {
            if (r == slow_rank && threadIdx.x == 0) usleep((useconds_t)slow_us);
            kry_emul_grid_sync();
        };
        for (int k = 0; k < K; ++k) {
            double* w = m.qreg + (k & 1) * qld;
            // "SpMV": w = A v_k.  The guard step makes w almost a combination of the basis (heavy cancellation).
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
                if (k == guard_step)
                    w[i] = 2.0 * m.V[(long long)k * ldv + i] + 0.5 * m.V[i] + 2e-3 * m.wgen[(long long)k * n + i] / sqrt((double)NG);
                else
                    w[i] = m.wgen[(long long)k * n + i];
            }
            boundary();                                   // kernel boundary
            dist_dot_kernel<double, 2>(n, m.V, ldv, k + 1, w, 1, m.partials, m.ticket, pa);
            boundary();
            UpdScaleArgs<double> a;
            a.n = n;
            a.V = m.V;
            a.ldv = ldv;
            a.nv = k + 1;
            a.q = w;
            a.vnext = m.V + (long long)(k + 1) * ldv;
            a.h_acc = m.hcol;
            a.nrm_out = m.hcol + (k + 1);
            a.nhalo = m.nhalo;
            a.peer_q = q_tab;
            a.q_elem_offset = (k & 1) * qld;
            a.halo_peer = m.halo_peer;
            a.halo_off = m.halo_off;
            a.halo_base = n;
            a.halo_dst = m.V + (long long)(k + 1) * ldv + n;
            a.k_givens = givens ? k : -1;
            a.rcol = m.Rt + (long long)k * mcol;
            a.cs = m.cs;
            a.y = m.y;
            a.mailbox = m.mailbox + (long long)k * 64;
            a.partials = m.partials + (size_t)KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS;
            a.ticket = m.ticket + 2;
            a.pa = pa;
            dist_update_scale_kernel<double, 2>(a);
            boundary();
            // the host's booking of the column (the Givens tail leaves it in the mailbox and zeroes the accumulator)
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                for (int i = 0; i < k + 2; ++i) {
                    m.Hrec[(long long)k * mcol + i] = givens ? m.mailbox[(long long)k * 64 + 1 + i] : m.hcol[i];
                    if (!givens) m.hcol[i] = 0.0;
                }
            }
            boundary();
        }
    };
    if (!emul_launch_ranks(R, G, KRY_THREADS, 0, body)) {
        printf("FAIL a CTA died\n");
        return 1;
    }

    // ---------------- reference: the same process on the global vectors, long double ----------------
    std::vector<std::vector<LD>> Vr(K + 1, std::vector<LD>(NG));
    for (long long i = 0; i < NG; ++i) Vr[0][i] = (LD)M[i / n].V[i % n];
    std::vector<std::vector<LD>> Hr(K, std::vector<LD>(mcol, 0.0L));
    for (int k = 0; k < K; ++k) {
        std::vector<LD> w(NG);
        for (long long i = 0; i < NG; ++i) {
            const LD z = (LD)M[i / n].wgen[(long long)k * n + i % n];
            w[i] = (k == guard_step) ? 2.0L * Vr[k][i] + 0.5L * Vr[0][i] + 2e-3L * z / sqrtl((LD)NG) : z;
        }
        for (int j = 0; j <= k; ++j) {
            LD c = 0;
            for (long long i = 0; i < NG; ++i) c += Vr[j][i] * w[i];
            Hr[k][j] = c;
        }
        for (int j = 0; j <= k; ++j)
            for (long long i = 0; i < NG; ++i) w[i] -= Hr[k][j] * Vr[j][i];
        LD s = 0;
        for (long long i = 0; i < NG; ++i) s += w[i] * w[i];
        Hr[k][k + 1] = sqrtl(s);
        for (long long i = 0; i < NG; ++i) Vr[k + 1][i] = w[i] / Hr[k][k + 1];
    }
    double ev = 0, eh = 0, eres = 0;
    bool halo_exact = true, ranks_same = true, epochs_ok = true, finite = true;
    for (int k = 0; k < K; ++k) {
        // the guard step amplifies rounding by ||w|| / ||w - V c|| (~1e3); elsewhere the norm comes from the
        // difference <w, w> - sum c^2 (relative error ~1e-13 away from cancellation)
        for (long long i = 0; i < NG; ++i) {
            const double got = M[i / n].V[(long long)(k + 1) * ldv + i % n];
            if (!std::isfinite(got)) finite = false;
            ev = fmax(ev, fabs(got - (double)Vr[k + 1][i]) * sqrt((double)NG));
        }
        for (int r = 0; r < R; ++r) {
            for (int i = 0; i < k + 2; ++i) {
                eh = fmax(eh, fabs(M[r].Hrec[(long long)k * mcol + i] - (double)Hr[k][i]) / fmax(1.0, fabs((double)Hr[k][k + 1])));
                if (M[r].Hrec[(long long)k * mcol + i] != M[0].Hrec[(long long)k * mcol + i]) ranks_same = false;
            }
            for (long long t = 0; t < M[r].nhalo; ++t) {
                const double own = M[M[r].halo_peer[t]].V[(long long)(k + 1) * ldv + M[r].halo_off[t]];
                if (M[r].V[(long long)(k + 1) * ldv + n + t] != own) halo_exact = false;
            }
        }
    }
    const unsigned long long want_epoch = (unsigned long long)K + (guard_step >= 0 && guard_step < K ? 1 : 0);
    for (int r = 0; r < R; ++r)
        if (M[r].epoch[0] != want_epoch) epochs_ok = false;
    if (givens) {
        // residual norms of the least-squares problems min || beta e_1 - H_k y || from the reference H
        std::vector<LD> cs(2 * K), y(K + 2, 0.0L);
        y[0] = 2.5L;
        for (int k = 0; k < K; ++k) {
            std::vector<LD> col(Hr[k].begin(), Hr[k].begin() + k + 2);
            for (int i = 0; i < k; ++i) {
                const LD t0 = cs[2 * i] * col[i] + cs[2 * i + 1] * col[i + 1], t1 = -cs[2 * i + 1] * col[i] + cs[2 * i] * col[i + 1];
                col[i] = t0;
                col[i + 1] = t1;
            }
            const LD rr = hypotl(col[k], col[k + 1]);
            cs[2 * k] = col[k] / rr;
            cs[2 * k + 1] = col[k + 1] / rr;
            const LD y1 = -cs[2 * k + 1] * y[k];
            y[k] = cs[2 * k] * y[k];
            y[k + 1] = y1;
            for (int r = 0; r < R; ++r) eres = fmax(eres, fabs(M[r].mailbox[(long long)k * 64] - (double)fabsl(y1)) / 2.5);
        }
    }
    const bool ok = finite && ev <= 2e-9 && eh <= 1e-11 && eres <= 1e-11 && halo_exact && ranks_same && epochs_ok;
    printf("%s dist ranks=%d sweepCTAs=%d steps=%d n=%lld givens=%d guard_step=%d: basis %.2e H %.2e residual %.2e | halo copies "
           "bitwise %d, H identical on all ranks %d, epochs %llu (want %llu)\n", ok ? "ok" : "FAIL", R, Gs, K, n, givens,
           guard_step, ev, eh, eres, (int)halo_exact, (int)ranks_same, M[0].epoch[0], want_epoch);
    return ok ? 0 : 1;
}
