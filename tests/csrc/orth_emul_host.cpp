// TEST INFRASTRUCTURE (CPU tier) -- runs the DEVICE code of the headline kernels, kry_orth_fused /
// kry_orth_fused_dist (orth_kernel) and kry_project (proj_kernel) of krypy_b200/csrc/kry_orth_kernels.cuh, included
// unchanged, on the host over the CUDA execution emulator of tests/csrc/cuda_emul and compares with
// extended-precision references.  Row-partitioned runs are emulated as R ranks of G CTAs each that share the peer
// slot / flag arrays (release / acquire on shared memory).  Driven by tests/test_orth_emul_cpu.py:
//     orth_emul_host orth <dtype f64|f32> <vec> <algo 0|1> <passes> <nv> <j0> <n> <grid> <ranks> <pre 0|1> <separate_P 0|1>
//     orth_emul_host proj <dtype f64|f32> <vec> <d> <iterations> <n> <grid> <with_QR 0|1>
// prints one line "ok ..." or "FAIL ..." and exits 0 / 1.
#define KRY_EMUL 1
#include <random>

#include "emul_runtime.h"

#include "kry_orth_kernels.cuh"

typedef long double LD;

template <typename T> struct Tol;
template <> struct Tol<double> { static constexpr double v = 1e-13; };
template <> struct Tol<float> { static constexpr double v = 3e-6; };

// ------------------------------------------------------------------ orth_kernel
template <typename T, int VEC>
static int run_orth(int algo, int passes, int nv, int j0, long long n, int G, int R, int pre, int separate_P) {
    std::mt19937_64 rng(4321 + 13 * nv + n + 7 * R);
    std::normal_distribution<double> nd;
    // VEC > 1 needs rows that start 16-byte aligned; VEC == 1 exercises an odd leading dimension
    const long long ldv = VEC > 1 ? (n + 7) / 8 * 8 + 8 : n + 3;
    const int nvs = nv > 0 ? nv : 1;
    const double sc = 1.0 / std::sqrt((double)n * R);
    std::vector<T*> V(R), P(R), q(R), vnext(R), prev(R);
    std::vector<double*> h(R), partials(R), precoef(R);
    for (int r = 0; r < R; ++r) {
        V[r] = dev_alloc<T>((size_t)nvs * ldv);
        P[r] = separate_P ? dev_alloc<T>((size_t)nvs * ldv) : V[r];
        q[r] = dev_alloc<T>(n + 8);
        vnext[r] = dev_alloc<T>(n + 8);
        prev[r] = dev_alloc<T>(n + 8);
        h[r] = dev_alloc<double>(nvs + 2);
        precoef[r] = dev_alloc<double>(1);
        partials[r] = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
        for (long long i = 0; i < (long long)nvs * ldv; ++i) {
            V[r][i] = (T)(nd(rng) * sc);
            if (separate_P) P[r][i] = (T)(nd(rng) * sc);
        }
        for (long long i = 0; i < n; ++i) {
            q[r][i] = (T)nd(rng);
            prev[r][i] = (T)nd(rng);
            vnext[r][i] = (T)NAN;
        }
        for (int j = 0; j < nvs + 2; ++j) h[r][j] = 0.5;
        precoef[r][0] = 0.37;
    }
    // peer tables (shared by all emulated ranks)
    unsigned long long** flag_tab = dev_alloc<unsigned long long*>(R);
    double** slot_tab = dev_alloc<double*>(R);
    std::vector<unsigned long long*> epoch(R);
    for (int r = 0; r < R; ++r) {
        flag_tab[r] = dev_alloc<unsigned long long>(PEER_MAX_RANKS);
        slot_tab[r] = dev_alloc<double>(2ull * PEER_MAX_RANKS * PEER_SLOT);
        epoch[r] = dev_alloc<unsigned long long>(1);
        epoch[r][0] = 6;                                   // a run in progress: the counter is not at zero
        for (int p = 0; p < R; ++p) flag_tab[r][p] = 6;
    }

    // reference on the global vectors (rank after rank), long double
    const long long NG = n * R;
    auto at = [&](const std::vector<T*>& B, int j, long long i) { return (LD)B[i / n][(long long)j * ldv + i % n]; };
    std::vector<LD> qr(NG), hr(nvs, 0.0L);
    for (long long i = 0; i < NG; ++i) qr[i] = (LD)q[i / n][i % n];
    if (pre)
        for (long long i = 0; i < NG; ++i) qr[i] -= 0.37L * (LD)prev[i / n][i % n];
    for (int p = 0; p < passes; ++p) {
        if (algo == KRY_ORTH_CGS) {
            std::vector<LD> c(nvs, 0.0L);
            for (int j = j0; j < nv; ++j)
                for (long long i = 0; i < NG; ++i) c[j] += at(V, j, i) * qr[i];
            for (int j = j0; j < nv; ++j) {
                hr[j] += c[j];
                for (long long i = 0; i < NG; ++i) qr[i] -= c[j] * at(P, j, i);
            }
        } else {
            for (int j = j0; j < nv; ++j) {
                LD c = 0;
                for (long long i = 0; i < NG; ++i) c += at(V, j, i) * qr[i];
                hr[j] += c;
                for (long long i = 0; i < NG; ++i) qr[i] -= c * at(P, j, i);
            }
        }
    }
    LD nr2 = 0;
    for (long long i = 0; i < NG; ++i) nr2 += qr[i] * qr[i];
    const LD nr = sqrtl(nr2);

    std::vector<OrthArgs<T>> args(R);
    for (int r = 0; r < R; ++r) {
        PeerArgs pa;
        pa.world = R;
        pa.rank = r;
        pa.epoch_dev = epoch[r];
        pa.slots = slot_tab;
        pa.flags = flag_tab;
        args[r] = OrthArgs<T>{n, V[r], P[r], ldv, j0, nv, passes, algo, q[r], pre ? prev[r] : nullptr,
                              pre ? precoef[r] : nullptr, h[r], h[r] + nvs, vnext[r], partials[r], pa};
    }
    bool ran;
    if (R > 1)
        ran = emul_launch_ranks(R, G, KRY_THREADS, 0, [&args](int r) { orth_kernel<T, VEC, true>(args[r]); });
    else
        ran = emul_launch(G, KRY_THREADS, 0, [&args]() { orth_kernel<T, VEC, false>(args[0]); });
    if (!ran) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    double eh = 0, eq = 0, ev = 0, en = 0;
    bool same = true;
    for (int r = 0; r < R; ++r) {
        for (int j = 0; j < nvs; ++j) {
            const double want = (j >= j0 && j < nv) ? 0.5 + (double)hr[j] : 0.5;
            eh = fmax(eh, fabs(h[r][j] - want));
            if (h[r][j] != h[0][j]) same = false;          // rank-order sums: bitwise identical on every rank
        }
        en = fmax(en, fabs(h[r][nvs] - (double)nr) / (double)nr);
        if (h[r][nvs] != h[0][nvs]) same = false;
        if (R > 1 && epoch[r][0] != epoch[0][0]) same = false;
    }
    double qmax = 0;
    for (long long i = 0; i < NG; ++i) qmax = fmax(qmax, fabs((double)qr[i]));
    for (long long i = 0; i < NG; ++i) {
        eq = fmax(eq, fabs((double)q[i / n][i % n] - (double)qr[i]) / qmax);
        ev = fmax(ev, fabs((double)vnext[i / n][i % n] - (double)(qr[i] / nr)));
    }
    const double tol = Tol<T>::v * (sizeof(T) == 4 ? std::sqrt((double)nvs) : 1.0);
    const bool ok = same && eh <= 10 * tol && eq <= tol && ev <= tol && en <= tol && (R == 1 || epoch[0][0] > 6);
    printf("%s orth T=%s VEC=%d algo=%d passes=%d nv=%d j0=%d n=%lld G=%d R=%d pre=%d: h %.2e q %.2e vnext %.2e nrm %.2e "
           "ranks identical %d epoch %llu\n", ok ? "ok" : "FAIL", sizeof(T) == 8 ? "f64" : "f32", VEC, algo, passes, nv, j0,
           n, G, R, pre, eh, eq, ev, en, (int)same, epoch[0][0]);
    return ok ? 0 : 1;
}

// ------------------------------------------------------------------ proj_kernel
template <typename T, int VEC>
static int run_proj(int d, int iterations, long long n, int G, int with_qr) {
    std::mt19937_64 rng(99 + d + n);
    std::normal_distribution<double> nd;
    const long long ld = VEC > 1 ? (n + 7) / 8 * 8 + 8 : n + 3;
    T* W = dev_alloc<T>((size_t)d * ld);
    T* V = dev_alloc<T>((size_t)d * ld);
    T* a = dev_alloc<T>(n + 8);
    double* Q = dev_alloc<double>(d * d);
    double* Rm = dev_alloc<double>(d * d);
    double* cfirst = dev_alloc<double>(d);
    double* partials = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
    const double sc = 1.0 / std::sqrt((double)n);
    for (long long i = 0; i < (long long)d * ld; ++i) {
        W[i] = (T)(nd(rng) * sc);
        V[i] = (T)(nd(rng) * sc);
    }
    for (long long i = 0; i < n; ++i) a[i] = (T)nd(rng);
    // a well conditioned small transform: Q a rotation-like dense matrix, R upper triangular with a strong diagonal
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            Q[i * d + j] = nd(rng) / std::sqrt((double)d);
            Rm[i * d + j] = j > i ? 0.3 * nd(rng) : (j == i ? 2.0 + 0.1 * i : 0.0);
        }
    std::vector<LD> ar(n), c0(d);
    for (long long i = 0; i < n; ++i) ar[i] = (LD)a[i];
    for (int it = 0; it < iterations; ++it) {
        std::vector<LD> c(d, 0.0L), t(d, 0.0L);
        for (int j = 0; j < d; ++j)
            for (long long i = 0; i < n; ++i) c[j] += (LD)W[(long long)j * ld + i] * ar[i];
        if (it == 0) c0 = c;
        if (with_qr) {
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) t[i] += (LD)Q[j * d + i] * c[j];            // Q^T c
            for (int j = d - 1; j >= 0; --j) {                                          // R x = t
                t[j] /= (LD)Rm[j * d + j];
                for (int i = 0; i < j; ++i) t[i] -= t[j] * (LD)Rm[i * d + j];
            }
        } else {
            t = c;
        }
        for (int j = 0; j < d; ++j)
            for (long long i = 0; i < n; ++i) ar[i] -= t[j] * (LD)V[(long long)j * ld + i];
    }
    ProjArgs<T> p = {n, W, ld, V, ld, d, iterations, a, with_qr ? Q : nullptr, with_qr ? Rm : nullptr, cfirst, partials};
    if (!emul_launch(G, KRY_THREADS, 0, [p]() { proj_kernel<T, VEC>(p); })) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    double ea = 0, ec = 0, amax = 0;
    for (long long i = 0; i < n; ++i) amax = fmax(amax, fabs((double)ar[i]));
    for (long long i = 0; i < n; ++i) ea = fmax(ea, fabs((double)a[i] - (double)ar[i]) / amax);
    for (int j = 0; j < d; ++j) ec = fmax(ec, fabs(cfirst[j] - (double)c0[j]));
    const double tol = Tol<T>::v * 10;
    const bool ok = ea <= tol * iterations && ec <= tol;
    printf("%s proj T=%s VEC=%d d=%d iterations=%d n=%lld G=%d QR=%d: a %.2e c_first %.2e\n", ok ? "ok" : "FAIL",
           sizeof(T) == 8 ? "f64" : "f32", VEC, d, iterations, n, G, with_qr, ea, ec);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 13 && !strcmp(argv[1], "orth")) {
        const bool f64 = !strcmp(argv[2], "f64");
        const int vec = atoi(argv[3]);
        const int algo = atoi(argv[4]), passes = atoi(argv[5]), nv = atoi(argv[6]), j0 = atoi(argv[7]);
        const long long n = atoll(argv[8]);
        const int G = atoi(argv[9]), R = atoi(argv[10]), pre = atoi(argv[11]), sp = atoi(argv[12]);
        if (f64 && vec == 2) return run_orth<double, 2>(algo, passes, nv, j0, n, G, R, pre, sp);
        if (f64 && vec == 1) return run_orth<double, 1>(algo, passes, nv, j0, n, G, R, pre, sp);
        if (!f64 && vec == 4) return run_orth<float, 4>(algo, passes, nv, j0, n, G, R, pre, sp);
        if (!f64 && vec == 1) return run_orth<float, 1>(algo, passes, nv, j0, n, G, R, pre, sp);
    }
    if (argc >= 9 && !strcmp(argv[1], "proj")) {
        const bool f64 = !strcmp(argv[2], "f64");
        const int vec = atoi(argv[3]), d = atoi(argv[4]), its = atoi(argv[5]);
        const long long n = atoll(argv[6]);
        const int G = atoi(argv[7]), qr = atoi(argv[8]);
        if (f64 && vec == 2) return run_proj<double, 2>(d, its, n, G, qr);
        if (f64 && vec == 1) return run_proj<double, 1>(d, its, n, G, qr);
        if (!f64 && vec == 4) return run_proj<float, 4>(d, its, n, G, qr);
    }
    fprintf(stderr, "usage: see the header of this file\n");
    return 2;
}
