// TEST INFRASTRUCTURE (CPU tier) -- runs the DEVICE code of kry_lanczos_diag / kry_lanczos_diag_dist
// (krypy_b200/csrc/kry_lanczos_kernels.cuh, included unchanged: the whole Lanczos step for a diagonal inner-product
// matrix in one cooperative kernel, BASELINE config C5) on the host over the CUDA execution emulator of
// tests/csrc/cuda_emul, single rank and row-partitioned over emulated ranks, against a long-double reference that
// applies the kernel's rounding points (B q rounded to the storage type before it enters a dot).
// Driven by tests/test_lanczos_emul_cpu.py:
//     lanczos_emul_host <dtype f64|f32> <vec> <n per rank> <grid> <ranks> <pre 0|1> <vnext 0|1>
#define KRY_EMUL 1
#include <random>

#include "emul_runtime.h"

#include "kry_lanczos_kernels.cuh"

typedef long double LD;

template <typename T, int VEC>
static int run(long long n, int G, int R, int pre, int want_next) {
    std::mt19937_64 rng(555 + n + R);
    std::normal_distribution<double> nd;
    std::uniform_real_distribution<double> ud(1.0, 2.0);
    std::vector<T*> vprev(R), vk(R), b(R), q(R), vnext(R);
    std::vector<double*> h3(R), precoef(R), partials(R);
    std::vector<unsigned long long*> epoch(R);
    unsigned long long** flag_tab = dev_alloc<unsigned long long*>(R);
    double** slot_tab = dev_alloc<double*>(R);
    const long long NG = n * R;
    const double sc = 1.0 / std::sqrt((double)NG);
    for (int r = 0; r < R; ++r) {
        vprev[r] = dev_alloc<T>(n + 8);
        vk[r] = dev_alloc<T>(n + 8);
        b[r] = dev_alloc<T>(n + 8);
        q[r] = dev_alloc<T>(n + 8);
        vnext[r] = dev_alloc<T>(n + 8);
        h3[r] = dev_alloc<double>(3);
        precoef[r] = dev_alloc<double>(1);
        partials[r] = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
        epoch[r] = dev_alloc<unsigned long long>(1);
        flag_tab[r] = dev_alloc<unsigned long long>(PEER_MAX_RANKS);
        slot_tab[r] = dev_alloc<double>(2ull * PEER_MAX_RANKS * PEER_SLOT);
        epoch[r][0] = 3;
        for (int p = 0; p < R; ++p) flag_tab[r][p] = 3;
        for (long long i = 0; i < n; ++i) {
            vprev[r][i] = (T)(nd(rng) * sc);
            vk[r][i] = (T)(nd(rng) * sc);
            b[r][i] = (T)ud(rng);
            q[r][i] = (T)nd(rng);
            vnext[r][i] = (T)NAN;
        }
        h3[r][0] = 0.4;       // H[k-1,k]: also the pre-subtraction coefficient
        h3[r][1] = 0.25;      // H[k,k] accumulates
        h3[r][2] = -1.0;
        precoef[r][0] = 0.4;
    }
    // reference with the kernel's rounding points
    auto rT = [](LD v) { return (LD)(T)(double)v; };
    std::vector<LD> qr(NG);
    LD alpha = 0, beta2 = 0;
    for (long long i = 0; i < NG; ++i) {
        const int r = (int)(i / n);
        const long long l = i % n;
        LD qe = (LD)q[r][l];
        if (pre) qe = rT(qe - 0.4L * (LD)vprev[r][l]);
        qr[i] = qe;
        alpha += (LD)vk[r][l] * rT((LD)b[r][l] * qe);
    }
    for (long long i = 0; i < NG; ++i) {
        const int r = (int)(i / n);
        const long long l = i % n;
        qr[i] = rT(qr[i] - alpha * (LD)vk[r][l]);
        beta2 += qr[i] * rT((LD)b[r][l] * qr[i]);
    }
    const LD beta = sqrtl(fabsl(beta2));

    std::vector<LanczosArgs<T>> args(R);
    for (int r = 0; r < R; ++r) {
        PeerArgs pa;
        pa.world = R;
        pa.rank = r;
        pa.epoch_dev = epoch[r];
        pa.slots = slot_tab;
        pa.flags = flag_tab;
        args[r] = LanczosArgs<T>{n, pre ? vprev[r] : nullptr, vk[r], b[r], q[r], pre ? precoef[r] : nullptr, h3[r],
                                 want_next ? vnext[r] : nullptr, partials[r], pa};
    }
    bool ran;
    if (R > 1)
        ran = emul_launch_ranks(R, G, KRY_THREADS, 0, [&args](int r) { lanczos_diag_kernel<T, VEC, true>(args[r]); });
    else
        ran = emul_launch(G, KRY_THREADS, 0, [&args]() { lanczos_diag_kernel<T, VEC, false>(args[0]); });
    if (!ran) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    const double eps = sizeof(T) == 8 ? 1e-13 : 2e-6;
    double ea = 0, eb = 0, eq = 0, ev = 0;
    bool same = true;
    for (int r = 0; r < R; ++r) {
        ea = fmax(ea, fabs(h3[r][1] - (0.25 + (double)alpha)));
        eb = fmax(eb, fabs(h3[r][2] - (double)beta) / (double)beta);
        if (h3[r][1] != h3[0][1] || h3[r][2] != h3[0][2] || h3[r][0] != 0.4) same = false;
        if (R > 1 && epoch[r][0] != 5) same = false;       // two exchanges
    }
    double qmax = 0;
    for (long long i = 0; i < NG; ++i) qmax = fmax(qmax, fabs((double)qr[i]));
    for (long long i = 0; i < NG; ++i) {
        eq = fmax(eq, fabs((double)q[i / n][i % n] - (double)qr[i]) / qmax);
        if (want_next) ev = fmax(ev, fabs((double)vnext[i / n][i % n] - (double)(qr[i] / beta)));
    }
    const bool ok = same && ea <= 10 * eps && eb <= 10 * eps && eq <= 10 * eps && ev <= 10 * eps;
    printf("%s lanczos T=%s VEC=%d n=%lld G=%d R=%d pre=%d next=%d: alpha %.2e beta %.2e q %.2e vnext %.2e ranks identical %d\n",
           ok ? "ok" : "FAIL", sizeof(T) == 8 ? "f64" : "f32", VEC, n, G, R, pre, want_next, ea, eb, eq, ev, (int)same);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc < 8) {
        fprintf(stderr, "usage: see the header of this file\n");
        return 2;
    }
    const bool f64 = !strcmp(argv[1], "f64");
    const int vec = atoi(argv[2]);
    const long long n = atoll(argv[3]);
    const int G = atoi(argv[4]), R = atoi(argv[5]), pre = atoi(argv[6]), nx = atoi(argv[7]);
    if (f64 && vec == 2) return run<double, 2>(n, G, R, pre, nx);
    if (f64 && vec == 1) return run<double, 1>(n, G, R, pre, nx);
    if (!f64 && vec == 4) return run<float, 4>(n, G, R, pre, nx);
    if (!f64 && vec == 1) return run<float, 1>(n, G, R, pre, nx);
    return 2;
}
