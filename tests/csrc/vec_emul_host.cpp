// TEST INFRASTRUCTURE (CPU tier) -- runs the DEVICE code of the N-sized streaming kernels
// (krypy_b200/csrc/kry_vec_kernels.cuh, included unchanged: kry_axpby, kry_axpy_dev, kry_scale_dev, kry_diag_mul,
// kry_rot90, kry_block_dot with its sqrt / accumulate epilogues, kry_block_axpy, kry_block_combine, kry_gemv_dense)
// on the host over the CUDA execution emulator of tests/csrc/cuda_emul against long-double references.
//     vec_emul_host <dtype f64|f32> <vec> <n> <nv> <grid>
#define KRY_EMUL 1
#include <random>

#include "emul_runtime.h"

#include "kry_vec_kernels.cuh"

typedef long double LD;

template <typename T, int VEC>
static int run(long long n, int nv, int G) {
    std::mt19937_64 rng(17 + n + nv);
    std::normal_distribution<double> nd;
    const double eps = sizeof(T) == 8 ? 1e-14 : 3e-6;
    const long long ldv = VEC > 1 ? (n + 7) / 8 * 8 + 8 : n + 3;
    T *x = dev_alloc<T>(n + 8), *y = dev_alloc<T>(n + 8), *z = dev_alloc<T>(n + 8), *d = dev_alloc<T>(n + 8);
    T* V = dev_alloc<T>((size_t)nv * ldv);
    double* coef = dev_alloc<double>(nv + 2);
    double* out = dev_alloc<double>(nv + 2);
    double* acc = dev_alloc<double>(nv + 2);
    double* partials = dev_alloc<double>(2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS);
    unsigned int* ticket = dev_alloc<unsigned int>(2);
    for (long long i = 0; i < n; ++i) {
        x[i] = (T)nd(rng);
        y[i] = (T)nd(rng);
        d[i] = (T)(1.0 + 0.5 * nd(rng));
    }
    for (long long i = 0; i < (long long)nv * ldv; ++i) V[i] = (T)nd(rng);
    for (int j = 0; j < nv + 2; ++j) {
        coef[j] = nd(rng);
        acc[j] = 0.25;
        out[j] = -9.0;
    }
    std::vector<LD> x0(n), y0(n);
    for (long long i = 0; i < n; ++i) {
        x0[i] = (LD)x[i];
        y0[i] = (LD)y[i];
    }
    double worst = 0;
    const char* bad = "";
    auto check = [&](const char* what, double e) {
        if (!(e <= eps)) bad = what;
        if (!(e <= worst)) worst = e;
    };
    bool ran = true;
    // axpby: z = a x + b y, and the y == NULL form
    ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() { axpby_kernel<T, VEC>(n, 0.75, x, -1.25, y, z); });
    {
        double e = 0;
        for (long long i = 0; i < n; ++i) e = fmax(e, fabs((double)z[i] - (double)(0.75L * x0[i] - 1.25L * y0[i])));
        check("axpby", e / 4);
    }
    ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() { axpby_kernel<T, VEC>(n, -2.0, x, 0.0, (const T*)nullptr, z); });
    {
        double e = 0;
        for (long long i = 0; i < n; ++i) e = fmax(e, fabs((double)z[i] - (double)(-2.0L * x0[i])));
        check("axpby without y", e / 4);
    }
    // axpy_dev: y += sign * coef[0] * x
    ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() { axpy_dev_kernel<T, VEC>(n, coef, -1.0, x, y); });
    {
        double e = 0;
        for (long long i = 0; i < n; ++i) e = fmax(e, fabs((double)y[i] - (double)(y0[i] - (LD)coef[0] * x0[i])));
        check("axpy_dev", e / 4);
        for (long long i = 0; i < n; ++i) y0[i] = (LD)y[i];
    }
    // scale_dev: divide and multiply
    coef[1] = 1.7;
    ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() { scale_dev_kernel<T, VEC>(n, coef + 1, 1, -1.0, x, z); });
    {
        double e = 0;
        for (long long i = 0; i < n; ++i) e = fmax(e, fabs((double)z[i] - (double)(-x0[i] / 1.7L)));
        check("scale_dev divide", e / 4);
    }
    ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() { scale_dev_kernel<T, VEC>(n, coef + 1, 0, 2.0, x, z); });
    {
        double e = 0;
        for (long long i = 0; i < n; ++i) e = fmax(e, fabs((double)z[i] - (double)(2.0L * x0[i] * 1.7L)));
        check("scale_dev multiply", e / 8);
    }
    // diag_mul
    ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() { diag_mul_kernel<T, VEC>(n, d, x, z); });
    {
        double e = 0;
        for (long long i = 0; i < n; ++i) e = fmax(e, fabs((double)z[i] - (double)((LD)d[i] * x0[i])));
        check("diag_mul", e / 8);
    }
    // rot90 on n / 2 complex numbers: exact (a swap and a sign)
    if (n >= 2) {
        const long long nc = n / 2;
        ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() { rot90_kernel<T>(nc, x, z); });
        bool exact = true;
        for (long long k = 0; k < nc; ++k)
            if (z[2 * k] != -x[2 * k + 1] || z[2 * k + 1] != x[2 * k]) exact = false;
        check("rot90", exact ? 0.0 : 1.0);
    }
    // block_dot: plain, sqrt epilogue on <x, x>, accumulate
    ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() {
        block_dot_kernel<T, VEC>(n, V, ldv, nv, x, partials, ticket, out, 0, acc);
    });
    {
        double e = 0;
        for (int j = 0; j < nv; ++j) {
            LD s = 0;
            for (long long i = 0; i < n; ++i) s += (LD)V[(long long)j * ldv + i] * x0[i];
            e = fmax(e, fabs(out[j] - (double)s) / std::sqrt((double)n));
            e = fmax(e, fabs(acc[j] - (0.25 + out[j])));
        }
        if (out[nv] != -9.0 || ticket[0] != 0) e = 1.0;
        check("block_dot", e / 8);
    }
    ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() {
        block_dot_kernel<T, VEC>(n, x, n, 1, x, partials, ticket, out, 1, (double*)nullptr);
    });
    {
        LD s = 0;
        for (long long i = 0; i < n; ++i) s += x0[i] * x0[i];
        check("block_dot sqrt", fabs(out[0] - (double)sqrtl(s)) / (double)sqrtl(s));
    }
    // block_axpy: y += sign * V^T coef ; block_combine: z = x + V^T coef and z = V^T coef
    ran = ran && emul_launch(G, KRY_THREADS, 64 * sizeof(double) + 1024, [=]() {
        block_axpy_kernel<T, VEC, false>(n, V, ldv, nv, coef, -1.0, (const T*)nullptr, y);
    });
    {
        double e = 0;
        for (long long i = 0; i < n; ++i) {
            LD s = y0[i];
            for (int j = 0; j < nv; ++j) s -= (LD)coef[j] * (LD)V[(long long)j * ldv + i];
            e = fmax(e, fabs((double)y[i] - (double)s));
        }
        check("block_axpy", e / (4.0 * std::sqrt((double)nv + 1)));
    }
    for (int with_x0 = 1; with_x0 >= 0; --with_x0) {
        const T* xx = with_x0 ? x : (const T*)nullptr;
        ran = ran && emul_launch(G, KRY_THREADS, 64 * sizeof(double) + 1024, [=]() {
            block_axpy_kernel<T, VEC, true>(n, V, ldv, nv, coef, 1.0, xx, z);
        });
        double e = 0;
        for (long long i = 0; i < n; ++i) {
            LD s = 0;
            for (int j = 0; j < nv; ++j) s += (LD)coef[j] * (LD)V[(long long)j * ldv + i];
            if (with_x0) s += x0[i];
            e = fmax(e, fabs((double)z[i] - (double)s));
        }
        check(with_x0 ? "block_combine" : "block_combine without x0", e / (4.0 * std::sqrt((double)nv + 1)));
    }
    // gemv: y = A x for the nv x n row-major matrix V (leading dimension ldv)
    {
        T* yy = dev_alloc<T>(nv + 2);
        ran = ran && emul_launch(G, KRY_THREADS, 0, [=]() { gemv_kernel<T>(nv, n, V, ldv, x, yy); });
        double e = 0;
        for (int j = 0; j < nv; ++j) {
            LD s = 0;
            for (long long i = 0; i < n; ++i) s += (LD)V[(long long)j * ldv + i] * x0[i];
            e = fmax(e, fabs((double)yy[j] - (double)s) / std::sqrt((double)n));
        }
        check("gemv", e / 8);
    }
    if (!ran) {
        printf("FAIL a CTA died\n");
        return 1;
    }
    const bool ok = bad[0] == 0;
    printf("%s vec kernels T=%s VEC=%d n=%lld nv=%d G=%d: worst scaled error %.2e%s%s\n", ok ? "ok" : "FAIL",
           sizeof(T) == 8 ? "f64" : "f32", VEC, n, nv, G, worst, ok ? "" : " in ", bad);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: see the header of this file\n");
        return 2;
    }
    const bool f64 = !strcmp(argv[1], "f64");
    const int vec = atoi(argv[2]);
    const long long n = atoll(argv[3]);
    const int nv = atoi(argv[4]), G = atoi(argv[5]);
    if (f64 && vec == 2) return run<double, 2>(n, nv, G);
    if (f64 && vec == 1) return run<double, 1>(n, nv, G);
    if (!f64 && vec == 4) return run<float, 4>(n, nv, G);
    if (!f64 && vec == 1) return run<float, 1>(n, nv, G);
    return 2;
}
