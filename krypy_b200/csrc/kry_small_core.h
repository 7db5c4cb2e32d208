// Serial cores of the complex (interleaved re/im) small recurrences: complex Givens / Hessenberg
// update and complex back substitution.  Plain C++ without CUDA types, compiled twice:
//   * by nvcc into the single-thread part of givens_z_kernel / tri_solve_z_kernel (kry_small.cu);
//   * by g++ into tests/_build/libkry_small_core_host.so, so the CPU test tier can run exactly
//     this code against scipy.linalg.blas.zrotg / numpy (tests/test_small_core_cpu.py).
// Reference semantics: krypy/utils.py:405-436 (Givens: drotg for real-valued input, zrotg
// otherwise; r = c*a + s*b; G = [[c, s], [-conj(s), c]]), krypy/linsys.py:982-993.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define KRY_HD __host__ __device__ __forceinline__
#else
#define KRY_HD static inline
#endif

// separately rounded products and sums (no FMA contraction), like numpy's complex arithmetic;
// the host build uses -ffp-contract=off
KRY_HD double kryc_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
KRY_HD double kryc_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// BLAS drotg (reference BLAS 3.10 algorithm; the twin of kry_drotg in kry_small.cu)
KRY_HD void kryc_drotg(double a, double b, double* c, double* s) {
    const double safmin = 2.2250738585072014e-308, safmax = 4.4942328371557898e+307;
    const double anorm = fabs(a), bnorm = fabs(b);
    if (bnorm == 0.0) {
        *c = 1.0;
        *s = 0.0;
    } else if (anorm == 0.0) {
        *c = 0.0;
        *s = 1.0;
    } else {
        const double scl = fmin(safmax, fmax(safmin, fmax(anorm, bnorm)));
        const double sigma = (anorm > bnorm) ? copysign(1.0, a) : copysign(1.0, b);
        const double as = a / scl, bs = b / scl;
        const double r = sigma * (scl * sqrt(kryc_add(kryc_mul(as, as), kryc_mul(bs, bs))));
        *c = a / r;
        *s = b / r;
    }
}

// BLAS zrotg (LAPACK 3.10 semantics: c real >= 0, s = conj(g) f / (|f| sqrt(|f|^2+|g|^2));
// f == 0: c = 0, s = conj(g)/|g|; g == 0: c = 1, s = 0).  Inputs are scaled by
// max(|re|,|im|) over both so the squares can neither overflow nor underflow.
KRY_HD void kryc_zrotg(double fr, double fi, double gr, double gi, double* c, double* sr, double* si) {
    if (gr == 0.0 && gi == 0.0) {
        *c = 1.0;
        *sr = 0.0;
        *si = 0.0;
        return;
    }
    if (fr == 0.0 && fi == 0.0) {
        const double u = fmax(fabs(gr), fabs(gi));
        const double ar = gr / u, ai = gi / u;
        const double d = sqrt(kryc_add(kryc_mul(ar, ar), kryc_mul(ai, ai)));
        *c = 0.0;
        *sr = ar / d;
        *si = -ai / d;
        return;
    }
    const double f1 = fmax(fabs(fr), fabs(fi)), g1 = fmax(fabs(gr), fabs(gi));
    const double u = fmax(f1, g1);
    const double ar = fr / u, ai = fi / u, br = gr / u, bi = gi / u;
    const double f2 = kryc_add(kryc_mul(ar, ar), kryc_mul(ai, ai));
    const double g2 = kryc_add(kryc_mul(br, br), kryc_mul(bi, bi));
    const double h2 = f2 + g2;
    if (f2 == 0.0) {
        // |f| underflowed against |g|: the rotation is (numerically) the f == 0 one with the phase of f
        const double fa = hypot(fr, fi), ga = sqrt(g2);
        const double pr = fr / fa, pi = fi / fa;                    // f/|f|
        const double qr = br / ga, qi = -bi / ga;                   // conj(g)/|g|
        *c = (fa / u) / ga;
        *sr = kryc_add(kryc_mul(pr, qr), -kryc_mul(pi, qi));
        *si = kryc_add(kryc_mul(pr, qi), kryc_mul(pi, qr));
        return;
    }
    *c = sqrt(f2 / h2);
    const double d = sqrt(kryc_mul(f2, h2));
    const double tr = ar / d, ti = ai / d;                          // f / sqrt(f2*h2)
    // s = conj(g) * t = (br - i bi)(tr + i ti)
    *sr = kryc_add(kryc_mul(br, tr), kryc_mul(bi, ti));
    *si = kryc_add(kryc_mul(br, ti), -kryc_mul(bi, tr));
}

// (x0, x1) <- G (x0, x1),  G = [[c, s], [-conj(s), c]], c real  (numpy.dot(G, x), utils.py:434-436)
KRY_HD void kryc_zrot(double c, double sr, double si, double* x0, double* x1) {
    const double x0r = x0[0], x0i = x0[1], x1r = x1[0], x1i = x1[1];
    // s*x1
    const double ar = kryc_add(kryc_mul(sr, x1r), -kryc_mul(si, x1i));
    const double ai = kryc_add(kryc_mul(sr, x1i), kryc_mul(si, x1r));
    // conj(s)*x0
    const double br = kryc_add(kryc_mul(sr, x0r), kryc_mul(si, x0i));
    const double bi = kryc_add(kryc_mul(sr, x0i), -kryc_mul(si, x0r));
    x0[0] = kryc_add(kryc_mul(c, x0r), ar);
    x0[1] = kryc_add(kryc_mul(c, x0i), ai);
    x1[0] = kryc_add(kryc_mul(c, x1r), -br);
    x1[1] = kryc_add(kryc_mul(c, x1i), -bi);
}

// Rotation record: 4 doubles [c, flag, s_re, s_im]; flag is informational (1: drotg branch).
// One GMRES step on column k (linsys.py:982-993):
//   r[0..k+1]   (complex, interleaved) column k of H on entry, column k of R on return
//   rot[0..k)   the stored rotations;  rot_new: receives rotation k
//   y2 = (y[k], y[k+1]) (4 doubles) rotated in place;  returns |y[k+1]|
KRY_HD double kryc_givens_step(int k, double* r, const double* rot, double* rot_new, double* y2) {
    for (int i = 0; i < k; ++i) kryc_zrot(rot[4 * i], rot[4 * i + 2], rot[4 * i + 3], r + 2 * i, r + 2 * i + 2);
    double c, sr, si = 0.0, flag;
    double* a = r + 2 * k;
    double* b = r + 2 * k + 2;
    if (a[1] == 0.0 && b[1] == 0.0) {          // numpy.isreal(x).all(): utils.py:419-424
        kryc_drotg(a[0], b[0], &c, &sr);
        flag = 1.0;
    } else {
        kryc_zrotg(a[0], a[1], b[0], b[1], &c, &sr, &si);
        flag = 0.0;
    }
    rot_new[0] = c;
    rot_new[1] = flag;
    rot_new[2] = sr;
    rot_new[3] = si;
    kryc_zrot(c, sr, si, a, b);
    kryc_zrot(c, sr, si, y2, y2 + 2);
    return hypot(y2[2], y2[3]);
}

// x <- R[:k,:k]^{-1} x, complex upper triangular R (row-major, leading dimension ldr complex
// elements), column-oriented back substitution (LAPACK trtrs order; linsys.py:946).
KRY_HD void kryc_tri_solve(int k, const double* R, long long ldr, double* x) {
    for (int j = k - 1; j >= 0; --j) {
        const double dr = R[2 * ((long long)j * ldr + j)], di = R[2 * ((long long)j * ldr + j) + 1];
        // Smith's complex division x[j] / R[j,j]
        double qr, qi;
        const double xr = x[2 * j], xi = x[2 * j + 1];
        if (fabs(dr) >= fabs(di)) {
            const double t = di / dr, den = dr + di * t;
            qr = (xr + xi * t) / den;
            qi = (xi - xr * t) / den;
        } else {
            const double t = dr / di, den = dr * t + di;
            qr = (xr * t + xi) / den;
            qi = (xi * t - xr) / den;
        }
        x[2 * j] = qr;
        x[2 * j + 1] = qi;
        for (int i = 0; i < j; ++i) {
            const double ar = R[2 * ((long long)i * ldr + j)], ai = R[2 * ((long long)i * ldr + j) + 1];
            x[2 * i] -= kryc_add(kryc_mul(ar, qr), -kryc_mul(ai, qi));
            x[2 * i + 1] -= kryc_add(kryc_mul(ar, qi), kryc_mul(ai, qr));
        }
    }
}
