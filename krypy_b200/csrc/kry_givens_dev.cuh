// Device code of the GMRES Givens / Hessenberg update (linsys.py:982-993), shared by the stand-alone
// kernel (kry_small.cu) and by the row-partitioned fused update kernel (kry_dist.cu), whose extra CTA runs
// it concurrently with the sweep instead of as a kernel of its own.
#pragma once
#include "kry_common.cuh"

// BLAS drotg (reference BLAS 3.10 / OpenBLAS >= 0.3.20 algorithm), returns c, s
// with r = sigma*hypot(a,b), sigma = sign of the larger-magnitude input.
// krypy/utils.py:421-424 takes (c, s) from scipy.linalg.blas.drotg.
__device__ __forceinline__ void kry_drotg(double a, double b, double& c, double& s) {
    const double safmin = 2.2250738585072014e-308, safmax = 4.4942328371557898e+307;
    const double anorm = fabs(a), bnorm = fabs(b);
    if (bnorm == 0.0) {
        c = 1.0;
        s = 0.0;
    } else if (anorm == 0.0) {
        c = 0.0;
        s = 1.0;
    } else {
        const double scl = fmin(safmax, fmax(safmin, fmax(anorm, bnorm)));
        const double sigma = (anorm > bnorm) ? copysign(1.0, a) : copysign(1.0, b);
        const double as = a / scl, bs = b / scl;
        const double r = sigma * (scl * sqrt(__dadd_rn(__dmul_rn(as, as), __dmul_rn(bs, bs))));
        c = a / r;
        s = b / r;
    }
}

// G = [[c, s], [-s, c]] applied to (x0, x1): numpy.dot(G, x), utils.py:434-436
__device__ __forceinline__ void kry_rot(double c, double s, double& x0, double& x1) {
    const double t0 = __dadd_rn(__dmul_rn(c, x0), __dmul_rn(s, x1));
    const double t1 = __dadd_rn(__dmul_rn(-s, x0), __dmul_rn(c, x1));
    x0 = t0;
    x1 = t1;
}

// One CTA.  sh: >= 3k + 2 doubles of shared memory.  hcol[0..k+1] is column k of H (raw), zeroed on exit
// (h accumulates with +=); rcol, cs, y: device state; mailbox: [0] |y[k+1]|, [1..k+2] H column,
// [k+3..2k+4] R column.
__device__ __forceinline__ void givens_body(int k, double* hcol, double* rcol, double* cs, double* y, double* mailbox,
                                            double* sh) {
    double* r = sh;              // k+2
    double* rot = sh + (k + 2);  // 2k
    for (int i = threadIdx.x; i < k + 2; i += blockDim.x) r[i] = hcol[i];
    for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) rot[i] = cs[i];
    __syncthreads();
    // raw Hessenberg column goes to the host (invariant-subspace test, H attribute)
    for (int i = threadIdx.x; i < k + 2; i += blockDim.x) {
        mailbox[1 + i] = r[i];
        hcol[i] = 0.0;   // h accumulates with += (reorthogonalisation): leave it zeroed
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < k; ++i) kry_rot(rot[2 * i], rot[2 * i + 1], r[i], r[i + 1]);   // linsys.py:985-986
        double c, s;
        kry_drotg(r[k], r[k + 1], c, s);                                                  // linsys.py:989
        cs[2 * k] = c;
        cs[2 * k + 1] = s;
        kry_rot(c, s, r[k], r[k + 1]);                                                    // linsys.py:990
        double y0 = y[k], y1 = y[k + 1];
        kry_rot(c, s, y0, y1);                                                            // linsys.py:991
        y[k] = y0;
        y[k + 1] = y1;
        mailbox[0] = fabs(y1);                                                            // linsys.py:993
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k + 2; i += blockDim.x) {
        rcol[i] = r[i];
        mailbox[k + 3 + i] = r[i];
    }
}
