// Device code of the small (latency-bound) recurrences (csrc/kry_small.cu): the GMRES Givens / Hessenberg update,
// the triangular solves, the MINRES sliding QR, the projector's small transform, and their complex twins.  A
// header of its own so that the CPU test tier can run them over the CUDA execution emulator (tests/csrc/cuda_emul,
// tests/test_small_emul_cpu.py).
#pragma once
#include "kry_common.cuh"
#include "kry_small_core.h"
#include "kry_givens_dev.cuh"

// the dynamic shared memory of a kernel as doubles
#ifdef KRY_EMUL
#define KRY_DYN_SMEM_DOUBLES(name) double* name = reinterpret_cast<double*>(kry_emul_dynamic_smem())
#else
#define KRY_DYN_SMEM_DOUBLES(name) extern __shared__ double name[]
#endif

__global__ void __launch_bounds__(128) givens_kernel(int k, double* hcol, double* rcol, double* cs, double* y,
                                                     double* mailbox) {
    KRY_DYN_SMEM_DOUBLES(sh);
    givens_body(k, hcol, rcol, cs, y, mailbox, sh);
}

// TR: R is stored column after column (entry (i, j) at R[j * ldr + i]) -- the layout the Givens kernel leaves
// behind when every step's rcol points at its own row of a device-resident array
template <bool TR>
__global__ void __launch_bounds__(128) tri_solve_kernel(int k, const double* R, long long ldr, const double* y,
                                                        double* out) {
    KRY_DYN_SMEM_DOUBLES(sh);
    double* x = sh;  // k
    for (int i = threadIdx.x; i < k; i += blockDim.x) x[i] = y[i];
    __syncthreads();
    // column-oriented back substitution (LAPACK trtrs order); the column update is parallel
    for (int j = k - 1; j >= 0; --j) {
        __shared__ double xj;
        if (threadIdx.x == 0) {
            xj = x[j] / R[(long long)j * ldr + j];
            x[j] = xj;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < j; i += blockDim.x)
            x[i] = fma(-xj, TR ? R[(long long)j * ldr + i] : R[(long long)i * ldr + j], x[i]);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = x[i];
}

// state: [0]G1c [1]G1s [2]G1valid [3]G2c [4]G2s [5]G2valid [6]y0 [7]unused
//        [8]R0 [9]R1 [10]R2 [11]ycoef
__global__ void minres_recur_kernel(int k, double* h3, double* st, int shift, double* mailbox) {
    if (threadIdx.x != 0) return;
    double R0 = 0.0, R1 = h3[0], R2, R3;                // linsys.py:827-828 (H[k-1,k]; 0 for k == 0)
    if (k == 0) R1 = 0.0;
    if (st[2] != 0.0) kry_rot(st[0], st[1], R0, R1);    // :829-830
    R2 = h3[1];                                         // :833
    R3 = h3[2];
    if (st[5] != 0.0) kry_rot(st[3], st[4], R1, R2);    // :834-835
    st[0] = st[3]; st[1] = st[4]; st[2] = st[5];        // :836
    double c, s;
    kry_drotg(R2, R3, c, s);                            // :838
    st[3] = c; st[4] = s; st[5] = 1.0;
    R2 = __dadd_rn(__dmul_rn(c, R2), __dmul_rn(s, R3)); // :839  r = c*a + s*b
    double y0 = st[6], y1 = 0.0;
    kry_rot(c, s, y0, y1);                              // :841
    st[8] = R0; st[9] = R1; st[10] = R2; st[11] = y0;   // :844, :846
    st[6] = y1;                                         // :847
    mailbox[0] = fabs(y1);                              // :849
    mailbox[1] = R0; mailbox[2] = R1; mailbox[3] = R2; mailbox[4] = y0;
    mailbox[5] = h3[0]; mailbox[6] = h3[1]; mailbox[7] = h3[2];
    if (shift) {
        h3[0] = h3[2];   // next step's H[k, k+1] = H[k+1, k]   (utils.py:1003)
        h3[1] = 0.0;     // alpha accumulates with +=
    }
}

__global__ void __launch_bounds__(128) small_qr_apply_kernel(int d, const double* Q, const double* R,
                                                             const double* c_in, double* c_out) {
    KRY_DYN_SMEM_DOUBLES(sh);
    double* c = sh;        // d
    double* t = sh + d;    // d
    for (int i = threadIdx.x; i < d; i += blockDim.x) c[i] = c_in[i];
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < d; ++j) s = fma(Q[(long long)j * d + i], c[j], s);     // Q^H c
        t[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = d - 1; j >= 0; --j) {   // column-oriented back substitution with R
            const double xj = t[j] / R[(long long)j * d + j];
            t[j] = xj;
            for (int i = 0; i < j; ++i) t[i] = fma(-xj, R[(long long)i * d + j], t[i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) c_out[i] = t[i];
}

// Complex twins of givens_kernel / tri_solve_kernel (complex numbers interleaved re/im in
// double arrays; the serial cores live in kry_small_core.h and are unit-tested on the host).
// cs: 4 doubles per rotation [c, flag, s_re, s_im].
// mailbox: [ |y[k+1]|, H[0..k+1,k] (2(k+2) doubles), R[0..k+1,k] (2(k+2) doubles) ].
__global__ void __launch_bounds__(128) givens_z_kernel(int k, double* hcol, double* rcol, double* cs, double* y,
                                                       double* mailbox) {
    KRY_DYN_SMEM_DOUBLES(sh);
    const int nr = 2 * (k + 2);
    double* r = sh;           // 2(k+2)
    double* rot = sh + nr;    // 4k
    for (int i = threadIdx.x; i < nr; i += blockDim.x) r[i] = hcol[i];
    for (int i = threadIdx.x; i < 4 * k; i += blockDim.x) rot[i] = cs[i];
    __syncthreads();
    for (int i = threadIdx.x; i < nr; i += blockDim.x) {
        mailbox[1 + i] = r[i];
        hcol[i] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double yy[4] = {y[2 * k], y[2 * k + 1], y[2 * k + 2], y[2 * k + 3]};
        double rn[4];
        const double res = kryc_givens_step(k, r, rot, rn, yy);
        for (int i = 0; i < 4; ++i) {
            cs[4 * k + i] = rn[i];
            y[2 * k + i] = yy[i];
        }
        mailbox[0] = res;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nr; i += blockDim.x) {
        rcol[i] = r[i];
        mailbox[1 + nr + i] = r[i];
    }
}

__global__ void tri_solve_z_kernel(int k, const double* R, long long ldr, const double* y, double* out) {
    KRY_DYN_SMEM_DOUBLES(sh);
    double* x = sh;   // 2k
    for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) x[i] = y[i];
    __syncthreads();
    if (threadIdx.x == 0) kryc_tri_solve(k, R, ldr, x);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) out[i] = x[i];
}

