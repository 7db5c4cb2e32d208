// Latency-bound device recurrences: GMRES Givens/Hessenberg update, triangular
// solve, MINRES sliding QR.  One small CTA each; inputs are staged into shared
// memory in parallel, one thread runs the (inherently serial) recurrence, the
// results go to device state and to the pinned host mailbox in parallel.
#include "kry_common.cuh"
#include "kry_small_core.h"
#include "kry_givens_dev.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

__global__ void __launch_bounds__(128) givens_kernel(int k, double* hcol, double* rcol, double* cs, double* y,
                                                     double* mailbox) {
    extern __shared__ double sh[];
    givens_body(k, hcol, rcol, cs, y, mailbox, sh);
}

// TR: R is stored column after column (entry (i, j) at R[j * ldr + i]) -- the layout the Givens kernel leaves
// behind when every step's rcol points at its own row of a device-resident array
template <bool TR>
__global__ void __launch_bounds__(128) tri_solve_kernel(int k, const double* R, long long ldr, const double* y,
                                                        double* out) {
    extern __shared__ double sh[];
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
    extern __shared__ double sh[];
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
    extern __shared__ double sh[];
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
    extern __shared__ double sh[];
    double* x = sh;   // 2k
    for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) x[i] = y[i];
    __syncthreads();
    if (threadIdx.x == 0) kryc_tri_solve(k, R, ldr, x);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) out[i] = x[i];
}

extern "C" {

int kry_small_qr_apply(kry_ctx* ctx, int d, const double* Q_dev, const double* R_dev, const double* c_in_dev,
                       double* c_out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(d >= 0 && d <= 2048, "bad d");
    if (d == 0) return KRY_OK;
    KRY_REQUIRE(Q_dev && R_dev && c_in_dev && c_out_dev, "NULL argument");
    small_qr_apply_kernel<<<1, 128, sizeof(double) * 2 * (size_t)d, ctx->stream>>>(d, Q_dev, R_dev, c_in_dev, c_out_dev);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_givens_update(kry_ctx* ctx, int k, double* hcol_dev, double* rcol_dev, double* cs_dev, double* y_dev,
                      int mailbox_off) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && hcol_dev && rcol_dev && cs_dev && y_dev, "bad arguments");
    KRY_REQUIRE(k <= 2000, "k too large for the single-CTA Givens update (use restarts)");
    KRY_REQUIRE(mailbox_off >= 0 && mailbox_off + 2 * k + 5 <= KRY_MAILBOX_DOUBLES, "mailbox overflow");
    size_t smem = sizeof(double) * (size_t)(3 * k + 2);
    givens_kernel<<<1, 128, smem, ctx->stream>>>(k, hcol_dev, rcol_dev, cs_dev, y_dev, ctx->d_mailbox + mailbox_off);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_tri_solve(kry_ctx* ctx, int k, const double* R_dev, long long ldr, const double* y_dev, double* out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && ldr >= k, "bad arguments");
    if (k == 0) return KRY_OK;
    KRY_REQUIRE(R_dev && y_dev && out_dev, "NULL argument");
    KRY_REQUIRE(k <= 6000, "k too large");
    tri_solve_kernel<false><<<1, 128, sizeof(double) * (size_t)k, ctx->stream>>>(k, R_dev, ldr, y_dev, out_dev);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_tri_solve_t(kry_ctx* ctx, int k, const double* Rt_dev, long long ldr, const double* y_dev, double* out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && ldr >= k, "bad arguments");
    if (k == 0) return KRY_OK;
    KRY_REQUIRE(Rt_dev && y_dev && out_dev, "NULL argument");
    KRY_REQUIRE(k <= 6000, "k too large");
    tri_solve_kernel<true><<<1, 128, sizeof(double) * (size_t)k, ctx->stream>>>(k, Rt_dev, ldr, y_dev, out_dev);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_minres_recur(kry_ctx* ctx, int k, double* h3_dev, double* st_dev, int shift, int mailbox_off) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && h3_dev && st_dev, "bad arguments");
    KRY_REQUIRE(mailbox_off >= 0 && mailbox_off + 8 <= KRY_MAILBOX_DOUBLES, "mailbox overflow");
    minres_recur_kernel<<<1, 32, 0, ctx->stream>>>(k, h3_dev, st_dev, shift, ctx->d_mailbox + mailbox_off);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_givens_update_z(kry_ctx* ctx, int k, double* hcol_dev, double* rcol_dev, double* cs_dev, double* y_dev,
                        int mailbox_off) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && hcol_dev && rcol_dev && cs_dev && y_dev, "bad arguments");
    KRY_REQUIRE(k <= 1000, "k too large for the single-CTA Givens update (use restarts)");
    KRY_REQUIRE(mailbox_off >= 0 && mailbox_off + 4 * k + 9 <= KRY_MAILBOX_DOUBLES, "mailbox overflow");
    size_t smem = sizeof(double) * (size_t)(2 * (k + 2) + 4 * k);
    givens_z_kernel<<<1, 128, smem, ctx->stream>>>(k, hcol_dev, rcol_dev, cs_dev, y_dev, ctx->d_mailbox + mailbox_off);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_tri_solve_z(kry_ctx* ctx, int k, const double* R_dev, long long ldr, const double* y_dev, double* out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(k >= 0 && ldr >= k, "bad arguments");
    if (k == 0) return KRY_OK;
    KRY_REQUIRE(R_dev && y_dev && out_dev, "NULL argument");
    KRY_REQUIRE(k <= 3000, "k too large");
    tri_solve_z_kernel<<<1, 128, sizeof(double) * 2 * (size_t)k, ctx->stream>>>(k, R_dev, ldr, y_dev, out_dev);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

}  // extern "C"
