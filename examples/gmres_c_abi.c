/* A plain C consumer of the krypy_b200 C ABI (include/krypy_b200.h): restarted GMRES(m) with the
 * fused block Gram-Schmidt kernel on the 2-D 5-point Laplacian, no Python and no torch -- device
 * memory comes from the CUDA runtime, the host does convergence control from the pinned mailbox.
 * It is what a non-Python binding (cgo / JNI / ...) would do with the same entry points.
 *
 * build:  gcc -O2 -I include examples/gmres_c_abi.c -o gmres_c_abi \
 *             -L krypy_b200 -lkrypy_b200 -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/krypy_b200
 * run  :  ./gmres_c_abi [n=512] [m=30] [cycles=5]        (needs a B200)
 *
 * The loop mirrors krypy/linsys.py:951-997 (Gmres._solve) + :1021-1072 (restarts):
 *   q = A v_k                          kry_spmv_csr            (utils.py:968)
 *   h = V^T q; q -= V h; v_{k+1}=q/|q| kry_orth_fused          (utils.py:996-1045)
 *   Givens update of the Hessenberg LS kry_givens_update       (linsys.py:982-993)
 *   x = x0 + V R^-1 y                  kry_tri_solve + kry_block_combine (linsys.py:941-949)
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "krypy_b200.h"

#define CK(call)                                                                        \
    do {                                                                                \
        int rc_ = (call);                                                               \
        if (rc_ != 0) {                                                                 \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, kry_last_error());            \
            return 1;                                                                   \
        }                                                                               \
    } while (0)
#define CU(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                 \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 512;
    const int m = argc > 2 ? atoi(argv[2]) : 30;
    const int cycles = argc > 3 ? atoi(argv[3]) : 5;
    const long long N = (long long)n * n;
    const double tol = 1e-10;

    /* ---- host CSR of the 5-point Laplacian (diag 4, off-diagonals -1), b = 1 ---- */
    int* rowptr = (int*)malloc(sizeof(int) * (N + 1));
    int* colidx = (int*)malloc(sizeof(int) * 5 * N);
    double* vals = (double*)malloc(sizeof(double) * 5 * N);
    long long nnz = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const long long r = (long long)i * n + j;
            rowptr[r] = (int)nnz;
            if (i > 0) { colidx[nnz] = (int)(r - n); vals[nnz++] = -1.0; }
            if (j > 0) { colidx[nnz] = (int)(r - 1); vals[nnz++] = -1.0; }
            colidx[nnz] = (int)r; vals[nnz++] = 4.0;
            if (j < n - 1) { colidx[nnz] = (int)(r + 1); vals[nnz++] = -1.0; }
            if (i < n - 1) { colidx[nnz] = (int)(r + n); vals[nnz++] = -1.0; }
        }
    rowptr[N] = (int)nnz;

    /* ---- device buffers (owned by the caller, as the ABI prescribes) ---- */
    const long long ld = (N + 31) / 32 * 32;          /* vector-major basis, padded leading dimension */
    int *d_rowptr, *d_colidx;
    double *d_vals, *d_b, *d_x, *d_r, *d_q, *d_V, *d_small;
    CU(cudaSetDevice(0));
    CU(cudaMalloc((void**)&d_rowptr, sizeof(int) * (N + 1)));
    CU(cudaMalloc((void**)&d_colidx, sizeof(int) * nnz));
    CU(cudaMalloc((void**)&d_vals, sizeof(double) * nnz));
    CU(cudaMalloc((void**)&d_b, sizeof(double) * N));
    CU(cudaMalloc((void**)&d_x, sizeof(double) * N));
    CU(cudaMalloc((void**)&d_r, sizeof(double) * N));
    CU(cudaMalloc((void**)&d_q, sizeof(double) * N));
    CU(cudaMalloc((void**)&d_V, sizeof(double) * ld * (m + 1)));
    /* small device state: hcol[m+2] | rcol[m+2] | cs[2m+2] | y[m+2] | yy[m] | R[m*m] | tmp[2] */
    const int o_h = 0, o_r = m + 2, o_cs = 2 * (m + 2), o_y = o_cs + 2 * m + 2, o_yy = o_y + m + 2,
              o_R = o_yy + m, o_tmp = o_R + m * m, nsmall = o_tmp + 2;
    CU(cudaMalloc((void**)&d_small, sizeof(double) * nsmall));
    CU(cudaMemcpy(d_rowptr, rowptr, sizeof(int) * (N + 1), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_colidx, colidx, sizeof(int) * nnz, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_vals, vals, sizeof(double) * nnz, cudaMemcpyHostToDevice));
    double* ones = (double*)malloc(sizeof(double) * N);
    for (long long i = 0; i < N; ++i) ones[i] = 1.0;
    CU(cudaMemcpy(d_b, ones, sizeof(double) * N, cudaMemcpyHostToDevice));
    CU(cudaMemset(d_x, 0, sizeof(double) * N));

    kry_ctx* ctx = NULL;
    CK(kry_ctx_create(0, NULL, &ctx));
    double* mb = kry_mailbox_host(ctx);
    double* R = (double*)calloc((size_t)m * m, sizeof(double));
    double* hR = (double*)malloc(sizeof(double) * m * m);

    /* ||b|| */
    CK(kry_block_dot(ctx, KRY_F64, N, d_b, ld, 1, d_b, d_small + o_tmp, 1, NULL));
    CK(kry_sync(ctx));
    double bnorm;
    CU(cudaMemcpy(&bnorm, d_small + o_tmp, sizeof(double), cudaMemcpyDeviceToHost));

    int total_its = 0;
    double relres = 1.0;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, 0));
    for (int c = 0; c < cycles && relres > tol; ++c) {
        /* r = b - A x ; beta = ||r|| ; v_0 = r / beta */
        CK(kry_spmv_csr(ctx, KRY_F64, N, N, nnz, d_rowptr, d_colidx, d_vals, d_x, d_r, NULL, NULL));
        CK(kry_axpby(ctx, KRY_F64, N, 1.0, d_b, -1.0, d_r, d_r));
        CK(kry_block_dot(ctx, KRY_F64, N, d_r, ld, 1, d_r, d_small + o_tmp, 1, NULL));
        CK(kry_scale_dev(ctx, KRY_F64, N, d_small + o_tmp, 1, 1.0, d_r, d_V));
        CU(cudaMemsetAsync(d_small, 0, sizeof(double) * o_tmp, 0));         /* h, r, cs, y, yy, R = 0 */
        CU(cudaMemcpyAsync(d_small + o_y, d_small + o_tmp, sizeof(double), cudaMemcpyDeviceToDevice, 0));
        memset(R, 0, sizeof(double) * m * m);
        int k = 0;
        for (; k < m; ++k) {
            CK(kry_spmv_csr(ctx, KRY_F64, N, N, nnz, d_rowptr, d_colidx, d_vals, d_V + (long long)k * ld, d_q,
                            NULL, NULL));
            /* ONE fused kernel: h[0..k] += V^T q ; q -= V h ; h[k+1] = ||q|| ; v_{k+1} = q / ||q|| */
            CK(kry_orth_fused(ctx, KRY_F64, N, d_V, d_V, ld, 0, k + 1, d_q, 1, KRY_ORTH_CGS, NULL, NULL,
                              d_small + o_h, d_small + o_h + k + 1, d_V + (long long)(k + 1) * ld));
            CK(kry_givens_update(ctx, k, d_small + o_h, d_small + o_r, d_small + o_cs, d_small + o_y, 0));
            CK(kry_sync(ctx));                        /* the one host synchronisation of the step */
            relres = mb[0] / bnorm;                   /* |y[k+1]| / ||b||, linsys.py:993 */
            for (int i = 0; i <= k; ++i) R[(size_t)i * m + k] = mb[k + 3 + i];   /* column k of R */
            ++total_its;
            if (relres <= tol) { ++k; break; }
        }
        /* x += V[:, :k] R^-1 y */
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) hR[(size_t)i * k + j] = R[(size_t)i * m + j];
        CU(cudaMemcpy(d_small + o_R, hR, sizeof(double) * k * k, cudaMemcpyHostToDevice));
        CK(kry_tri_solve(ctx, k, d_small + o_R, k, d_small + o_y, d_small + o_yy));
        CK(kry_block_combine(ctx, KRY_F64, N, d_V, ld, k, d_small + o_yy, d_x, d_q));
        CU(cudaMemcpyAsync(d_x, d_q, sizeof(double) * N, cudaMemcpyDeviceToDevice, 0));
        printf("cycle %d: %d iterations, updated relative residual %.3e\n", c, k, relres);
    }
    CU(cudaEventRecord(e1, 0));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));

    /* explicit residual of the result */
    CK(kry_spmv_csr(ctx, KRY_F64, N, N, nnz, d_rowptr, d_colidx, d_vals, d_x, d_r, NULL, NULL));
    CK(kry_axpby(ctx, KRY_F64, N, 1.0, d_b, -1.0, d_r, d_r));
    CK(kry_block_dot(ctx, KRY_F64, N, d_r, ld, 1, d_r, d_small + o_tmp, 1, NULL));
    CK(kry_sync(ctx));
    double rn;
    CU(cudaMemcpy(&rn, d_small + o_tmp, sizeof(double), cudaMemcpyDeviceToHost));
    printf("N=%lld  %d iterations in %.2f ms (%.0f it/s), %lld kernel launches, explicit ||b-Ax||/||b|| = %.3e\n",
           N, total_its, ms, total_its / (ms * 1e-3), kry_launch_count(ctx), rn / bnorm);
    const int ok = fabs(rn / bnorm - relres) <= 1e-6 * (relres + 1e-12) + 1e-12;
    printf("%s\n", ok ? "OK: explicit and updated residual agree" : "MISMATCH");
    kry_ctx_destroy(ctx);
    return ok ? 0 : 2;
}
