// Native complex128 kernels of the Arnoldi hot loop (one translation unit, nothing else is touched):
//
//   kry_orth_fused_z : the fused Gram-Schmidt step of kry_orth.cu on COMPLEX vectors.  One 16-byte pack
//                      is one complex number; a basis vector is read ONCE per sweep and feeds both the
//                      real and the imaginary accumulator (dots) / both components of q (update).  The
//                      real-embedding path (krypy_b200/_cplx.py, twin storage v_j, i v_j) reads every
//                      basis vector twice for the same arithmetic.
//   kry_spmv_csr_z   : CSR SpMV on complex vectors with complex (20 bytes per entry) or real (12 bytes)
//                      matrix values; the embedded real CSR moves 48 bytes per complex entry.  Same
//                      TMA-staged, warp-specialised design as kry_spmv.cu.
//
// Coefficients are interleaved (re, im) doubles, exactly the arrays the twin path produces, so the small
// complex recurrences (kry_givens_update_z, kry_tri_solve_z) consume them unchanged.
#include <stdlib.h>
#include "kry_common.cuh"
#include "kry_sweeps.cuh"
#include "kry_zorth.cuh"
#include "kry_zspmv.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

template <typename TV, int CPR, int STAGES>
static int zlaunch_staged(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                          const TV* vals, const Z* x, Z* y) {
    typedef ZSpmvCfg<TV, CPR, STAGES> Cfg;
    auto kern = zspmv_staged_kernel<TV, CPR, STAGES>;
    static thread_local int occ[16] = {0};
    int& o = occ[ctx->device & 15];
    if (o == 0) {
        KRY_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        int nb = 0;
        KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, ZSPMV_THREADS, Cfg::SMEM_BYTES));
        KRY_REQUIRE(nb >= 1, "staged complex SpMV kernel does not fit on an SM");
        o = nb;
    }
    long long ntiles = (nrows + ZSPMV_R - 1) / ZSPMV_R;
    long long cap = (long long)ctx->sm_count * o;
    int g = (int)(ntiles < cap ? ntiles : cap);
    if (g < 1) g = 1;
    kern<<<g, ZSPMV_THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(nrows, nnz, rowptr, colidx, vals, x, y);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename TV>
static int zspmv_dispatch(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                          const TV* vals, const Z* x, Z* y) {
    const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    const bool al = kry_aligned16(vals) && kry_aligned16(colidx);
    if (al && avg <= 5.5) return zlaunch_staged<TV, 6, 2>(ctx, nrows, nnz, rowptr, colidx, vals, x, y);
    if (al && avg <= 7.5) return zlaunch_staged<TV, 8, 2>(ctx, nrows, nnz, rowptr, colidx, vals, x, y);
    if (al && avg <= 15.0) return zlaunch_staged<TV, 16, 2>(ctx, nrows, nnz, rowptr, colidx, vals, x, y);
    long long need = (nrows * 32 + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * 8;
    int g = (int)(need < cap ? need : cap);
    if (g < 1) g = 1;
    zspmv_warp_kernel<TV><<<g, KRY_THREADS, 0, ctx->stream>>>(nrows, rowptr, colidx, vals, x, y);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int zorth_blocks(kry_ctx* ctx, int* out) {     // co-resident grid size of zorth_kernel, cached per device
    static thread_local int cached[16] = {0};
    int& c = cached[ctx->device & 15];
    if (c == 0) {
        int nb = 0;
        KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, zorth_kernel, KRY_THREADS, 0));
        KRY_REQUIRE(nb >= 1, "complex orth kernel does not fit");
        c = nb * ctx->sm_count;
    }
    *out = c;
    return KRY_OK;
}

extern "C" {

int kry_orth_fused_z(kry_ctx* ctx, long long n, const void* Vdot, const void* Vsub, long long ldv, int j0, int nv,
                     void* q, int passes, int algo, double* h_dev, double* nrm_dev, void* vnext) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && q, "bad arguments");
    KRY_REQUIRE(j0 >= 0 && nv >= j0, "bad basis range");
    KRY_REQUIRE(nv == j0 || (Vdot && Vsub && h_dev), "NULL basis / h");
    KRY_REQUIRE(passes == 1 || passes == 2, "passes must be 1 or 2");
    KRY_REQUIRE(algo == KRY_ORTH_CGS || algo == KRY_ORTH_MGS, "unknown algo");
    KRY_REQUIRE(!vnext || nrm_dev, "vnext requires nrm_dev");
    KRY_REQUIRE(algo != KRY_ORTH_CGS || 2 * (nv - j0) <= KRY_MAX_SLOTS, "CGS: too many vectors in one call");
    if (!Vdot) Vdot = q;   // never dereferenced when nv == j0
    if (!Vsub) Vsub = q;
    KRY_REQUIRE(kry_aligned16(Vdot) && kry_aligned16(Vsub) && kry_aligned16(q) && (!vnext || kry_aligned16(vnext)),
                "complex vectors must be 16-byte aligned");
    int max_blocks = 0, rc;
    if ((rc = zorth_blocks(ctx, &max_blocks))) return rc;
    ZOrthArgs a = {n, (const Z*)Vdot, (const Z*)Vsub, ldv, j0, nv, passes, algo, (Z*)q, h_dev, nrm_dev, (Z*)vnext,
                   ctx->d_partials};
    long long need = (n + KRY_THREADS - 1) / KRY_THREADS;
    if (need < 1) need = 1;
    long long cap = max_blocks < KRY_MAX_PARTIAL_BLOCKS ? max_blocks : KRY_MAX_PARTIAL_BLOCKS;
    const int g = (int)(need < cap ? need : cap);
    void* args[] = {&a};
    KRY_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)zorth_kernel, dim3(g), dim3(KRY_THREADS), args, 0, ctx->stream));
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_spmv_csr_z(kry_ctx* ctx, int vals_complex, long long nrows, long long ncols, long long nnz, const int* rowptr,
                   const int* colidx, const void* vals, const void* x, void* y) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(nrows >= 0 && ncols >= 0 && nnz >= 0, "negative size");
    KRY_REQUIRE(nnz < 2147483647LL, "nnz must fit int32 row pointers");
    KRY_REQUIRE(rowptr && x && y, "NULL argument");
    KRY_REQUIRE(nnz == 0 || (colidx && vals), "NULL matrix arrays");
    KRY_REQUIRE(kry_aligned16(x) && kry_aligned16(y), "complex vectors must be 16-byte aligned");
    if (nrows == 0) return KRY_OK;
    if (vals_complex) {
        KRY_REQUIRE(nnz == 0 || kry_aligned16(vals), "complex matrix values must be 16-byte aligned");
        return zspmv_dispatch<double2>(ctx, nrows, nnz, rowptr, colidx, (const double2*)vals, (const Z*)x, (Z*)y);
    }
    return zspmv_dispatch<double>(ctx, nrows, nnz, rowptr, colidx, (const double*)vals, (const Z*)x, (Z*)y);
}

}  // extern "C"
