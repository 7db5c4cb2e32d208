// Block (tall-skinny) kernels for the set-up of the deflation projector (krypy/utils.py:680-707 qr,
// :440-520 Projection.__init__, deflation.py:33-56): the reference orthonormalises an (N, d) block
// column by column (LAPACK QR or Python MGS: d^2 dependent reductions).  Here a block is
// orthonormalised by CholQR2 -- two rounds of  G = X^H X (ONE pass over the block),  R = chol(G) on
// the host (d x d),  X <- X R^-1 (one read + one write of the block) -- i.e. six block passes and
// two host synchronisations instead of d(d+1) dependent sweeps.
//
//   kry_gram        C = X^H Y for kx x ky <= 512 outputs in one pass over X and Y: 128-row chunks of
//                   all vectors are staged in shared memory (cp.async, double buffered), every warp owns up to two 4x4 output
//                   blocks in registers (two rows per lane and step: 8 LDS.128 feed 32 FMAs),
//                   deterministic reduction (shuffle tree, per-CTA partials, last CTA sums in order)
//   kry_block_trsm  Q = X R^-1, R upper triangular d x d (d <= 32): R^-1 is formed once per CTA in shared
//                   memory, then one thread per row: the row of X in registers times R^-1
// Both are bandwidth-sized (3.2 GFLOP on 0.64 GB at N = 4M, d = 20); no tensor cores: fp64.
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#include "kry_block_kernels.cuh"

template <typename T>
static int gram_launch(kry_ctx* ctx, long long n, const T* X, long long ldx, int kx, const T* Y, long long ldy,
                       int ky, int same, double* out) {
    const int nvec = same ? kx : kx + ky;
    const size_t smem = 2 * sizeof(T) * (size_t)nvec * GRAM_LD;          // two chunk buffers
    auto kern = gram_kernel<T>;
    static thread_local size_t smem_set[16] = {0};
    size_t& cur = smem_set[ctx->device & 15];
    if (smem > cur) {
        KRY_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    int occ = 0;
    KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GRAM_THREADS, smem));
    KRY_REQUIRE(occ >= 1, "gram kernel does not fit on an SM");
    long long nchunks = (n + GRAM_E - 1) / GRAM_E;
    long long cap = (long long)ctx->sm_count * occ;
    const long long pcap = (2LL * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS) / ((long long)kx * ky);   // scratch doubles
    if (cap > pcap) cap = pcap;
    int g = (int)(nchunks < cap ? nchunks : cap);
    if (g < 1) g = 1;
    kern<<<g, GRAM_THREADS, smem, ctx->stream>>>(n, X, ldx, kx, Y, ldy, ky, same, ctx->d_partials, ctx->d_ticket + 12, out);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename T>
static int trsm_launch(kry_ctx* ctx, long long n, const T* X, long long ldx, int d, const double* R, T* Q, long long ldq) {
    long long need = (n + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * 4;
    int g = (int)(need < cap ? need : cap);
    if (g < 1) g = 1;
    if (d <= 8) block_trsm_kernel<T, 8><<<g, KRY_THREADS, 0, ctx->stream>>>(n, X, ldx, d, R, Q, ldq);
    else if (d <= 16) block_trsm_kernel<T, 16><<<g, KRY_THREADS, 0, ctx->stream>>>(n, X, ldx, d, R, Q, ldq);
    else if (d <= 24) block_trsm_kernel<T, 24><<<g, KRY_THREADS, 0, ctx->stream>>>(n, X, ldx, d, R, Q, ldq);
    else block_trsm_kernel<T, 32><<<g, KRY_THREADS, 0, ctx->stream>>>(n, X, ldx, d, R, Q, ldq);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

extern "C" {

int kry_gram(kry_ctx* ctx, int dtype, long long n, const void* X, long long ldx, int kx, const void* Y, long long ldy,
             int ky, double* out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && kx >= 1 && ky >= 1 && X && Y && out_dev, "bad arguments");
    const int same = (X == Y && ldx == ldy && kx == ky) ? 1 : 0;
    KRY_REQUIRE((same ? kx : kx + ky) <= GRAM_MAXV, "too many vectors for one call (kx + ky <= 64)");
    KRY_REQUIRE(((kx + 3) / 4) * ((ky + 3) / 4) <= GRAM_BPW * (GRAM_THREADS / 32), "kx x ky too large for one call");
    if (dtype == KRY_F64)
        return gram_launch<double>(ctx, n, (const double*)X, ldx, kx, (const double*)Y, ldy, ky, same, out_dev);
    if (dtype == KRY_F32)
        return gram_launch<float>(ctx, n, (const float*)X, ldx, kx, (const float*)Y, ldy, ky, same, out_dev);
    kry_set_error("kry_gram: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_block_trsm(kry_ctx* ctx, int dtype, long long n, const void* X, long long ldx, int d, const double* R_dev,
                   void* Q, long long ldq) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && d >= 1 && d <= 32 && X && R_dev && Q, "bad arguments (1 <= d <= 32)");
    if (dtype == KRY_F64) return trsm_launch<double>(ctx, n, (const double*)X, ldx, d, R_dev, (double*)Q, ldq);
    if (dtype == KRY_F32) return trsm_launch<float>(ctx, n, (const float*)X, ldx, d, R_dev, (float*)Q, ldq);
    kry_set_error("kry_block_trsm: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

}  // extern "C"
