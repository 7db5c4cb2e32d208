// Fused Lanczos step for a DIAGONAL inner-product matrix B = diag(b) (BASELINE config C5:
// MINRES with ip_B, krypy/utils.py:1000-1045 with inner(X, Y, ip_B) = X^H (B Y), utils.py:190-193):
//
//   q -= H[k-1,k] v_{k-1}                       (three-term recurrence, utils.py:1003-1009)
//   alpha = <v_k, q>_B = sum_i v_k[i] (b[i] q[i])  ; H[k,k] += alpha ; q -= alpha v_k
//   beta  = sqrt(<q, q>_B)                        (utils.py:1034, norm(Av, ip_B))
//   v_{k+1} = q / beta                            (utils.py:1045)
//
// as ONE cooperative kernel (two grid-wide reductions) instead of the seven launches of the
// generic-inner-product path (axpy, diag_mul, block_dot, axpy, diag_mul, block_dot, scale): 11
// instead of 21 vector passes.  Default for a diagonal ip_B since round 2 (KRY_LANCZOS_DIAGB=0 in the
// host layer selects the generic sequence).  Row-partitioned runs (kry_lanczos_diag_dist) complete
// both reductions over NVLink peer memory inside the kernel.
//
// Rounding mirrors the unfused path: b*q is rounded to the storage type before it enters the dot
// (kry_diag_mul stores B q), q is used as stored.  Reductions are deterministic for a fixed grid.
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

template <typename T> __device__ __forceinline__ double lz_round_as(double v) { return (double)(T)v; }

template <typename T>
struct LanczosArgs {
    long long n;
    const T* vprev;          // v_{k-1} (NULL for k == 0)
    const T* vk;             // v_k
    const T* b;              // diagonal of B
    T* q;                    // in: A v_k, out: the orthogonalised vector
    const double* pre_coef;  // &H[k-1,k] (device)
    double* h3;              // [H[k-1,k], H[k,k] (+=), H[k+1,k]] (device), as kry_minres_recur reads it
    T* vnext;                // v_{k+1}
    double* partials;        // [2][KRY_MAX_SLOTS][KRY_MAX_PARTIAL_BLOCKS] scratch of the context
    PeerArgs peer;           // world == 1: single GPU; otherwise both reductions are completed over NVLink
};

__device__ __forceinline__ double* lz_slot(double* partials, int buf) {
    return partials + (size_t)buf * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS;
}

// fixed-order sum of the per-CTA partials of buffer `buf`; identical in every CTA
__device__ __forceinline__ double lz_reduce(double* partials, int buf, double* sm) {
    const double* p = lz_slot(partials, buf);
    double v = 0.0;
    for (int c = threadIdx.x; c < (int)gridDim.x; c += blockDim.x) v += __ldcg(p + c);
    return kry_block_sum(v, sm);
}

template <typename T, int VEC, bool PEER>
__global__ void __launch_bounds__(KRY_THREADS, 4) lanczos_diag_kernel(LanczosArgs<T> a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sm[32];
    __shared__ double c_s[2];
    __shared__ double stage[PEER ? PEER_MAX_RANKS * PEER_SLOT : 1];
    __shared__ int okflag;
    unsigned long long epoch = PEER ? dld_volatile_u64(a.peer.epoch_dev) : 0ull;
    const long long n = a.n;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tail0 = nvec * VEC + threadIdx.x;      // scalar tail handled by CTA 0
    const bool tail_cta = (blockIdx.x == 0);
    const bool pre = (a.vprev != nullptr);
    const double pre_c = pre ? a.pre_coef[0] : 0.0;
    T* q = a.q;

    // ---- phase A: pre-subtraction, alpha = <v_k, B q> ----
    double acc = 0.0;
    for (long long i = i0; i < nvec; i += stride) {
        double qv[VEC], vv[VEC], bv[VEC];
        VecIO<T, VEC>::loadrw(q, i, qv);
        VecIO<T, VEC>::load(a.vk, i, vv);
        VecIO<T, VEC>::load(a.b, i, bv);
        if (pre) {
            double pv[VEC];
            VecIO<T, VEC>::load(a.vprev, i, pv);
#pragma unroll
            for (int u = 0; u < VEC; ++u) qv[u] = fma(-pre_c, pv[u], qv[u]);
            VecIO<T, VEC>::store(q, i, qv);
#pragma unroll
            for (int u = 0; u < VEC; ++u) qv[u] = lz_round_as<T>(qv[u]);     // value as stored
        }
#pragma unroll
        for (int u = 0; u < VEC; ++u) acc = fma(vv[u], lz_round_as<T>(bv[u] * qv[u]), acc);
    }
    if (tail_cta) {
        for (long long i = tail0; i < n; i += blockDim.x) {
            double qe = (double)q[i];
            if (pre) {
                qe = fma(-pre_c, (double)a.vprev[i], qe);
                q[i] = (T)qe;
                qe = (double)q[i];
            }
            acc = fma((double)a.vk[i], lz_round_as<T>((double)a.b[i] * qe), acc);
        }
    }
    {
        const double s = kry_block_sum(acc, sm);
        if (threadIdx.x == 0) lz_slot(a.partials, 0)[blockIdx.x] = s;
    }
    grid.sync();
    double alpha = lz_reduce(a.partials, 0, sm);
    if (PEER) {   // local sum (identical in every CTA) -> global sum in rank order
        __syncthreads();
        if (threadIdx.x == 0) c_s[0] = alpha;
        __syncthreads();
        peer_exchange(a.peer, ++epoch, c_s, 1, stage, &okflag);
        alpha = c_s[0];
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.h3[1] += alpha;

    // ---- phase B: q -= alpha v_k, beta^2 = <q, B q> ----
    double nrm2 = 0.0;
    for (long long i = i0; i < nvec; i += stride) {
        double qv[VEC], vv[VEC], bv[VEC];
        VecIO<T, VEC>::loadrw(q, i, qv);
        VecIO<T, VEC>::load(a.vk, i, vv);
        VecIO<T, VEC>::load(a.b, i, bv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) qv[u] = fma(-alpha, vv[u], qv[u]);
        VecIO<T, VEC>::store(q, i, qv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            const double r = lz_round_as<T>(qv[u]);
            nrm2 = fma(r, lz_round_as<T>(bv[u] * r), nrm2);
        }
    }
    if (tail_cta) {
        for (long long i = tail0; i < n; i += blockDim.x) {
            double qe = fma(-alpha, (double)a.vk[i], (double)q[i]);
            q[i] = (T)qe;
            qe = (double)q[i];
            nrm2 = fma(qe, lz_round_as<T>((double)a.b[i] * qe), nrm2);
        }
    }
    {
        const double s = kry_block_sum(nrm2, sm);
        if (threadIdx.x == 0) lz_slot(a.partials, 1)[blockIdx.x] = s;
    }
    grid.sync();
    double beta2 = lz_reduce(a.partials, 1, sm);
    if (PEER) {
        __syncthreads();
        if (threadIdx.x == 0) c_s[0] = beta2;
        __syncthreads();
        peer_exchange(a.peer, ++epoch, c_s, 1, stage, &okflag);
        beta2 = c_s[0];
        __syncthreads();
    }
    const double beta = sqrt(fabs(beta2));      // sqrt(|ip|), utils.py:238
    if (blockIdx.x == 0 && threadIdx.x == 0) a.h3[2] = beta;

    // ---- phase C: v_{k+1} = q / beta ----
    if (a.vnext != nullptr) {
        for (long long i = i0; i < nvec; i += stride) {
            double qv[VEC];
            VecIO<T, VEC>::loadrw(q, i, qv);
#pragma unroll
            for (int u = 0; u < VEC; ++u) qv[u] = beta > 0.0 ? qv[u] / beta : 0.0;
            VecIO<T, VEC>::store(a.vnext, i, qv);
        }
        if (tail_cta)
            for (long long i = tail0; i < n; i += blockDim.x)
                a.vnext[i] = (T)(beta > 0.0 ? (double)q[i] / beta : 0.0);
    }
    if (PEER) {
        // every CTA read epoch_dev before the first grid.sync: order the write-back after those reads
        grid.sync();
        if (blockIdx.x == 0 && threadIdx.x == 0) *a.peer.epoch_dev = epoch;
    }
}

template <typename T>
static int lanczos_launch(kry_ctx* ctx, LanczosArgs<T>& a) {
    const int W = VecWidth<T>::value;
    const bool al = kry_aligned16(a.vk) && kry_aligned16(a.b) && kry_aligned16(a.q) &&
                    (!a.vprev || kry_aligned16(a.vprev)) && (!a.vnext || kry_aligned16(a.vnext));
    static int blocks_per_sm[2] = {0, 0};
    const int which = al ? 0 : 1;
    const bool peer = a.peer.world > 1;
    if (blocks_per_sm[which] == 0) {
        // (the PEER instantiation needs more shared memory: size the grid for it, it fits both)
        int nb = 0;
        if (al) KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lanczos_diag_kernel<T, W, true>, KRY_THREADS, 0));
        else KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lanczos_diag_kernel<T, 1, true>, KRY_THREADS, 0));
        KRY_REQUIRE(nb >= 1, "lanczos kernel does not fit");
        blocks_per_sm[which] = nb;
    }
    long long need = ((al ? a.n / W : a.n) + KRY_THREADS - 1) / KRY_THREADS;
    if (need < 1) need = 1;
    long long cap = (long long)blocks_per_sm[which] * ctx->sm_count;
    if (cap > KRY_MAX_PARTIAL_BLOCKS) cap = KRY_MAX_PARTIAL_BLOCKS;
    const int g = (int)(need < cap ? need : cap);
    void* args[] = {&a};
    void* k = peer ? (al ? (void*)lanczos_diag_kernel<T, W, true> : (void*)lanczos_diag_kernel<T, 1, true>)
                   : (al ? (void*)lanczos_diag_kernel<T, W, false> : (void*)lanczos_diag_kernel<T, 1, false>);
    KRY_CHECK_CUDA(cudaLaunchCooperativeKernel(k, dim3(g), dim3(KRY_THREADS), args, 0, ctx->stream));
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

extern "C" {

static int lanczos_impl(kry_ctx* ctx, int dtype, long long n, const void* vprev, const void* vk, const void* bdiag,
                        void* q, const double* pre_coef_dev, double* h3_dev, void* vnext, PeerArgs peer) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && vk && bdiag && q && h3_dev, "bad arguments");
    KRY_REQUIRE(!vprev || pre_coef_dev, "vprev without pre_coef_dev");
    if (dtype == KRY_F64) {
        LanczosArgs<double> a = {n, (const double*)vprev, (const double*)vk, (const double*)bdiag, (double*)q,
                                 pre_coef_dev, h3_dev, (double*)vnext, ctx->d_partials, peer};
        return lanczos_launch<double>(ctx, a);
    }
    if (dtype == KRY_F32) {
        LanczosArgs<float> a = {n, (const float*)vprev, (const float*)vk, (const float*)bdiag, (float*)q,
                                pre_coef_dev, h3_dev, (float*)vnext, ctx->d_partials, peer};
        return lanczos_launch<float>(ctx, a);
    }
    kry_set_error("kry_lanczos_diag: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_lanczos_diag(kry_ctx* ctx, int dtype, long long n, const void* vprev, const void* vk, const void* bdiag,
                     void* q, const double* pre_coef_dev, double* h3_dev, void* vnext) {
    PeerArgs peer;
    memset(&peer, 0, sizeof(peer));
    peer.world = 1;
    return lanczos_impl(ctx, dtype, n, vprev, vk, bdiag, q, pre_coef_dev, h3_dev, vnext, peer);
}

int kry_lanczos_diag_dist(kry_ctx* ctx, int dtype, long long n, const void* vprev, const void* vk, const void* bdiag,
                          void* q, const double* pre_coef_dev, double* h3_dev, void* vnext, int world, int rank,
                          unsigned long long* epoch_dev, double* const* peer_slots_dev,
                          unsigned long long* const* peer_flags_dev) {
    KRY_REQUIRE(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world, "bad world/rank");
    KRY_REQUIRE(epoch_dev && peer_slots_dev && peer_flags_dev, "NULL peer argument");
    PeerArgs peer;
    peer.world = world;
    peer.rank = rank;
    peer.epoch_dev = epoch_dev;
    peer.slots = peer_slots_dev;
    peer.flags = peer_flags_dev;
    return lanczos_impl(ctx, dtype, n, vprev, vk, bdiag, q, pre_coef_dev, h3_dev, vnext, peer);
}

}  // extern "C"
