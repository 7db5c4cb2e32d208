// Row-partitioned Gram-Schmidt step with the NVLink exchange FUSED into the compute kernels
// (one process per GPU, peers mapped with CUDA IPC, see kry_peer.cu):
//
//   kry_dist_dot    : c_local = Vdot^H q      ; the last CTA stores the partial sums straight into
//                     every peer's slot array (P2P stores) and releases its flag
//   kry_dist_update : every CTA acquires the peers' flags, sums the partials in rank order
//                     (bitwise identical everywhere), q -= Vsub c, ||q||^2 partials; the last CTA
//                     publishes the local ||q||^2 the same way
//   kry_dist_scale  : acquires, nrm = sqrt(sum), v_next = q / nrm
//   kry_dist_halo   : flag handshake (all ranks' basis rows are complete) + P2P gather of the
//                     remote entries of v_k that the local CSR rows reference
//
// Compared with separate all-reduce kernels this removes four launches and four kernel
// boundaries per Arnoldi step; no NCCL call and no host synchronisation is involved.
// Protocol: one monotone operation counter per rank in device memory (epoch_dev, identical
// sequence on all ranks); slot arrays are double buffered by epoch parity; a rank publishes
// epoch e only after it has observed every peer's epoch e-1.
#include "kry_common.cuh"
#include "kry_sweeps.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#include "kry_dist_kernels.cuh"

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static inline int dgrid(const kry_ctx* ctx, long long nvec, int per_sm) {
    long long need = (nvec + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

static int make_peer(PeerArgs& pa, int world, int rank, unsigned long long* epoch_dev, double* const* slots,
                     unsigned long long* const* flags) {
    KRY_REQUIRE(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world, "bad world/rank");
    KRY_REQUIRE(epoch_dev && slots && flags, "NULL peer argument");
    pa.world = world;
    pa.rank = rank;
    pa.epoch_dev = epoch_dev;
    pa.slots = slots;
    pa.flags = flags;
    return KRY_OK;
}

template <typename T>
static int dist_dot_launch(kry_ctx* ctx, long long n, const T* V, long long ldv, int nv, const T* q, int want_sq,
                           PeerArgs pa) {
    const int W = VecWidth<T>::value;
    bool al = kry_aligned16(V) && kry_aligned16(q) && (ldv % W == 0);
    if (al)
        dist_dot_kernel<T, W><<<dgrid(ctx, n / W, 2), KRY_THREADS, 0, ctx->stream>>>(
            n, V, ldv, nv, q, want_sq, ctx->d_partials, ctx->d_ticket + 4, pa);
    else
        dist_dot_kernel<T, 1><<<dgrid(ctx, n, 2), KRY_THREADS, 0, ctx->stream>>>(
            n, V, ldv, nv, q, want_sq, ctx->d_partials, ctx->d_ticket + 4, pa);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename T>
static int dist_update_launch(kry_ctx* ctx, long long n, const T* V, long long ldv, int nv, T* q, double* h_acc,
                              int want_nrm, PeerArgs pa) {
    const int W = VecWidth<T>::value;
    bool al = kry_aligned16(V) && kry_aligned16(q) && (ldv % W == 0);
    // partials of the norm live behind the dot partials' slot 0..63 region: use slot 64's row
    double* part = ctx->d_partials + (size_t)KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS;
    if (al)
        dist_update_kernel<T, W><<<dgrid(ctx, n / W, 2), KRY_THREADS, 0, ctx->stream>>>(
            n, V, ldv, nv, q, h_acc, want_nrm, part, ctx->d_ticket + 5, pa);
    else
        dist_update_kernel<T, 1><<<dgrid(ctx, n, 2), KRY_THREADS, 0, ctx->stream>>>(
            n, V, ldv, nv, q, h_acc, want_nrm, part, ctx->d_ticket + 5, pa);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename T>
static int dist_scale_launch(kry_ctx* ctx, long long n, const T* q, T* vnext, double* nrm_out, PeerArgs pa) {
    const int W = VecWidth<T>::value;
    bool al = kry_aligned16(q) && (!vnext || kry_aligned16(vnext));
    if (al)
        dist_scale_kernel<T, W><<<dgrid(ctx, n / W, 4), KRY_THREADS, 0, ctx->stream>>>(n, q, vnext, nrm_out, pa);
    else
        dist_scale_kernel<T, 1><<<dgrid(ctx, n, 4), KRY_THREADS, 0, ctx->stream>>>(n, q, vnext, nrm_out, pa);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename T, int VEC>
static int update_scale_launch(kry_ctx* ctx, UpdScaleArgs<T>& a) {
    auto kern = dist_update_scale_kernel<T, VEC>;
    static thread_local int occ[16] = {0};
    int& o = occ[ctx->device & 15];
    if (o == 0) {
        int nb = 0;
        KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, KRY_THREADS, 0));
        KRY_REQUIRE(nb >= 1, "kernel does not fit on an SM");
        o = nb > 2 ? 2 : nb;       // two CTAs per SM carry the sweep (as in the cooperative kernels)
    }
    const int extra = a.k_givens >= 0 ? 1 : 0;
    // all CTAs co-resident (the guard path waits for the other CTAs of the grid through the peers)
    long long cap = (long long)ctx->sm_count * o - extra;
    if (cap > KRY_MAX_PARTIAL_BLOCKS) cap = KRY_MAX_PARTIAL_BLOCKS;
    long long need = (a.n / VEC + KRY_THREADS - 1) / KRY_THREADS;
    if (need < 1) need = 1;
    const int g = (int)(need < cap ? need : cap);
    kern<<<g + extra, KRY_THREADS, 0, ctx->stream>>>(a);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename T>
static int update_scale_dtype(kry_ctx* ctx, long long n, const void* V, long long ldv, int nv, const void* q,
                              void* vnext, double* h_acc_dev, double* nrm_out_dev, long long nhalo,
                              const void* const* peer_q_dev, long long q_elem_offset, const int* halo_peer,
                              const int* halo_off, long long halo_base, void* halo_dst, int k_givens, double* rcol_dev,
                              double* cs_dev, double* y_dev, long long mailbox_off, const PeerArgs& pa) {
    UpdScaleArgs<T> a;
    a.n = n;
    a.V = (const T*)V;
    a.ldv = ldv;
    a.nv = nv;
    a.q = (const T*)q;
    a.vnext = (T*)vnext;
    a.h_acc = h_acc_dev;
    a.nrm_out = nrm_out_dev;
    a.nhalo = nhalo;
    a.peer_q = (const T* const*)peer_q_dev;
    a.q_elem_offset = q_elem_offset;
    a.halo_peer = halo_peer;
    a.halo_off = halo_off;
    a.halo_base = halo_base;
    a.halo_dst = (T*)halo_dst;
    a.k_givens = k_givens;
    a.rcol = rcol_dev;
    a.cs = cs_dev;
    a.y = y_dev;
    a.mailbox = ctx->d_mailbox + mailbox_off;
    a.partials = ctx->d_partials + (size_t)KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS;
    a.ticket = ctx->d_ticket + 14;
    a.pa = pa;
    const int W = VecWidth<T>::value;
    const bool al = kry_aligned16(V) && kry_aligned16(q) && kry_aligned16(vnext) && (ldv % W == 0);
    if (al) return update_scale_launch<T, VecWidth<T>::value>(ctx, a);
    return update_scale_launch<T, 1>(ctx, a);
}

extern "C" {

int kry_dist_dot(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv, const void* q,
                 int want_sq, int world, int rank, unsigned long long* epoch_dev, double* const* peer_slots_dev,
                 unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && nv >= 1 && nv + (want_sq ? 1 : 0) <= PEER_SLOT && V && q,
                "bad arguments (1 <= nv, nv + want_sq <= 64)");
    PeerArgs pa;
    int rc = make_peer(pa, world, rank, epoch_dev, peer_slots_dev, peer_flags_dev);
    if (rc) return rc;
    if (dtype == KRY_F64)
        return dist_dot_launch<double>(ctx, n, (const double*)V, ldv, nv, (const double*)q, want_sq ? 1 : 0, pa);
    if (dtype == KRY_F32)
        return dist_dot_launch<float>(ctx, n, (const float*)V, ldv, nv, (const float*)q, want_sq ? 1 : 0, pa);
    kry_set_error("kry_dist_dot: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_dist_update(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv, void* q,
                    double* h_acc_dev, int want_nrm, int world, int rank, unsigned long long* epoch_dev,
                    double* const* peer_slots_dev, unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && nv >= 1 && nv <= PEER_SLOT && V && q, "bad arguments (1 <= nv <= 64)");
    PeerArgs pa;
    int rc = make_peer(pa, world, rank, epoch_dev, peer_slots_dev, peer_flags_dev);
    if (rc) return rc;
    if (dtype == KRY_F64)
        return dist_update_launch<double>(ctx, n, (const double*)V, ldv, nv, (double*)q, h_acc_dev, want_nrm, pa);
    if (dtype == KRY_F32)
        return dist_update_launch<float>(ctx, n, (const float*)V, ldv, nv, (float*)q, h_acc_dev, want_nrm, pa);
    kry_set_error("kry_dist_update: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_dist_scale(kry_ctx* ctx, int dtype, long long n, const void* q, void* vnext, double* nrm_out_dev, int world,
                   int rank, unsigned long long* epoch_dev, double* const* peer_slots_dev,
                   unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && q && nrm_out_dev, "bad arguments");
    PeerArgs pa;
    int rc = make_peer(pa, world, rank, epoch_dev, peer_slots_dev, peer_flags_dev);
    if (rc) return rc;
    if (dtype == KRY_F64) return dist_scale_launch<double>(ctx, n, (const double*)q, (double*)vnext, nrm_out_dev, pa);
    if (dtype == KRY_F32) return dist_scale_launch<float>(ctx, n, (const float*)q, (float*)vnext, nrm_out_dev, pa);
    kry_set_error("kry_dist_scale: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_dist_scale_halo(kry_ctx* ctx, int dtype, long long n, const void* q, void* vnext, double* nrm_out_dev,
                        long long nhalo, const void* const* peer_bases_dev, long long elem_offset,
                        const int* halo_peer, const int* halo_off, void* halo_dst, int world, int rank,
                        unsigned long long* epoch_dev, double* const* peer_slots_dev,
                        unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && q && vnext && nrm_out_dev && nhalo >= 0, "bad arguments");
    KRY_REQUIRE(nhalo == 0 || (peer_bases_dev && halo_peer && halo_off && halo_dst), "NULL halo argument");
    PeerArgs pa;
    int rc = make_peer(pa, world, rank, epoch_dev, peer_slots_dev, peer_flags_dev);
    if (rc) return rc;
    unsigned int* tk = ctx->d_ticket + 8;
    if (dtype == KRY_F64) {
        const double* qq = (const double*)q;
        double* vn = (double*)vnext;
        if (kry_aligned16(qq) && kry_aligned16(vn))
            dist_scale_halo_kernel<double, 2><<<dgrid(ctx, n / 2, 4), KRY_THREADS, 0, ctx->stream>>>(
                n, qq, vn, nrm_out_dev, nhalo, (const double* const*)peer_bases_dev, elem_offset, halo_peer, halo_off,
                (double*)halo_dst, tk, pa);
        else
            dist_scale_halo_kernel<double, 1><<<dgrid(ctx, n, 4), KRY_THREADS, 0, ctx->stream>>>(
                n, qq, vn, nrm_out_dev, nhalo, (const double* const*)peer_bases_dev, elem_offset, halo_peer, halo_off,
                (double*)halo_dst, tk, pa);
    } else if (dtype == KRY_F32) {
        const float* qq = (const float*)q;
        float* vn = (float*)vnext;
        if (kry_aligned16(qq) && kry_aligned16(vn))
            dist_scale_halo_kernel<float, 4><<<dgrid(ctx, n / 4, 4), KRY_THREADS, 0, ctx->stream>>>(
                n, qq, vn, nrm_out_dev, nhalo, (const float* const*)peer_bases_dev, elem_offset, halo_peer, halo_off,
                (float*)halo_dst, tk, pa);
        else
            dist_scale_halo_kernel<float, 1><<<dgrid(ctx, n, 4), KRY_THREADS, 0, ctx->stream>>>(
                n, qq, vn, nrm_out_dev, nhalo, (const float* const*)peer_bases_dev, elem_offset, halo_peer, halo_off,
                (float*)halo_dst, tk, pa);
    } else {
        kry_set_error("kry_dist_scale_halo: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_dist_scale_haloq(kry_ctx* ctx, int dtype, long long n, const void* q, void* vnext, double* nrm_out_dev,
                         long long nhalo, const void* const* peer_bases_dev, long long q_elem_offset,
                         const int* halo_peer, const int* halo_off, void* halo_dst, int world, int rank,
                         unsigned long long* epoch_dev, double* const* peer_slots_dev,
                         unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && q && vnext && nrm_out_dev && nhalo >= 0, "bad arguments");
    KRY_REQUIRE(nhalo == 0 || (peer_bases_dev && halo_peer && halo_off && halo_dst), "NULL halo argument");
    PeerArgs pa;
    int rc = make_peer(pa, world, rank, epoch_dev, peer_slots_dev, peer_flags_dev);
    if (rc) return rc;
    if (dtype == KRY_F64) {
        const double* qq = (const double*)q;
        double* vn = (double*)vnext;
        if (kry_aligned16(qq) && kry_aligned16(vn))
            dist_scale_haloq_kernel<double, 2><<<dgrid(ctx, n / 2, 4), KRY_THREADS, 0, ctx->stream>>>(
                n, qq, vn, nrm_out_dev, nhalo, (const double* const*)peer_bases_dev, q_elem_offset, halo_peer,
                halo_off, (double*)halo_dst, pa);
        else
            dist_scale_haloq_kernel<double, 1><<<dgrid(ctx, n, 4), KRY_THREADS, 0, ctx->stream>>>(
                n, qq, vn, nrm_out_dev, nhalo, (const double* const*)peer_bases_dev, q_elem_offset, halo_peer,
                halo_off, (double*)halo_dst, pa);
    } else if (dtype == KRY_F32) {
        const float* qq = (const float*)q;
        float* vn = (float*)vnext;
        if (kry_aligned16(qq) && kry_aligned16(vn))
            dist_scale_haloq_kernel<float, 4><<<dgrid(ctx, n / 4, 4), KRY_THREADS, 0, ctx->stream>>>(
                n, qq, vn, nrm_out_dev, nhalo, (const float* const*)peer_bases_dev, q_elem_offset, halo_peer,
                halo_off, (float*)halo_dst, pa);
        else
            dist_scale_haloq_kernel<float, 1><<<dgrid(ctx, n, 4), KRY_THREADS, 0, ctx->stream>>>(
                n, qq, vn, nrm_out_dev, nhalo, (const float* const*)peer_bases_dev, q_elem_offset, halo_peer,
                halo_off, (float*)halo_dst, pa);
    } else {
        kry_set_error("kry_dist_scale_haloq: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_dist_update_scale(kry_ctx* ctx, int dtype, long long n, const void* V, long long ldv, int nv, const void* q,
                          void* vnext, double* h_acc_dev, double* nrm_out_dev, long long nhalo,
                          const void* const* peer_q_dev, long long q_elem_offset, const int* halo_peer,
                          const int* halo_off, long long halo_base, void* halo_dst, int k_givens, double* rcol_dev,
                          double* cs_dev, double* y_dev, long long mailbox_off, int world, int rank,
                          unsigned long long* epoch_dev, double* const* peer_slots_dev,
                          unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(n >= 0 && nv >= 1 && nv < PEER_SLOT && V && q && vnext && nrm_out_dev, "bad arguments (1 <= nv <= 63)");
    KRY_REQUIRE(nhalo >= 0, "negative halo size");
    KRY_REQUIRE(nhalo == 0 || (peer_q_dev && halo_peer && halo_off && halo_dst), "NULL halo argument");
    KRY_REQUIRE(k_givens < 0 || (k_givens + 1 == nv && h_acc_dev && nrm_out_dev == h_acc_dev + nv && rcol_dev &&
                                 cs_dev && y_dev && mailbox_off >= 0 && mailbox_off + 2 * (long long)k_givens + 5 <= KRY_MAILBOX_DOUBLES),
                "Givens tail: k + 1 == nv, nrm_out == h_acc + nv, state arrays and a mailbox window required");
    PeerArgs pa;
    int rc = make_peer(pa, world, rank, epoch_dev, peer_slots_dev, peer_flags_dev);
    if (rc) return rc;
    if (dtype == KRY_F64)
        return update_scale_dtype<double>(ctx, n, V, ldv, nv, q, vnext, h_acc_dev, nrm_out_dev, nhalo, peer_q_dev,
                                          q_elem_offset, halo_peer, halo_off, halo_base, halo_dst, k_givens, rcol_dev,
                                          cs_dev, y_dev, mailbox_off, pa);
    if (dtype == KRY_F32)
        return update_scale_dtype<float>(ctx, n, V, ldv, nv, q, vnext, h_acc_dev, nrm_out_dev, nhalo, peer_q_dev,
                                         q_elem_offset, halo_peer, halo_off, halo_base, halo_dst, k_givens, rcol_dev,
                                         cs_dev, y_dev, mailbox_off, pa);
    kry_set_error("kry_dist_update_scale: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

int kry_dist_halo(kry_ctx* ctx, int dtype, long long nhalo, const void* const* peer_bases_dev, long long elem_offset,
                  const int* halo_peer, const int* halo_off, void* dst, int world, int rank,
                  unsigned long long* epoch_dev, double* const* peer_slots_dev,
                  unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(nhalo >= 0, "negative size");
    KRY_REQUIRE(nhalo == 0 || (peer_bases_dev && halo_peer && halo_off && dst), "NULL argument");
    PeerArgs pa;
    int rc = make_peer(pa, world, rank, epoch_dev, peer_slots_dev, peer_flags_dev);
    if (rc) return rc;
    long long need = (nhalo + KRY_THREADS - 1) / KRY_THREADS;
    if (need < 1) need = 1;
    long long cap = (long long)ctx->sm_count * 4;
    int g = (int)(need < cap ? need : cap);
    if (dtype == KRY_F64)
        dist_halo_kernel<double><<<g, KRY_THREADS, 0, ctx->stream>>>(
            nhalo, (const double* const*)peer_bases_dev, elem_offset, halo_peer, halo_off, (double*)dst,
            ctx->d_ticket + 6, pa);
    else if (dtype == KRY_F32)
        dist_halo_kernel<float><<<g, KRY_THREADS, 0, ctx->stream>>>(
            nhalo, (const float* const*)peer_bases_dev, elem_offset, halo_peer, halo_off, (float*)dst,
            ctx->d_ticket + 6, pa);
    else {
        kry_set_error("kry_dist_halo: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

}  // extern "C"
