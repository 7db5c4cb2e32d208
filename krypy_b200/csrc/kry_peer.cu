// Multi-GPU exchange over NVLink 5 / NVSwitch peer memory (one process per GPU).
//
// Buffers that peers must see are cudaMalloc'ed here and exported with CUDA IPC; every
// rank opens its peers' handles once at set-up and from then on the kernels below read
// and write remote HBM directly with ordinary ld/st on the mapped peer pointers:
//   * kry_halo_gather   : x_halo[i] = peer_x[owner(i)][offset(i)]  (P2P loads of exactly the
//                         remote vector entries the local CSR rows reference -- the row-
//                         partitioned SpMV's exchange step; replaces a full all-gather)
//   * kry_peer_allreduce: sum of <= 64 doubles across ranks: P2P stores of the partials into
//                         every peer's slot array + release flag, acquire-spin on the local
//                         flags, fixed rank-order sum (bitwise identical on every rank)
//   * kry_peer_barrier  : the same handshake without payload
// No NCCL call is needed on the iteration path.
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#include "kry_peer_kernels.cuh"

extern "C" {

int kry_peer_alloc(kry_ctx* ctx, long long bytes, void** out) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(bytes > 0 && out, "bad arguments");
    KRY_CHECK_CUDA(cudaMalloc(out, (size_t)bytes));
    KRY_CHECK_CUDA(cudaMemsetAsync(*out, 0, (size_t)bytes, ctx->stream));
    KRY_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return KRY_OK;
}

int kry_peer_free(kry_ctx* ctx, void* p) {
    KRY_ENTER(ctx);
    if (p) KRY_CHECK_CUDA(cudaFree(p));
    return KRY_OK;
}

int kry_ipc_export(kry_ctx* ctx, const void* p, unsigned char handle[64]) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(p && handle, "NULL argument");
    cudaIpcMemHandle_t h;
    KRY_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(p)));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle, &h, 64);
    return KRY_OK;
}

int kry_ipc_open(kry_ctx* ctx, const unsigned char handle[64], void** out) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(handle && out, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    KRY_CHECK_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return KRY_OK;
}

int kry_ipc_close(kry_ctx* ctx, void* p) {
    KRY_ENTER(ctx);
    if (p) KRY_CHECK_CUDA(cudaIpcCloseMemHandle(p));
    return KRY_OK;
}

int kry_halo_gather(kry_ctx* ctx, int dtype, long long nhalo, const void* const* peer_bases_dev,
                    long long elem_offset, const int* halo_peer, const int* halo_off, const double* div_dev,
                    void* dst) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(nhalo >= 0, "negative size");
    if (nhalo == 0) return KRY_OK;
    KRY_REQUIRE(peer_bases_dev && halo_peer && halo_off && dst, "NULL argument");
    long long need = (nhalo + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * 8;
    int g = (int)(need < cap ? need : cap);
    if (dtype == KRY_F64)
        halo_gather_kernel<double><<<g, KRY_THREADS, 0, ctx->stream>>>(
            nhalo, (const double* const*)peer_bases_dev, elem_offset, halo_peer, halo_off, div_dev, (double*)dst);
    else if (dtype == KRY_F32)
        halo_gather_kernel<float><<<g, KRY_THREADS, 0, ctx->stream>>>(
            nhalo, (const float* const*)peer_bases_dev, elem_offset, halo_peer, halo_off, div_dev, (float*)dst);
    else {
        kry_set_error("kry_halo_gather: unsupported dtype %d", dtype);
        return KRY_ERR_UNSUPPORTED;
    }
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_peer_allreduce(kry_ctx* ctx, int world, int rank, unsigned long long* epoch_dev, int n, double* inout_dev,
                       double* const* peer_slots_dev, unsigned long long* const* peer_flags_dev, int post,
                       double* acc_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world, "bad world/rank");
    KRY_REQUIRE(n >= 0 && n <= PEER_SLOT, "n must be <= 64");
    KRY_REQUIRE(epoch_dev != nullptr, "epoch_dev is NULL");
    KRY_REQUIRE(peer_slots_dev && peer_flags_dev && (n == 0 || inout_dev), "NULL argument");
    peer_allreduce_kernel<<<1, 128, 0, ctx->stream>>>(world, rank, epoch_dev, n, inout_dev, peer_slots_dev,
                                                      peer_flags_dev, post, acc_dev);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

int kry_peer_barrier(kry_ctx* ctx, int world, int rank, unsigned long long* epoch_dev,
                     double* const* peer_slots_dev, unsigned long long* const* peer_flags_dev) {
    return kry_peer_allreduce(ctx, world, rank, epoch_dev, 0, nullptr, peer_slots_dev, peer_flags_dev, 0, nullptr);
}

}  // extern "C"
