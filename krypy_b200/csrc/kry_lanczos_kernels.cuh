// Device code of kry_lanczos_diag / kry_lanczos_diag_dist (csrc/kry_lanczos.cu).  A header of its own so that the
// CPU test tier can compile exactly this kernel for the host over the CUDA execution emulator
// (tests/csrc/cuda_emul, tests/test_lanczos_emul_cpu.py).
#pragma once
#include "kry_common.cuh"

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

