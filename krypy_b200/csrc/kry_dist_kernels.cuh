// Device code of the row-partitioned Gram-Schmidt / halo kernels (csrc/kry_dist.cu).  A header of its own so that
// the CPU test tier can compile exactly these kernels for the host and run a row-partitioned Arnoldi process over
// emulated ranks (tests/csrc/cuda_emul, tests/csrc/dist_emul_host.cpp, tests/test_dist_emul_cpu.py).
#pragma once
#include "kry_common.cuh"
#include "kry_sweeps.cuh"

// ---------------------------------------------------------------------------
// K1: local block dot + publish.  Up to 16 vectors per pass over q, sweeps specialised on the exact
// vector count (kry_sweeps.cuh), one barrier pair per CTA reduction.
// ---------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 2)
dist_dot_kernel(long long n, const T* __restrict__ V, long long ldv, int nv, const T* q, int want_sq, double* partials,
                unsigned int* ticket, PeerArgs pa) {
    __shared__ double red[(ORTH_JT + 1) * 8];
    __shared__ double fin[PEER_SLOT];
    __shared__ bool last;
    const unsigned long long E = dld_volatile_u64(pa.epoch_dev);
    for (int jb = 0; jb < nv; jb += ORTH_JT) {
        const int nt = nv - jb < ORTH_JT ? nv - jb : ORTH_JT;
        // <q, q> rides along with the last tile as one more sum (slot nv): no pass of its own
        if (want_sq && jb + nt == nv)
            dots_dispatch<T, VEC, true>(nt, V + (long long)jb * ldv, ldv, q, n, red, partials, 0, jb);
        else
            dots_dispatch<T, VEC, false>(nt, V + (long long)jb * ldv, ldv, q, n, red, partials, 0, jb);
    }
    const int nred = nv + (want_sq ? 1 : 0);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        {   // final local sums: one warp per basis vector, lanes stride over the CTAs (fixed order)
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
            for (int j = w; j < nred; j += nw) {
                double v = 0.0;
                for (int b = lane; b < (int)gridDim.x; b += 32)
                    v += __ldcg(partials + (long long)j * KRY_MAX_PARTIAL_BLOCKS + b);
                v = kry_warp_sum(v);
                if (lane == 0) fin[j] = v;
            }
        }
        __syncthreads();
        peer_publish(pa, E + 1ull, fin, nred);
        if (threadIdx.x == 0) {
            *pa.epoch_dev = E + 1ull;
            *ticket = 0u;
        }
    }
}

// ---------------------------------------------------------------------------
// K2: acquire + global sum, q -= Vsub c, ||q||^2 (+ publish)
// ---------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 2)
dist_update_kernel(long long n, const T* __restrict__ V, long long ldv, int nv, T* q, double* h_acc, int want_nrm,
                   double* partials, unsigned int* ticket, PeerArgs pa) {
    __shared__ double sm[32];
    __shared__ double c_s[PEER_SLOT];
    __shared__ int okflag;
    __shared__ bool last;
    const unsigned long long E = dld_volatile_u64(pa.epoch_dev);     // the epoch kry_dist_dot published
    __shared__ double stage[PEER_MAX_RANKS * PEER_SLOT];
    const bool ok = peer_wait(pa, E, &okflag);
    {   // all world*nv partials are fetched in parallel, then summed in rank order
        const double* mine = pa.slots[pa.rank] + (size_t)(E & 1ull) * (size_t)pa.world * PEER_SLOT;
        for (int idx = threadIdx.x; idx < pa.world * nv; idx += blockDim.x) {
            const int r = idx / nv, j = idx - r * nv;
            stage[r * PEER_SLOT + j] = dld_volatile_f64(mine + (size_t)r * PEER_SLOT + j);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < nv; j += blockDim.x) {
            double sum = 0.0;
            for (int r = 0; r < pa.world; ++r) sum += stage[r * PEER_SLOT + j];
            c_s[j] = ok ? sum : nan_f64();
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && h_acc)
        for (int j = threadIdx.x; j < nv; j += blockDim.x) h_acc[j] += c_s[j];
    const double nrm2 = update_dispatch<T, VEC, false>(V, ldv, nv, c_s, q, n, want_nrm != 0);
    if (!want_nrm) return;
    double s = kry_block_sum(nrm2, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double v = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) v += __ldcg(partials + b);
        double tot = kry_block_sum(v, sm);
        __syncthreads();
        if (threadIdx.x == 0) c_s[0] = tot;
        __syncthreads();
        peer_publish(pa, E + 1ull, c_s, 1);
        if (threadIdx.x == 0) {
            *pa.epoch_dev = E + 1ull;
            *ticket = 0u;
        }
    }
}

// ---------------------------------------------------------------------------
// K3q: acquire ||q||^2, v_next = q / nrm, and the halo of v_next gathered from the peers'
// UN-NORMALISED q (divided by the same nrm here: bitwise the value the owner stores).  A peer
// publishes its ||q||^2 partial only after its q segment is complete, so the norm's flag is also
// the "segment complete" handshake: one cross-GPU wait instead of two.  The peers rewrite their q
// with the next SpMV, therefore the host alternates between two q buffers (step parity): a rank
// cannot reach the SpMV after next before every neighbour has passed this kernel.
// ---------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 4)
dist_scale_haloq_kernel(long long n, const T* q, T* vnext, double* nrm_out, long long nhalo,
                        const T* const* peer_bases, long long q_elem_offset, const int* __restrict__ halo_peer,
                        const int* __restrict__ halo_off, T* halo_dst, PeerArgs pa) {
    __shared__ int okflag;
    __shared__ double nrm_s;
    const unsigned long long E = dld_volatile_u64(pa.epoch_dev);
    const bool ok = peer_wait(pa, E, &okflag);
    if (threadIdx.x == 0) nrm_s = ok ? sqrt(fabs(peer_sum(pa, E, 0))) : nan_f64();
    __syncthreads();
    const double nrm = nrm_s;
    if (blockIdx.x == 0 && threadIdx.x == 0) nrm_out[0] = nrm;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // halo first: the remote loads' latency overlaps with the local sweep
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nhalo; i += stride) {
        const T* src = peer_bases[__ldg(halo_peer + i)] + q_elem_offset;
        const T v = *(const volatile T*)(src + __ldg(halo_off + i));
        halo_dst[i] = ok ? (T)(nrm > 0.0 ? (double)v / nrm : 0.0) : (T)nan_f64();
    }
    scale_pass<T, VEC>(q, vnext, n, nrm);
}

// ---------------------------------------------------------------------------
// K3: acquire ||q||^2, v_next = q / nrm
// ---------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 2)
dist_scale_kernel(long long n, const T* q, T* vnext, double* nrm_out, PeerArgs pa) {
    __shared__ int okflag;
    __shared__ double nrm_s;
    const unsigned long long E = dld_volatile_u64(pa.epoch_dev);
    const bool ok = peer_wait(pa, E, &okflag);
    if (threadIdx.x == 0) nrm_s = ok ? sqrt(fabs(peer_sum(pa, E, 0))) : nan_f64();
    __syncthreads();
    const double nrm = nrm_s;
    if (blockIdx.x == 0 && threadIdx.x == 0) nrm_out[0] = nrm;
    if (vnext == nullptr) return;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double qv[VEC];
        VecIO<T, VEC>::loadrw(q, i, qv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) qv[u] = nrm > 0.0 ? qv[u] / nrm : 0.0;
        VecIO<T, VEC>::store(vnext, i, qv);
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x)
            vnext[i] = (T)(nrm > 0.0 ? (double)q[i] / nrm : 0.0);
}

// ---------------------------------------------------------------------------
// K2f: the rest of a block-CGS Arnoldi step after kry_spmv_csr_mdot, with ONE cross-GPU wait:
//   acquire the peers' partials of [V^H w, <w, w>]  (w = A v_k, untouched in q),
//   c = rank-order sums,  ||w - V c||^2 = <w, w> - sum c_j^2  (V orthonormal),
//   v_next = (w - V c) / nrm in one sweep (w is read once, nothing is written back to q),
//   halo of v_next = (w_halo - V_halo c) / nrm from the peers' w (complete since they published) and the
//   halo entries of v_0..v_k this rank already holds behind its basis rows -- the same fma sequence as
//   the owner's sweep, so the copy is bitwise the owner's value -- no second handshake,
//   and (k_givens >= 0) the GMRES Givens / Hessenberg update in one EXTRA CTA (the last one), which runs
//   beside the sweep instead of as a kernel of its own on the critical path.
// Guard: the difference above cancels when w lies almost in span(V).  If it keeps less than 1e-3 of
// <w, w> (or is not finite) every CTA of every rank -- the decision is taken on bitwise identical numbers
// -- computes the local ||w - V c||^2 exactly, the last one publishes it (epoch + 1), all acquire, and the
// sweep runs with the exact norm.  The sweep CTAs wait for each other only in that case, therefore the grid
// is sized to be co-resident.
// ---------------------------------------------------------------------------
#include "kry_givens_dev.cuh"

template <typename T>
struct UpdScaleArgs {
    long long n;
    const T* V;
    long long ldv;
    int nv;
    const T* q;
    T* vnext;
    double* h_acc;        // h[0..nv) += c, h[nv] = nrm   (nrm_out == h_acc + nv for the solvers)
    double* nrm_out;
    long long nhalo;
    const T* const* peer_q;        // peer pointer table of the region q lives in
    long long q_elem_offset;
    const int* halo_peer;
    const int* halo_off;
    long long halo_base;           // element offset of the halo part inside a basis row (== block)
    T* halo_dst;
    int k_givens;                  // >= 0: Givens update of column k in the extra CTA
    double *rcol, *cs, *y, *mailbox;
    double* partials;
    unsigned int* ticket;
    PeerArgs pa;
};

#define DUS_GUARD 1e-3

template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 2) dist_update_scale_kernel(UpdScaleArgs<T> a) {
    __shared__ double sm[32];
    __shared__ double c_s[PEER_SLOT];
    __shared__ double stage[PEER_MAX_RANKS * PEER_SLOT];
    __shared__ double gsh[3 * PEER_SLOT + 8];
    __shared__ double nrm_s;
    __shared__ int okflag, guard_s;
    __shared__ bool last;
    const PeerArgs& pa = a.pa;
    const int nv = a.nv;
    const unsigned long long E = dld_volatile_u64(pa.epoch_dev);     // the epoch kry_spmv_csr_mdot published
    const bool ok = peer_wait(pa, E, &okflag);
    {
        const double* mine = pa.slots[pa.rank] + (size_t)(E & 1ull) * (size_t)pa.world * PEER_SLOT;
        for (int idx = threadIdx.x; idx < pa.world * (nv + 1); idx += blockDim.x) {
            const int r = idx / (nv + 1), j = idx - r * (nv + 1);
            stage[r * PEER_SLOT + j] = dld_volatile_f64(mine + (size_t)r * PEER_SLOT + j);
        }
        __syncthreads();
        for (int j = threadIdx.x; j <= nv; j += blockDim.x) {
            double sum = 0.0;
            for (int r = 0; r < pa.world; ++r) sum += stage[r * PEER_SLOT + j];        // rank order
            c_s[j] = ok ? sum : nan_f64();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double ww = c_s[nv];
            double est = ww;
            for (int j = 0; j < nv; ++j) est = fma(-c_s[j], c_s[j], est);              // fixed order
            const bool fine = (est >= DUS_GUARD * ww) && (ww >= 0.0) && (est <= ww);   // (false for NaN)
            guard_s = (fine || !ok) ? 0 : 1;
            nrm_s = ok ? sqrt(est > 0.0 ? est : 0.0) : nan_f64();
        }
        __syncthreads();
    }
    const int extra = a.k_givens >= 0 ? 1 : 0;
    const int nsweep = (int)gridDim.x - extra;
    const bool sweeper = (int)blockIdx.x < nsweep;
    // (the extra CTA is the LAST one and takes no elements: the sweeps stride over nsweep CTAs)
    double nrm = nrm_s;
    if (guard_s) {
        // ---- rare: heavy cancellation, take the exact norm with one more exchange ----
        if (sweeper) {
            const double part = update_scale_dispatch<T, VEC>(a.V, a.ldv, nv, c_s, a.q, (T*)nullptr, a.n, 1.0, false, nsweep);
            const double s = kry_block_sum(part, sm);
            if (threadIdx.x == 0) a.partials[blockIdx.x] = s;
        }
        // every CTA of the grid (the extra one too: it has read the epoch by now) takes a ticket; the last
        // one sums the sweepers' partials, publishes and advances the epoch
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int t = atomicAdd(a.ticket, 1u);
            last = (t == gridDim.x - 1u);
        }
        __syncthreads();
        if (last) {
            __threadfence();
            double v = 0.0;
            for (int b = threadIdx.x; b < nsweep; b += blockDim.x) v += __ldcg(a.partials + b);
            const double tot = kry_block_sum(v, sm);
            __syncthreads();
            if (threadIdx.x == 0) gsh[0] = tot;
            __syncthreads();
            peer_publish(pa, E + 1ull, gsh, 1);
            if (threadIdx.x == 0) {
                *pa.epoch_dev = E + 1ull;
                *a.ticket = 0u;
            }
        }
        const bool ok2 = peer_wait(pa, E + 1ull, &okflag);
        if (threadIdx.x == 0) nrm_s = ok2 ? sqrt(fabs(peer_sum(pa, E + 1ull, 0))) : nan_f64();
        __syncthreads();
        nrm = nrm_s;
    }
    if (!sweeper || (extra == 0 && blockIdx.x == 0)) {
        // coefficients and norm of the step: h += c, h[nv] = nrm
        if (a.h_acc)
            for (int j = threadIdx.x; j < nv; j += blockDim.x) a.h_acc[j] += c_s[j];
        if (threadIdx.x == 0) a.nrm_out[0] = nrm;
    }
    if (!sweeper) {
        __syncthreads();
        givens_body(a.k_givens, a.h_acc, a.rcol, a.cs, a.y, a.mailbox, gsh);
        return;
    }
    // ---- halo of v_next first, dealt round-robin to the sweep CTAs (entry i -> CTA i % nsweep): a few
    //      threads per CTA issue the remote loads and the sweep of the other warps hides their latency ----
    for (long long i = (long long)threadIdx.x * nsweep + blockIdx.x; i < a.nhalo; i += (long long)blockDim.x * nsweep) {
        const T* src = a.peer_q[__ldg(a.halo_peer + i)] + a.q_elem_offset;
        double w = (double)*(const volatile T*)(src + __ldg(a.halo_off + i));
        const T* vh = a.V + a.halo_base + i;
        int j = 0;
        for (; j + 8 <= nv; j += 8) {
            double vv[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) vv[t] = (double)__ldg(vh + (long long)(j + t) * a.ldv);
#pragma unroll
            for (int t = 0; t < 8; ++t) w = fma(-c_s[j + t], vv[t], w);        // the owner's fma order
        }
        for (; j < nv; ++j) w = fma(-c_s[j], (double)__ldg(vh + (long long)j * a.ldv), w);
        w = round_as<T>(w);
        a.halo_dst[i] = ok ? (T)(nrm > 0.0 ? w / nrm : 0.0) : (T)nan_f64();
    }
    update_scale_dispatch<T, VEC>(a.V, a.ldv, nv, c_s, a.q, a.vnext, a.n, nrm, true, nsweep);
}

// ---------------------------------------------------------------------------
// K3+K4 fused: acquire ||q||^2, v_next = q / nrm, publish "my segment of v_next is complete",
// acquire the peers' flags, gather the halo of v_next -- the next SpMV starts without any
// further handshake.  All CTAs are co-resident (grid <= 4 CTAs/SM), so waiting on the peers
// inside the kernel cannot starve the local publisher.
// ---------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(KRY_THREADS, 4)
dist_scale_halo_kernel(long long n, const T* q, T* vnext, double* nrm_out, long long nhalo,
                       const T* const* peer_bases, long long elem_offset, const int* __restrict__ halo_peer,
                       const int* __restrict__ halo_off, T* halo_dst, unsigned int* ticket, PeerArgs pa) {
    __shared__ int okflag;
    __shared__ double nrm_s;
    __shared__ bool last;
    const unsigned long long E = dld_volatile_u64(pa.epoch_dev);
    const bool ok = peer_wait(pa, E, &okflag);
    if (threadIdx.x == 0) nrm_s = ok ? sqrt(fabs(peer_sum(pa, E, 0))) : nan_f64();
    __syncthreads();
    const double nrm = nrm_s;
    if (blockIdx.x == 0 && threadIdx.x == 0) nrm_out[0] = nrm;
    const long long nvec = n / VEC;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        double qv[VEC];
        VecIO<T, VEC>::loadrw(q, i, qv);
#pragma unroll
        for (int u = 0; u < VEC; ++u) qv[u] = nrm > 0.0 ? qv[u] / nrm : 0.0;
        VecIO<T, VEC>::store(vnext, i, qv);
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x)
            vnext[i] = (T)(nrm > 0.0 ? (double)q[i] / nrm : 0.0);
    // my segment is complete once every CTA is here: the last one releases the flag to all peers
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < pa.world) dst_release_sys(pa.flags[threadIdx.x] + pa.rank, E + 1ull);
    }
    const bool ok2 = peer_wait(pa, E + 1ull, &okflag);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nhalo; i += stride) {
        const T* src = peer_bases[__ldg(halo_peer + i)] + elem_offset;
        const T v = *(const volatile T*)(src + __ldg(halo_off + i));
        halo_dst[i] = ok2 ? v : (T)nan_f64();
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket + 1, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *pa.epoch_dev = E + 1ull;
        ticket[0] = 0u;
        ticket[1] = 0u;
    }
}

// ---------------------------------------------------------------------------
// K4: handshake + halo gather
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(KRY_THREADS)
dist_halo_kernel(long long nhalo, const T* const* peer_bases, long long elem_offset,
                 const int* __restrict__ halo_peer, const int* __restrict__ halo_off, T* dst,
                 unsigned int* ticket, PeerArgs pa) {
    __shared__ int okflag;
    __shared__ bool last;
    const unsigned long long E = dld_volatile_u64(pa.epoch_dev);
    if (blockIdx.x == 0) {
        // everything this rank wrote before this kernel (its segment of v_k) is complete
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < pa.world) dst_release_sys(pa.flags[threadIdx.x] + pa.rank, E + 1ull);
    }
    const bool ok = peer_wait(pa, E + 1ull, &okflag);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nhalo; i += stride) {
        const T* src = peer_bases[__ldg(halo_peer + i)] + elem_offset;
        const T v = *(const volatile T*)(src + __ldg(halo_off + i));
        dst[i] = ok ? v : (T)nan_f64();
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *pa.epoch_dev = E + 1ull;
        *ticket = 0u;
    }
}

