// Device code of kry_spmv_csr / kry_spmv_csr_mdot (csrc/kry_spmv.cu): the TMA-staged warp-specialised kernel and
// the warp-per-row kernel.  A header of its own so that the CPU test tier can compile exactly these kernels for
// the host over the CUDA execution emulator (tests/csrc/cuda_emul, tests/test_spmv_emul_cpu.py).
#pragma once
#include "kry_common.cuh"

#define SPMV_R 256                    // rows per tile == consumer threads per CTA
#define SPMV_THREADS (SPMV_R + 32)    // 8 consumer warps + 1 producer warp

#ifndef KRY_EMUL   // (the CPU tier's execution emulator supplies host versions of these wrappers)
// ---- mbarrier / bulk-copy PTX ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() {   // named barrier 1 over the consumer threads
    asm volatile("bar.sync 1, %0;" ::"n"(SPMV_R) : "memory");
}

#endif

template <typename T, int CPR, int STAGES>
struct SpmvCfg {
    static const int CAP = SPMV_R * CPR + 8;                       // entries per stage (multiple of 4)
    static const int STAGE_BYTES = CAP * (int)(sizeof(T) + sizeof(int));
    static const int SMEM_BYTES = 128 + STAGES * STAGE_BYTES;
};

// finish a CTA-partial dot: write partial, last CTA reduces in fixed order
__device__ __forceinline__ void finish_dot(double acc, double* partials, unsigned int* ticket, double* dot_out,
                                           double* sm, bool* last_flag) {
    double s = kry_block_sum(acc, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        *last_flag = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (*last_flag) {
        __threadfence();
        double v = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) v += __ldcg(partials + b);
        double r = kry_block_sum(v, sm);
        if (threadIdx.x == 0) {
            dot_out[0] = r;
            *ticket = 0u;
        }
    }
}

struct TileRows {   // row-pointer values a consumer thread needs for one tile
    int s, e, a, b;
};

__device__ __forceinline__ TileRows load_tile_rows(const int* __restrict__ rowptr, long long nrows, long long t,
                                                   int tid) {
    TileRows r;
    const long long r0 = t * SPMV_R;
    const long long r1 = (r0 + SPMV_R < nrows) ? r0 + SPMV_R : nrows;
    r.s = __ldg(rowptr + r0);
    r.e = __ldg(rowptr + r1);
    const long long row = r0 + tid;
    if (row < r1) {
        r.a = __ldg(rowptr + row);
        r.b = __ldg(rowptr + row + 1);
    } else {
        r.a = r.b = 0;
    }
    return r;
}

// Multi-vector dot epilogue (NACC > 0: up to NACC vectors): c[j] = <B[j], y> for j < nb and, when want_sq,
// c[nb] = <y, y>, accumulated by the thread that produces y[row] while the row's result is still in a
// register -- the Arnoldi step's V^H (A v) / the deflation projector's W^H (A v) without re-reading A v.
// The basis entries of the row are requested BEFORE the thread waits for the tile's matrix data, so their
// latency overlaps with the TMA stage and the x gathers.  Deterministic: fixed shuffle tree, fixed warp
// order, per-CTA partials summed in CTA order by the last CTA.  Row-partitioned runs (pa.world > 1): the
// last CTA stores the local sums straight into every peer's slot array and releases its flag (epoch + 1),
// exactly as kry_dist_dot does.
// MEASURED (B200, profiles/r2_mdot_kernel.txt): the accumulators cost the occupancy the x gathers live on
// (72 / 96 / 168 registers against 32), so the fused kernel only ties SpMV + block dot for >= 16 vectors
// and loses below; a second design (y tile in shared memory, one warp per vector, 56-72 registers) was
// slower still.  The solvers therefore keep the two-kernel form; this entry point stays for callers whose
// dot basis is wide and for the record.
template <typename T>
struct MDotArgs {
    const T* B;
    long long ldb;
    int nb, want_sq;
    double* out;        // pa.world == 1: the nb (+1) sums
    PeerArgs pa;
};

template <typename T, int CPR, int STAGES, bool DOT, int NACC>
__global__ void __launch_bounds__(SPMV_THREADS, (NACC == 16 ? 2 : 0))
spmv_staged_kernel(long long nrows, long long nnz, const int* __restrict__ rowptr,
                   const int* __restrict__ colidx, const T* __restrict__ vals, const T* __restrict__ x, T* y,
                   const T* __restrict__ w, double* partials, unsigned int* ticket, double* dot_out,
                   MDotArgs<T> md) {
    typedef SpmvCfg<T, CPR, STAGES> Cfg;
    const int CAP = Cfg::CAP;
#ifdef KRY_EMUL
    unsigned char* smem = kry_emul_dynamic_smem();
#else
    extern __shared__ __align__(128) unsigned char smem[];
#endif
    __shared__ double red_sm[32];
    __shared__ bool last_flag;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // [STAGES] producer -> consumers (tx bytes)
    uint64_t* empty = full + STAGES;                            // [STAGES] consumers -> producer
    unsigned char* stage_base = smem + 128;

    const int tid = threadIdx.x;
    const long long ntiles = (nrows + SPMV_R - 1) / SPMV_R;
    const long long G = gridDim.x;
    const long long nmine = ((long long)blockIdx.x < ntiles) ? (ntiles - blockIdx.x + G - 1) / G : 0;
    const int nnz_al = (int)(nnz & ~3LL);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SPMV_R / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    double dot_acc = 0.0;
    double macc[NACC > 0 ? NACC : 1], msq = 0.0;
#pragma unroll
    for (int t = 0; t < (NACC > 0 ? NACC : 1); ++t) macc[t] = 0.0;
    if (tid >= SPMV_R) {
        // ---------------- producer warp: one elected lane drives the TMA ring ----------------
        if (tid == SPMV_R && nmine > 0) {
            long long t = blockIdx.x;
            long long r0 = t * SPMV_R;
            long long r1 = (r0 + SPMV_R < nrows) ? r0 + SPMV_R : nrows;
            int s = __ldg(rowptr + r0), e = __ldg(rowptr + r1);
            for (long long it = 0; it < nmine; ++it) {
                // prefetch the next tile's extent before blocking on the ring slot
                int s_n = 0, e_n = 0;
                if (it + 1 < nmine) {
                    const long long tn = blockIdx.x + (it + 1) * G;
                    const long long q0 = tn * SPMV_R;
                    const long long q1 = (q0 + SPMV_R < nrows) ? q0 + SPMV_R : nrows;
                    s_n = __ldg(rowptr + q0);
                    e_n = __ldg(rowptr + q1);
                }
                const int st = (int)(it % STAGES);
                if (it >= STAGES) mbar_wait(&empty[st], (uint32_t)(((it / STAGES) - 1) & 1));
                const int s_al = s & ~3;
                const int e_al = (e + 3) & ~3;
                const int e_bulk = e_al < nnz_al ? e_al : nnz_al;
                const int cnt = e_bulk - s_al;
                T* sv = reinterpret_cast<T*>(stage_base + (size_t)st * Cfg::STAGE_BYTES);
                int* sc = reinterpret_cast<int*>(stage_base + (size_t)st * Cfg::STAGE_BYTES + (size_t)CAP * sizeof(T));
                if (e_al - s_al <= CAP && cnt > 0) {
                    mbar_expect_tx(&full[st], (uint32_t)cnt * (uint32_t)(sizeof(T) + sizeof(int)));
                    bulk_g2s(sv, vals + s_al, (uint32_t)cnt * (uint32_t)sizeof(T), &full[st]);
                    bulk_g2s(sc, colidx + s_al, (uint32_t)cnt * (uint32_t)sizeof(int), &full[st]);
                } else {
                    mbar_expect_tx(&full[st], 0u);  // nothing staged: complete the phase at once
                }
                s = s_n;
                e = e_n;
            }
        }
    } else if (nmine > 0) {
        // ---------------- consumers: thread per row out of shared memory ----------------
        TileRows cur = load_tile_rows(rowptr, nrows, blockIdx.x, tid);
        for (long long it = 0; it < nmine; ++it) {
            const long long t = blockIdx.x + it * G;
            // next tile's row pointers: issued now, consumed next iteration
            TileRows nxt = cur;
            if (it + 1 < nmine) nxt = load_tile_rows(rowptr, nrows, t + G, tid);
            const int st = (int)(it % STAGES);
            const uint32_t parity = (uint32_t)((it / STAGES) & 1);
            const long long r0 = t * SPMV_R;
            const long long r1 = (r0 + SPMV_R < nrows) ? r0 + SPMV_R : nrows;
            const int s = cur.s, e = cur.e, a = cur.a, b = cur.b;
            const int s_al = s & ~3;
            const int e_al = (e + 3) & ~3;
            const bool staged = (e_al - s_al) <= CAP;
            const long long row = r0 + tid;
            T* sv = reinterpret_cast<T*>(stage_base + (size_t)st * Cfg::STAGE_BYTES);
            int* sc = reinterpret_cast<int*>(stage_base + (size_t)st * Cfg::STAGE_BYTES + (size_t)CAP * sizeof(T));
            double sum = 0.0;
            double bv[NACC > 0 ? NACC : 1];
            if (NACC > 0) {
#pragma unroll
                for (int t = 0; t < NACC; ++t)
                    bv[t] = (t < md.nb && row < r1) ? (double)__ldg(md.B + (long long)t * md.ldb + row) : 0.0;
            }
            mbar_wait(&full[st], parity);
            if (staged) {
                const int e_bulk = e_al < nnz_al ? e_al : nnz_al;
                if (e > e_bulk) {
                    // the last (<4) entries of the matrix are not 16-byte coverable: plain copy
                    for (int jj = e_bulk + tid; jj < e; jj += SPMV_R) {
                        sv[jj - s_al] = vals[jj];
                        sc[jj - s_al] = colidx[jj];
                    }
                    consumer_bar_sync();
                }
                int jj = a - s_al;
                const int end = b - s_al;
                for (; jj + 4 <= end; jj += 4) {
                    const int c0 = sc[jj], c1 = sc[jj + 1], c2 = sc[jj + 2], c3 = sc[jj + 3];
                    const double x0 = (double)__ldg(x + c0), x1 = (double)__ldg(x + c1);
                    const double x2 = (double)__ldg(x + c2), x3 = (double)__ldg(x + c3);
                    sum = __dadd_rn(sum, __dmul_rn((double)sv[jj], x0));
                    sum = __dadd_rn(sum, __dmul_rn((double)sv[jj + 1], x1));
                    sum = __dadd_rn(sum, __dmul_rn((double)sv[jj + 2], x2));
                    sum = __dadd_rn(sum, __dmul_rn((double)sv[jj + 3], x3));
                }
                if (jj < end) {   // 1..3 remaining entries: gather first, then the ordered sum
                    const int n = end - jj;
                    const int c0 = sc[jj];
                    const int c1 = n > 1 ? sc[jj + 1] : c0;
                    const int c2 = n > 2 ? sc[jj + 2] : c0;
                    const double x0 = (double)__ldg(x + c0), x1 = (double)__ldg(x + c1), x2 = (double)__ldg(x + c2);
                    sum = __dadd_rn(sum, __dmul_rn((double)sv[jj], x0));
                    if (n > 1) sum = __dadd_rn(sum, __dmul_rn((double)sv[jj + 1], x1));
                    if (n > 2) sum = __dadd_rn(sum, __dmul_rn((double)sv[jj + 2], x2));
                }
            } else {
                for (int jj = a; jj < b; ++jj)
                    sum = __dadd_rn(sum, __dmul_rn((double)__ldg(vals + jj), (double)__ldg(x + __ldg(colidx + jj))));
            }
            if (row < r1) {
                if (y) y[row] = (T)sum;
                if (DOT) dot_acc = fma((double)__ldg(w + row), (double)(T)sum, dot_acc);
                if (NACC > 0) {
                    const double ys = (double)(T)sum;           // the value as stored
#pragma unroll
                    for (int t = 0; t < NACC; ++t) macc[t] = fma(bv[t], ys, macc[t]);
                    msq = fma(ys, ys, msq);
                }
            }
            // this warp is done with slot st: let the producer refill it
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[st]);
            cur = nxt;
        }
    }
    if (DOT) finish_dot(dot_acc, partials, ticket, dot_out, red_sm, &last_flag);
    if (NACC > 0) {
        __shared__ double mred[(NACC > 0 ? NACC + 1 : 1) * 8];
        __shared__ double mfin[PEER_SLOT];
        const int nred = md.nb + (md.want_sq ? 1 : 0);
        const int lane = tid & 31, wp = tid >> 5;
        if (tid < SPMV_R) {
#pragma unroll
            for (int t = 0; t < NACC; ++t) {
                if (t < md.nb) {                              // uniform
                    const double sj = kry_warp_sum(macc[t]);
                    if (lane == 0) mred[t * 8 + wp] = sj;
                }
            }
            if (md.want_sq) {
                const double sq = kry_warp_sum(msq);
                if (lane == 0) mred[md.nb * 8 + wp] = sq;
            }
        }
        __syncthreads();
        if (tid < nred) {
            double sj = 0.0;
            for (int ww = 0; ww < SPMV_R / 32; ++ww) sj += mred[tid * 8 + ww];      // fixed warp order
            partials[(size_t)tid * KRY_MAX_PARTIAL_BLOCKS + blockIdx.x] = sj;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            unsigned int tk = atomicAdd(ticket, 1u);
            last_flag = (tk == gridDim.x - 1);
        }
        __syncthreads();
        if (last_flag) {
            __threadfence();
            const int nw = SPMV_THREADS >> 5;
            for (int j = wp; j < nred; j += nw) {             // one warp per sum, lanes stride over the CTAs
                double v = 0.0;
                for (int b = lane; b < (int)gridDim.x; b += 32)
                    v += __ldcg(partials + (size_t)j * KRY_MAX_PARTIAL_BLOCKS + b);
                v = kry_warp_sum(v);
                if (lane == 0) mfin[j] = v;
            }
            __syncthreads();
            if (md.pa.world > 1) {
                const unsigned long long E = dld_volatile_u64(md.pa.epoch_dev);
                peer_publish(md.pa, E + 1ull, mfin, nred);
                if (tid == 0) *md.pa.epoch_dev = E + 1ull;
            } else {
                if (tid < nred) md.out[tid] = mfin[tid];
            }
            if (tid == 0) *ticket = 0u;
        }
    }
}

template <typename T, bool DOT>
__global__ void __launch_bounds__(KRY_THREADS)
spmv_warp_kernel(long long nrows, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                 const T* __restrict__ vals, const T* __restrict__ x, T* y, const T* __restrict__ w,
                 double* partials, unsigned int* ticket, double* dot_out) {
    __shared__ double red_sm[32];
    __shared__ bool last_flag;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    double dot_acc = 0.0;
    for (long long r = warp; r < nrows; r += nwarps) {
        const int a = __ldg(rowptr + r), b = __ldg(rowptr + r + 1);
        double acc = 0.0;
        for (int jj = a + lane; jj < b; jj += 32)
            acc = fma((double)__ldg(vals + jj), (double)__ldg(x + __ldg(colidx + jj)), acc);
        acc = kry_warp_sum(acc);
        if (lane == 0) {
            if (y) y[r] = (T)acc;
            if (DOT) dot_acc = fma((double)__ldg(w + r), (double)(T)acc, dot_acc);
        }
    }
    if (DOT) finish_dot(dot_acc, partials, ticket, dot_out, red_sm, &last_flag);
}

