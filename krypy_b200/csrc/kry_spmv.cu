// CSR SpMV for sm_100a.
//
// Short-row matrices (stencils: 5/7 nnz per row) use a persistent, warp-specialised
// kernel: one producer warp stages each 256-row block's contiguous vals[] and colidx[]
// ranges into shared memory with 1-D TMA bulk copies (cp.async.bulk.shared::cluster.global
// + mbarrier complete_tx, SASS UBLKCP) through a STAGES-deep ring (full/empty mbarriers),
// so tens of KB per SM are in flight without holding registers.  Eight consumer warps run
// thread-per-row out of shared memory: for banded matrices the x gathers of neighbouring
// threads are contiguous (coalesced ld.global.nc), the next tile's row pointers are
// prefetched into registers while the current tile is computed, and each row is summed
// sequentially left to right with separately rounded multiply and add -- the same order and
// rounding as scipy's csr_matvec, so y is bit-identical to the reference's SpMV in fp64.
//
// Long-row matrices (or unaligned arrays) use a warp-per-row kernel with coalesced direct
// loads and a shuffle reduction.
#include <stdlib.h>
#include "kry_common.cuh"

#define KRY_ENTER(ctx)                                                         \
    KRY_REQUIRE((ctx) != nullptr, "ctx is NULL");                              \
    KRY_CHECK_CUDA(cudaSetDevice((ctx)->device))

#define SPMV_R 256                    // rows per tile == consumer threads per CTA
#define SPMV_THREADS (SPMV_R + 32)    // 8 consumer warps + 1 producer warp

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
    extern __shared__ __align__(128) unsigned char smem[];
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

static int spmv_stage_override() {   // tuning knob: KRY_SPMV_STAGES=2|3|4
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("KRY_SPMV_STAGES");
        v = s ? atoi(s) : 0;
        if (v != 2 && v != 3 && v != 4) v = 0;
    }
    return v;
}

template <typename T, int CPR, int STAGES, bool DOT, int NACC>
static int launch_staged_md(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                            const T* vals, const T* x, T* y, const T* w, double* dot_out, const MDotArgs<T>& md) {
    typedef SpmvCfg<T, CPR, STAGES> Cfg;
    auto kern = spmv_staged_kernel<T, CPR, STAGES, DOT, NACC>;
    static thread_local int occ[16] = {0};
    int& o = occ[ctx->device & 15];
    if (o == 0) {
        KRY_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        int nb = 0;
        KRY_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, SPMV_THREADS, Cfg::SMEM_BYTES));
        KRY_REQUIRE(nb >= 1, "staged SpMV kernel does not fit on an SM");
        o = nb;
    }
    long long ntiles = (nrows + SPMV_R - 1) / SPMV_R;
    long long cap = (long long)ctx->sm_count * o;
    if (cap > KRY_MAX_PARTIAL_BLOCKS) cap = KRY_MAX_PARTIAL_BLOCKS;
    int g = (int)(ntiles < cap ? ntiles : cap);
    if (g < 1) g = 1;
    kern<<<g, SPMV_THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(nrows, nnz, rowptr, colidx, vals, x, y, w,
                                                            ctx->d_partials, ctx->d_ticket + (NACC > 0 ? 13 : 1),
                                                            dot_out, md);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

template <typename T, int CPR, int STAGES, bool DOT>
static int launch_staged(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                         const T* vals, const T* x, T* y, const T* w, double* dot_out) {
    MDotArgs<T> md;
    memset(&md, 0, sizeof(md));
    return launch_staged_md<T, CPR, STAGES, DOT, 0>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out, md);
}

// multi-dot epilogue: accumulator tile (8 / 16 / 32 registers) by the number of vectors; deeper TMA ring
// where the register tile leaves a single CTA per SM
template <typename T, int CPR>
static int launch_mdot_cpr(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                           const T* vals, const T* x, T* y, const MDotArgs<T>& md) {
    if (md.nb <= 8)
        return launch_staged_md<T, CPR, 2, false, 8>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, nullptr, nullptr, md);
    if (md.nb <= 16)
        return launch_staged_md<T, CPR, 3, false, 16>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, nullptr, nullptr, md);
    return launch_staged_md<T, CPR, 4, false, 32>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, nullptr, nullptr, md);
}

template <typename T>
static int spmv_mdot_dispatch(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                              const T* vals, const T* x, T* y, const MDotArgs<T>& md) {
    const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    const bool al = kry_aligned16(vals) && kry_aligned16(colidx);
    if (al && avg <= 5.5) return launch_mdot_cpr<T, 6>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, md);
    if (al && avg <= 7.5) return launch_mdot_cpr<T, 8>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, md);
    if (al && avg <= 15.0) return launch_mdot_cpr<T, 16>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, md);
    kry_set_error("kry_spmv_csr_mdot: only the staged short-row path (<= 15 entries per row on average, 16-byte "
                  "aligned matrix arrays) carries the multi-dot epilogue");
    return KRY_ERR_UNSUPPORTED;
}

template <typename T, int CPR, bool DOT>
static int launch_staged_cpr(kry_ctx* ctx, int stages, long long nrows, long long nnz, const int* rowptr,
                             const int* colidx, const T* vals, const T* x, T* y, const T* w, double* dot_out) {
    if (stages == 2) return launch_staged<T, CPR, 2, DOT>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    if (stages == 4) return launch_staged<T, CPR, 4, DOT>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    return launch_staged<T, CPR, 3, DOT>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
}

template <typename T, bool DOT>
static int spmv_dispatch(kry_ctx* ctx, long long nrows, long long nnz, const int* rowptr, const int* colidx,
                         const T* vals, const T* x, T* y, const T* w, double* dot_out) {
    const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
    const bool al = kry_aligned16(vals) && kry_aligned16(colidx);
    const int ov = spmv_stage_override();
    // stage capacity (entries per row) just above the mean row length; a 2-deep ring measured
    // fastest on B200 (118 us vs 121/133 us for 3/4 stages on config C2: occupancy beats depth)
    if (al && avg <= 5.5)
        return launch_staged_cpr<T, 6, DOT>(ctx, ov ? ov : 2, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    if (al && avg <= 7.5)
        return launch_staged_cpr<T, 8, DOT>(ctx, ov ? ov : 2, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    if (al && avg <= 15.0)
        return launch_staged_cpr<T, 16, DOT>(ctx, ov ? ov : 2, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    if (al && avg <= 30.0)
        return launch_staged<T, 32, 2, DOT>(ctx, nrows, nnz, rowptr, colidx, vals, x, y, w, dot_out);
    long long need = (nrows * 32 + KRY_THREADS - 1) / KRY_THREADS;
    long long cap = (long long)ctx->sm_count * 8;
    int g = (int)(need < cap ? need : cap);
    if (g < 1) g = 1;
    spmv_warp_kernel<T, DOT><<<g, KRY_THREADS, 0, ctx->stream>>>(nrows, rowptr, colidx, vals, x, y, w,
                                                                 ctx->d_partials, ctx->d_ticket + 1, dot_out);
    KRY_LAUNCHED(ctx);
    return KRY_OK;
}

extern "C" int kry_spmv_csr(kry_ctx* ctx, int dtype, long long nrows, long long ncols, long long nnz,
                            const int* rowptr, const int* colidx, const void* vals, const void* x, void* y,
                            const void* w_dev, double* dot_out_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(nrows >= 0 && ncols >= 0 && nnz >= 0, "negative size");
    KRY_REQUIRE(nnz < 2147483647LL, "nnz must fit int32 row pointers");
    KRY_REQUIRE(rowptr && x, "NULL argument");
    KRY_REQUIRE(nnz == 0 || (colidx && vals), "NULL matrix arrays");
    KRY_REQUIRE(y || w_dev, "neither y nor a dot epilogue requested");
    KRY_REQUIRE(!w_dev || dot_out_dev, "w given without dot_out");
    if (nrows == 0) return KRY_OK;
    if (dtype == KRY_F64) {
        if (w_dev)
            return spmv_dispatch<double, true>(ctx, nrows, nnz, rowptr, colidx, (const double*)vals,
                                               (const double*)x, (double*)y, (const double*)w_dev, dot_out_dev);
        return spmv_dispatch<double, false>(ctx, nrows, nnz, rowptr, colidx, (const double*)vals, (const double*)x,
                                            (double*)y, nullptr, nullptr);
    }
    if (dtype == KRY_F32) {
        if (w_dev)
            return spmv_dispatch<float, true>(ctx, nrows, nnz, rowptr, colidx, (const float*)vals, (const float*)x,
                                              (float*)y, (const float*)w_dev, dot_out_dev);
        return spmv_dispatch<float, false>(ctx, nrows, nnz, rowptr, colidx, (const float*)vals, (const float*)x,
                                           (float*)y, nullptr, nullptr);
    }
    kry_set_error("kry_spmv_csr: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}

// y = A x and c[j] = <B[j], y> (j < nb), c[nb] = <y, y> (want_sq) in ONE pass: the dots are taken while each
// row's result is still in the register of the thread that computed it.
//   world == 1: the sums go to out_dev[0 .. nb + want_sq)
//   world  > 1: row-partitioned run; the local sums are stored into every peer's slot array and this
//               rank's flag is released with epoch + 1 (no wait here: the consumer, kry_dist_update_scale
//               or kry_dist_update, acquires)
// Replaces utils.py:968 (A v_k) + the k+1 inner products of utils.py:1015 (block classical Gram-Schmidt),
// and deflation.py:135-143 / utils.py:604-627 (<W, A v> of the projector) when B = W.
extern "C" int kry_spmv_csr_mdot(kry_ctx* ctx, int dtype, long long nrows, long long ncols, long long nnz,
                                 const int* rowptr, const int* colidx, const void* vals, const void* x, void* y,
                                 const void* B, long long ldb, int nb, int want_sq, double* out_dev, int world,
                                 int rank, unsigned long long* epoch_dev, double* const* peer_slots_dev,
                                 unsigned long long* const* peer_flags_dev) {
    KRY_ENTER(ctx);
    KRY_REQUIRE(nrows >= 1 && ncols >= 0 && nnz >= 0, "bad size");
    KRY_REQUIRE(nnz < 2147483647LL, "nnz must fit int32 row pointers");
    KRY_REQUIRE(rowptr && x && y, "NULL argument");
    KRY_REQUIRE(nnz == 0 || (colidx && vals), "NULL matrix arrays");
    KRY_REQUIRE(nb >= 0 && nb <= 32 && nb + (want_sq ? 1 : 0) >= 1, "0 <= nb <= 32 dot vectors (+ the square) per call");
    KRY_REQUIRE(nb == 0 || B, "NULL dot basis");
    KRY_REQUIRE(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world, "bad world/rank");
    KRY_REQUIRE(world > 1 ? (epoch_dev && peer_slots_dev && peer_flags_dev) : (out_dev != nullptr),
                "world > 1 needs the peer tables, world == 1 an output array");
    PeerArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.world = world;
    pa.rank = rank;
    pa.epoch_dev = epoch_dev;
    pa.slots = peer_slots_dev;
    pa.flags = peer_flags_dev;
    if (dtype == KRY_F64) {
        MDotArgs<double> md = {(const double*)B, ldb, nb, want_sq ? 1 : 0, out_dev, pa};
        return spmv_mdot_dispatch<double>(ctx, nrows, nnz, rowptr, colidx, (const double*)vals, (const double*)x,
                                          (double*)y, md);
    }
    if (dtype == KRY_F32) {
        MDotArgs<float> md = {(const float*)B, ldb, nb, want_sq ? 1 : 0, out_dev, pa};
        return spmv_mdot_dispatch<float>(ctx, nrows, nnz, rowptr, colidx, (const float*)vals, (const float*)x,
                                         (float*)y, md);
    }
    kry_set_error("kry_spmv_csr_mdot: unsupported dtype %d", dtype);
    return KRY_ERR_UNSUPPORTED;
}
