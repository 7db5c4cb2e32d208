// Device code of kry_spmv_csr_z (csrc/kry_cplx.cu).  A header of its own so that the CPU test tier can compile
// exactly these kernels for the host over a small CUDA execution emulator (tests/csrc/cuda_emul), which supplies
// host versions of the mbarrier / bulk-copy wrappers below (KRY_EMUL).
#pragma once
#include "kry_common.cuh"

#ifndef KRY_ZTYPE
#define KRY_ZTYPE
typedef double2 Z;
#endif

// ---------------------------------------------------------------------------
// CSR SpMV on complex vectors (structure of spmv_staged_kernel, kry_spmv.cu)
// ---------------------------------------------------------------------------
#define ZSPMV_R 256
#define ZSPMV_THREADS (ZSPMV_R + 32)

#ifndef KRY_EMUL
__device__ __forceinline__ uint32_t z_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void z_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(z_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void z_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void z_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(z_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void z_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(z_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void z_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t addr = z_smem_u32(bar);
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
__device__ __forceinline__ void z_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            z_smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(z_smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void z_consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ZSPMV_R) : "memory"); }

#endif   // KRY_EMUL (the emulator defines the same functions)

// sum += a * x with separately rounded products and sums, entry after entry in storage order (the order of
// scipy's csr_matvec on complex data, krypy/utils.py:1593-1594)
template <typename TV> struct ZMul;
template <> struct ZMul<double2> {
    static __device__ __forceinline__ void madd(const double2 a, const double2 x, double& sr, double& si) {
        sr = __dadd_rn(sr, __dsub_rn(__dmul_rn(a.x, x.x), __dmul_rn(a.y, x.y)));
        si = __dadd_rn(si, __dadd_rn(__dmul_rn(a.x, x.y), __dmul_rn(a.y, x.x)));
    }
};
template <> struct ZMul<double> {
    static __device__ __forceinline__ void madd(const double a, const double2 x, double& sr, double& si) {
        sr = __dadd_rn(sr, __dmul_rn(a, x.x));
        si = __dadd_rn(si, __dmul_rn(a, x.y));
    }
};

template <typename TV, int CPR, int STAGES>
struct ZSpmvCfg {
    static const int CAP = ZSPMV_R * CPR + 8;                      // entries per stage (multiple of 4)
    static const int STAGE_BYTES = CAP * (int)(sizeof(TV) + sizeof(int));
    static const int SMEM_BYTES = 128 + STAGES * STAGE_BYTES;
};

struct ZTileRows {
    int s, e, a, b;
};

__device__ __forceinline__ ZTileRows z_load_tile_rows(const int* __restrict__ rowptr, long long nrows, long long t,
                                                      int tid) {
    ZTileRows r;
    const long long r0 = t * ZSPMV_R;
    const long long r1 = (r0 + ZSPMV_R < nrows) ? r0 + ZSPMV_R : nrows;
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

template <typename TV, int CPR, int STAGES>
__global__ void __launch_bounds__(ZSPMV_THREADS)
zspmv_staged_kernel(long long nrows, long long nnz, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                    const TV* __restrict__ vals, const Z* __restrict__ x, Z* y) {
    typedef ZSpmvCfg<TV, CPR, STAGES> Cfg;
    const int CAP = Cfg::CAP;
#ifdef KRY_EMUL
    unsigned char* smem = kry_emul_dynamic_smem();
#else
    extern __shared__ __align__(128) unsigned char smem[];
#endif
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // [STAGES] producer -> consumers (tx bytes)
    uint64_t* empty = full + STAGES;                            // [STAGES] consumers -> producer
    unsigned char* stage_base = smem + 128;

    const int tid = threadIdx.x;
    const long long ntiles = (nrows + ZSPMV_R - 1) / ZSPMV_R;
    const long long G = gridDim.x;
    const long long nmine = ((long long)blockIdx.x < ntiles) ? (ntiles - blockIdx.x + G - 1) / G : 0;
    const int nnz_al = (int)(nnz & ~3LL);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            z_mbar_init(&full[s], 1);
            z_mbar_init(&empty[s], ZSPMV_R / 32);
        }
        z_mbar_fence_init();
    }
    __syncthreads();

    if (tid >= ZSPMV_R) {
        // ---------------- producer warp: one elected lane drives the TMA ring ----------------
        if (tid == ZSPMV_R && nmine > 0) {
            long long t = blockIdx.x;
            long long r0 = t * ZSPMV_R;
            long long r1 = (r0 + ZSPMV_R < nrows) ? r0 + ZSPMV_R : nrows;
            int s = __ldg(rowptr + r0), e = __ldg(rowptr + r1);
            for (long long it = 0; it < nmine; ++it) {
                int s_n = 0, e_n = 0;
                if (it + 1 < nmine) {
                    const long long tn = blockIdx.x + (it + 1) * G;
                    const long long q0 = tn * ZSPMV_R;
                    const long long q1 = (q0 + ZSPMV_R < nrows) ? q0 + ZSPMV_R : nrows;
                    s_n = __ldg(rowptr + q0);
                    e_n = __ldg(rowptr + q1);
                }
                const int st = (int)(it % STAGES);
                if (it >= STAGES) z_mbar_wait(&empty[st], (uint32_t)(((it / STAGES) - 1) & 1));
                const int s_al = s & ~3;
                const int e_al = (e + 3) & ~3;
                const int e_bulk = e_al < nnz_al ? e_al : nnz_al;
                const int cnt = e_bulk - s_al;
                TV* sv = reinterpret_cast<TV*>(stage_base + (size_t)st * Cfg::STAGE_BYTES);
                int* sc = reinterpret_cast<int*>(stage_base + (size_t)st * Cfg::STAGE_BYTES + (size_t)CAP * sizeof(TV));
                if (e_al - s_al <= CAP && cnt > 0) {
                    z_mbar_expect_tx(&full[st], (uint32_t)cnt * (uint32_t)(sizeof(TV) + sizeof(int)));
                    z_bulk_g2s(sv, vals + s_al, (uint32_t)cnt * (uint32_t)sizeof(TV), &full[st]);
                    z_bulk_g2s(sc, colidx + s_al, (uint32_t)cnt * (uint32_t)sizeof(int), &full[st]);
                } else {
                    z_mbar_expect_tx(&full[st], 0u);   // nothing staged: complete the phase at once
                }
                s = s_n;
                e = e_n;
            }
        }
    } else if (nmine > 0) {
        // ---------------- consumers: thread per row out of shared memory ----------------
        ZTileRows cur = z_load_tile_rows(rowptr, nrows, blockIdx.x, tid);
        for (long long it = 0; it < nmine; ++it) {
            const long long t = blockIdx.x + it * G;
            ZTileRows nxt = cur;
            if (it + 1 < nmine) nxt = z_load_tile_rows(rowptr, nrows, t + G, tid);
            const int st = (int)(it % STAGES);
            const uint32_t parity = (uint32_t)((it / STAGES) & 1);
            const long long r0 = t * ZSPMV_R;
            const long long r1 = (r0 + ZSPMV_R < nrows) ? r0 + ZSPMV_R : nrows;
            const int s = cur.s, e = cur.e, a = cur.a, b = cur.b;
            const int s_al = s & ~3;
            const int e_al = (e + 3) & ~3;
            const bool staged = (e_al - s_al) <= CAP;
            const long long row = r0 + tid;
            TV* sv = reinterpret_cast<TV*>(stage_base + (size_t)st * Cfg::STAGE_BYTES);
            int* sc = reinterpret_cast<int*>(stage_base + (size_t)st * Cfg::STAGE_BYTES + (size_t)CAP * sizeof(TV));
            double sr = 0.0, si = 0.0;
            z_mbar_wait(&full[st], parity);
            if (staged) {
                const int e_bulk = e_al < nnz_al ? e_al : nnz_al;
                if (e > e_bulk) {
                    // the last (<4) entries of the matrix are not 16-byte coverable in colidx: plain copy
                    for (int jj = e_bulk + tid; jj < e; jj += ZSPMV_R) {
                        sv[jj - s_al] = vals[jj];
                        sc[jj - s_al] = colidx[jj];
                    }
                    z_consumer_bar_sync();
                }
                int jj = a - s_al;
                const int end = b - s_al;
                for (; jj + 4 <= end; jj += 4) {
                    const int c0 = sc[jj], c1 = sc[jj + 1], c2 = sc[jj + 2], c3 = sc[jj + 3];
                    const Z x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
                    ZMul<TV>::madd(sv[jj], x0, sr, si);
                    ZMul<TV>::madd(sv[jj + 1], x1, sr, si);
                    ZMul<TV>::madd(sv[jj + 2], x2, sr, si);
                    ZMul<TV>::madd(sv[jj + 3], x3, sr, si);
                }
                if (jj < end) {   // 1..3 remaining entries: gather first, then the ordered sum
                    const int m = end - jj;
                    const int c0 = sc[jj];
                    const int c1 = m > 1 ? sc[jj + 1] : c0;
                    const int c2 = m > 2 ? sc[jj + 2] : c0;
                    const Z x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2);
                    ZMul<TV>::madd(sv[jj], x0, sr, si);
                    if (m > 1) ZMul<TV>::madd(sv[jj + 1], x1, sr, si);
                    if (m > 2) ZMul<TV>::madd(sv[jj + 2], x2, sr, si);
                }
            } else {
                for (int jj = a; jj < b; ++jj) ZMul<TV>::madd(__ldg(vals + jj), __ldg(x + __ldg(colidx + jj)), sr, si);
            }
            if (row < r1) y[row] = make_double2(sr, si);
            // this warp is done with slot st: let the producer refill it
            __syncwarp();
            if ((tid & 31) == 0) z_mbar_arrive(&empty[st]);
            cur = nxt;
        }
    }
}

// long rows / unaligned arrays: warp per row, coalesced loads, shuffle reduction
template <typename TV>
__global__ void __launch_bounds__(KRY_THREADS)
zspmv_warp_kernel(long long nrows, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                  const TV* __restrict__ vals, const Z* __restrict__ x, Z* y) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < nrows; r += nwarps) {
        const int a = __ldg(rowptr + r), b = __ldg(rowptr + r + 1);
        double sr = 0.0, si = 0.0;
        for (int jj = a + lane; jj < b; jj += 32) ZMul<TV>::madd(__ldg(vals + jj), __ldg(x + __ldg(colidx + jj)), sr, si);
        sr = kry_warp_sum(sr);
        si = kry_warp_sum(si);
        if (lane == 0) y[r] = make_double2(sr, si);
    }
}

