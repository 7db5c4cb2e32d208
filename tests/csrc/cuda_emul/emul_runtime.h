// TEST INFRASTRUCTURE -- runtime of the CUDA execution emulator (see cuda_runtime.h in this directory): include
// ONCE per harness program, before the device headers.  One OS thread per CUDA thread, one process per CTA, pthread
// barriers for __syncthreads / warp shuffles / grid.sync, a model of mbarriers + bulk copies that checks alignment,
// byte counts and the ring protocol, and host versions of the peer-exchange PTX helpers of kry_common.cuh
// (release / acquire on memory shared by the emulated ranks).
#pragma once
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <functional>
#include <map>
#include <vector>

#include <cuda_runtime.h>

// ------------------------------------------------------------------ emulator runtime
thread_local uint3 threadIdx;
uint3 blockIdx;
dim3 blockDim, gridDim;

struct WarpX {
    pthread_barrier_t bar;
    double buf[32];
};
static pthread_barrier_t g_block_bar, g_named_bar;
static WarpX g_warps[32];
static pthread_barrier_t* g_grid_bar;
static unsigned char* g_dyn_smem;

void __syncthreads() { pthread_barrier_wait(&g_block_bar); }
void __syncwarp() { pthread_barrier_wait(&g_warps[threadIdx.x >> 5].bar); }
void __threadfence() { __sync_synchronize(); }
void __threadfence_system() { __sync_synchronize(); }
unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
unsigned char* kry_emul_dynamic_smem() { return g_dyn_smem; }
double __shfl_xor_sync(unsigned int, double v, int lane_mask) {
    WarpX& w = g_warps[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    w.buf[lane] = v;
    pthread_barrier_wait(&w.bar);
    const double r = w.buf[lane ^ lane_mask];
    pthread_barrier_wait(&w.bar);
    return r;
}
static inline void emul_jitter(int one_in);
void kry_emul_grid_sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
        emul_jitter(2);
        pthread_barrier_wait(g_grid_bar);
    }
    __syncthreads();
}

// mbarrier + bulk copy (the PTX wrappers of kry_zspmv.cuh)
struct EmBar {
    uint32_t count;
    int32_t pending;
    int64_t tx;
    uint32_t phase;
};
static std::map<uintptr_t, EmBar> g_bars;
static pthread_mutex_t g_bar_mu = PTHREAD_MUTEX_INITIALIZER;
static void embar_check(EmBar& b) {
    if (b.tx < 0) {
        fprintf(stderr, "emulated mbarrier: more bytes completed than expected\n");
        _exit(3);
    }
    if (b.pending == 0 && b.tx == 0) {
        b.phase++;
        b.pending = (int32_t)b.count;
    }
}
static EmBar& embar(uint64_t* bar) {
    auto it = g_bars.find((uintptr_t)bar);
    if (it == g_bars.end()) {
        fprintf(stderr, "emulated mbarrier used before init\n");
        _exit(3);
    }
    return it->second;
}
void z_mbar_init(uint64_t* bar, uint32_t count) {
    pthread_mutex_lock(&g_bar_mu);
    g_bars[(uintptr_t)bar] = EmBar{count, (int32_t)count, 0, 0};
    pthread_mutex_unlock(&g_bar_mu);
}
void z_mbar_fence_init() {}
void z_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    pthread_mutex_lock(&g_bar_mu);
    EmBar& b = embar(bar);
    b.tx += bytes;
    b.pending--;
    embar_check(b);
    pthread_mutex_unlock(&g_bar_mu);
}
void z_mbar_arrive(uint64_t* bar) {
    pthread_mutex_lock(&g_bar_mu);
    EmBar& b = embar(bar);
    b.pending--;
    if (b.pending < 0) {
        fprintf(stderr, "emulated mbarrier: too many arrivals\n");
        _exit(3);
    }
    embar_check(b);
    pthread_mutex_unlock(&g_bar_mu);
}
void z_mbar_wait(uint64_t* bar, uint32_t parity) {
    for (long spins = 0;; ++spins) {
        pthread_mutex_lock(&g_bar_mu);
        const bool done = (embar(bar).phase & 1u) != parity;
        pthread_mutex_unlock(&g_bar_mu);
        if (done) return;
        if (spins > 4000000) {
            fprintf(stderr, "emulated mbarrier: wait timed out (deadlock in the ring protocol)\n");
            _exit(4);
        }
        usleep(20);
    }
}
static size_t g_smem_bytes = 0;
void z_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    // the hardware requires 16-byte aligned addresses and sizes; the destination must lie in the CTA's window
    if (((uintptr_t)dst_smem & 15) || ((uintptr_t)src_gmem & 15) || (bytes & 15)) {
        fprintf(stderr, "bulk copy: misaligned address or size (%p %p %u)\n", dst_smem, src_gmem, bytes);
        _exit(5);
    }
    if ((unsigned char*)dst_smem < g_dyn_smem || (unsigned char*)dst_smem + bytes > g_dyn_smem + g_smem_bytes) {
        fprintf(stderr, "bulk copy: destination outside the dynamic shared memory window\n");
        _exit(5);
    }
    memcpy(dst_smem, src_gmem, bytes);
    pthread_mutex_lock(&g_bar_mu);
    EmBar& b = embar(bar);
    b.tx -= bytes;
    embar_check(b);
    pthread_mutex_unlock(&g_bar_mu);
}
void z_consumer_bar_sync() { pthread_barrier_wait(&g_named_bar); }


// Schedule jitter (EMUL_JITTER=<max microseconds>): random sleeps in front of every release, in a fraction of the
// acquires and at kernel boundaries, drawn per process, so that repeated runs see different interleavings of
// the emulated ranks and CTAs (a rank a whole step ahead of its neighbours, a publisher that is late, ...).
static inline void emul_jitter(int one_in) {
    static int max_us = -1;
    static thread_local unsigned long long st = 0;
    if (max_us < 0) {
        const char* e = getenv("EMUL_JITTER");
        max_us = e ? atoi(e) : 0;
    }
    if (max_us <= 0) return;
    if (st == 0) st = 0x9E3779B97F4A7C15ull * (unsigned long long)(getpid() * 1000003 + (int)threadIdx.x + 1);
    st ^= st << 13;
    st ^= st >> 7;
    st ^= st << 17;
    if ((st >> 20) % (unsigned)one_in == 0) usleep((useconds_t)((st >> 33) % (unsigned)max_us));
}

// peer-exchange helpers (kry_common.cuh, #ifndef KRY_EMUL): the emulated ranks share their memory
static inline void dst_release_sys(unsigned long long* p, unsigned long long v) {
    emul_jitter(1);
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
}
static inline unsigned long long dld_acquire_sys(const unsigned long long* p) {
    emul_jitter(64);
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
static inline unsigned long long dld_volatile_u64(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
static inline double dld_volatile_f64(const double* p) { return *(const volatile double*)p; }
static inline unsigned long long dglobal_timer_ns() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

template <typename T> static T* dev_alloc(size_t count) {
    void* p = mmap(nullptr, count * sizeof(T) + 64, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) {
        perror("mmap");
        exit(2);
    }
    return (T*)p;
}

struct ThreadArg {
    int t;
    const std::function<void()>* body;
};
static void* thread_main(void* p) {
    ThreadArg* a = (ThreadArg*)p;
    threadIdx.x = (unsigned)a->t;
    threadIdx.y = threadIdx.z = 0;
    (*a->body)();
    return nullptr;
}

// R emulated ranks x G CTAs: one process per CTA, one OS thread per CUDA thread; every rank has its own grid
// barrier; body(rank) runs in every thread.  Returns false if any CTA died.
static bool emul_launch_ranks(int R, int G, int nthreads, size_t smem_bytes, const std::function<void(int)>& body) {
    pthread_barrier_t* bars = dev_alloc<pthread_barrier_t>((size_t)R);
    pthread_barrierattr_t ba;
    pthread_barrierattr_init(&ba);
    pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
    for (int r = 0; r < R; ++r) pthread_barrier_init(&bars[r], &ba, (unsigned)G);
    gridDim = dim3(G);
    blockDim = dim3(nthreads);
    std::vector<pid_t> pids;
    for (int r = 0; r < R; ++r) {
        for (int b = 0; b < G; ++b) {
            pid_t pid = fork();
            if (pid == 0) {
                alarm(600);                               // a deadlocked kernel must not hang the test tier
                g_grid_bar = &bars[r];
                blockIdx.x = (unsigned)b;
                blockIdx.y = blockIdx.z = 0;
                pthread_barrier_init(&g_block_bar, nullptr, (unsigned)nthreads);
                pthread_barrier_init(&g_named_bar, nullptr, 256u);
                for (int w = 0; w < (nthreads + 31) / 32; ++w) pthread_barrier_init(&g_warps[w].bar, nullptr, 32u);
                g_smem_bytes = smem_bytes;
                g_dyn_smem = (unsigned char*)aligned_alloc(128, (smem_bytes + 127) / 128 * 128 + 128);
                std::function<void()> fn = [&body, r]() { body(r); };
                std::vector<pthread_t> th(nthreads);
                std::vector<ThreadArg> args(nthreads);
                pthread_attr_t at;
                pthread_attr_init(&at);
                pthread_attr_setstacksize(&at, 1 << 20);
                for (int t = 0; t < nthreads; ++t) {
                    args[t] = ThreadArg{t, &fn};
                    if (pthread_create(&th[t], &at, thread_main, &args[t])) {
                        perror("pthread_create");
                        _exit(6);
                    }
                }
                for (int t = 0; t < nthreads; ++t) pthread_join(th[t], nullptr);
                _exit(0);
            }
            pids.push_back(pid);
        }
    }
    bool ok = true;
    for (pid_t p : pids) {
        int st = 0;
        waitpid(p, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ok = false;
    }
    return ok;
}

static bool emul_launch(int G, int nthreads, size_t smem_bytes, const std::function<void()>& body) {
    return emul_launch_ranks(1, G, nthreads, smem_bytes, [&body](int) { body(); });
}
