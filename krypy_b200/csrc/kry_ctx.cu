// Context, error reporting, pinned mailbox.
#include <stdarg.h>
#include <stdlib.h>
#include "kry_common.cuh"

static thread_local char g_err[512] = "";

void kry_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int kry_orth_max_blocks(int dtype, int* out);   // kry_orth.cu
int kry_proj_max_blocks(int dtype, int* out);   // kry_orth.cu

extern "C" {

int kry_version(void) { return KRY_ABI_VERSION; }

const char* kry_last_error(void) { return g_err; }

int kry_ctx_create(int device, void* stream, kry_ctx** out) {
    KRY_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        kry_set_error("no CUDA device available (%s): krypy_b200 has no CPU fallback",
                      cudaGetErrorString(e));
        return KRY_ERR_CUDA;
    }
    KRY_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    KRY_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    KRY_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        kry_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                      prop.major, prop.minor);
        return KRY_ERR_UNSUPPORTED;
    }
    kry_ctx* c = new kry_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->stream = (cudaStream_t)stream;
    c->sm_count = prop.multiProcessorCount;
    c->cc = prop.major * 10 + prop.minor;
    c->l2_bytes = prop.l2CacheSize;
    c->smem_optin = (long long)prop.sharedMemPerBlockOptin;
    c->coop = prop.cooperativeLaunch;
    size_t pbytes = sizeof(double) * 2ull * KRY_MAX_SLOTS * KRY_MAX_PARTIAL_BLOCKS;
    KRY_CHECK_CUDA(cudaMalloc(&c->d_partials, pbytes));
    KRY_CHECK_CUDA(cudaMemset(c->d_partials, 0, pbytes));
    KRY_CHECK_CUDA(cudaMalloc(&c->d_ticket, 64 * sizeof(unsigned int)));
    KRY_CHECK_CUDA(cudaMemset(c->d_ticket, 0, 64 * sizeof(unsigned int)));
    KRY_CHECK_CUDA(cudaHostAlloc(&c->h_mailbox, sizeof(double) * KRY_MAILBOX_DOUBLES,
                                 cudaHostAllocMapped | cudaHostAllocPortable));
    memset(c->h_mailbox, 0, sizeof(double) * KRY_MAILBOX_DOUBLES);
    KRY_CHECK_CUDA(cudaHostGetDevicePointer(&c->d_mailbox, c->h_mailbox, 0));
    if (!c->coop) {
        kry_set_error("device does not support cooperative launches");
        return KRY_ERR_UNSUPPORTED;
    }
    int rc;
    if ((rc = kry_orth_max_blocks(KRY_F64, &c->orth_blocks_f64))) return rc;
    if ((rc = kry_orth_max_blocks(KRY_F32, &c->orth_blocks_f32))) return rc;
    if ((rc = kry_proj_max_blocks(KRY_F64, &c->proj_blocks_f64))) return rc;
    if ((rc = kry_proj_max_blocks(KRY_F32, &c->proj_blocks_f32))) return rc;
    c->orth_blocks_f64 *= c->sm_count;
    c->orth_blocks_f32 *= c->sm_count;
    c->proj_blocks_f64 *= c->sm_count;
    c->proj_blocks_f32 *= c->sm_count;
    *out = c;
    return KRY_OK;
}

int kry_ctx_destroy(kry_ctx* ctx) {
    if (!ctx) return KRY_OK;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_partials);
    cudaFree(ctx->d_ticket);
    cudaFreeHost(ctx->h_mailbox);
    delete ctx;
    return KRY_OK;
}

int kry_ctx_set_stream(kry_ctx* ctx, void* stream) {
    KRY_REQUIRE(ctx != nullptr, "ctx is NULL");
    ctx->stream = (cudaStream_t)stream;
    return KRY_OK;
}

int kry_device_info(kry_ctx* ctx, long long info[8]) {
    KRY_REQUIRE(ctx != nullptr && info != nullptr, "NULL argument");
    info[0] = ctx->sm_count;
    info[1] = ctx->cc;
    info[2] = ctx->l2_bytes;
    info[3] = ctx->smem_optin;
    info[4] = ctx->coop;
    info[5] = ctx->orth_blocks_f64;
    info[6] = ctx->proj_blocks_f64;
    info[7] = 0;
    return KRY_OK;
}

// L2 residency window on the context's stream (cudaStreamAttributeAccessPolicyWindow): accesses of kernels
// launched afterwards to [base, base + bytes) are "persisting" (kept in the L2 set-aside), the rest of the window
// beyond the set-aside "streaming".  base == NULL or bytes <= 0 removes the window and demotes the persisting lines.
// info: [0] persistingL2CacheMaxSize, [1] accessPolicyMaxWindowSize, [2] set-aside in effect, [3] window bytes,
// [4] hit ratio * 1e6.  Returns KRY_ERR_UNSUPPORTED (not an error of the solve) when the device or the stream
// does not take the hint.
// (a failing hint must not poison the error state the next launch check reads: clear it)
#define KRY_L2_TRY(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            kry_set_error("kry_l2_window: %s failed: %s", #expr, cudaGetErrorString(_e));      \
            (void)cudaGetLastError();                                                         \
            return KRY_ERR_UNSUPPORTED;                                                       \
        }                                                                                     \
    } while (0)

static long long carve_set[16] = {0};       // per device: the set-aside limit is only touched when it changes

int kry_l2_window(kry_ctx* ctx, const void* base, long long bytes, long long info[5]) {
    KRY_REQUIRE(ctx != nullptr, "ctx is NULL");
    KRY_CHECK_CUDA(cudaSetDevice(ctx->device));
    int max_persist = 0, max_window = 0;
    KRY_L2_TRY(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device));
    KRY_L2_TRY(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device));
    if (info) {
        info[0] = max_persist;
        info[1] = max_window;
        info[2] = info[3] = info[4] = 0;
    }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (base == nullptr || bytes <= 0) {
        attr.accessPolicyWindow.base_ptr = nullptr;
        attr.accessPolicyWindow.num_bytes = 0;
        attr.accessPolicyWindow.hitRatio = 0.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        KRY_L2_TRY(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        KRY_L2_TRY(cudaCtxResetPersistingL2Cache());
        // give the set-aside back: measured on a B200, an unused set-aside is NOT available to normal accesses
        // (C2 with an idle 80 MB set-aside: Gram-Schmidt kernel 509 instead of 443 us, SpMV 162 instead of 138 us)
        if (carve_set[ctx->device & 15] != 0) {
            KRY_L2_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
            carve_set[ctx->device & 15] = 0;
        }
        return KRY_OK;
    }
    if (max_persist <= 0 || max_window <= 0) {
        kry_set_error("kry_l2_window: the device has no persisting L2 set-aside");
        return KRY_ERR_UNSUPPORTED;
    }
    const long long carve = bytes < (long long)max_persist ? bytes : (long long)max_persist;
    if (carve_set[ctx->device & 15] != carve) {
        KRY_L2_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)carve));
        carve_set[ctx->device & 15] = carve;
    }
    size_t got = 0;
    KRY_L2_TRY(cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize));
    const long long win = bytes < (long long)max_window ? bytes : (long long)max_window;
    double ratio = win > 0 ? (double)got / (double)win : 0.0;
    if (ratio > 1.0) ratio = 1.0;
    if (const char* e = getenv("KRY_L2_WINDOW_RATIO")) {   // measurement knob: fraction of the window that persists
        const double r = atof(e);
        if (r > 0.0 && r < ratio) ratio = r;
    }
    attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    attr.accessPolicyWindow.num_bytes = (size_t)win;
    attr.accessPolicyWindow.hitRatio = (float)ratio;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    KRY_L2_TRY(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    if (info) {
        info[2] = (long long)got;
        info[3] = win;
        info[4] = (long long)(ratio * 1e6);
    }
    return KRY_OK;
}

double* kry_mailbox_host(kry_ctx* ctx) { return ctx ? ctx->h_mailbox : nullptr; }
double* kry_mailbox_dev(kry_ctx* ctx) { return ctx ? ctx->d_mailbox : nullptr; }

int kry_sync(kry_ctx* ctx) {
    KRY_REQUIRE(ctx != nullptr, "ctx is NULL");
    KRY_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return KRY_OK;
}

long long kry_launch_count(kry_ctx* ctx) { return ctx ? ctx->launches : -1; }
void kry_reset_launch_count(kry_ctx* ctx) {
    if (ctx) ctx->launches = 0;
}

}  // extern "C"
