// runtime.cu -- process-wide runtime + per-thread contexts + the MoB200_* residency extension.
//
// Replaces the reference CUDA shim's "global CUcontext at static-init, NULL stream, alloc/copy/free per call"
// (cgo/cuda/cuda.cpp:63-219) with: lazy init, one stream + scratch arena per calling thread, pointer
// classification so resident (device) columns are used in place, and pinned host allocation for the caller.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace mob {

std::atomic<uint64_t> g_launches{0};
static std::once_flag g_once;
static int g_device = -1;
static int g_init_rc = MO_RC_INTERNAL_ERROR;
static int g_sms = kSMs;
static char g_init_err[256] = "runtime not initialised";

void set_error(const char *fmt, ...) {
    // the thread context may not exist yet (init failure): keep a thread-local buffer of its own
    static thread_local char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    ThreadCtx &t = tctx();
    memcpy(t.err, buf, sizeof buf);
}

int num_sms() { return g_sms; }

int runtime_init(int device) {
    std::call_once(g_once, [device]() {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            snprintf(g_init_err, sizeof g_init_err, "libmo_b200: no CUDA device available (%s); this library has no CPU path",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
            return;
        }
        int d = device;
        if (d < 0) {
            const char *s = getenv("MO_B200_DEVICE");
            if (!s) s = getenv("LOCAL_RANK");
            d = s ? atoi(s) : 0;
        }
        if (d >= n) d = d % n;
        e = cudaSetDevice(d);
        if (e != cudaSuccess) {
            snprintf(g_init_err, sizeof g_init_err, "cudaSetDevice(%d) failed: %s", d, cudaGetErrorString(e));
            return;
        }
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess) g_sms = p.multiProcessorCount;
        g_device = d;
        g_init_rc = MO_RC_SUCCESS;
        g_init_err[0] = 0;
    });
    return g_init_rc;
}

ThreadCtx &tctx() {
    static thread_local ThreadCtx t;
    if (!t.ready) {
        if (runtime_init(-1) != MO_RC_SUCCESS) {
            snprintf(t.err, sizeof t.err, "%s", g_init_err);
            return t;
        }
        if (cudaSetDevice(g_device) != cudaSuccess) { snprintf(t.err, sizeof t.err, "cudaSetDevice failed"); return t; }
        if (cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&t.ev0) != cudaSuccess || cudaEventCreate(&t.ev1) != cudaSuccess ||
            cudaEventCreate(&t.kev0) != cudaSuccess || cudaEventCreate(&t.kev1) != cudaSuccess) {
            snprintf(t.err, sizeof t.err, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            return t;
        }
        t.pinned_sz = 1 << 16;
        if (cudaHostAlloc((void **)&t.pinned, t.pinned_sz, cudaHostAllocDefault) != cudaSuccess) {
            snprintf(t.err, sizeof t.err, "pinned staging allocation failed");
            return t;
        }
        if (cudaMalloc((void **)&t.ctrl, 256) != cudaSuccess || cudaMemset(t.ctrl, 0, 256) != cudaSuccess) {
            snprintf(t.err, sizeof t.err, "control block allocation failed");
            return t;
        }
        t.own_stream = true;
        t.ready = true;
    }
    return t;
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

void *arena_alloc(ThreadCtx &t, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    while (t.cur_block < t.blocks.size()) {
        ArenaBlock &b = t.blocks[t.cur_block];
        if (t.cur_off + bytes <= b.size) { void *p = b.p + t.cur_off; t.cur_off += bytes; return p; }
        t.cur_block++; t.cur_off = 0;
    }
    size_t sz = bytes > (size_t)(64u << 20) ? bytes : (size_t)(64u << 20);
    char *p = nullptr;
    cudaError_t e = cudaMalloc((void **)&p, sz);
    if (e != cudaSuccess) { set_error("device scratch allocation of %zu bytes failed: %s", sz, cudaGetErrorString(e)); return nullptr; }
    t.blocks.push_back({p, sz});
    t.cur_block = t.blocks.size() - 1;
    t.cur_off = bytes;
    return p;
}

void arena_reset(ThreadCtx &t) {
    // consolidate a fragmented arena into one block so the steady state is a single bump allocator
    if (t.blocks.size() > 1) {
        size_t total = 0;
        for (auto &b : t.blocks) { total += b.size; cudaFree(b.p); }
        t.blocks.clear();
        char *p = nullptr;
        if (cudaMalloc((void **)&p, total) == cudaSuccess) t.blocks.push_back({p, total});
        else cudaGetLastError();
    }
    t.cur_block = 0; t.cur_off = 0; t.arena_epoch++; t.kev_prio = 0;
}

const void *Stager::in(const void *p, size_t bytes) {
    if (!p || bytes == 0) return p;
    if (is_device_ptr(p)) return p;
    void *d = arena_alloc(t, bytes);
    if (!d) { failed = true; return nullptr; }
    cudaError_t e = cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, t.stream);
    if (e != cudaSuccess) { set_error("H2D copy of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); failed = true; return nullptr; }
    return d;
}

void *Stager::out(void *p, size_t bytes, bool preload) {
    if (!p || bytes == 0) return p;
    if (is_device_ptr(p)) return p;
    void *d = arena_alloc(t, bytes);
    if (!d) { failed = true; return nullptr; }
    if (preload) {
        cudaError_t e = cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, t.stream);
        if (e != cudaSuccess) { set_error("H2D preload failed: %s", cudaGetErrorString(e)); failed = true; return nullptr; }
    }
    backs.push_back({p, d, bytes});
    return d;
}

int Stager::finish() {
    int rc = failed ? MO_RC_INTERNAL_ERROR : MO_RC_SUCCESS;
    if (!failed) {
        for (auto &b : backs) {
            cudaError_t e = cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, t.stream);
            if (e != cudaSuccess) { set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = MO_RC_INTERNAL_ERROR; break; }
        }
    }
    cudaError_t e = cudaStreamSynchronize(t.stream);
    if (e != cudaSuccess) { set_error("stream synchronize failed: %s", cudaGetErrorString(e)); rc = MO_RC_INTERNAL_ERROR; }
    backs.clear();
    arena_reset(t);
    return rc;
}

Stager::~Stager() {
    if (!backs.empty()) { cudaStreamSynchronize(t.stream); arena_reset(t); }
}

int read_back(ThreadCtx &t, void *host_dst, const void *dev_src, size_t bytes) {
    if (bytes > t.pinned_sz) { set_error("read_back too large"); return MO_RC_INTERNAL_ERROR; }
    MOB_CUDA_TRY(cudaMemcpyAsync(t.pinned, dev_src, bytes, cudaMemcpyDeviceToHost, t.stream));
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    memcpy(host_dst, t.pinned, bytes);
    return MO_RC_SUCCESS;
}

}  // namespace mob

using namespace mob;

#define REQUIRE_CTX(t)                               \
    ThreadCtx &t = tctx();                           \
    if (!t.ready) return MO_RC_INTERNAL_ERROR

__global__ void flush_l2_kernel(int4 *p, size_t n16) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) p[i] = make_int4((int)i, 0, 0, 0);
}

extern "C" {

int32_t MoB200_Init(int32_t device) {
    int rc = runtime_init(device);
    if (rc != MO_RC_SUCCESS) return rc;
    REQUIRE_CTX(t);
    return MO_RC_SUCCESS;
}

const char *MoB200_Version(void) { return "mo_b200 0.1 (sm_100a)"; }

int32_t MoB200_DeviceCount(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int32_t MoB200_DeviceAlloc(uint64_t bytes, void **dptr) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaMalloc(dptr, bytes ? bytes : 1));
    return MO_RC_SUCCESS;
}
int32_t MoB200_DeviceFree(void *dptr) {
    REQUIRE_CTX(t);
    MoB200_SearchRelease(dptr);   // a prepared search operand of (or built against) this buffer must not outlive it
    MOB_CUDA_TRY(cudaFree(dptr));
    return MO_RC_SUCCESS;
}
int32_t MoB200_HostAlloc(uint64_t bytes, void **hptr) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return MO_RC_SUCCESS;
}
int32_t MoB200_HostFree(void *hptr) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaFreeHost(hptr));
    return MO_RC_SUCCESS;
}
int32_t MoB200_HostRegister(void *hptr, uint64_t bytes) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaHostRegister(hptr, bytes, cudaHostRegisterDefault));
    return MO_RC_SUCCESS;
}
int32_t MoB200_HostUnregister(void *hptr) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaHostUnregister(hptr));
    return MO_RC_SUCCESS;
}
int32_t MoB200_Upload(void *dst_dev, const void *src_host, uint64_t bytes) {
    REQUIRE_CTX(t);
    search_invalidate(dst_dev, bytes);
    MOB_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, t.stream));
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    return MO_RC_SUCCESS;
}
int32_t MoB200_Download(void *dst_host, const void *src_dev, uint64_t bytes) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, t.stream));
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    return MO_RC_SUCCESS;
}
// stream-ordered copies without a synchronise: the host buffer should be pinned (MoB200_HostAlloc / HostRegister) and must not be
// read (download) or reused (upload) before MoB200_Sync
int32_t MoB200_DownloadAsync(void *dst_host, const void *src_dev, uint64_t bytes) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, t.stream));
    return MO_RC_SUCCESS;
}
int32_t MoB200_UploadAsync(void *dst_dev, const void *src_host, uint64_t bytes) {
    REQUIRE_CTX(t);
    search_invalidate(dst_dev, bytes);
    MOB_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, t.stream));
    return MO_RC_SUCCESS;
}
int32_t MoB200_Memset(void *dst_dev, int32_t value, uint64_t bytes) {
    REQUIRE_CTX(t);
    search_invalidate(dst_dev, bytes);
    MOB_CUDA_TRY(cudaMemsetAsync(dst_dev, value, bytes, t.stream));
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    return MO_RC_SUCCESS;
}
int32_t MoB200_Sync(void) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    return MO_RC_SUCCESS;
}
int32_t MoB200_SetStream(void *cuda_stream) {
    REQUIRE_CTX(t);
    static thread_local cudaStream_t own = nullptr;
    if (t.own_stream) own = t.stream;
    if (cuda_stream) { t.stream = (cudaStream_t)cuda_stream; t.own_stream = false; }
    else { t.stream = own; t.own_stream = true; }
    return MO_RC_SUCCESS;
}
int32_t MoB200_TimerStart(void) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaEventRecord(t.ev0, t.stream));
    return MO_RC_SUCCESS;
}
int32_t MoB200_TimerStop(float *ms) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaEventRecord(t.ev1, t.stream));
    MOB_CUDA_TRY(cudaEventSynchronize(t.ev1));
    MOB_CUDA_TRY(cudaEventElapsedTime(ms, t.ev0, t.ev1));
    return MO_RC_SUCCESS;
}
uint64_t MoB200_KernelLaunchCount(void) { return g_launches.load(); }

int32_t MoB200_LastKernelMs(float *ms) {
    REQUIRE_CTX(t);
    MOB_CUDA_TRY(cudaEventSynchronize(t.kev1));
    MOB_CUDA_TRY(cudaEventElapsedTime(ms, t.kev0, t.kev1));
    return MO_RC_SUCCESS;
}

int32_t MoB200_LastError(char *buf, uint64_t buflen) {
    ThreadCtx &t = tctx();
    if (buf && buflen) { snprintf(buf, buflen, "%s", t.err); }
    return (int32_t)strlen(t.err);
}

int32_t MoB200_FlushL2(void) {
    REQUIRE_CTX(t);
    static thread_local void *buf = nullptr;
    const size_t bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    if (!buf) MOB_CUDA_TRY(cudaMalloc(&buf, bytes));
    flush_l2_kernel<<<num_sms() * 4, 256, 0, t.stream>>>((int4 *)buf, bytes / 16);
    MOB_LAUNCH_CHECK();
    return MO_RC_SUCCESS;
}

}  // extern "C"
