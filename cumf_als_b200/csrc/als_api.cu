// als_api.cu -- host side of the B200 ALS hot path: work plans, the resident solver
// handle, the doALS driver and the C ABI declared in include/cumf_als.h.
//
// Mirrors the reference's driver doALS (als.cu:662-1035) for the path
//   per-row Gram (+RHS) formation -> batched f x f solve -> RMSE observable,
// re-designed for one B200: CSR/CSC/COO are uploaded once and stay resident (the
// reference re-uploads CSR every iteration, als.cu:734-739), the X_BATCH /
// THETA_BATCH loop that bounded a 12 GB card's Gram buffer (als.cu:768-777) is
// replaced by an internal workspace cap (unfused path) or by never materialising
// A at all (fused path), and long rows are split deterministically across CTAs.
#include <cublas_v2.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace cumf {

// ---------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

// OPT-IN buffer cache.  By default (CUMF_CACHE_MB unset or 0) every device buffer of a solver is cudaFree'd when the
// solver is destroyed, i.e. doALS returns with all its device memory released, like the reference (als.cu:1026-1033).
// With CUMF_CACHE_MB=<n> up to n MiB of a destroyed solver's buffers are kept for the next one (a caller that runs doALS
// repeatedly on inputs of the same shape saves the cudaMalloc/cudaFree round trips); only buffers that went through
// release_to_cache() -- whose work is known to be complete -- are reused, cumf_release_cached_memory() returns them.
namespace {
struct CachedBuf { void* p; size_t bytes; int device; };
std::vector<CachedBuf> g_buf_cache;
size_t g_buf_cache_bytes = 0;
std::mutex g_buf_cache_mutex;
constexpr size_t kCacheMinBytes = 256;     // (the 16-byte scalars are not worth a list entry)
}  // namespace

extern "C" int cumf_release_cached_memory(void);

thread_local DevArena* t_arena = nullptr;
// set by cumf_group_create* around the creation of its shards: their factor replicas and barrier flags are carved out of the
// shard's arena too (a stand-alone solver keeps them as allocations of their own: cumf_als_ipc_export hands them to peers)
static thread_local bool t_group_member = false;

int DevBuf::alloc_own(size_t n) {
    DevArena* keep = t_arena;
    t_arena = nullptr;
    const int rc = alloc(n);
    t_arena = keep;
    return rc;
}

namespace {
int* g_pinned_ints = nullptr;
int g_pinned_next = 0;
constexpr int kPinnedInts = 1024;
}  // namespace
int* DevBuf::pinned_int() {
    std::lock_guard<std::mutex> lock(g_buf_cache_mutex);
    unsigned int flags = 0;
    if (g_pinned_ints && cudaHostGetFlags(&flags, g_pinned_ints) != cudaSuccess) {      // the context was reset under us
        cudaGetLastError();
        g_pinned_ints = nullptr;
    }
    if (!g_pinned_ints) {
        if (cudaHostAlloc(reinterpret_cast<void**>(&g_pinned_ints), sizeof(int) * kPinnedInts, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            g_pinned_ints = nullptr;
            return nullptr;
        }
        g_pinned_next = 0;
    }
    int* p = g_pinned_ints + (g_pinned_next++ % kPinnedInts);
    *p = 0;
    return p;
}

int DevBuf::alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    if (t_arena && t_arena->base) {
        const size_t off = (t_arena->used + 1023) & ~(size_t)1023;
        if (off + n <= t_arena->cap) {
            p = t_arena->base + off;
            bytes = n;
            in_arena = true;
            t_arena->used = off + n;
            return CUMF_OK;
        }
        t_arena->overflow_bytes += n;
        t_arena->overflow_count += 1;
    }
    if (n >= kCacheMinBytes) {
        std::lock_guard<std::mutex> lock(g_buf_cache_mutex);
        int dev = 0;
        cudaGetDevice(&dev);
        size_t best = g_buf_cache.size();
        for (size_t i = 0; i < g_buf_cache.size(); ++i) {
            const CachedBuf& c = g_buf_cache[i];
            if (c.device == dev && c.bytes >= n && c.bytes <= n + n / 4 && (best == g_buf_cache.size() || c.bytes < g_buf_cache[best].bytes)) best = i;
        }
        if (best < g_buf_cache.size()) {
            p = g_buf_cache[best].p;
            bytes = g_buf_cache[best].bytes;
            g_buf_cache_bytes -= bytes;
            g_buf_cache.erase(g_buf_cache.begin() + (long)best);
            return CUMF_OK;
        }
    }
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess && g_buf_cache_bytes > 0) {    // make room: drop what is cached and retry once
        cudaGetLastError();
        cumf_release_cached_memory();
        e = cudaMalloc(&p, n);
    }
    if (e != cudaSuccess) {
        p = nullptr;
        set_last_error(std::string("cudaMalloc(") + std::to_string(n) + "): " + cudaGetErrorString(e));
        return CUMF_ECUDA;
    }
    bytes = n;
    return CUMF_OK;
}
void DevBuf::release() {
    if (p && !borrowed && !in_arena) cudaFree(p);
    p = nullptr;
    bytes = 0;
    borrowed = false;
    in_arena = false;
}
// the caller guarantees that no work touching the buffer is in flight
void DevBuf::release_to_cache() {
    const char* v = getenv("CUMF_CACHE_MB");
    const size_t cap = (size_t)((v && *v) ? std::max(0L, atol(v)) : 0) << 20;
    std::unique_lock<std::mutex> lock(g_buf_cache_mutex);
    if (p && !borrowed && !in_arena && bytes >= kCacheMinBytes && g_buf_cache_bytes + bytes <= cap) {
        int dev = 0;
        cudaGetDevice(&dev);
        g_buf_cache.push_back(CachedBuf{p, bytes, dev});
        g_buf_cache_bytes += bytes;
        p = nullptr;
        bytes = 0;
        return;
    }
    lock.unlock();
    release();
}
// ---- pinned staging arenas + copy kernel (see common.cuh) --------------------------------------------
namespace {
constexpr int kMaxStageDevices = 16;
struct Stage {
    unsigned char* p = nullptr;
    size_t bytes = 0, used = 0;
    std::vector<unsigned char*> retired;     // arenas outgrown while copies from them may still be queued
    std::mutex mutex;
};
Stage g_stages[kMaxStageDevices];            // one per device: the shards of a group build and upload their plans in parallel
thread_local Stage* t_stage = nullptr;       // the arena of the StagingSection this thread holds
__global__ void copy_words_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t words) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
Stage& stage_of_current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return g_stages[(dev >= 0 && dev < kMaxStageDevices) ? dev : 0];
}
}  // namespace

// A cudaDeviceReset by the caller (the reference CLI ends with one, main.cpp:168; die() below) invalidates the arena:
// probe it before reuse and drop a dangling pointer instead of writing through it.
static bool stage_alive(Stage& g) {
    if (!g.p) return false;
    unsigned int flags = 0;
    if (cudaHostGetFlags(&flags, g.p) == cudaSuccess) return true;
    cudaGetLastError();
    g.p = nullptr;
    g.bytes = g.used = 0;
    g.retired.clear();
    return false;
}

static void stage_reset(Stage& g) {
    g.used = 0;
    if (!stage_alive(g)) return;
    for (unsigned char* p : g.retired) cudaFreeHost(p);
    g.retired.clear();
}
void staging_reset() { stage_reset(t_stage ? *t_stage : stage_of_current_device()); }

StagingSection::StagingSection() {
    Stage& g = stage_of_current_device();
    g.mutex.lock();
    t_stage = &g;
    held = true;
    stage_reset(g);
}
int StagingSection::finish(cudaStream_t st) {
    if (!held) return CUMF_OK;
    const cudaError_t e = cudaStreamSynchronize(st);
    held = false;
    Stage* g = t_stage;
    t_stage = nullptr;
    if (g) g->mutex.unlock();
    if (e != cudaSuccess) {
        set_last_error(std::string("plan upload: ") + cudaGetErrorString(e));
        return CUMF_ECUDA;
    }
    return CUMF_OK;
}
StagingSection::~StagingSection() {
    if (held) {
        cudaStreamSynchronize(last);
        Stage* g = t_stage;
        t_stage = nullptr;
        if (g) g->mutex.unlock();
    }
}

static void staging_release() {
    for (Stage& g : g_stages) {
        std::lock_guard<std::mutex> lock(g.mutex);
        if (stage_alive(g)) {
            for (unsigned char* p : g.retired) cudaFreeHost(p);
            cudaFreeHost(g.p);
        }
        g.retired.clear();
        g.p = nullptr;
        g.bytes = g.used = 0;
    }
}

int upload_via_kernel(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return CUMF_OK;
    CUMF_REQUIRE((bytes & 3u) == 0, "upload_via_kernel: size must be a multiple of 4");
    Stage& g = t_stage ? *t_stage : stage_of_current_device();
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (!stage_alive(g) || g.used + need > g.bytes) {
        const size_t grow = std::max<size_t>(std::max<size_t>(g.bytes * 2, (size_t)16 << 20), g.used + need);
        unsigned char* fresh = nullptr;
        if (cudaHostAlloc(reinterpret_cast<void**>(&fresh), grow, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            // no pinned memory to be had: fall back to the (blocking) copy engine path
            CUMF_CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
            return CUMF_OK;
        }
        if (g.p) g.retired.push_back(g.p);     // queued copies may still read it: freed at the next reset
        g.p = fresh;
        g.bytes = grow;
        g.used = 0;
    }
    unsigned char* slot = g.p + g.used;
    g.used += need;
    memcpy(slot, h_src, bytes);
    const size_t words = bytes / 4;
    const int blocks = (int)std::min<size_t>((words + 255) / 256, 592);
    copy_words_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<uint32_t*>(d_dst), reinterpret_cast<const uint32_t*>(slot), words);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

static double wall_seconds() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static long env_long(const char* name, long dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return atol(v);
}
static int env_choice(const char* name, const char* a, int va, const char* b, int vb, const char* c, int vc, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    if (a && !strcasecmp(v, a)) return va;
    if (b && !strcasecmp(v, b)) return vb;
    if (c && !strcasecmp(v, c)) return vc;
    return dflt;
}

static int check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_last_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
        return CUMF_ENOGPU;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        set_last_error("this library contains sm_100a code only; found compute capability major " +
                       std::to_string(major));
        return CUMF_ENOGPU;
    }
    return CUMF_OK;
}

static int check_f(int f) {
    // main.cpp:33-36 rejects f % 10 != 0; the kernels additionally view rows as float2 (als.cu:805)
    CUMF_REQUIRE(f >= 10 && f <= 200 && f % 10 == 0, "f must be a multiple of 10 in [10, 200]");
    return CUMF_OK;
}

// ---------------------------------------------------------------------------------
// LU oracle mode: cuBLAS batched LU without pivoting (als.cu:58-122).
// ---------------------------------------------------------------------------------
namespace {
__global__ void fill_ptrs_kernel(float** Ap, float** bp, float* A, float* b, int batch, int f) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < batch) {
        Ap[k] = A + (size_t)k * f * f;   // devPtrTTHost[k] = &tt[k*f*f]            (als.cu:70-74)
        bp[k] = b + (size_t)k * f;       // devPtrYthetaTHost[k] = &ythetaT[... k*f] (als.cu:91-95)
    }
}
struct CublasHandle {      // created per call on the current device (oracle mode only: not a hot path), destroyed on every exit path
    cublasHandle_t h = nullptr;
    ~CublasHandle() { if (h) cublasDestroy(h); }
};
}  // namespace

// returns everything this library keeps between calls to the driver: cached device buffers and the pinned staging arena
extern "C" int cumf_release_cached_memory(void) {
    {
        std::lock_guard<std::mutex> lock(g_buf_cache_mutex);
        int dev = 0;
        cudaGetDevice(&dev);
        for (auto& c : g_buf_cache) {
            if (dev != c.device) cudaSetDevice(c.device);
            cudaFree(c.p);
            if (dev != c.device) cudaSetDevice(dev);
        }
        g_buf_cache.clear();
        g_buf_cache_bytes = 0;
    }
    staging_release();
    return CUMF_OK;
}

int launch_lu(float* d_A, float* d_x, float* d_b, int batch, int f, cudaStream_t st) {
    if (batch <= 0) return CUMF_OK;
    CublasHandle cb;
    if (cublasCreate(&cb.h) != CUBLAS_STATUS_SUCCESS) {
        set_last_error("cublasCreate failed");
        return CUMF_ECUDA;
    }
    cublasHandle_t g_cublas = cb.h;
    cublasSetStream(g_cublas, st);
    DevBuf ptrs, info;
    CUMF_TRY(ptrs.alloc(sizeof(float*) * 2 * (size_t)batch));
    CUMF_TRY(info.alloc(sizeof(int) * (size_t)batch));
    float** Ap = ptrs.as<float*>();
    float** bp = Ap + batch;
    fill_ptrs_kernel<<<(batch + 255) / 256, 256, 0, st>>>(Ap, bp, d_A, d_b, batch, f);
    CUMF_CUDA_TRY(cudaGetLastError());
    // A is symmetric, so its row-major storage is also its column-major storage.
    if (cublasSgetrfBatched(g_cublas, f, Ap, f, nullptr, info.as<int>(), batch) != CUBLAS_STATUS_SUCCESS) {
        set_last_error("cublasSgetrfBatched failed");
        return CUMF_ECUDA;
    }
    int info2 = 0;
    if (cublasSgetrsBatched(g_cublas, CUBLAS_OP_N, f, 1, (const float**)Ap, f, nullptr, bp, f, &info2, batch) !=
        CUBLAS_STATUS_SUCCESS) {
        set_last_error("cublasSgetrsBatched failed");
        return CUMF_ECUDA;
    }
    // als.cu:108: the solution (left in the RHS) is copied into the factor
    CUMF_CUDA_TRY(cudaMemcpyAsync(d_x, d_b, sizeof(float) * (size_t)batch * f, cudaMemcpyDeviceToDevice, st));
    CUMF_CUDA_TRY(cudaStreamSynchronize(st));
    return CUMF_OK;
}

}  // namespace cumf

using namespace cumf;

// ---------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------
struct cumf_plan {
    int rows = 0, row_begin = 0, row_end = 0, f = 0, path = CUMF_PATH_SIMT;
    long long base = 0;                 // h_rowptr[row_begin]: chunk offsets are relative to it
    std::vector<Chunk> chunks;
    std::vector<SplitRow> splits;
    std::vector<int> row_chunk_ptr;     // first chunk of each owned row (+1 sentinel)
    DevBuf d_chunks, d_splits;
    DevBuf scratchA, scratchB;          // split partials
    DevBuf tt, rhs;                     // materialised batch workspace (unfused path)
    int batch_rows = 0;
    int last_launches = 0;
    TcWork* tc = nullptr;
    // Optional by-product of a fused CG half-step: per solver warpgroup (2 x CTAs) and per split row the sum of
    // x^T b + x^T r + reg x^T x, from which the squared error of the rows' ratings follows (gram_tc.cu).  Enabled by
    // the resident solver for the train RMSE; valid until the opposing factor changes.
    bool collect_sse = false, sse_terms_valid = false;
    DevBuf sse_terms;
    int sse_terms_count = 0;
    // optional timing of the dominant (Gram) kernel
    bool time_kernel = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kernel_events;
    double kernel_ms_total = 0.0;
};

// cache = true only when the device is known to be idle (cumf_als_destroy): the buffers are kept for the next plan
static void plan_free(cumf_plan* p, bool cache = false) {
    if (!p) return;
    auto rel = [cache](DevBuf& b) { if (cache) b.release_to_cache(); else b.release(); };
    rel(p->d_chunks); rel(p->d_splits);
    rel(p->scratchA); rel(p->scratchB); rel(p->tt); rel(p->rhs); rel(p->sse_terms);
    if (p->tc) tc_plan_destroy(p->tc, cache);
    for (auto& e : p->kernel_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    delete p;
}

// force_slots: treat every row as "split" (each chunk stores its partial [A|b]); used by cumf_gram
// to materialise A through the fused kernel.
// The rating range of row r is [h_begin[r], h_end[r]) (absolute positions in colidx/val).  With a CSR row
// pointer that is (rowptr[r], rowptr[r+1]); the partial-Gram scheme passes narrower per-row ranges.
// h_begin / h_end are indexed by (row - index_base): callers that hold pointers for their own rows only (a shard of a
// 50 M-row matrix) do not have to materialise arrays over all rows.
static int plan_create_core(cumf_plan** out, const long long* h_begin_in, const long long* h_end_in, int rows, int row_begin,
                            int row_end, int f, int path, bool alloc_workspace, bool force_slots, int index_base = 0) {
    const long long* h_begin = h_begin_in ? h_begin_in - index_base : nullptr;
    const long long* h_end = h_end_in ? h_end_in - index_base : nullptr;
    CUMF_REQUIRE(out && h_begin_in && h_end_in, "null pointer");
    CUMF_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= rows, "bad row range");
    CUMF_TRY(check_f(f));
    CUMF_TRY(check_device());
    cumf_plan* p = new cumf_plan();
    p->rows = rows; p->row_begin = row_begin; p->row_end = row_end; p->f = f;
    if (path == CUMF_PATH_AUTO) path = tc_path_supports(f) ? CUMF_PATH_TC : CUMF_PATH_SIMT;
    if (path == CUMF_PATH_TC && !tc_path_supports(f)) {
        delete p;
        set_last_error("the fused tcgen05 path does not handle f = " + std::to_string(f));
        return CUMF_EUNSUPPORTED;
    }
    p->path = path;
    // chunk offsets are relative to the smallest position the plan touches
    p->base = 0;
    for (int r = row_begin; r < row_end; ++r)
        if (h_end[r] > h_begin[r]) { p->base = h_begin[r]; break; }
    for (int r = row_begin; r < row_end; ++r)
        if (h_end[r] > h_begin[r]) p->base = std::min<long long>(p->base, h_begin[r]);

    const int max_chunk = (int)env_long("CUMF_SPLIT_NNZ", path == CUMF_PATH_TC ? 8192 : 4096);
    const int owned = row_end - row_begin;
    p->row_chunk_ptr.resize(owned + 1);
    int slot = 0;
    for (int r = row_begin; r < row_end; ++r) {
        p->row_chunk_ptr[r - row_begin] = (int)p->chunks.size();
        const long long n = h_end[r] - h_begin[r];
        // an empty row keeps a (0,0) chunk so that it still gets its lambda*0 system / zero partial
        const long long s = n > 0 ? h_begin[r] - p->base : 0, e = s + n;
        if (n < 0 || e > 0x7fffffffLL) {
            plan_free(p);
            set_last_error("row pointers must be non-decreasing and a shard must hold < 2^31 ratings");
            return CUMF_EINVAL;
        }
        if (n <= max_chunk && !force_slots) {
            p->chunks.push_back(Chunk{r, (int)s, (int)e, -1});
        } else if (n <= max_chunk) {
            p->splits.push_back(SplitRow{r, slot, 1, (int)n});
            p->chunks.push_back(Chunk{r, (int)s, (int)e, slot++});
        } else {
            const int parts = (int)((n + max_chunk - 1) / max_chunk);
            const long long per = (n + parts - 1) / parts;
            p->splits.push_back(SplitRow{r, slot, parts, (int)n});
            for (int q = 0; q < parts; ++q) {
                const long long b = s + q * per, en = std::min(e, b + per);
                p->chunks.push_back(Chunk{r, (int)b, (int)en, slot++});
            }
        }
    }
    p->row_chunk_ptr[owned] = (int)p->chunks.size();

    int rc = p->d_chunks.alloc(sizeof(Chunk) * std::max<size_t>(1, p->chunks.size()));
    if (rc == CUMF_OK) rc = p->d_splits.alloc(sizeof(SplitRow) * std::max<size_t>(1, p->splits.size()));
    if (rc == CUMF_OK) {
        // plan metadata goes to the device through the pinned staging arena (process-wide, one section at a time)
        StagingSection sec;
        if (!p->chunks.empty())
            rc = upload_via_kernel(p->d_chunks.p, p->chunks.data(), sizeof(Chunk) * p->chunks.size(), 0);
        if (rc == CUMF_OK && !p->splits.empty())
            rc = upload_via_kernel(p->d_splits.p, p->splits.data(), sizeof(SplitRow) * p->splits.size(), 0);
        const int rc2 = sec.finish(0);
        if (rc == CUMF_OK) rc = rc2;
    }
    const size_t ff = (size_t)f * f;
    if (rc == CUMF_OK && slot > 0) {
        rc = p->scratchA.alloc(sizeof(float) * ff * slot);
        if (rc == CUMF_OK) rc = p->scratchB.alloc(sizeof(float) * (size_t)f * slot);
    }
    if (rc == CUMF_OK && path == CUMF_PATH_SIMT && alloc_workspace) {
        const size_t cap = (size_t)env_long("CUMF_WORKSPACE_MB", 8192) << 20;
        long long br = (long long)(cap / (ff * sizeof(float)));
        if (br < 1) br = 1;
        p->batch_rows = (int)std::min<long long>(br, std::max(1, owned));
        rc = p->tt.alloc(sizeof(float) * ff * p->batch_rows);
        if (rc == CUMF_OK) rc = p->rhs.alloc(sizeof(float) * (size_t)f * p->batch_rows);
    }
    if (rc == CUMF_OK && path == CUMF_PATH_TC) {
        rc = tc_plan_create(&p->tc, p->chunks, p->d_chunks.as<Chunk>(), p->splits, owned, f);
        // rows split across CTAs are reduced and solved through the materialised path
        if (rc == CUMF_OK && !p->splits.empty() && alloc_workspace) {
            p->batch_rows = (int)p->splits.size();
            rc = p->tt.alloc(sizeof(float) * ff * p->batch_rows);
            if (rc == CUMF_OK) rc = p->rhs.alloc(sizeof(float) * (size_t)f * p->batch_rows);
        }
    }
    if (rc != CUMF_OK) {
        if (rc == CUMF_ECUDA && g_last_error.empty()) set_last_error("plan upload failed");
        plan_free(p);
        return rc;
    }
    *out = p;
    return CUMF_OK;
}


extern "C" int cumf_plan_destroy(cumf_plan* plan) {
    plan_free(plan);
    return CUMF_OK;
}
extern "C" int cumf_plan_last_launches(const cumf_plan* plan) { return plan ? plan->last_launches : 0; }
extern "C" int cumf_plan_set_factor_rows(cumf_plan* plan, int rows) {
    CUMF_REQUIRE(plan && rows > 0, "cumf_plan_set_factor_rows: bad argument");
    if (plan->tc) tc_plan_set_factor_rows(plan->tc, rows);
    return CUMF_OK;
}

static void plan_time_begin(cumf_plan* p, cudaStream_t st, cudaEvent_t* e0, cudaEvent_t* e1) {
    *e0 = *e1 = nullptr;
    if (!p->time_kernel) return;
    cudaEventCreate(e0);
    cudaEventCreate(e1);
    cudaEventRecord(*e0, st);
}
static void plan_time_end(cumf_plan* p, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1) {
    if (!p->time_kernel || !e0) return;
    cudaEventRecord(e1, st);
    p->kernel_events.emplace_back(e0, e1);
}
static int plan_create_impl(cumf_plan** out, const int* h_rowptr, int rows, int row_begin, int row_end, int f,
                            int path, bool alloc_workspace, bool force_slots = false) {
    CUMF_REQUIRE(out && h_rowptr, "null pointer");
    CUMF_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= rows, "bad row range");
    std::vector<long long> ptr((size_t)(row_end - row_begin) + 1);
    for (int r = row_begin; r <= row_end; ++r) ptr[r - row_begin] = h_rowptr[r];
    return plan_create_core(out, ptr.data(), ptr.data() + 1, rows, row_begin, row_end, f, path, alloc_workspace, force_slots, row_begin);
}

extern "C" int cumf_plan_create(cumf_plan** out, const int* h_rowptr, int rows, int row_begin, int row_end, int f,
                                int path) {
    return plan_create_impl(out, h_rowptr, rows, row_begin, row_end, f, path, true);
}

// 64-bit row pointers (hugewiki.cu:2266 squeezes its 3.1 G ratings through `unsigned int`; here the pointer array is int64
// and only the ratings of ONE plan -- one shard -- must number < 2^31)
extern "C" int cumf_plan_create64(cumf_plan** out, const long long* h_rowptr, int rows, int row_begin, int row_end, int f,
                                  int path) {
    CUMF_REQUIRE(out && h_rowptr, "null pointer");
    CUMF_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= rows, "bad row range");
    return plan_create_core(out, h_rowptr + row_begin, h_rowptr + row_begin + 1, rows, row_begin, row_end, f, path, true, false, row_begin);
}

// Partial-Gram plan: all `rows` rows, row r restricted to ratings [h_begin[r], h_end[r]).  On the fused path
// every chunk stores its partial [A|b] (nothing is solved in-kernel), so the plan carries one slot per chunk.
extern "C" int cumf_plan_create_ranges(cumf_plan** out, const long long* h_begin, const long long* h_end, int rows,
                                       int f, int path) {
    CUMF_REQUIRE(path == CUMF_PATH_SIMT || path == CUMF_PATH_TC || path == CUMF_PATH_AUTO, "path");
    if (path == CUMF_PATH_AUTO) path = tc_path_supports(f) ? CUMF_PATH_TC : CUMF_PATH_SIMT;
    return plan_create_core(out, h_begin, h_end, rows, 0, rows, f, path, false, path == CUMF_PATH_TC);
}

// [A|b] of every row of the plan over the plan's rating ranges, lambda * (ratings in range) on the diagonal
// (the per-GPU partial of hugewiki.cu:1675-1678 when the ranges are one GPU's share); tt is [rows][f*f],
// rhs is [rows][f].  Asynchronous on `stream`.
extern "C" int cumf_plan_gram(cumf_plan* p, const int* d_colidx, const float* d_val, const float* d_factor,
                              float lambda, float* d_tt, float* d_rhs, void* stream) {
    CUMF_REQUIRE(p && d_colidx && d_val && d_factor && d_tt && d_rhs, "null pointer");
    CUMF_REQUIRE(p->row_begin == 0 && p->row_end == p->rows, "cumf_plan_gram needs a plan over all its rows");
    cudaStream_t st = (cudaStream_t)stream;
    const int f = p->f;
    const long long base = p->base;
    p->last_launches = 0;
    cudaEvent_t e0, e1;
    plan_time_begin(p, st, &e0, &e1);
    if (p->path == CUMF_PATH_TC) {
        CUMF_REQUIRE(p->splits.size() == (size_t)p->rows, "plan was not created by cumf_plan_create_ranges");
        int launches = 0;
        CUMF_TRY(tc_update_factor(p->tc, p->d_chunks.as<Chunk>(), (int)p->chunks.size(), d_colidx + base, d_val + base,
                                  d_factor, nullptr, f, lambda, 0.f, p->scratchA.as<float>(), p->scratchB.as<float>(),
                                  st, &launches));
        p->last_launches += launches;
        CUMF_TRY(launch_split_reduce(p->d_splits.as<SplitRow>(), 0, (int)p->splits.size(), f, lambda, 0, 0, d_tt, d_rhs,
                                     p->scratchA.as<float>(), p->scratchB.as<float>(), st));
        p->last_launches += 1;
    } else {
        CUMF_TRY(launch_gram_simt(p->d_chunks.as<Chunk>(), 0, (int)p->chunks.size(), d_colidx + base, d_val + base,
                                  d_factor, f, lambda, 0, d_tt, d_rhs, p->scratchA.as<float>(),
                                  p->scratchB.as<float>(), st));
        p->last_launches += 1;
        if (!p->splits.empty()) {
            CUMF_TRY(launch_split_reduce(p->d_splits.as<SplitRow>(), 0, (int)p->splits.size(), f, lambda, 0, 0, d_tt,
                                         d_rhs, p->scratchA.as<float>(), p->scratchB.as<float>(), st));
            p->last_launches += 1;
        }
    }
    plan_time_end(p, st, e0, e1);
    return CUMF_OK;
}

static double plan_collect_kernel_ms(cumf_plan* p) {
    for (auto& e : p->kernel_events) {
        float ms = 0.f;
        cudaEventSynchronize(e.second);
        if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) p->kernel_ms_total += ms;
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    p->kernel_events.clear();
    return p->kernel_ms_total;
}

// x = d_out + row_base * f: the batch's rows of the factor; peers (optional) receive the same rows
static int solve_batch(float* tt, float* x, float* rhs, int batch, int f, int solver, float cg_iter, cudaStream_t st,
                       int* launches, const PeerOut* peers = nullptr, int row_base = 0) {
    if (solver == CUMF_SOLVER_LU) {
        CUMF_TRY(launch_lu(tt, x, rhs, batch, f, st));
        *launches += 1;   // our pointer-fill kernel; the cuBLAS kernels are library code
        if (peers)        // oracle mode: plain peer copies of the solved block
            for (int k = 0; k < peers->n; ++k)
                CUMF_CUDA_TRY(cudaMemcpyAsync(peers->p[k] + (size_t)row_base * f, x, sizeof(float) * (size_t)batch * f, cudaMemcpyDefault, st));
    } else {
        CUMF_TRY(launch_cg(tt, x, rhs, batch, f, cg_iter, nullptr, st, 0.f, nullptr, peers, row_base));
        *launches += 1;
    }
    return CUMF_OK;
}

static int update_factor_impl(cumf_plan* p, const int* d_colidx, const float* d_val, const float* d_factor,
                              float* d_out, float lambda, int solver, float cgIter, void* stream, const PeerOut* peers,
                              const SplitOut* split_out = nullptr, bool table_current = false, bool* wrote_split = nullptr);

extern "C" int cumf_update_factor(cumf_plan* p, const int* d_colidx, const float* d_val, const float* d_factor,
                                  float* d_out, float lambda, int solver, float cgIter, void* stream) {
    return update_factor_impl(p, d_colidx, d_val, d_factor, d_out, lambda, solver, cgIter, stream, nullptr);
}

// peers (optional): replicas of d_out on other GPUs that receive every updated row (the row-block exchange of the sharded
// half-step, done by the solver epilogues themselves)
// split_out (optional): f = 100 gather tables that receive the split form of every row this half-step solves; *wrote_split says
// whether every row went there (fused generic kernel + CG tail); table_current: this plan's own gather table is up to date
static int update_factor_impl(cumf_plan* p, const int* d_colidx, const float* d_val, const float* d_factor,
                              float* d_out, float lambda, int solver, float cgIter, void* stream, const PeerOut* peers,
                              const SplitOut* split_out, bool table_current, bool* wrote_split) {
    CUMF_REQUIRE(p && d_colidx && d_val && d_factor && d_out, "null pointer");
    if (peers && peers->n == 0) peers = nullptr;
    if (split_out && split_out->n == 0) split_out = nullptr;
    if (wrote_split) *wrote_split = false;
    CUMF_REQUIRE(solver == CUMF_SOLVER_CG || solver == CUMF_SOLVER_LU, "unknown solver");
    cudaStream_t st = (cudaStream_t)stream;
    const int f = p->f;
    int launches = 0;
    const Chunk* d_chunks = p->d_chunks.as<Chunk>();
    const SplitRow* d_splits = p->d_splits.as<SplitRow>();

    if (p->path == CUMF_PATH_TC && solver == CUMF_SOLVER_CG) {
        const int ns = (int)p->splits.size();
        double* terms = nullptr;
        p->sse_terms_valid = false;
        if (p->collect_sse) {
            const int want = tc_sse_terms_per_cta() * tc_plan_grid(p->tc) + ns;
            if (p->sse_terms_count != want) {
                p->sse_terms.release();
                CUMF_TRY(p->sse_terms.alloc(sizeof(double) * std::max(1, want)));
                p->sse_terms_count = want;
            }
            terms = p->sse_terms.as<double>();
            CUMF_CUDA_TRY(cudaMemsetAsync(terms, 0, sizeof(double) * std::max(1, want), st));
        }
        cudaEvent_t e0, e1;
        plan_time_begin(p, st, &e0, &e1);
        if (split_out && !(f == 100 && tc_plan_impl(p->tc) == 2)) split_out = nullptr;     // only the generic kernel writes split rows
        const TcExtra extra = tc_extra_from(peers, split_out, table_current);
        CUMF_TRY(tc_update_factor(p->tc, d_chunks, (int)p->chunks.size(), d_colidx, d_val, d_factor, d_out,
                                  f, lambda, cgIter, p->scratchA.as<float>(), p->scratchB.as<float>(), st, &launches, terms,
                                  (peers || split_out || table_current) ? &extra : nullptr));
        plan_time_end(p, st, e0, e1);
        // rows that were split across CTAs: reduce their partials into a compact batch, solve it
        if (ns > 0) {
            CUMF_TRY(launch_split_reduce(d_splits, 0, ns, f, lambda, /*compact=*/1, 0, p->tt.as<float>(),
                                         p->rhs.as<float>(), p->scratchA.as<float>(), p->scratchB.as<float>(), st));
            CUMF_TRY(launch_cg(p->tt.as<float>(), d_out, p->rhs.as<float>(), ns, f, cgIter, d_splits, st, lambda,
                               terms ? terms + tc_sse_terms_per_cta() * tc_plan_grid(p->tc) : nullptr, peers, 0, split_out));
            launches += 2;
        }
        if (wrote_split) *wrote_split = (split_out != nullptr);
        p->sse_terms_valid = (terms != nullptr);
        p->last_launches = launches;
        return CUMF_OK;
    }

    // unfused path: materialise A for a batch of rows, then solve the batch
    if (p->path != CUMF_PATH_SIMT || p->batch_rows <= 0) {
        set_last_error("this plan was built for the fused path; the LU oracle needs a CUMF_PATH_SIMT plan");
        return CUMF_EUNSUPPORTED;
    }
    const int owned = p->row_end - p->row_begin;
    size_t s_lo = 0;
    for (int b0 = 0; b0 < owned; b0 += p->batch_rows) {
        const int b1 = std::min(owned, b0 + p->batch_rows);
        const int c0 = p->row_chunk_ptr[b0], c1 = p->row_chunk_ptr[b1];
        const int row_base = p->row_begin + b0;
        cudaEvent_t e0, e1;
        plan_time_begin(p, st, &e0, &e1);
        CUMF_TRY(launch_gram_simt(d_chunks, c0, c1, d_colidx, d_val, d_factor, f, lambda, row_base, p->tt.as<float>(),
                                  p->rhs.as<float>(), p->scratchA.as<float>(), p->scratchB.as<float>(), st));
        plan_time_end(p, st, e0, e1);
        launches += (c1 > c0);
        // split rows inside this batch
        size_t s_hi = s_lo;
        while (s_hi < p->splits.size() && p->splits[s_hi].row < p->row_begin + b1) ++s_hi;
        if (s_hi > s_lo) {
            CUMF_TRY(launch_split_reduce(d_splits, (int)s_lo, (int)s_hi, f, lambda, /*compact=*/0, row_base,
                                         p->tt.as<float>(), p->rhs.as<float>(), p->scratchA.as<float>(),
                                         p->scratchB.as<float>(), st));
            launches += 1;
        }
        s_lo = s_hi;
        CUMF_TRY(solve_batch(p->tt.as<float>(), d_out + (size_t)row_base * f, p->rhs.as<float>(), b1 - b0, f, solver,
                             cgIter, st, &launches, peers, row_base));
    }
    p->last_launches = launches;
    return CUMF_OK;
}

// ---------------------------------------------------------------------------------
// misc C ABI
// ---------------------------------------------------------------------------------
extern "C" const char* cumf_last_error(void) { return g_last_error.c_str(); }
extern "C" int cumf_version(void) { return 100; }

// ---------------------------------------------------------------------------------
// b4 seams
// ---------------------------------------------------------------------------------
extern "C" int cumf_gram(int batch_offset, int batch_size, float* d_tt, float* d_rhs, const int* d_rowptr,
                         const int* d_colidx, const float* d_val, float lambda, int m, int f, const float* d_factor,
                         int path, void* stream) {
    CUMF_REQUIRE(d_tt && d_rowptr && d_colidx && d_factor, "null pointer");
    CUMF_REQUIRE(!d_rhs || d_val, "d_val is required when d_rhs is requested");
    CUMF_REQUIRE(path != CUMF_PATH_TC || (d_rhs && d_val), "the fused path always forms the RHS: pass d_rhs and d_val");
    CUMF_REQUIRE(batch_offset >= 0 && batch_size >= 0, "negative batch");
    CUMF_TRY(check_f(f));
    CUMF_TRY(check_device());
    cudaStream_t st = (cudaStream_t)stream;
    // rows beyond m are skipped like `if (row < m)` in the reference kernels (als.cu:449-450)
    const int row_end = std::min(m, batch_offset + batch_size);
    if (row_end <= batch_offset) return CUMF_OK;
    std::vector<int> h_rowptr(m + 1);
    CUMF_CUDA_TRY(cudaMemcpyAsync(h_rowptr.data(), d_rowptr, sizeof(int) * (m + 1), cudaMemcpyDeviceToHost, st));
    CUMF_CUDA_TRY(cudaStreamSynchronize(st));
    if (path == CUMF_PATH_AUTO) path = CUMF_PATH_SIMT;
    cumf_plan* p = nullptr;
    int rc = CUMF_OK;
    if (path == CUMF_PATH_TC) {
        // materialise A through the fused tensor-core kernel: every chunk stores its partial
        // [A|b] (split-row mode), then the deterministic reduce adds lambda*n_u and writes tt/rhs.
        CUMF_TRY(plan_create_impl(&p, h_rowptr.data(), m, batch_offset, row_end, f, CUMF_PATH_TC, false, true));
        const long long base = p->base;
        int launches = 0;
        rc = tc_update_factor(p->tc, p->d_chunks.as<Chunk>(), (int)p->chunks.size(), d_colidx + base,
                              d_val ? d_val + base : nullptr, d_factor, nullptr, f, lambda, 0.f,
                              p->scratchA.as<float>(), p->scratchB.as<float>(), st, &launches);
        if (rc == CUMF_OK)
            rc = launch_split_reduce(p->d_splits.as<SplitRow>(), 0, (int)p->splits.size(), f, lambda, 0, batch_offset,
                                     d_tt, d_rhs, p->scratchA.as<float>(), p->scratchB.as<float>(), st);
    } else {
        // the plan's own workspace is not needed here: aim the kernel at the caller's buffers
        CUMF_TRY(plan_create_impl(&p, h_rowptr.data(), m, batch_offset, row_end, f, CUMF_PATH_SIMT, false));
        const long long base = p->base;
        rc = launch_gram_simt(p->d_chunks.as<Chunk>(), 0, (int)p->chunks.size(), d_colidx + base,
                              d_val ? d_val + base : nullptr, d_factor, f, lambda, batch_offset, d_tt, d_rhs,
                              p->scratchA.as<float>(), p->scratchB.as<float>(), st);
        if (rc == CUMF_OK && !p->splits.empty())
            rc = launch_split_reduce(p->d_splits.as<SplitRow>(), 0, (int)p->splits.size(), f, lambda, 0, batch_offset,
                                     d_tt, d_rhs, p->scratchA.as<float>(), p->scratchB.as<float>(), st);
    }
    if (rc == CUMF_OK && cudaStreamSynchronize(st) != cudaSuccess) {
        set_last_error(std::string("cumf_gram: ") + cudaGetErrorString(cudaGetLastError()));
        rc = CUMF_ECUDA;
    }
    plan_free(p);
    return rc;
}

extern "C" int cumf_cg(const float* d_A, float* d_x, const float* d_b, int batchSize, int f, float cgIter,
                       void* stream) {
    CUMF_REQUIRE(d_A && d_x && d_b, "null pointer");
    CUMF_TRY(check_f(f));
    CUMF_TRY(check_device());
    CUMF_TRY(launch_cg(d_A, d_x, d_b, batchSize, f, cgIter, nullptr, (cudaStream_t)stream));
    // updateXWithCGHost synchronises and checks (cg.cu:686-687)
    CUMF_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return CUMF_OK;
}

extern "C" int cumf_lu(float* d_A, float* d_x, float* d_b, int batchSize, int f, void* stream) {
    CUMF_REQUIRE(d_A && d_x && d_b, "null pointer");
    CUMF_REQUIRE(f >= 1, "f");
    CUMF_TRY(check_device());
    return launch_lu(d_A, d_x, d_b, batchSize, f, (cudaStream_t)stream);
}

extern "C" int cumf_rmse(const float* d_val, const int* d_row, const int* d_col, const float* d_thetaT,
                         const float* d_XT, long count, int f, int drop_tail, float* rmse_out, double* sse_out,
                         void* stream) {
    CUMF_REQUIRE(d_val && d_row && d_col && d_thetaT && d_XT, "null pointer");
    CUMF_REQUIRE(f >= 2 && f % 2 == 0, "f must be even");
    CUMF_TRY(check_device());
    cudaStream_t st = (cudaStream_t)stream;
    // the test-set launch has (count-1)/256 blocks of 256 threads (als.cu:1006)
    long launched = drop_tail ? ((count - 1) / 256) * 256 : count;
    if (launched < 0) launched = 0;
    DevBuf part, out;
    CUMF_TRY(part.alloc(sizeof(double) * sse_partial_capacity()));
    CUMF_TRY(out.alloc(sizeof(double)));
    int rc = launch_sse(d_val, d_row, d_col, d_thetaT, d_XT, launched, f, out.as<double>(), part.as<double>(),
                        sse_partial_capacity(), st);
    double sse = 0.0;
    if (rc == CUMF_OK && cudaMemcpyAsync(&sse, out.p, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = CUMF_ECUDA;
    if (rc == CUMF_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = CUMF_ECUDA;
    if (rc != CUMF_OK) return rc;
    if (sse_out) *sse_out = sse;
    if (rmse_out) *rmse_out = sqrtf((float)sse / (float)count);   // als.cu:991, 1018
    return CUMF_OK;
}

// ---------------------------------------------------------------------------------
// resident solver handle
// ---------------------------------------------------------------------------------
namespace cumf_multi { struct SameDeviceSync; }
struct cumf_als_solver {
    int m = 0, n = 0, f = 0;
    long nnz = 0, nnz_test = 0;
    float lambda = 0.f;
    int xb = 0, xe = 0, tb = 0, te = 0;
    int device = 0, solver = CUMF_SOLVER_CG, path = CUMF_PATH_AUTO;
    float cg_iter = 6.0f;               // CG_ITER, als.cu:32
    // owned slices (rebased): CSR rows [xb,xe), CSC columns [tb,te)
    DevBuf csr_col, csr_val, csc_row, csc_val;
    DevBuf coo_row;                     // cooRowIndex for the owned CSR slice (train RMSE pairing, als.cu:979-980)
    DevBuf test_row, test_col, test_val;
    long train_cnt = 0, test_cnt = 0, csc_cnt = 0;
    DevBuf theta, x;                    // full replicas
    DevBuf sse, partials;
    // uploads run on their own (non-blocking) stream; the first use of each group waits on its event, so the
    // first X half-step overlaps the CSC upload and the first theta half-step the COO/test upload
    cudaStream_t up_stream = nullptr;
    cudaEvent_t ev_csr = nullptr, ev_csc = nullptr, ev_rmse = nullptr;
    // train RMSE walk: -1 undecided, 0 literal (cooRow[i], csrCol[i], csrVal[i]) pairs, 1 by CSC columns (theta row
    // in registers, X gathered), 2 by CSR rows (X row in registers, theta gathered); 1/2 need cooRow == CSR rows
    int train_mode = -1;
    // train RMSE from the theta half-step's by-product (plan.sse_terms): needs the whole matrix on this GPU, the fused
    // CG path, cooRow == CSR rows, and X untouched since that half-step
    bool theta_fresh = false, can_collect_sse = false;
    double sum_r2 = -1.0;               // sum of squared train ratings (computed on first use)
    // one-time RMSE preparation enqueued on the upload stream right behind the data it reads (COO-vs-CSR check flag,
    // sum of squared ratings), so that it runs while the first half-steps do: prep[0] = flag (as double), prep[1] = sum r^2
    DevBuf prep, prep_partials;
    bool prep_coo = false, prep_r2 = false;
    cumf_plan* px = nullptr;
    cumf_plan* pt = nullptr;
    // multi-GPU (row sharding): this shard is rank `rank` of `nranks`; peer_x / peer_theta are the other
    // ranks' factor replicas (peer access in one process, CUDA IPC across processes), flags / peer_flags the epoch words of
    // the device-side barrier that ends every half-step
    int rank = 0, nranks = 1;
    PeerOut peer_x{}, peer_theta{};
    DevBuf flags;                                   // [8] unsigned long long, written by the peers
    unsigned long long* peer_flags[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [r] = rank r's flags (own: local)
    unsigned long long epoch = 0;
    SplitOut theta_split_out{};                     // X-side gather tables (own + peers') that theta half-steps keep up to date
    bool theta_table_current = false;               // ... and whether the last theta half-step did (any other write to theta resets it)
    bool theta_ptr_exposed = false;                 // cumf_als_theta_ptr was handed out: never trust the table
    cumf_multi::SameDeviceSync* same_dev = nullptr; // set when the shards of a group share one device (test mode): event barrier
    std::vector<void*> ipc_opened;
    cudaStream_t run_stream = nullptr;              // a group gives every shard its own stream (two shards may share a device in tests)
    DevArena arena;                                 // everything sized at build time lives in ONE device allocation
    DevBuf arena_block;
    // timers
    double ms_x = 0, ms_theta = 0;
    long launches = 0, iterations = 0;
};

// CUMF_FUSED_SPLIT=0: theta half-steps do not write split rows, every X half-step re-splits all of theta (round 1 behaviour)
static bool fused_split_enabled() { return env_long("CUMF_FUSED_SPLIT", 1) != 0; }
static unsigned short* own_gather_table(cumf_als_solver* s) {
    return (s && s->px && s->px->tc && fused_split_enabled()) ? tc_plan_split_table_f100(s->px->tc) : nullptr;
}

extern "C" int cumf_als_destroy(cumf_als_solver* s) {
    if (!s) return CUMF_OK;
    const bool debug = env_long("CUMF_DEBUG", 0) != 0;
    double t[5] = {wall_seconds(), 0, 0, 0, 0};
    cudaSetDevice(s->device);
    if (s->up_stream) cudaStreamSynchronize(s->up_stream);
    cudaDeviceSynchronize();            // what cudaFree would do implicitly: nothing may still use the buffers kept for reuse
    t[1] = wall_seconds();
    for (void* p : s->ipc_opened) cudaIpcCloseMemHandle(p);
    s->ipc_opened.clear();
    if (s->run_stream) { cudaStreamDestroy(s->run_stream); s->run_stream = nullptr; }
    s->csr_col.release_to_cache(); s->csr_val.release_to_cache(); s->csc_row.release_to_cache(); s->csc_val.release_to_cache();
    s->coo_row.release_to_cache(); s->test_row.release_to_cache(); s->test_col.release_to_cache(); s->test_val.release_to_cache();
    const double t_theta0 = wall_seconds();
    s->theta.release_to_cache();
    const double t_theta1 = wall_seconds();
    s->x.release_to_cache(); s->sse.release(); s->partials.release();
    s->prep.release(); s->prep_partials.release();
    t[2] = wall_seconds();
    if (debug && t[2] - t[1] > 0.005)
        printf("\trelease: of which theta replica %.4f s, X replica and scalars %.4f s\n", t_theta1 - t_theta0, t[2] - t_theta1);
    plan_free(s->px, true);
    plan_free(s->pt, true);
    if (s->up_stream) { cudaStreamSynchronize(s->up_stream); cudaStreamDestroy(s->up_stream); }
    if (s->ev_csr) cudaEventDestroy(s->ev_csr);
    if (s->ev_csc) cudaEventDestroy(s->ev_csc);
    if (s->ev_rmse) cudaEventDestroy(s->ev_rmse);
    s->flags.release();
    t[3] = wall_seconds();
    const size_t arena_bytes = s->arena_block.bytes;
    s->arena_block.release_to_cache();      // one cudaFree for everything that was carved out of the arena
    t[4] = wall_seconds();
    if (debug)
        printf("\trelease: device idle after %.4f s, rating/factor buffers %.4f s, plans/streams/events %.4f s, arena (%.0f MB) %.4f s\n",
               t[1] - t[0], t[2] - t[1], t[3] - t[2], arena_bytes / 1048576.0, t[4] - t[3]);
    delete s;
    return CUMF_OK;
}

// Asynchronous when `host` is pinned (the reference CLI's buffers are, main.cpp:50-69); from pageable memory the
// runtime stages the copy and the call returns when the source has been read -- same result either way.
template <typename T>
static int upload(DevBuf& buf, const T* host, size_t count, cudaStream_t st) {
    CUMF_TRY(buf.alloc(sizeof(T) * count));
    if (count) CUMF_CUDA_TRY(cudaMemcpyAsync(buf.p, host, sizeof(T) * count, cudaMemcpyHostToDevice, st));
    return CUMF_OK;
}

// What one shard is built from: the row pointers of its OWN rows / columns, rebased to 0 (int64: a whole matrix may hold
// more than 2^31 ratings, one shard may not), and its slices of the rating arrays -- in host memory (uploaded here) or
// already on the device (borrowed: the caller keeps them alive; the sharded generator / loader path).
struct ShardSource {
    std::vector<long long> x_ptr, t_ptr;      // (x_end - x_begin + 1), (t_end - t_begin + 1) entries
    bool on_device = false;
    const int* csr_col = nullptr; const float* csr_val = nullptr;
    const int* csc_row = nullptr; const float* csc_val = nullptr;
    const int* coo_row = nullptr;             // host mode: this shard's slice of cooRowIndex, or null
    bool coo_is_csr = false;                  // device mode: the train samples ARE the CSR entries (no cooRowIndex array)
    const int* test_row = nullptr; const int* test_col = nullptr; const float* test_val = nullptr;
    long test_cnt = 0;                        // this shard's share of the launched test samples
};

template <typename T>
static int adopt(DevBuf& buf, const T* src, size_t count, bool on_device, cudaStream_t st) {
    if (on_device) { buf.borrow(const_cast<T*>(src), sizeof(T) * count); return CUMF_OK; }
    return upload(buf, src, count, st);
}

// wait_uploads = false leaves the rating uploads in flight when it returns (the half-steps and the RMSE wait on
// their events): only for callers that keep the host arrays alive until the first RMSE, i.e. cumf_doALS.
static int als_create_core(cumf_als_solver** out, const ShardSource& src, int m, int n, int f, long nnz, long nnz_test,
                           float lambda, int x_begin, int x_end, int t_begin, int t_end, int device, int solver, int path,
                           bool wait_uploads, const float* thetaTHost, const float* XTHost) {
    CUMF_REQUIRE(out, "null pointer");
    CUMF_REQUIRE(m > 0 && n > 0 && nnz >= 0 && nnz_test >= 0, "bad sizes");
    CUMF_REQUIRE(0 <= x_begin && x_begin <= x_end && x_end <= m, "bad X row range");
    CUMF_REQUIRE(0 <= t_begin && t_begin <= t_end && t_end <= n, "bad theta row range");
    CUMF_REQUIRE(src.x_ptr.size() == (size_t)(x_end - x_begin) + 1 && src.t_ptr.size() == (size_t)(t_end - t_begin) + 1, "row pointer slices");
    CUMF_TRY(check_f(f));
    CUMF_CUDA_TRY(cudaSetDevice(device));
    CUMF_TRY(check_device());
    if (solver == CUMF_SOLVER_LU) path = CUMF_PATH_SIMT;   // the LU oracle needs a materialised A
    cumf_als_solver* s = new cumf_als_solver();
    s->m = m; s->n = n; s->f = f; s->nnz = nnz; s->nnz_test = nnz_test; s->lambda = lambda;
    s->xb = x_begin; s->xe = x_end; s->tb = t_begin; s->te = t_end;
    s->device = device; s->solver = solver; s->path = path;
    int rc = CUMF_OK;
    const long long xn = src.x_ptr.back() - src.x_ptr.front(), tn = src.t_ptr.back() - src.t_ptr.front();
    // One device allocation for everything sized here (DevArena, common.cuh): rating slices, plans, stage tables, the split
    // tables of both sides, RMSE scratch.  What does not fit (split-row scratch of unusual shapes) falls back to cudaMalloc;
    // the factor replicas and the barrier flags stay allocations of their own (CUDA IPC exports whole allocations).
    struct ArenaScope {
        DevArena* prev;
        explicit ArenaScope(DevArena* a) : prev(t_arena) { t_arena = a; }
        ~ArenaScope() { t_arena = prev; }
    };
    const bool member = t_group_member;      // shard of a one-process group: nothing of it is ever exported over CUDA IPC
    if (env_long("CUMF_ARENA", 1) != 0) {
        const size_t owned = (size_t)(x_end - x_begin) + (size_t)(t_end - t_begin);
        size_t est = (size_t)64 << 20;
        if (member) est += ((size_t)m + (size_t)n) * f * sizeof(float) + 4096;    // the factor replicas and the flags live here too
        if (!src.on_device) est += (size_t)xn * 12 + (size_t)tn * 8 + (size_t)src.test_cnt * 12;
        est += owned * 64 + ((size_t)(xn + tn) / 32 + owned) * 8;                 // chunks, meta, stage tables
        est += ((size_t)m + (size_t)n + 2) * 512 * (f > 127 ? 2 : 1);              // pre-split fp16 tables of both sides
        // partial [A | b] slots of rows that are cut into several chunks (plan_create_core: more than CUMF_SPLIT_NNZ ratings), and
        // the compact batch their sums are solved from: ~0.5 GB on the Netflix X side, so it must be part of the estimate
        if (path != CUMF_PATH_SIMT) {
            const long long max_chunk = env_long("CUMF_SPLIT_NNZ", 8192);
            auto scratch = [&](const std::vector<long long>& ptr) {
                size_t slots = 0, rows_split = 0;
                for (size_t r = 0; r + 1 < ptr.size(); ++r) {
                    const long long cnt = ptr[r + 1] - ptr[r];
                    if (cnt > max_chunk) { slots += (size_t)((cnt + max_chunk - 1) / max_chunk); rows_split += 1; }
                }
                return (slots + rows_split) * ((size_t)f * f + f) * sizeof(float) + (slots + rows_split) * 2048;
            };
            est += scratch(src.x_ptr) + scratch(src.t_ptr);
        }
        est += est / 16;
        if (s->arena_block.alloc(est) == CUMF_OK) { s->arena.base = s->arena_block.as<unsigned char>(); s->arena.cap = est; s->arena.used = 0; }
        else cudaGetLastError();                                                  // no arena: every buffer gets its own allocation
    }
    ArenaScope arena_scope(&s->arena);
    auto fail = [&](int code) { t_arena = arena_scope.prev; cumf_als_destroy(s); return code; };
    if (member && s->arena.base) {
        // barrier flags: zeroed now, synchronously, while nothing is queued on this device (a peer may write them as soon as
        // ITS uploads are done)
        if ((rc = s->flags.alloc(8 * sizeof(unsigned long long))) != CUMF_OK) return fail(rc);
        if (cudaMemset(s->flags.p, 0, 8 * sizeof(unsigned long long)) != cudaSuccess) { set_last_error("cumf_als_create: flags"); return fail(CUMF_ECUDA); }
    }
    CUMF_REQUIRE(xn == 0 || (src.csr_col && src.csr_val), "CSR slice");
    CUMF_REQUIRE(tn == 0 || (src.csc_row && src.csc_val), "CSC slice");
    // Order: what the first X half-step needs goes to the copy engine first (factors, CSR), the work plans are built on
    // the host while those 1 GB are in flight, then the CSC / COO / test uploads follow; each group has its event.
    const bool debug = env_long("CUMF_DEBUG", 0) != 0;
    const double t_begin_wall = wall_seconds();
    if (cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_csr, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_csc, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_rmse, cudaEventDisableTiming) != cudaSuccess) {
        set_last_error("cumf_als_create: cannot create the upload stream");
        return fail(CUMF_ECUDA);
    }
    cudaStream_t up = s->up_stream;
    // the plan of the first half-step before anything is on the copy engine (its small synchronous copies would queue
    // behind the uploads); the theta-side plan is built while factors + CSR are in flight
    if ((rc = plan_create_core(&s->px, src.x_ptr.data(), src.x_ptr.data() + 1, m, x_begin, x_end, f, path, true, false, x_begin)) != CUMF_OK)
        return fail(rc);
    if ((rc = member ? s->theta.alloc(sizeof(float) * (size_t)n * f) : s->theta.alloc_own(sizeof(float) * (size_t)n * f)) != CUMF_OK) return fail(rc);
    if ((rc = member ? s->x.alloc(sizeof(float) * (size_t)m * f) : s->x.alloc_own(sizeof(float) * (size_t)m * f)) != CUMF_OK) return fail(rc);
    // initial factors (optional here; cumf_als_set_factors otherwise) go first: the X half-step needs them
    if ((thetaTHost && cudaMemcpyAsync(s->theta.p, thetaTHost, sizeof(float) * (size_t)n * f, cudaMemcpyHostToDevice, up) != cudaSuccess) ||
        (XTHost && cudaMemcpyAsync(s->x.p, XTHost, sizeof(float) * (size_t)m * f, cudaMemcpyHostToDevice, up) != cudaSuccess)) {
        set_last_error("cumf_als_create: factor upload failed");
        return fail(CUMF_ECUDA);
    }
    if ((rc = adopt(s->csr_col, src.csr_col, (size_t)xn, src.on_device, up)) != CUMF_OK) return fail(rc);
    if ((rc = adopt(s->csr_val, src.csr_val, (size_t)xn, src.on_device, up)) != CUMF_OK) return fail(rc);
    cudaEventRecord(s->ev_csr, up);
    if ((rc = adopt(s->csc_row, src.csc_row, (size_t)tn, src.on_device, up)) != CUMF_OK) return fail(rc);
    if ((rc = adopt(s->csc_val, src.csc_val, (size_t)tn, src.on_device, up)) != CUMF_OK) return fail(rc);
    cudaEventRecord(s->ev_csc, up);
    const double t_plans = wall_seconds();
    if ((rc = plan_create_core(&s->pt, src.t_ptr.data(), src.t_ptr.data() + 1, n, t_begin, t_end, f, path, true, false, t_begin)) != CUMF_OK)
        return fail(rc);
    const double t_plans_end = wall_seconds();
    s->px->time_kernel = s->pt->time_kernel = (env_long("CUMF_TIME_KERNELS", 0) != 0);
    cumf_plan_set_factor_rows(s->px, n);     // X rows gather theta rows, and vice versa
    cumf_plan_set_factor_rows(s->pt, m);
    const bool have_train = src.coo_row != nullptr || src.coo_is_csr;
    // train RMSE as a by-product of the theta half-step (see cumf_als_sse); CUMF_SSE_DIRECT=1 keeps the streaming kernel
    // (a shard's by-product covers the ratings of its theta rows = its CSC slice; over all shards that is every rating once)
    s->can_collect_sse = (s->pt->path == CUMF_PATH_TC && solver == CUMF_SOLVER_CG && have_train && env_long("CUMF_SSE_DIRECT", 0) == 0);
    if ((rc = s->sse.alloc(sizeof(double) * 2)) != CUMF_OK) return fail(rc);
    if ((rc = s->partials.alloc(sizeof(double) * sse_partial_capacity())) != CUMF_OK) return fail(rc);
    if ((rc = s->prep.alloc(sizeof(double) * 2)) != CUMF_OK) return fail(rc);
    if ((rc = s->prep_partials.alloc(sizeof(double) * sse_partial_capacity())) != CUMF_OK) return fail(rc);
    cudaMemsetAsync(s->prep.p, 0, sizeof(double) * 2, up);
    if (src.coo_row) {
        if ((rc = upload(s->coo_row, src.coo_row, (size_t)xn, up)) != CUMF_OK) return fail(rc);
        s->train_cnt = (long)xn;
    } else if (src.coo_is_csr) {
        s->train_cnt = (long)xn;
        s->train_mode = 2;                  // walk the CSR rows: the samples are the matrix entries by construction
    }
    s->csc_cnt = (long)tn;
    if (src.test_row && src.test_col && src.test_val && src.test_cnt > 0) {
        if ((rc = adopt(s->test_row, src.test_row, (size_t)src.test_cnt, src.on_device, up)) != CUMF_OK) return fail(rc);
        if ((rc = adopt(s->test_col, src.test_col, (size_t)src.test_cnt, src.on_device, up)) != CUMF_OK) return fail(rc);
        if ((rc = adopt(s->test_val, src.test_val, (size_t)src.test_cnt, src.on_device, up)) != CUMF_OK) return fail(rc);
        s->test_cnt = src.test_cnt;
    }
    // One-time RMSE preparation behind the last upload (kernels on this stream queue behind the persistent half-step
    // kernel that owns every SM, so anything placed before an upload would hold that upload back):
    if (s->can_collect_sse && tn > 0) {
        // sum of squared ratings, the constant of the by-product train RMSE
        if ((rc = launch_sumsq(s->csc_val.as<float>(), (long)tn, s->prep.as<double>() + 1, s->prep_partials.as<double>(),
                               sse_partial_capacity(), up)) != CUMF_OK) return fail(rc);
        s->prep_r2 = true;
    }
    if (src.coo_row && env_long("CUMF_SSE_LITERAL", 0) == 0 && xn > 0) {
        // is cooRowIndex the CSR row expansion?  (decides the train-RMSE walk, see cumf_als_sse)
        if ((rc = launch_coo_check(s->px->d_chunks.as<Chunk>(), (int)s->px->chunks.size(), s->coo_row.as<int>(),
                                   reinterpret_cast<int*>(s->prep.p), up)) != CUMF_OK) return fail(rc);
        s->prep_coo = true;
    }
    if (debug) {
        printf("\tsetup: X plan + factor/CSR/CSC uploads enqueued %.4f s, theta plan %.4f s, rest %.4f s\n", t_plans - t_begin_wall,
               t_plans_end - t_plans, wall_seconds() - t_plans_end);
        printf("\tsetup: arena %.0f of %.0f MB used, %d requests (%.0f MB) did not fit and are allocations of their own\n",
               s->arena.used / 1048576.0, s->arena.cap / 1048576.0, s->arena.overflow_count, s->arena.overflow_bytes / 1048576.0);
    }
    cudaEventRecord(s->ev_rmse, up);
    if (wait_uploads && cudaStreamSynchronize(s->up_stream) != cudaSuccess) {
        set_last_error(std::string("cumf_als_create: upload failed: ") + cudaGetErrorString(cudaGetLastError()));
        return fail(CUMF_ECUDA);
    }
    if (unsigned short* tab = own_gather_table(s)) { s->theta_split_out.p[0] = tab; s->theta_split_out.n = 1; }
    *out = s;
    return CUMF_OK;
}

// host arrays over the WHOLE matrix (the reference's ten arrays), row pointers int32 or int64
template <typename PtrT>
static int als_create_host(cumf_als_solver** out, const PtrT* csrRowIndexHostPtr, const int* csrColIndexHostPtr,
                           const float* csrValHostPtr, const int* cscRowIndexHostPtr, const PtrT* cscColIndexHostPtr,
                           const float* cscValHostPtr, const int* cooRowIndexHostPtr, const int* cooRowIndexTestHostPtr,
                           const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, int m, int n, int f, long nnz,
                           long nnz_test, float lambda, int x_begin, int x_end, int t_begin, int t_end, int device, int solver,
                           int path, bool wait_uploads, const float* thetaTHost, const float* XTHost) {
    CUMF_REQUIRE(out && csrRowIndexHostPtr && csrColIndexHostPtr && csrValHostPtr && cscRowIndexHostPtr &&
                     cscColIndexHostPtr && cscValHostPtr, "null pointer");
    CUMF_REQUIRE(m > 0 && n > 0, "bad sizes");
    CUMF_REQUIRE(0 <= x_begin && x_begin <= x_end && x_end <= m, "bad X row range");
    CUMF_REQUIRE(0 <= t_begin && t_begin <= t_end && t_end <= n, "bad theta row range");
    // note the reference's argument order for CSC (main.cpp:99-101, als.cu:867-869):
    // cscColIndex is the pointer array (n+1), cscRowIndex the row ids (nnz).
    ShardSource src;
    const long long xo = (long long)csrRowIndexHostPtr[x_begin], to = (long long)cscColIndexHostPtr[t_begin];
    src.x_ptr.resize((size_t)(x_end - x_begin) + 1);
    for (int r = x_begin; r <= x_end; ++r) src.x_ptr[r - x_begin] = (long long)csrRowIndexHostPtr[r] - xo;
    src.t_ptr.resize((size_t)(t_end - t_begin) + 1);
    for (int r = t_begin; r <= t_end; ++r) src.t_ptr[r - t_begin] = (long long)cscColIndexHostPtr[r] - to;
    src.csr_col = csrColIndexHostPtr + xo; src.csr_val = csrValHostPtr + xo;
    src.csc_row = cscRowIndexHostPtr + to; src.csc_val = cscValHostPtr + to;
    src.coo_row = cooRowIndexHostPtr ? cooRowIndexHostPtr + xo : nullptr;
    if (cooRowIndexTestHostPtr && cooColIndexTestHostPtr && cooValHostTestPtr && nnz_test > 0) {
        // samples the reference's launch covers: 256*((nnz_test-1)/256) (als.cu:1006); this
        // shard's share is the contiguous slice proportional to its X row range.
        const long long eff = ((long long)(nnz_test - 1) / 256) * 256;
        const long long t0 = eff * x_begin / m, t1 = eff * x_end / m;
        src.test_row = cooRowIndexTestHostPtr + t0; src.test_col = cooColIndexTestHostPtr + t0; src.test_val = cooValHostTestPtr + t0;
        src.test_cnt = (long)(t1 - t0);
    }
    return als_create_core(out, src, m, n, f, nnz, nnz_test, lambda, x_begin, x_end, t_begin, t_end, device, solver, path,
                           wait_uploads, thetaTHost, XTHost);
}

static int als_create_impl(cumf_als_solver** out, const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr,
                           const float* csrValHostPtr, const int* cscRowIndexHostPtr,
                           const int* cscColIndexHostPtr, const float* cscValHostPtr,
                           const int* cooRowIndexHostPtr, const int* cooRowIndexTestHostPtr,
                           const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, int m, int n, int f,
                           long nnz, long nnz_test, float lambda, int x_begin, int x_end, int t_begin, int t_end,
                           int device, int solver, int path, bool wait_uploads,
                           const float* thetaTHost = nullptr, const float* XTHost = nullptr) {
    return als_create_host<int>(out, csrRowIndexHostPtr, csrColIndexHostPtr, csrValHostPtr, cscRowIndexHostPtr,
                                cscColIndexHostPtr, cscValHostPtr, cooRowIndexHostPtr, cooRowIndexTestHostPtr,
                                cooColIndexTestHostPtr, cooValHostTestPtr, m, n, f, nnz, nnz_test, lambda, x_begin, x_end,
                                t_begin, t_end, device, solver, path, wait_uploads, thetaTHost, XTHost);
}

extern "C" int cumf_als_create(cumf_als_solver** out, const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr,
                               const float* csrValHostPtr, const int* cscRowIndexHostPtr,
                               const int* cscColIndexHostPtr, const float* cscValHostPtr,
                               const int* cooRowIndexHostPtr, const int* cooRowIndexTestHostPtr,
                               const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, int m, int n, int f,
                               long nnz, long nnz_test, float lambda, int x_begin, int x_end, int t_begin, int t_end,
                               int device, int solver, int path) {
    return als_create_impl(out, csrRowIndexHostPtr, csrColIndexHostPtr, csrValHostPtr, cscRowIndexHostPtr,
                           cscColIndexHostPtr, cscValHostPtr, cooRowIndexHostPtr, cooRowIndexTestHostPtr,
                           cooColIndexTestHostPtr, cooValHostTestPtr, m, n, f, nnz, nnz_test, lambda, x_begin, x_end,
                           t_begin, t_end, device, solver, path, /*wait_uploads=*/true);
}

// Same with int64 row / column pointer arrays: matrices beyond 2^31 ratings (hugewiki.cu:27-42: 3.1 G), of which this
// shard's slices must hold < 2^31 each.
extern "C" int cumf_als_create64(cumf_als_solver** out, const long long* csrRowIndexHostPtr, const int* csrColIndexHostPtr,
                                 const float* csrValHostPtr, const int* cscRowIndexHostPtr,
                                 const long long* cscColIndexHostPtr, const float* cscValHostPtr,
                                 const int* cooRowIndexHostPtr, const int* cooRowIndexTestHostPtr,
                                 const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, int m, int n, int f,
                                 long nnz, long nnz_test, float lambda, int x_begin, int x_end, int t_begin, int t_end,
                                 int device, int solver, int path) {
    return als_create_host<long long>(out, csrRowIndexHostPtr, csrColIndexHostPtr, csrValHostPtr, cscRowIndexHostPtr,
                                      cscColIndexHostPtr, cscValHostPtr, cooRowIndexHostPtr, cooRowIndexTestHostPtr,
                                      cooColIndexTestHostPtr, cooValHostTestPtr, m, n, f, nnz, nnz_test, lambda, x_begin,
                                      x_end, t_begin, t_end, device, solver, path, /*wait_uploads=*/true, nullptr, nullptr);
}

// A shard whose rating slices are ALREADY on `device` (generated there, or loaded there shard by shard: no pass through a
// host copy of the whole matrix).  h_csr_ptr / h_csc_ptr: the shard's own row / column pointers, rebased to 0
// (x_end - x_begin + 1 and t_end - t_begin + 1 entries, host).  The device arrays are borrowed: they must outlive the
// solver.  The train samples are the CSR entries themselves; test samples (optional) are this shard's own list.
extern "C" int cumf_als_create_device(cumf_als_solver** out, const long long* h_csr_ptr, const int* d_csr_col,
                                      const float* d_csr_val, const long long* h_csc_ptr, const int* d_csc_row,
                                      const float* d_csc_val, const int* d_test_row, const int* d_test_col,
                                      const float* d_test_val, long test_cnt, int m, int n, int f, long nnz, long nnz_test,
                                      float lambda, int x_begin, int x_end, int t_begin, int t_end, int device, int solver,
                                      int path) {
    CUMF_REQUIRE(out && h_csr_ptr && h_csc_ptr, "null pointer");
    CUMF_REQUIRE(0 <= x_begin && x_begin <= x_end && 0 <= t_begin && t_begin <= t_end, "bad ranges");
    ShardSource src;
    src.on_device = true;
    src.coo_is_csr = true;
    src.x_ptr.assign(h_csr_ptr, h_csr_ptr + (size_t)(x_end - x_begin) + 1);
    src.t_ptr.assign(h_csc_ptr, h_csc_ptr + (size_t)(t_end - t_begin) + 1);
    CUMF_REQUIRE(src.x_ptr.front() == 0 && src.t_ptr.front() == 0, "shard pointers must be rebased to 0");
    src.csr_col = d_csr_col; src.csr_val = d_csr_val; src.csc_row = d_csc_row; src.csc_val = d_csc_val;
    src.test_row = d_test_row; src.test_col = d_test_col; src.test_val = d_test_val; src.test_cnt = test_cnt;
    return als_create_core(out, src, m, n, f, nnz, nnz_test, lambda, x_begin, x_end, t_begin, t_end, device, solver, path,
                           /*wait_uploads=*/true, nullptr, nullptr);
}

// Ask the theta half-steps to leave the train-SSE by-product (costs one more block reduction per row, about 4 % of
// that half-step; saves the streaming pass over all ratings in cumf_als_sse).  cumf_doALS turns it on because it
// evaluates the RMSE after every iteration; a caller that only looks at the RMSE now and then leaves it off.
extern "C" int cumf_als_collect_train_sse(cumf_als_solver* s, int on) {
    CUMF_REQUIRE(s && s->pt, "null pointer");
    s->pt->collect_sse = (on != 0) && s->can_collect_sse;
    if (s->pt->collect_sse && s->pt->tc) {
        // the per-CTA / per-split-row terms are allocated now: the half-steps themselves must not allocate (a device
        // allocation synchronises the device, see tc_plan_set_factor_rows)
        cudaSetDevice(s->device);
        const int want = tc_sse_terms_per_cta() * tc_plan_grid(s->pt->tc) + (int)s->pt->splits.size();
        if (s->pt->sse_terms_count != want) {
            s->pt->sse_terms.release();
            if (s->pt->sse_terms.alloc(sizeof(double) * std::max(1, want)) == CUMF_OK) s->pt->sse_terms_count = want;
        }
    }
    if (!s->pt->collect_sse) { s->pt->sse_terms_valid = false; s->theta_fresh = false; }
    return s->pt->collect_sse ? 1 : 0;
}

extern "C" int cumf_als_set_factors(cumf_als_solver* s, const float* thetaTHost, const float* XTHost) {
    CUMF_REQUIRE(s && thetaTHost && XTHost, "null pointer");
    CUMF_CUDA_TRY(cudaSetDevice(s->device));
    CUMF_CUDA_TRY(cudaMemcpy(s->theta.p, thetaTHost, sizeof(float) * (size_t)s->n * s->f, cudaMemcpyHostToDevice));
    CUMF_CUDA_TRY(cudaMemcpy(s->x.p, XTHost, sizeof(float) * (size_t)s->m * s->f, cudaMemcpyHostToDevice));
    s->theta_fresh = false;
    s->theta_table_current = false;
    return CUMF_OK;
}
extern "C" int cumf_als_get_factors(cumf_als_solver* s, float* thetaTHost, float* XTHost) {
    CUMF_REQUIRE(s, "null pointer");
    CUMF_CUDA_TRY(cudaSetDevice(s->device));
    if (thetaTHost) CUMF_CUDA_TRY(cudaMemcpy(thetaTHost, s->theta.p, sizeof(float) * (size_t)s->n * s->f, cudaMemcpyDeviceToHost));
    if (XTHost) CUMF_CUDA_TRY(cudaMemcpy(XTHost, s->x.p, sizeof(float) * (size_t)s->m * s->f, cudaMemcpyDeviceToHost));
    return CUMF_OK;
}
extern "C" int cumf_als_shape(const cumf_als_solver* s, int* m, int* n, int* f) {
    CUMF_REQUIRE(s, "null pointer");
    if (m) *m = s->m;
    if (n) *n = s->n;
    if (f) *f = s->f;
    return CUMF_OK;
}
// (a caller that holds the raw pointer may rewrite theta behind the solver's back: the X side's gather table is then rebuilt by
// a full split pass before every X half-step, as in round 1)
extern "C" float* cumf_als_theta_ptr(cumf_als_solver* s) {
    if (!s) return nullptr;
    s->theta_ptr_exposed = true;
    s->theta_table_current = false;
    return s->theta.as<float>();
}
// library-internal (synth.cu): both replicas for an in-library rewrite; the caller does not keep the pointers
extern "C" void cumf_als_internal_rewrite_factors(cumf_als_solver* s, float** theta, float** x) {
    s->theta_fresh = false;
    s->theta_table_current = false;
    *theta = s->theta.as<float>();
    *x = s->x.as<float>();
}
extern "C" float* cumf_als_x_ptr(cumf_als_solver* s) { return s ? s->x.as<float>() : nullptr; }

extern "C" int cumf_als_update_x(cumf_als_solver* s, void* stream) {
    CUMF_REQUIRE(s, "null pointer");
    CUMF_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s->ev_csr, 0));
    s->theta_fresh = false;
    CUMF_TRY(update_factor_impl(s->px, s->csr_col.as<int>(), s->csr_val.as<float>(), s->theta.as<float>(),
                                s->x.as<float>(), s->lambda, s->solver, s->cg_iter, stream, &s->peer_x, nullptr,
                                /*table_current=*/s->theta_table_current));
    s->launches += s->px->last_launches;
    return CUMF_OK;
}
extern "C" int cumf_als_update_theta(cumf_als_solver* s, void* stream) {
    CUMF_REQUIRE(s, "null pointer");
    CUMF_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s->ev_csc, 0));
    // the X side's gather table (f = 100 long-row kernel) gets the split form of every theta row from this half-step's solver
    // epilogues, on every replica: the next X half-step skips its split pass over all of theta
    s->theta_table_current = false;
    bool wrote = false;
    // (the table is allocated with the plan and never moves while the row hint holds; if it ever did, the peers would be
    // writing into the old one: stop pushing split rows rather than trust stale pointers)
    if (s->theta_split_out.n > 0 && s->theta_split_out.p[0] != own_gather_table(s)) s->theta_split_out.n = 0;
    CUMF_TRY(update_factor_impl(s->pt, s->csc_row.as<int>(), s->csc_val.as<float>(), s->x.as<float>(),
                                s->theta.as<float>(), s->lambda, s->solver, s->cg_iter, stream, &s->peer_theta,
                                s->theta_split_out.n > 0 ? &s->theta_split_out : nullptr, false, &wrote));
    s->theta_table_current = wrote && !s->theta_ptr_exposed;
    s->launches += s->pt->last_launches;
    s->theta_fresh = s->pt->sse_terms_valid;
    return CUMF_OK;
}

extern "C" int cumf_als_sse(cumf_als_solver* s, double* train_sse, double* test_sse, void* stream) {
    CUMF_REQUIRE(s, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    double h[2] = {0.0, 0.0};
    double* d = s->sse.as<double>();
    CUMF_CUDA_TRY(cudaStreamWaitEvent(st, s->ev_csr, 0));
    CUMF_CUDA_TRY(cudaStreamWaitEvent(st, s->ev_csc, 0));
    CUMF_CUDA_TRY(cudaStreamWaitEvent(st, s->ev_rmse, 0));
    CUMF_CUDA_TRY(cudaMemsetAsync(d, 0, 2 * sizeof(double), st));
    if (train_sse && s->train_cnt > 0 && s->train_mode < 0) {
        // The reference pairs cooRowIndex[i] with csrColIndex[i], csrVal[i] (als.cu:979-980).  When
        // cooRowIndex is what the CSR row pointer says (every loader of the reference produces that), those samples
        // are the matrix entries and can be walked row by row or column by column; otherwise keep the literal pairs.
        s->train_mode = 0;
        if (env_long("CUMF_SSE_LITERAL", 0) == 0) {
            int h_flag = 1;
            if (s->prep_coo) {
                // checked on the upload stream while the first half-steps ran (ev_rmse is behind it)
                CUMF_CUDA_TRY(cudaMemcpyAsync(&h_flag, s->prep.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                CUMF_CUDA_TRY(cudaStreamSynchronize(st));
            } else {
                int* flag = reinterpret_cast<int*>(d);      // scratch: re-zeroed below
                CUMF_TRY(launch_coo_check(s->px->d_chunks.as<Chunk>(), (int)s->px->chunks.size(), s->coo_row.as<int>(), flag, st));
                CUMF_CUDA_TRY(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
                CUMF_CUDA_TRY(cudaStreamSynchronize(st));
                CUMF_CUDA_TRY(cudaMemsetAsync(d, 0, 2 * sizeof(double), st));
                s->launches += 1;
            }
            if (h_flag == 0) {
                // gather from the smaller factor.  A row shard keeps the CSR walk: its sample set (the owned CSR rows)
                // is then the same whatever the other ranks decide.
                const bool whole = (s->xb == 0 && s->xe == s->m && s->tb == 0 && s->te == s->n);
                s->train_mode = (whole && s->m <= s->n) ? 1 : 2;
                // (cooRowIndex is not read again; its buffer goes with the solver -- cudaFree would synchronise the device here)
            }
        }
    }
    bool train_done = false;
    if (train_sse && s->train_cnt > 0 && s->train_mode != 0 && s->theta_fresh && s->pt->sse_terms_valid) {
        // By-product path: the theta half-step left  T = sum_rows (x^T b + x^T r + reg x^T x)  and X has not changed
        // since, so  train SSE = sum r^2 - T  (gram_tc.cu) with no further pass over the ratings.  The subtraction
        // loses log10(sum r^2 / SSE) digits of the fp32 row terms (about 1.6 on rating data); the streaming kernel
        // takes over when the fit is so tight that less than three digits would be left, or on any non-finite term.
        if (s->sum_r2 < 0.0 && s->prep_r2) {
            double h_r2 = 0.0;
            CUMF_CUDA_TRY(cudaMemcpyAsync(&h_r2, s->prep.as<double>() + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
            CUMF_CUDA_TRY(cudaStreamSynchronize(st));
            s->sum_r2 = h_r2;
        }
        if (s->sum_r2 < 0.0) {
            CUMF_TRY(launch_sumsq(s->csc_val.as<float>(), s->csc_cnt, d, s->partials.as<double>(), sse_partial_capacity(), st));
            double h_r2 = 0.0;
            CUMF_CUDA_TRY(cudaMemcpyAsync(&h_r2, d, sizeof(double), cudaMemcpyDeviceToHost, st));
            CUMF_CUDA_TRY(cudaStreamSynchronize(st));
            s->sum_r2 = h_r2;
            s->launches += 2;
        }
        CUMF_TRY(launch_sum_doubles(s->pt->sse_terms.as<double>(), s->pt->sse_terms_count, d, st));
        double h_t = 0.0;
        CUMF_CUDA_TRY(cudaMemcpyAsync(&h_t, d, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUMF_CUDA_TRY(cudaStreamSynchronize(st));
        s->launches += 1;
        const double sse_alg = s->sum_r2 - h_t;
        if (std::isfinite(sse_alg) && sse_alg > 1e-3 * s->sum_r2) {
            h[0] = sse_alg;
            train_done = true;
            if (env_long("CUMF_SSE_CHECK", 0) != 0) {
                CUMF_TRY(launch_sse_chunks(s->pt->d_chunks.as<Chunk>(), (int)s->pt->chunks.size(), s->csc_row.as<int>(),
                                           s->csc_val.as<float>(), s->theta.as<float>(), s->x.as<float>(), 1, s->f, d,
                                           s->partials.as<double>(), sse_partial_capacity(), st));
                double h_d = 0.0;
                CUMF_CUDA_TRY(cudaMemcpyAsync(&h_d, d, sizeof(double), cudaMemcpyDeviceToHost, st));
                CUMF_CUDA_TRY(cudaStreamSynchronize(st));
                printf("[CUMF_SSE_CHECK] train SSE by-product %.9e  streaming %.9e  rel diff %.3e  (sum r^2 %.6e)\n", sse_alg, h_d,
                       (sse_alg - h_d) / h_d, s->sum_r2);
            }
        }
        CUMF_CUDA_TRY(cudaMemsetAsync(d, 0, 2 * sizeof(double), st));
    }
    if (train_sse && s->train_cnt > 0 && !train_done) {
        // a shard that was asked for the by-product counts the ratings of its CSC slice: its streaming fall-back must walk
        // the same set (by columns), or the sum over the shards would count ratings twice
        const bool whole_matrix = (s->xb == 0 && s->xe == s->m && s->tb == 0 && s->te == s->n);
        const int mode = (s->pt->collect_sse && !whole_matrix && s->train_mode != 0) ? 1 : s->train_mode;
        if (mode == 1) {
            CUMF_TRY(launch_sse_chunks(s->pt->d_chunks.as<Chunk>(), (int)s->pt->chunks.size(), s->csc_row.as<int>(),
                                       s->csc_val.as<float>(), s->theta.as<float>(), s->x.as<float>(), 1, s->f, d,
                                       s->partials.as<double>(), sse_partial_capacity(), st));
        } else if (mode == 2) {
            CUMF_TRY(launch_sse_chunks(s->px->d_chunks.as<Chunk>(), (int)s->px->chunks.size(), s->csr_col.as<int>(),
                                       s->csr_val.as<float>(), s->x.as<float>(), s->theta.as<float>(), 0, s->f, d,
                                       s->partials.as<double>(), sse_partial_capacity(), st));
        } else {
            CUMF_TRY(launch_sse(s->csr_val.as<float>(), s->coo_row.as<int>(), s->csr_col.as<int>(), s->theta.as<float>(),
                                s->x.as<float>(), s->train_cnt, s->f, d, s->partials.as<double>(), sse_partial_capacity(), st));
        }
        s->launches += 2;
    }
    if (test_sse && s->test_cnt > 0) {
        CUMF_TRY(launch_sse(s->test_val.as<float>(), s->test_row.as<int>(), s->test_col.as<int>(), s->theta.as<float>(),
                            s->x.as<float>(), s->test_cnt, s->f, d + 1, s->partials.as<double>(), sse_partial_capacity(), st));
        s->launches += 2;
    }
    double hd[2] = {0.0, 0.0};
    CUMF_CUDA_TRY(cudaMemcpyAsync(hd, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUMF_CUDA_TRY(cudaStreamSynchronize(st));
    if (!train_done) h[0] = hd[0];
    h[1] = hd[1];
    if (train_sse) *train_sse = h[0];
    if (test_sse) *test_sse = h[1];
    return CUMF_OK;
}

extern "C" int cumf_als_peer_barrier(cumf_als_solver* s, void* stream);

extern "C" int cumf_als_iterate(cumf_als_solver* s, int iters, float* ms_out, void* stream) {
    CUMF_REQUIRE(s && iters >= 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    struct Events {
        std::vector<cudaEvent_t> v;
        ~Events() { for (auto e : v) if (e) cudaEventDestroy(e); }
    } evs;
    evs.v.assign(2 * (size_t)iters + 1, nullptr);
    std::vector<cudaEvent_t>& ev = evs.v;
    for (auto& e : ev) CUMF_CUDA_TRY(cudaEventCreate(&e));
    CUMF_CUDA_TRY(cudaEventRecord(ev[0], st));
    int rc = CUMF_OK;
    // a shard of a same-device group keeps meeting its peers on the host after a failure (see SameDeviceSync)
    auto barrier = [&]() {
        if (rc != CUMF_OK && !s->same_dev) return;
        const std::string first = rc != CUMF_OK ? cumf_last_error() : "";
        const int b = cumf_als_peer_barrier(s, stream);                  // no-op for a single rank
        if (rc == CUMF_OK) rc = b; else set_last_error(first);
    };
    for (int it = 0; it < iters && (rc == CUMF_OK || s->same_dev); ++it) {
        if (rc == CUMF_OK) rc = cumf_als_update_x(s, stream);
        barrier();
        if (rc == CUMF_OK && cudaEventRecord(ev[2 * it + 1], st) != cudaSuccess) rc = CUMF_ECUDA;
        if (rc == CUMF_OK) rc = cumf_als_update_theta(s, stream);
        barrier();
        if (rc == CUMF_OK && cudaEventRecord(ev[2 * it + 2], st) != cudaSuccess) rc = CUMF_ECUDA;
    }
    if (rc == CUMF_OK && cudaStreamSynchronize(st) != cudaSuccess) {
        set_last_error(std::string("cumf_als_iterate: ") + cudaGetErrorString(cudaGetLastError()));
        rc = CUMF_ECUDA;
    }
    if (rc == CUMF_OK) {
        float total = 0.f;
        for (int it = 0; it < iters; ++it) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, ev[2 * it], ev[2 * it + 1]);
            cudaEventElapsedTime(&b, ev[2 * it + 1], ev[2 * it + 2]);
            s->ms_x += a; s->ms_theta += b; total += a + b;
        }
        s->iterations += iters;
        if (ms_out) *ms_out = total;
    }
    return rc;
}

extern "C" int cumf_als_timers(cumf_als_solver* s, double* out6, int reset) {
    CUMF_REQUIRE(s && out6, "null pointer");
    out6[0] = s->ms_x; out6[1] = s->ms_theta;
    out6[2] = plan_collect_kernel_ms(s->px);
    out6[3] = plan_collect_kernel_ms(s->pt);
    out6[4] = (double)s->launches; out6[5] = (double)s->iterations;
    if (reset) {
        s->ms_x = s->ms_theta = 0; s->launches = 0; s->iterations = 0;
        s->px->kernel_ms_total = s->pt->kernel_ms_total = 0;
    }
    return CUMF_OK;
}


// ---------------------------------------------------------------------------------
// Multi-GPU row sharding inside the library (replaces the X_BATCH / THETA_BATCH loop of als.cu:768-777
// and the peer-copy scheme of hugewiki.cu:2562-2572, 2744-2745).  Rank g owns rating-balanced row ranges of X and theta,
// holds full replicas of both factors, and its solver epilogues store every updated row into all replicas (PeerOut), so
// the exchange overlaps the half-step; a device-side flag barrier (one tiny kernel per half-step) orders the half-steps.
// Two ways to connect the ranks:
//   * cumf_group_* / doALS with CUMF_GPUS=n : one process, n devices, one host thread per device
//   * cumf_als_ipc_export / _import        : one process per GPU (torchrun, MPI), CUDA IPC handles exchanged by the caller
// ---------------------------------------------------------------------------------
namespace cumf_multi {
struct HostBarrier {
    std::atomic<int> count{0}, sense{0};
    int n = 1;
    void wait() {
        const int s = sense.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
            count.store(0, std::memory_order_relaxed);
            sense.store(s ^ 1, std::memory_order_release);
        } else {
            while (sense.load(std::memory_order_acquire) == s) std::this_thread::yield();
        }
    }
};
// Test mode CUMF_GROUP_SAME_DEVICE=1 (every shard on one GPU): a spinning barrier kernel is not safe there -- streams of one
// device share its few hardware queues, so shard A's spinning barrier can sit in FRONT of the very kernel of shard B it waits
// for (seen when one host thread ran half a step ahead of the other).  Such a group orders its half-steps with events instead:
// each shard records, the host threads meet, each shard's stream waits for the others' records.  Every shard thread must call
// the barrier the same number of times, failed or not.
struct SameDeviceSync {
    HostBarrier hb;
    cudaEvent_t ev[8][2] = {};
    int n = 0;
    explicit SameDeviceSync(int shards) : n(shards) {
        hb.n = shards;
        for (int r = 0; r < shards; ++r)
            for (int p = 0; p < 2; ++p) cudaEventCreateWithFlags(&ev[r][p], cudaEventDisableTiming);
    }
    ~SameDeviceSync() {
        for (int r = 0; r < n; ++r)
            for (int p = 0; p < 2; ++p) if (ev[r][p]) cudaEventDestroy(ev[r][p]);
    }
};
struct FlagPtrs { unsigned long long* p[8]; };
// thread r of rank `me`: publish my epoch in rank r's flag word [me], then wait until rank r has published its own in mine
__global__ void peer_barrier_kernel(FlagPtrs flags, int me, int nranks, unsigned long long epoch) {
    const int r = threadIdx.x;
    if (r >= nranks || r == me) return;
    __threadfence_system();                                       // this GPU's earlier writes (peer rows) before the flag
    volatile unsigned long long* theirs = flags.p[r] + me;
    *theirs = epoch;
    __threadfence_system();
    volatile unsigned long long* mine = flags.p[me] + r;
    unsigned long long spins = 0;
    while (*mine < epoch) {
        __nanosleep(200);
        if (++spins > (1ull << 26)) __trap();                     // ~13 s: a peer died; fail the launch instead of hanging
    }
    __threadfence_system();
}
}  // namespace cumf_multi

extern "C" int cumf_als_peer_barrier(cumf_als_solver* s, void* stream) {
    CUMF_REQUIRE(s, "null pointer");
    if (s->nranks <= 1) return CUMF_OK;
    if (s->same_dev) {
        ++s->epoch;
        const int par = (int)(s->epoch & 1);
        const cudaError_t e = cudaEventRecord(s->same_dev->ev[s->rank][par], (cudaStream_t)stream);
        s->same_dev->hb.wait();                                   // every shard's record of this epoch is enqueued
        CUMF_CUDA_TRY(e);
        for (int r = 0; r < s->nranks; ++r)
            if (r != s->rank) CUMF_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s->same_dev->ev[r][par], 0));
        return CUMF_OK;
    }
    cumf_multi::FlagPtrs fp;
    for (int r = 0; r < 8; ++r) fp.p[r] = s->peer_flags[r];
    ++s->epoch;
    cumf_multi::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fp, s->rank, s->nranks, s->epoch);
    CUMF_CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    return CUMF_OK;
}

// blob = 4 CUDA IPC handles of 64 bytes (X replica, theta replica, barrier flags, the allocation that holds the X side's gather
// table) + two 8-byte words (offset of the table in that allocation; 1 if there is such a table, else the 4th handle is unused)
namespace { constexpr int kIpcHandles = 4; }
extern "C" int cumf_als_ipc_blob_bytes(void) { return kIpcHandles * (int)sizeof(cudaIpcMemHandle_t) + 16; }


static int solver_alloc_flags(cumf_als_solver* s) {
    if (s->flags.p) return CUMF_OK;
    CUMF_TRY(s->flags.alloc_own(8 * sizeof(unsigned long long)));
    CUMF_CUDA_TRY(cudaMemset(s->flags.p, 0, 8 * sizeof(unsigned long long)));
    return CUMF_OK;
}

extern "C" int cumf_als_ipc_export(cumf_als_solver* s, void* blob) {
    CUMF_REQUIRE(s && blob, "null pointer");
    CUMF_CUDA_TRY(cudaSetDevice(s->device));
    CUMF_REQUIRE(!s->x.in_arena && !s->theta.in_arena, "a shard of a one-process group cannot be exported (its replicas are part of a larger allocation)");
    CUMF_TRY(solver_alloc_flags(s));
    cudaIpcMemHandle_t* h = reinterpret_cast<cudaIpcMemHandle_t*>(blob);
    CUMF_CUDA_TRY(cudaIpcGetMemHandle(&h[0], s->x.p));
    CUMF_CUDA_TRY(cudaIpcGetMemHandle(&h[1], s->theta.p));
    CUMF_CUDA_TRY(cudaIpcGetMemHandle(&h[2], s->flags.p));
    unsigned long long* words = reinterpret_cast<unsigned long long*>(h + kIpcHandles);
    words[0] = words[1] = 0;
    memset(&h[3], 0, sizeof(h[3]));
    if (unsigned short* tab = own_gather_table(s)) {
        // the table is usually carved out of the solver's arena: export the whole allocation and say where the table starts
        unsigned char* base = reinterpret_cast<unsigned char*>(tab);
        if (s->arena.base && base >= s->arena.base && base < s->arena.base + s->arena.cap) base = s->arena.base;
        if (cudaIpcGetMemHandle(&h[3], base) == cudaSuccess) {
            words[0] = (unsigned long long)(reinterpret_cast<unsigned char*>(tab) - base);
            words[1] = 1;
        } else {
            cudaGetLastError();
        }
    }
    return CUMF_OK;
}

// blobs: nranks blobs in rank order (this rank's own entry is ignored).  After this call every half-step of this solver
// also updates the peers' replicas, and cumf_als_iterate ends every half-step with the device-side barrier; every rank must
// call it, with the same nranks, before any of them runs a half-step.
extern "C" int cumf_als_ipc_import(cumf_als_solver* s, const void* blobs, int nranks, int my_rank) {
    CUMF_REQUIRE(s && blobs, "null pointer");
    CUMF_REQUIRE(nranks >= 1 && nranks <= 8 && my_rank >= 0 && my_rank < nranks, "1 <= nranks <= 8 ranks are supported");
    CUMF_CUDA_TRY(cudaSetDevice(s->device));
    CUMF_TRY(solver_alloc_flags(s));
    s->rank = my_rank;
    s->nranks = nranks;
    s->peer_x.n = s->peer_theta.n = 0;
    s->peer_flags[my_rank] = s->flags.as<unsigned long long>();
    const size_t stride = (size_t)cumf_als_ipc_blob_bytes();
    auto blob_of = [&](int r) { return reinterpret_cast<const cudaIpcMemHandle_t*>(static_cast<const unsigned char*>(blobs) + stride * r); };
    auto words_of = [&](int r) { return reinterpret_cast<const unsigned long long*>(blob_of(r) + kIpcHandles); };
    // split rows are pushed only if EVERY rank has a gather table to receive them (all ranks see the same blobs, so they agree)
    bool tables = own_gather_table(s) != nullptr;
    for (int r = 0; r < nranks; ++r) tables = tables && (r == my_rank || words_of(r)[1] == 1);
    s->theta_split_out.n = 0;
    if (tables) s->theta_split_out.p[s->theta_split_out.n++] = own_gather_table(s);
    for (int r = 0; r < nranks; ++r) {
        if (r == my_rank) continue;
        const cudaIpcMemHandle_t* h = blob_of(r);
        void* p[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < 3; ++k) {
            CUMF_CUDA_TRY(cudaIpcOpenMemHandle(&p[k], h[k], cudaIpcMemLazyEnablePeerAccess));
            s->ipc_opened.push_back(p[k]);
        }
        s->peer_x.p[s->peer_x.n++] = reinterpret_cast<float*>(p[0]);
        s->peer_theta.p[s->peer_theta.n++] = reinterpret_cast<float*>(p[1]);
        s->peer_flags[r] = reinterpret_cast<unsigned long long*>(p[2]);
        if (tables) {
            void* base = nullptr;
            CUMF_CUDA_TRY(cudaIpcOpenMemHandle(&base, h[3], cudaIpcMemLazyEnablePeerAccess));
            s->ipc_opened.push_back(base);
            s->theta_split_out.p[s->theta_split_out.n++] = reinterpret_cast<unsigned short*>(static_cast<unsigned char*>(base) + words_of(r)[0]);
        }
    }
    s->theta_table_current = false;
    return CUMF_OK;
}

// ---- one process, n devices -------------------------------------------------------------------------------------------
struct cumf_synth_shard;
extern "C" int cumf_synth_create(cumf_synth_shard** out, long long m, int n, float avg_deg, unsigned long long seed, int x0, int x1,
                                 int t0, int t1, long test_cnt, int device);
extern "C" int cumf_synth_destroy(cumf_synth_shard* sh);
extern "C" int cumf_synth_solver(cumf_synth_shard* sh, cumf_als_solver** out, int f, float lambda, long nnz_test_total, int solver, int path);
extern "C" long long cumf_synth_total_nnz(const cumf_synth_shard* sh);
extern "C" int cumf_als_init_factors_device(cumf_als_solver* s, unsigned long long seed, float scale);

struct cumf_als_group {
    std::vector<cumf_als_solver*> s;
    std::vector<cumf_synth_shard*> synth;     // device-resident shards the solvers borrow (cumf_group_create_synth)
    int m = 0, n = 0, f = 0;
    long nnz = 0, nnz_test = 0;
    std::vector<std::string> errors;        // per shard (g_last_error is thread-local)
    std::unique_ptr<cumf_multi::SameDeviceSync> same_dev;
};

// contiguous row ranges with (nearly) equal numbers of ratings: the deterministic replacement of hugewiki's dynamic batch
// queue (hugewiki.cu:2446-2496); same integer arithmetic as cumf_als_b200.data.nnz_balanced_ranges
static std::vector<std::pair<int, int>> nnz_balanced_ranges(const int* ptr, int rows, int parts) {
    std::vector<int> bounds(1, 0);
    const long long total = ptr[rows];
    for (int g = 1; g < parts; ++g) {
        const long long target = total * g / parts;
        int b = (int)(std::lower_bound(ptr, ptr + rows + 1, target, [](int a, long long t) { return (long long)a < t; }) - ptr);
        b = std::min(std::max(b, bounds.back()), rows);
        bounds.push_back(b);
    }
    bounds.push_back(rows);
    std::vector<std::pair<int, int>> out;
    for (int g = 0; g < parts; ++g) out.emplace_back(bounds[g], bounds[g + 1]);
    return out;
}

template <typename Fn> static void for_each_shard_parallel(int n, Fn fn) {
    std::vector<std::thread> th;
    for (int g = 1; g < n; ++g) th.emplace_back([=]() { fn(g); });
    fn(0);
    for (auto& t : th) t.join();
}

extern "C" int cumf_group_destroy(cumf_als_group* g) {
    if (!g) return CUMF_OK;
    // every device must be idle before any replica goes away: the peers' epilogues write into it
    for (auto* s : g->s) if (s) { cudaSetDevice(s->device); cudaDeviceSynchronize(); }
    for_each_shard_parallel((int)g->s.size(), [&](int k) {
        if (g->s[k]) cumf_als_destroy(g->s[k]);
        if (k < (int)g->synth.size() && g->synth[k]) cumf_synth_destroy(g->synth[k]);
    });
    delete g;
    return CUMF_OK;
}

// rank / peer tables of a fully created group (same process: plain device pointers, peer access enabled by the caller)
static void group_connect(cumf_als_group* g) {
    const int n = (int)g->s.size();
    bool one_device = n > 1;
    for (int k = 1; k < n; ++k) one_device = one_device && g->s[k]->device == g->s[0]->device;
    if (one_device) {
        cudaSetDevice(g->s[0]->device);
        g->same_dev.reset(new cumf_multi::SameDeviceSync(n));
    }
    bool all_tables = true;       // split rows are pushed only if every shard has a gather table to receive them
    for (int k = 0; k < n; ++k) all_tables = all_tables && own_gather_table(g->s[k]) != nullptr;
    for (int k = 0; k < n; ++k) {
        cumf_als_solver* s = g->s[k];
        s->rank = k;
        s->nranks = n;
        s->same_dev = g->same_dev.get();
        s->peer_x.n = s->peer_theta.n = 0;
        s->theta_split_out.n = 0;
        s->theta_table_current = false;
        if (all_tables) s->theta_split_out.p[s->theta_split_out.n++] = own_gather_table(s);
        for (int j = 0; j < n; ++j) {
            s->peer_flags[j] = g->s[j]->flags.as<unsigned long long>();
            if (j == k) continue;
            s->peer_x.p[s->peer_x.n++] = g->s[j]->x.as<float>();
            s->peer_theta.p[s->peer_theta.n++] = g->s[j]->theta.as<float>();
            if (all_tables) s->theta_split_out.p[s->theta_split_out.n++] = own_gather_table(g->s[j]);
        }
    }
}
static int shard_runtime_setup(cumf_als_solver* s, int first_device, int n_devices, bool same_device) {
    CUMF_TRY(solver_alloc_flags(s));
    if (cudaStreamCreateWithFlags(&s->run_stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_last_error("cumf_group_create: cannot create a stream");
        return CUMF_ECUDA;
    }
    if (!same_device)
        for (int j = 0; j < n_devices; ++j) {
            if (first_device + j == s->device) continue;
            const cudaError_t e = cudaDeviceEnablePeerAccess(first_device + j, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { set_last_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return CUMF_ECUDA; }
        }
    return CUMF_OK;
}

// A group over a synthetic matrix generated shard by shard ON the devices (synth.cu): the Hugewiki-scale configuration of
// BASELINE.json (m ~ 50 M, n ~ 40 K, 3.1 G ratings over 8 GPUs) never exists as one host copy.  Rows and columns are split
// evenly (the generator's degrees are i.i.d., so equal row counts are rating-balanced); theta0 = 0.2 * uniform, X0 = 0.
extern "C" int cumf_group_create_synth(cumf_als_group** out, long long m, int n, float avg_deg, unsigned long long seed,
                                       long test_per_shard, int f, float lambda, int first_device, int n_devices, int solver,
                                       int path) {
    CUMF_REQUIRE(out && m > 0 && m < (1ll << 31) && n > 0, "bad matrix spec");
    CUMF_REQUIRE(n_devices >= 1 && n_devices <= 8, "1 .. 8 devices");
    int have = 0;
    CUMF_CUDA_TRY(cudaGetDeviceCount(&have));
    const bool same_device = env_long("CUMF_GROUP_SAME_DEVICE", 0) != 0;
    CUMF_REQUIRE(first_device >= 0 && first_device + (same_device ? 1 : n_devices) <= have, "not that many devices on this node");
    cumf_als_group* g = new cumf_als_group();
    g->m = (int)m; g->n = n; g->f = f; g->nnz_test = test_per_shard * n_devices;
    g->s.assign(n_devices, nullptr);
    g->synth.assign(n_devices, nullptr);
    g->errors.assign(n_devices, std::string());
    std::vector<int> rcs(n_devices, CUMF_OK);
    for_each_shard_parallel(n_devices, [&](int k) {
        const int dev = same_device ? first_device : first_device + k;
        const int x0 = (int)(m * k / n_devices), x1 = (int)(m * (k + 1) / n_devices);
        const int t0 = (int)((long long)n * k / n_devices), t1 = (int)((long long)n * (k + 1) / n_devices);
        rcs[k] = cumf_synth_create(&g->synth[k], m, n, avg_deg, seed, x0, x1, t0, t1, test_per_shard, dev);
        if (rcs[k] == CUMF_OK) rcs[k] = cumf_synth_solver(g->synth[k], &g->s[k], f, lambda, (long)g->nnz_test, solver, path);
        if (rcs[k] == CUMF_OK) rcs[k] = shard_runtime_setup(g->s[k], first_device, n_devices, same_device);
        if (rcs[k] == CUMF_OK) rcs[k] = cumf_als_init_factors_device(g->s[k], seed ^ 0xFAC70125ull, 0.2f);
        if (rcs[k] != CUMF_OK) g->errors[k] = cumf_last_error();
    });
    for (int k = 0; k < n_devices; ++k)
        if (rcs[k] != CUMF_OK) {
            set_last_error("shard " + std::to_string(k) + ": " + g->errors[k]);
            const int rc = rcs[k];
            cumf_group_destroy(g);
            return rc;
        }
    g->nnz = (long)cumf_synth_total_nnz(g->synth[0]);
    group_connect(g);
    *out = g;
    return CUMF_OK;
}
extern "C" long cumf_group_nnz(const cumf_als_group* g) { return g ? g->nnz : -1; }
extern "C" long cumf_group_nnz_test(const cumf_als_group* g) { return g ? g->nnz_test : -1; }

static int group_create_impl(cumf_als_group** out, const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr,
                             const float* csrValHostPtr, const int* cscRowIndexHostPtr, const int* cscColIndexHostPtr,
                             const float* cscValHostPtr, const int* cooRowIndexHostPtr, const int* cooRowIndexTestHostPtr,
                             const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, int m, int n, int f, long nnz,
                             long nnz_test, float lambda, int first_device, int n_devices, int solver, int path,
                             bool wait_uploads, const float* thetaTHost, const float* XTHost) {
    CUMF_REQUIRE(out && csrRowIndexHostPtr && cscColIndexHostPtr, "null pointer");
    CUMF_REQUIRE(n_devices >= 1 && n_devices <= 8, "1 .. 8 devices");
    int have = 0;
    CUMF_CUDA_TRY(cudaGetDeviceCount(&have));
    // test hook: every shard on first_device (own stream each), so one GPU exercises the peer stores and the barrier
    const bool same_device = env_long("CUMF_GROUP_SAME_DEVICE", 0) != 0;
    CUMF_REQUIRE(first_device >= 0 && first_device + (same_device ? 1 : n_devices) <= have, "not that many devices on this node");
    auto device_of = [=](int k) { return same_device ? first_device : first_device + k; };
    cumf_als_group* g = new cumf_als_group();
    g->m = m; g->n = n; g->f = f; g->nnz = nnz; g->nnz_test = nnz_test;
    g->s.assign(n_devices, nullptr);
    g->errors.assign(n_devices, std::string());
    const auto xr = nnz_balanced_ranges(csrRowIndexHostPtr, m, n_devices);
    const auto tr = nnz_balanced_ranges(cscColIndexHostPtr, n, n_devices);
    std::vector<int> rcs(n_devices, CUMF_OK);
    // one host thread per device: work plans, allocations and the (asynchronous) uploads of all shards proceed in parallel,
    // every GPU pulls its slice over its own PCIe link
    const bool debug = env_long("CUMF_DEBUG", 0) != 0;
    for_each_shard_parallel(n_devices, [&](int k) {
        const double t0 = wall_seconds();
        t_group_member = true;
        struct Reset { ~Reset() { t_group_member = false; } } reset_member;
        rcs[k] = als_create_impl(&g->s[k], csrRowIndexHostPtr, csrColIndexHostPtr, csrValHostPtr, cscRowIndexHostPtr,
                                 cscColIndexHostPtr, cscValHostPtr, cooRowIndexHostPtr, cooRowIndexTestHostPtr,
                                 cooColIndexTestHostPtr, cooValHostTestPtr, m, n, f, nnz, nnz_test, lambda, xr[k].first,
                                 xr[k].second, tr[k].first, tr[k].second, device_of(k), solver, path, wait_uploads,
                                 thetaTHost, XTHost);
        const double t1 = wall_seconds();
        if (rcs[k] == CUMF_OK) rcs[k] = shard_runtime_setup(g->s[k], first_device, n_devices, same_device);
        if (rcs[k] != CUMF_OK) g->errors[k] = cumf_last_error();
        if (debug) printf("\tshard %d: solver %.4f s, stream + peer access %.4f s\n", k, t1 - t0, wall_seconds() - t1);
    });
    for (int k = 0; k < n_devices; ++k)
        if (rcs[k] != CUMF_OK) {
            set_last_error("shard " + std::to_string(k) + ": " + g->errors[k]);
            const int rc = rcs[k];
            cumf_group_destroy(g);
            return rc;
        }
    group_connect(g);
    *out = g;
    return CUMF_OK;
}

extern "C" int cumf_group_create(cumf_als_group** out, const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr,
                                 const float* csrValHostPtr, const int* cscRowIndexHostPtr, const int* cscColIndexHostPtr,
                                 const float* cscValHostPtr, const int* cooRowIndexHostPtr, const int* cooRowIndexTestHostPtr,
                                 const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, int m, int n, int f, long nnz,
                                 long nnz_test, float lambda, int first_device, int n_devices, int solver, int path) {
    return group_create_impl(out, csrRowIndexHostPtr, csrColIndexHostPtr, csrValHostPtr, cscRowIndexHostPtr, cscColIndexHostPtr,
                             cscValHostPtr, cooRowIndexHostPtr, cooRowIndexTestHostPtr, cooColIndexTestHostPtr, cooValHostTestPtr,
                             m, n, f, nnz, nnz_test, lambda, first_device, n_devices, solver, path, /*wait_uploads=*/true, nullptr,
                             nullptr);
}

extern "C" int cumf_group_size(const cumf_als_group* g) { return g ? (int)g->s.size() : 0; }
extern "C" cumf_als_solver* cumf_group_shard(cumf_als_group* g, int k) { return (g && k >= 0 && k < (int)g->s.size()) ? g->s[k] : nullptr; }

extern "C" int cumf_group_set_factors(cumf_als_group* g, const float* thetaTHost, const float* XTHost) {
    CUMF_REQUIRE(g, "null pointer");
    for (auto* s : g->s) CUMF_TRY(cumf_als_set_factors(s, thetaTHost, XTHost));
    return CUMF_OK;
}
// every replica holds both full factors after a half-step's barrier: one download
extern "C" int cumf_group_get_factors(cumf_als_group* g, float* thetaTHost, float* XTHost) {
    CUMF_REQUIRE(g && !g->s.empty(), "null pointer");
    for (auto* s : g->s) { CUMF_CUDA_TRY(cudaSetDevice(s->device)); CUMF_CUDA_TRY(cudaDeviceSynchronize()); }
    return cumf_als_get_factors(g->s[0], thetaTHost, XTHost);
}

// `iters` iterations on every shard (asynchronous launches device by device: the devices only meet in the barrier kernels);
// ms_out = device time of the slowest shard
extern "C" int cumf_group_iterate(cumf_als_group* g, int iters, float* ms_out) {
    CUMF_REQUIRE(g && iters >= 0, "bad argument");
    const int n = (int)g->s.size();
    std::vector<float> ms(n, 0.f);
    std::vector<int> rcs(n, CUMF_OK);
    for_each_shard_parallel(n, [&](int k) {
        cudaSetDevice(g->s[k]->device);
        rcs[k] = cumf_als_iterate(g->s[k], iters, &ms[k], g->s[k]->run_stream);
        if (rcs[k] != CUMF_OK) g->errors[k] = cumf_last_error();
    });
    for (int k = 0; k < n; ++k)
        if (rcs[k] != CUMF_OK) { set_last_error("shard " + std::to_string(k) + ": " + g->errors[k]); return rcs[k]; }
    if (ms_out) *ms_out = *std::max_element(ms.begin(), ms.end());
    return CUMF_OK;
}

extern "C" int cumf_group_collect_train_sse(cumf_als_group* g, int on) {
    CUMF_REQUIRE(g, "null pointer");
    int all = 1;
    for (auto* s : g->s) all &= cumf_als_collect_train_sse(s, on);
    if (!all) for (auto* s : g->s) cumf_als_collect_train_sse(s, 0);       // all shards or none: they must count the same sample set
    return all;
}

extern "C" int cumf_group_sse(cumf_als_group* g, double* train_sse, double* test_sse) {
    CUMF_REQUIRE(g, "null pointer");
    const int n = (int)g->s.size();
    std::vector<double> tr(n, 0.0), te(n, 0.0);
    std::vector<int> rcs(n, CUMF_OK);
    for_each_shard_parallel(n, [&](int k) {
        cudaSetDevice(g->s[k]->device);
        rcs[k] = cumf_als_sse(g->s[k], train_sse ? &tr[k] : nullptr, test_sse ? &te[k] : nullptr, g->s[k]->run_stream);
        if (rcs[k] != CUMF_OK) g->errors[k] = cumf_last_error();
    });
    double a = 0.0, b = 0.0;
    for (int k = 0; k < n; ++k) {
        if (rcs[k] != CUMF_OK) { set_last_error("shard " + std::to_string(k) + ": " + g->errors[k]); return rcs[k]; }
        a += tr[k]; b += te[k];          // rank order: deterministic
    }
    if (train_sse) *train_sse = a;
    if (test_sse) *test_sse = b;
    return CUMF_OK;
}

// ---------------------------------------------------------------------------------
// b1: doALS.  Same signature and observable behaviour as als.cu:662-1035: blocking,
// host pointers in, factors written back, last test RMSE returned, progress on stdout
// in the line formats the reference's scripts parse (print-test-result.sh:8-12).
// On an unrecoverable error it prints and exits like the cudacall macro (als.h:628-640).
// ---------------------------------------------------------------------------------
[[noreturn]] static void die(const char* where) {
    fprintf(stderr, "cumf_als_b200 error in %s: %s\n", where, cumf_last_error());
    cudaDeviceReset();
    exit(EXIT_FAILURE);
}

// doALS over CUMF_GPUS devices of this node (DEVICEID = the first): one host thread per device runs the iteration loop of its
// shard; the devices meet in the barrier kernels, the host threads only to add up the RMSE sums.  Same stdout contract.

static float doALS_multi(const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr, const float* csrValHostPtr,
                         const int* cscRowIndexHostPtr, const int* cscColIndexHostPtr, const float* cscValHostPtr,
                         const int* cooRowIndexHostPtr, float* thetaTHost, float* XTHost, const int* cooRowIndexTestHostPtr,
                         const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, const int m, const int n, const int f,
                         const long nnz, const long nnz_test, const float lambda, const int ITERS, const int DEVICEID,
                         const int gpus, const int solver, const int path, const bool quiet, const bool debug) {
    cumf_als_group* g = nullptr;
    const double t_setup = wall_seconds();
    if (group_create_impl(&g, csrRowIndexHostPtr, csrColIndexHostPtr, csrValHostPtr, cscRowIndexHostPtr, cscColIndexHostPtr,
                          cscValHostPtr, cooRowIndexHostPtr, cooRowIndexTestHostPtr, cooColIndexTestHostPtr, cooValHostTestPtr, m, n,
                          f, nnz, nnz_test, lambda, DEVICEID, gpus, solver, path, /*wait_uploads=*/false, thetaTHost, XTHost) != CUMF_OK)
        die("cumf_group_create");
    cumf_group_collect_train_sse(g, 1);
    if (debug) printf("\tsetup of %d shards (work plans; uploads continue in the background) run %f seconds.\n", gpus, wall_seconds() - t_setup);
    if (!quiet) printf("*******start iterations on %d GPUs (rows sharded by rating count)...\n", gpus);
    cumf_multi::HostBarrier hb;
    hb.n = gpus;
    std::vector<double> tr(2 * (size_t)gpus, 0.0), te(2 * (size_t)gpus, 0.0);
    std::vector<int> rcs(gpus, CUMF_OK);
    std::vector<std::string> errs(gpus);
    float final_rmse = 0.f;
    auto run = [&](int k) {
        cumf_als_solver* s = g->s[k];
        cudaSetDevice(s->device);
        auto step = [&](int rc) { if (rc != CUMF_OK && rcs[k] == CUMF_OK) { rcs[k] = rc; errs[k] = cumf_last_error(); } return rcs[k] == CUMF_OK; };
        // nobody pushes rows into a replica whose initial upload is still in flight
        void* st = s->run_stream;
        cudaStreamWaitEvent(s->run_stream, s->ev_csr, 0);
        step(cumf_als_peer_barrier(s, st));
        for (int iter = 0; iter < ITERS; ++iter) {
            double t0 = wall_seconds();
            // the barriers are taken by every shard, failed or not: the others are waiting in theirs
            auto barrier = [&]() { const int b = cumf_als_peer_barrier(s, st); if (rcs[k] == CUMF_OK) step(b); };
            {
                if (rcs[k] == CUMF_OK) step(cumf_als_update_x(s, st));
                barrier();
                if (debug && k == 0) {
                    cudaStreamSynchronize(s->run_stream);
                    printf("---------------------------ALS iteration %d, update X.----------------------------------\n", iter);
                    printf("update X run %f seconds, gridSize: %d, blockSize %d.\n", wall_seconds() - t0, m, f);
                    t0 = wall_seconds();
                }
                if (rcs[k] == CUMF_OK) step(cumf_als_update_theta(s, st));
                barrier();
                if (debug && k == 0) {
                    cudaStreamSynchronize(s->run_stream);
                    printf("---------------------------------- ALS iteration %d, update theta ----------------------------------\n", iter);
                    printf("update theta run %f seconds, gridSize: %d, blockSize %d.\n", wall_seconds() - t0, n, f);
                    printf("Calculate RMSE.\n");
                }
                if (rcs[k] == CUMF_OK)
                    step(cumf_als_sse(s, cooRowIndexHostPtr ? &tr[(iter & 1) * gpus + k] : nullptr, &te[(iter & 1) * gpus + k], st));
            }
            hb.wait();          // every shard's sums of this iteration are in (a failed shard still takes part)
            if (k == 0) {
                double a = 0.0, b = 0.0;
                for (int j = 0; j < gpus; ++j) { a += tr[(iter & 1) * gpus + j]; b += te[(iter & 1) * gpus + j]; }
                const float rmse_train = sqrtf((float)a / (float)nnz);        // als.cu:991
                final_rmse = sqrtf((float)b / (float)nnz_test);                // als.cu:1018
                if (!quiet) {
                    printf("--------- Train RMSE in iter %d: %f\n", iter, rmse_train);
                    printf("--------- Test RMSE in iter %d: %f\n", iter, final_rmse);
                }
            }
        }
    };
    for_each_shard_parallel(gpus, run);
    {
        std::string all;
        for (int k = 0; k < gpus; ++k)
            if (rcs[k] != CUMF_OK) all += "shard " + std::to_string(k) + " (status " + std::to_string(rcs[k]) + "): " + errs[k] + "; ";
        if (!all.empty()) { set_last_error(all); die("multi-GPU iteration"); }
    }
    const double t_down = wall_seconds();
    if (cumf_group_get_factors(g, thetaTHost, XTHost) != CUMF_OK) die("cumf_group_get_factors");
    const double t_free = wall_seconds();
    cumf_group_destroy(g);
    if (debug) printf("\tfactor download run %f seconds, release of the device buffers %f seconds.\n", t_free - t_down, wall_seconds() - t_free);
    return final_rmse;
}

float doALS(const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr, const float* csrValHostPtr,
            const int* cscRowIndexHostPtr, const int* cscColIndexHostPtr, const float* cscValHostPtr,
            const int* cooRowIndexHostPtr, float* thetaTHost, float* XTHost, const int* cooRowIndexTestHostPtr,
            const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, const int m, const int n, const int f,
            const long nnz, const long nnz_test, const float lambda, const int ITERS, const int X_BATCH,
            const int THETA_BATCH, const int DEVICEID) {
    const bool quiet = env_long("CUMF_QUIET", 0) != 0;
    const bool debug = env_long("CUMF_DEBUG", 0) != 0;
    const int solver = env_choice("CUMF_SOLVER", "cg", CUMF_SOLVER_CG, "lu", CUMF_SOLVER_LU, nullptr, 0, CUMF_SOLVER_CG);
    const int path = env_choice("CUMF_PATH", "auto", CUMF_PATH_AUTO, "simt", CUMF_PATH_SIMT, "tc", CUMF_PATH_TC, CUMF_PATH_AUTO);
    (void)X_BATCH; (void)THETA_BATCH;   // advisory: nothing is batched to fit a 12 GB card any more
    if (!quiet) {
        printf("*******parameters: m: %d, n:  %d, f: %d, nnz: %ld \n", m, n, f, nnz);
        printf("*******B200 path: solver %s, kernels %s; uploading CSR/CSC/COO once (resident)...\n",
               solver == CUMF_SOLVER_CG ? "CG" : "LU(cuBLAS oracle)",
               path == CUMF_PATH_SIMT ? "simt" : (path == CUMF_PATH_TC ? "tcgen05" : "auto"));
    }
    // CUMF_GPUS=n: the X_BATCH / THETA_BATCH model-parallel split (als.cu:768-777) becomes a row sharding over n GPUs
    const int gpus = (int)env_long("CUMF_GPUS", 1);
    if (gpus > 1)
        return doALS_multi(csrRowIndexHostPtr, csrColIndexHostPtr, csrValHostPtr, cscRowIndexHostPtr, cscColIndexHostPtr,
                           cscValHostPtr, cooRowIndexHostPtr, thetaTHost, XTHost, cooRowIndexTestHostPtr, cooColIndexTestHostPtr,
                           cooValHostTestPtr, m, n, f, nnz, nnz_test, lambda, ITERS, DEVICEID, gpus, solver, path, quiet, debug);
    cumf_als_solver* s = nullptr;
    const double t_setup = wall_seconds();
    // the host arrays outlive this call, so the uploads may still be in flight when the iterations start
    if (als_create_impl(&s, csrRowIndexHostPtr, csrColIndexHostPtr, csrValHostPtr, cscRowIndexHostPtr,
                        cscColIndexHostPtr, cscValHostPtr, cooRowIndexHostPtr, cooRowIndexTestHostPtr,
                        cooColIndexTestHostPtr, cooValHostTestPtr, m, n, f, nnz, nnz_test, lambda, 0, m, 0, n,
                        DEVICEID, solver, path, /*wait_uploads=*/false, thetaTHost, XTHost) != CUMF_OK)
        die("cumf_als_create");
    cumf_als_collect_train_sse(s, 1);
    if (debug) {
        printf("\tsetup (work plans; uploads continue in the background) run %f seconds.\n", wall_seconds() - t_setup);
        s->px->time_kernel = s->pt->time_kernel = true;     // CUDA events around the Gram(+fused solve) kernel
    }
    if (!quiet) printf("*******start iterations...\n");
    float final_rmse = 0.f;
    double kx_ms = 0.0, kt_ms = 0.0;      // running totals of the timed kernels (debug)
    for (int iter = 0; iter < ITERS; ++iter) {
        double t0 = wall_seconds();
        if (debug) printf("---------------------------ALS iteration %d, update X.----------------------------------\n", iter);
        if (cumf_als_update_x(s, nullptr) != CUMF_OK) die("update X");
        if (debug) {
            // the reference's -DDEBUG line set (als.cu:821, 827/830, 844-845, 850), one batch: hermitiantime.sh sums
            // field 5 of the "kernel run" lines, solvertime.sh field 5 of the "solver run" lines, print-test-result.sh
            // field 4 of the "update X/theta run" lines.  "kernel" = the Gram kernel (fused path: Gram + RHS + CG in one
            // launch), "solver" = whatever ran after it (the CG / LU kernels of the unfused path, the split-row tail).
            cudaStreamSynchronize(0);
            const double wall = wall_seconds() - t0;
            const double k = plan_collect_kernel_ms(s->px);
            const double kern = (k - kx_ms) * 1e-3;
            kx_ms = k;
            printf("\tupdate X kernel run %f seconds, gridSize: %d, blockSize %d.\n", kern, m, f);
            printf(solver == CUMF_SOLVER_CG ? "\tCG solver with fp32.\n" : "\tLU solver (cuBLAS getrfBatched).\n");
            printf("\tinvoke updateX with batch_size: %d, batch_offset: %d..\n", m, 0);
            printf("\tupdateX solver run seconds: %f \n", std::max(0.0, wall - kern));
            printf("update X run %f seconds, gridSize: %d, blockSize %d.\n", wall, m, f);
            t0 = wall_seconds();
            printf("---------------------------------- ALS iteration %d, update theta ----------------------------------\n", iter);
        }
        if (cumf_als_update_theta(s, nullptr) != CUMF_OK) die("update theta");
        if (debug) {
            cudaStreamSynchronize(0);      // als.cu:928, 931, 942/945, 957, 961-963
            const double wall = wall_seconds() - t0;
            const double k = plan_collect_kernel_ms(s->pt);
            const double kern = (k - kt_ms) * 1e-3;
            kt_ms = k;
            printf("\tupdate Theta kernel run %f seconds, gridSize: %d, blockSize %d.\n", kern, n, f);
            printf("*******invoke updateTheta with batch_size: %d, batch_offset: %d.\n", n, 0);
            printf(solver == CUMF_SOLVER_CG ? "\tCG solver with fp32.\n" : "\tLU solver (cuBLAS getrfBatched).\n");
            printf("\tupdateTheta solver run seconds: %f \n", std::max(0.0, wall - kern));
            printf("update theta run %f seconds, gridSize: %d, blockSize %d.\n", wall, n, f);
            printf("Calculate RMSE.\n");
        }
        double tr = 0.0, te = 0.0;
        t0 = wall_seconds();
        if (cumf_als_sse(s, cooRowIndexHostPtr ? &tr : nullptr, &te, nullptr) != CUMF_OK) die("RMSE");
        if (debug) printf("RMSE run %f seconds (train walk mode %d).\n", wall_seconds() - t0, s->train_mode);
        const float rmse_train = sqrtf((float)tr / (float)nnz);          // als.cu:991
        final_rmse = sqrtf((float)te / (float)nnz_test);                  // als.cu:1018
        if (!quiet) {
            printf("--------- Train RMSE in iter %d: %f\n", iter, rmse_train);
            printf("--------- Test RMSE in iter %d: %f\n", iter, final_rmse);
        }
    }
    const double t_down = wall_seconds();
    if (cumf_als_get_factors(s, thetaTHost, XTHost) != CUMF_OK) die("cumf_als_get_factors");   // als.cu:1024-1025
    const double t_free = wall_seconds();
    cumf_als_destroy(s);
    if (debug) printf("\tfactor download run %f seconds, release of the device buffers %f seconds.\n", t_free - t_down, wall_seconds() - t_free);
    // like the reference, do NOT cudaDeviceReset here: the caller owns the context (als.cu:1031-1033)
    return final_rmse;
}

extern "C" float cumf_doALS(const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr, const float* csrValHostPtr,
                            const int* cscRowIndexHostPtr, const int* cscColIndexHostPtr, const float* cscValHostPtr,
                            const int* cooRowIndexHostPtr, float* thetaTHost, float* XTHost,
                            const int* cooRowIndexTestHostPtr, const int* cooColIndexTestHostPtr,
                            const float* cooValHostTestPtr, const int m, const int n, const int f, const long nnz,
                            const long nnz_test, const float lambda, const int ITERS, const int X_BATCH,
                            const int THETA_BATCH, const int DEVICEID) {
    return doALS(csrRowIndexHostPtr, csrColIndexHostPtr, csrValHostPtr, cscRowIndexHostPtr, cscColIndexHostPtr,
                 cscValHostPtr, cooRowIndexHostPtr, thetaTHost, XTHost, cooRowIndexTestHostPtr, cooColIndexTestHostPtr,
                 cooValHostTestPtr, m, n, f, nnz, nnz_test, lambda, ITERS, X_BATCH, THETA_BATCH, DEVICEID);
}
