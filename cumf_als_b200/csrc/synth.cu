// synth.cu -- on-device, shard-by-shard synthetic rating matrices (bench / test tooling behind the C ABI).
//
// BASELINE.json's largest configuration (Hugewiki-scale: m ~ 50 M, n ~ 40 K, 3.1 G ratings over 8 GPUs; hugewiki.cu:27-42)
// cannot pass through one host copy of the matrix the way the reference's loaders do (host_utilities.cpp:19-98,
// hugewiki.cu:2332-2340 reads per-GPU batch files).  Here every GPU derives ITS slices of one global matrix directly in
// device memory, with no communication:
//   * the matrix is a pure function of (seed, row, column): row u holds d_u ratings (exponential degree distribution with the
//     requested mean, >= 1), its k-th column lies in the k-th of d_u equal strata of [0, n) (ascending, unique), the value is
//     a clamped low-rank score + noise in {1..5};
//   * CSR slice of rows [x0, x1): degrees -> exclusive scan -> one warp per row fills columns / values;
//   * CSC slice of columns [t0, t1): every row's strata that intersect [t0, t1) are enumerated (O(d_u (t1-t0)/n) per row),
//     compacted row by row (deterministic offsets from a scan of the per-row counts), and sorted by (column, row) with one
//     radix sort -- the order a host transposition gives, so the device path and a host-built CSC agree bit for bit.
// Everything the ALS path needs afterwards is the slices themselves (cumf_als_create_device borrows them).
#include <cub/cub.cuh>

#include "common.cuh"

namespace cumf {
namespace {

struct Spec {
    long long m;
    int n;
    float avg_deg;
    unsigned long long seed;
};

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {      // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ float u01(unsigned long long h) { return (float)((h >> 40) + 0.5) * (1.0f / 16777216.0f); }

__host__ __device__ __forceinline__ int degree_of(const Spec& s, long long u) {
    const float h = u01(mix64(s.seed ^ (unsigned long long)u * 0xD6E8FEB86659FD93ull));
    // exponential with mean avg_deg - 1, plus one: a long tail like rating data, every row non-empty (README.md:113)
    const float d = 1.0f + (s.avg_deg - 1.0f) * (-logf(1.0f - h));
    int di = (int)d;
    if (di < 1) di = 1;
    if (di > s.n) di = s.n;
    if (di > 65535) di = 65535;
    return di;
}
// k-th column of row u (d = degree_of(u)): uniform inside stratum [lo_k, lo_{k+1}), lo_k = floor(k n / d)
__host__ __device__ __forceinline__ int column_of(const Spec& s, long long u, int d, int k) {
    const long long lo = (long long)k * s.n / d, hi = (long long)(k + 1) * s.n / d;
    const unsigned long long h = mix64(s.seed ^ ((unsigned long long)u << 20) ^ (unsigned long long)k ^ 0xA5A5A5A5ull);
    return (int)(lo + (long long)(h % (unsigned long long)(hi - lo)));
}
__host__ __device__ __forceinline__ float rating_of(const Spec& s, long long u, int j) {
    float score = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float a = 2.f * u01(mix64(s.seed ^ (unsigned long long)u * 31ull + q + 0x1000ull)) - 1.f;
        const float b = 2.f * u01(mix64(s.seed ^ (unsigned long long)j * 131ull + q + 0x2000000ull)) - 1.f;
        score += a * b;
    }
    const float noise = u01(mix64(s.seed ^ ((unsigned long long)u * 1000003ull + (unsigned long long)j))) - 0.5f;
    float r = rintf(3.4f + 1.6f * score + 1.2f * noise);
    return fminf(5.f, fmaxf(1.f, r));
}

__global__ void degrees_kernel(Spec s, long long row0, int rows, int* __restrict__ deg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) deg[i] = degree_of(s, row0 + i);
}
// strata of row u that can hold a column in [t0, t1): k in [k0, k1)
__device__ __forceinline__ void strata_range(const Spec& s, int d, int t0, int t1, int* k0, int* k1) {
    long long a = (long long)t0 * d / s.n - 1, b = ((long long)t1 * d + s.n - 1) / s.n + 1;
    if (a < 0) a = 0;
    if (b > d) b = d;
    *k0 = (int)a; *k1 = (int)b;
}
__global__ void col_counts_kernel(Spec s, int t0, int t1, int* __restrict__ cnt) {      // per row: entries with column in [t0, t1)
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= s.m) return;
    const int d = degree_of(s, u);
    int k0, k1, c = 0;
    strata_range(s, d, t0, t1, &k0, &k1);
    for (int k = k0; k < k1; ++k) {
        const int j = column_of(s, u, d, k);
        c += (j >= t0 && j < t1);
    }
    cnt[u] = c;
}
__global__ void fill_csr_kernel(Spec s, long long row0, int rows, const long long* __restrict__ ptr, int* __restrict__ col,
                                float* __restrict__ val) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const long long u = row0 + warp;
    const int d = (int)(ptr[warp + 1] - ptr[warp]);
    for (int k = lane; k < d; k += 32) {
        const int j = column_of(s, u, d, k);
        col[ptr[warp] + k] = j;
        val[ptr[warp] + k] = rating_of(s, u, j);
    }
}
// key = (column - t0) << 32 | row ; m < 2^32 rows
__global__ void fill_csc_keys_kernel(Spec s, int t0, int t1, const long long* __restrict__ off, unsigned long long* __restrict__ keys,
                                     float* __restrict__ vals) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= s.m) return;
    const int d = degree_of(s, u);
    int k0, k1;
    strata_range(s, d, t0, t1, &k0, &k1);
    long long o = off[u];
    for (int k = k0; k < k1; ++k) {
        const int j = column_of(s, u, d, k);
        if (j >= t0 && j < t1) {
            keys[o] = ((unsigned long long)(j - t0) << 32) | (unsigned long long)u;
            vals[o] = rating_of(s, u, j);
            ++o;
        }
    }
}
__global__ void split_keys_kernel(const unsigned long long* __restrict__ keys, long long n, int* __restrict__ rows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rows[i] = (int)(keys[i] & 0xffffffffull);
}
__global__ void col_ptr_kernel(const unsigned long long* __restrict__ keys, long long n, int cols, long long* __restrict__ ptr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > cols) return;
    // first position whose column id is >= c
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)(keys[mid] >> 32) < (long long)c) lo = mid + 1; else hi = mid;
    }
    ptr[c] = lo;
}
__global__ void test_samples_kernel(Spec s, long long row0, int rows, long count, unsigned long long salt, int* __restrict__ trow,
                                    int* __restrict__ tcol, float* __restrict__ tval) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned long long h = mix64(s.seed ^ salt ^ (unsigned long long)i * 0x9E3779B1ull);
    const long long u = row0 + (long long)(h % (unsigned long long)rows);
    const int j = (int)(mix64(h) % (unsigned long long)s.n);
    trow[i] = (int)u; tcol[i] = j; tval[i] = rating_of(s, u, j);
}

template <typename T> struct CastFromInt { __host__ __device__ T operator()(int v) const { return (T)v; } };

template <typename T> int scan_exclusive(const int* d_in, T* d_out, long long count, cudaStream_t st) {      // d_out has count + 1 entries
    // sum as T: out[i] = sum_{k<i} in[k], out[count] = total
    cub::TransformInputIterator<T, CastFromInt<T>, const int*> it(d_in, CastFromInt<T>());
    void* tmp = nullptr;
    size_t bytes = 0;
    CUMF_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, it, d_out, (long long)count + 1, st));
    DevBuf t;
    CUMF_TRY(t.alloc(bytes));
    CUMF_CUDA_TRY(cub::DeviceScan::ExclusiveSum(t.p, bytes, it, d_out, (long long)count + 1, st));
    CUMF_CUDA_TRY(cudaStreamSynchronize(st));
    return CUMF_OK;
}

}  // namespace
}  // namespace cumf

using namespace cumf;

struct cumf_synth_shard {
    Spec spec{};
    int device = 0;
    int x0 = 0, x1 = 0, t0 = 0, t1 = 0;
    std::vector<long long> h_csr_ptr, h_csc_ptr;     // rebased to 0
    DevBuf csr_ptr, csr_col, csr_val, csc_row, csc_val, test_row, test_col, test_val;
    long test_cnt = 0;
    long long total_nnz = 0;                          // ratings of the WHOLE matrix (sum of all degrees)
};

extern "C" int cumf_synth_destroy(cumf_synth_shard* sh) {
    if (sh) { cudaSetDevice(sh->device); delete sh; }
    return CUMF_OK;
}

// Slices [x0, x1) x all columns (CSR) and all rows x [t0, t1) (CSC) of the matrix (m, n, avg_deg, seed), generated on `device`;
// test_cnt test samples with rows in [x0, x1).  The degree array of all m rows is scanned once per shard to learn nnz.
extern "C" int cumf_synth_create(cumf_synth_shard** out, long long m, int n, float avg_deg, unsigned long long seed, int x0, int x1,
                                 int t0, int t1, long test_cnt, int device) {
    CUMF_REQUIRE(out && m > 0 && m < (1ll << 31) && n > 0 && avg_deg >= 1.f, "bad matrix spec (m < 2^31 rows)");
    CUMF_REQUIRE(0 <= x0 && x0 <= x1 && x1 <= m && 0 <= t0 && t0 <= t1 && t1 <= n, "bad shard ranges");
    CUMF_CUDA_TRY(cudaSetDevice(device));
    cumf_synth_shard* sh = new cumf_synth_shard();
    sh->spec = Spec{m, n, avg_deg, seed};
    sh->device = device; sh->x0 = x0; sh->x1 = x1; sh->t0 = t0; sh->t1 = t1;
    auto fail = [&](int rc) { delete sh; return rc; };
    cudaStream_t st = nullptr;
    const int rows = x1 - x0, cols = t1 - t0;
    int rc;
    // ---- CSR slice
    {
        DevBuf deg;
        if ((rc = deg.alloc(sizeof(int) * (size_t)(rows + 1))) != CUMF_OK) return fail(rc);
        cudaMemsetAsync(deg.p, 0, sizeof(int) * (size_t)(rows + 1), st);
        if (rows) degrees_kernel<<<(rows + 255) / 256, 256, 0, st>>>(sh->spec, x0, rows, deg.as<int>());
        if ((rc = sh->csr_ptr.alloc(sizeof(long long) * (size_t)(rows + 1))) != CUMF_OK) return fail(rc);
        if ((rc = scan_exclusive<long long>(deg.as<int>(), sh->csr_ptr.as<long long>(), rows, st)) != CUMF_OK) return fail(rc);
        sh->h_csr_ptr.resize((size_t)rows + 1);
        if (cudaMemcpy(sh->h_csr_ptr.data(), sh->csr_ptr.p, sizeof(long long) * (size_t)(rows + 1), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(CUMF_ECUDA);
        const long long xn = sh->h_csr_ptr.back();
        if (xn >= (1ll << 31)) { set_last_error("a shard must hold < 2^31 ratings: use more shards"); return fail(CUMF_EINVAL); }
        if ((rc = sh->csr_col.alloc(sizeof(int) * (size_t)std::max<long long>(xn, 1))) != CUMF_OK) return fail(rc);
        if ((rc = sh->csr_val.alloc(sizeof(float) * (size_t)std::max<long long>(xn, 1))) != CUMF_OK) return fail(rc);
        if (rows) fill_csr_kernel<<<(unsigned)(((size_t)rows * 32 + 255) / 256), 256, 0, st>>>(sh->spec, x0, rows, sh->csr_ptr.as<long long>(),
                                                                                             sh->csr_col.as<int>(), sh->csr_val.as<float>());
    }
    // ---- CSC slice: per-row counts over ALL rows -> offsets -> (column, row) keys -> radix sort
    {
        DevBuf cnt, off, keys_a, keys_b, vals_a, tmp;
        if ((rc = cnt.alloc(sizeof(int) * (size_t)(m + 1))) != CUMF_OK) return fail(rc);
        cudaMemsetAsync(cnt.p, 0, sizeof(int) * (size_t)(m + 1), st);
        col_counts_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(sh->spec, t0, t1, cnt.as<int>());
        if ((rc = off.alloc(sizeof(long long) * (size_t)(m + 1))) != CUMF_OK) return fail(rc);
        if ((rc = scan_exclusive<long long>(cnt.as<int>(), off.as<long long>(), m, st)) != CUMF_OK) return fail(rc);
        long long tn = 0;
        if (cudaMemcpy(&tn, off.as<long long>() + m, sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(CUMF_ECUDA);
        if (tn >= (1ll << 31)) { set_last_error("a shard must hold < 2^31 ratings: use more shards"); return fail(CUMF_EINVAL); }
        const size_t cap = (size_t)std::max<long long>(tn, 1);
        if ((rc = keys_a.alloc(sizeof(unsigned long long) * cap)) != CUMF_OK) return fail(rc);
        if ((rc = keys_b.alloc(sizeof(unsigned long long) * cap)) != CUMF_OK) return fail(rc);
        if ((rc = vals_a.alloc(sizeof(float) * cap)) != CUMF_OK) return fail(rc);
        if ((rc = sh->csc_val.alloc(sizeof(float) * cap)) != CUMF_OK) return fail(rc);
        fill_csc_keys_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(sh->spec, t0, t1, off.as<long long>(), keys_a.as<unsigned long long>(),
                                                                          vals_a.as<float>());
        int col_bits = 1;
        while ((1ll << col_bits) < std::max(cols, 2)) ++col_bits;
        size_t bytes = 0;
        if (cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_a.as<unsigned long long>(), keys_b.as<unsigned long long>(), vals_a.as<float>(),
                                            sh->csc_val.as<float>(), (long long)tn, 0, 32 + col_bits, st) != cudaSuccess) return fail(CUMF_ECUDA);
        if ((rc = tmp.alloc(bytes)) != CUMF_OK) return fail(rc);
        if (cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys_a.as<unsigned long long>(), keys_b.as<unsigned long long>(), vals_a.as<float>(),
                                            sh->csc_val.as<float>(), (long long)tn, 0, 32 + col_bits, st) != cudaSuccess) return fail(CUMF_ECUDA);
        if ((rc = sh->csc_row.alloc(sizeof(int) * cap)) != CUMF_OK) return fail(rc);
        if (tn) split_keys_kernel<<<(unsigned)((tn + 255) / 256), 256, 0, st>>>(keys_b.as<unsigned long long>(), tn, sh->csc_row.as<int>());
        DevBuf cptr;
        if ((rc = cptr.alloc(sizeof(long long) * (size_t)(cols + 1))) != CUMF_OK) return fail(rc);
        col_ptr_kernel<<<(cols + 1 + 255) / 256, 256, 0, st>>>(keys_b.as<unsigned long long>(), tn, cols, cptr.as<long long>());
        sh->h_csc_ptr.resize((size_t)cols + 1);
        if (cudaMemcpy(sh->h_csc_ptr.data(), cptr.p, sizeof(long long) * (size_t)(cols + 1), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(CUMF_ECUDA);
        // total ratings of the matrix: sum of all degrees (every shard computes the same number)
        DevBuf degall, ptrall;
        if ((rc = degall.alloc(sizeof(int) * (size_t)(m + 1))) != CUMF_OK) return fail(rc);
        cudaMemsetAsync(degall.p, 0, sizeof(int) * (size_t)(m + 1), st);
        degrees_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(sh->spec, 0, (int)m, degall.as<int>());
        if ((rc = scan_exclusive<long long>(degall.as<int>(), off.as<long long>(), m, st)) != CUMF_OK) return fail(rc);
        if (cudaMemcpy(&sh->total_nnz, off.as<long long>() + m, sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(CUMF_ECUDA);
    }
    // ---- test samples of this shard
    if (test_cnt > 0 && rows > 0) {
        if ((rc = sh->test_row.alloc(sizeof(int) * (size_t)test_cnt)) != CUMF_OK) return fail(rc);
        if ((rc = sh->test_col.alloc(sizeof(int) * (size_t)test_cnt)) != CUMF_OK) return fail(rc);
        if ((rc = sh->test_val.alloc(sizeof(float) * (size_t)test_cnt)) != CUMF_OK) return fail(rc);
        test_samples_kernel<<<(unsigned)((test_cnt + 255) / 256), 256, 0, st>>>(sh->spec, x0, rows, test_cnt, 0x7e57ull + (unsigned long long)x0,
                                                                               sh->test_row.as<int>(), sh->test_col.as<int>(),
                                                                               sh->test_val.as<float>());
        sh->test_cnt = test_cnt;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
        set_last_error(std::string("cumf_synth_create: ") + cudaGetErrorString(cudaGetLastError()));
        return fail(CUMF_ECUDA);
    }
    *out = sh;
    return CUMF_OK;
}

// what = 0 CSR slice, 1 CSC slice: entries in the slice; ptr_out (optional, host): rebased int64 pointers (rows + 1 / cols + 1)
extern "C" long long cumf_synth_slice(const cumf_synth_shard* sh, int what, long long* ptr_out) {
    if (!sh) return -1;
    const std::vector<long long>& p = what == 0 ? sh->h_csr_ptr : sh->h_csc_ptr;
    if (ptr_out) std::copy(p.begin(), p.end(), ptr_out);
    return p.empty() ? 0 : p.back();
}
extern "C" long long cumf_synth_total_nnz(const cumf_synth_shard* sh) { return sh ? sh->total_nnz : -1; }
// copies of a slice's arrays to the host (tests: the replica is compared with the generic host path)
extern "C" int cumf_synth_download(const cumf_synth_shard* sh, int what, int* idx_out, float* val_out) {
    CUMF_REQUIRE(sh && idx_out && val_out, "null pointer");
    CUMF_CUDA_TRY(cudaSetDevice(sh->device));
    const long long n = cumf_synth_slice(sh, what, nullptr);
    const DevBuf& i = what == 0 ? sh->csr_col : sh->csc_row;
    const DevBuf& v = what == 0 ? sh->csr_val : sh->csc_val;
    CUMF_CUDA_TRY(cudaMemcpy(idx_out, i.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    CUMF_CUDA_TRY(cudaMemcpy(val_out, v.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    return CUMF_OK;
}
extern "C" int cumf_synth_download_test(const cumf_synth_shard* sh, int* row_out, int* col_out, float* val_out) {
    CUMF_REQUIRE(sh && row_out && col_out && val_out, "null pointer");
    CUMF_CUDA_TRY(cudaSetDevice(sh->device));
    if (sh->test_cnt == 0) return CUMF_OK;
    CUMF_CUDA_TRY(cudaMemcpy(row_out, sh->test_row.p, sizeof(int) * (size_t)sh->test_cnt, cudaMemcpyDeviceToHost));
    CUMF_CUDA_TRY(cudaMemcpy(col_out, sh->test_col.p, sizeof(int) * (size_t)sh->test_cnt, cudaMemcpyDeviceToHost));
    CUMF_CUDA_TRY(cudaMemcpy(val_out, sh->test_val.p, sizeof(float) * (size_t)sh->test_cnt, cudaMemcpyDeviceToHost));
    return CUMF_OK;
}

namespace cumf {
namespace {
// one warp per row: key = column << 32 | row for every entry
__global__ void csr_keys_kernel(const long long* __restrict__ ptr, const int* __restrict__ col, int rows, unsigned long long* __restrict__ keys) {
    const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= rows) return;
    for (long long k = ptr[warp] + lane; k < ptr[warp + 1]; k += 32) keys[k] = ((unsigned long long)(unsigned)col[k] << 32) | (unsigned long long)warp;
}
}  // namespace
}  // namespace cumf

// CSR -> CSC on the device (the reference ships both orientations as files, prepare_netflix_data.py:98-110):
// one radix sort of (column, row) keys, rows ascending inside every column like scipy's tocsc.  All pointers are device
// pointers; d_rowptr / d_colptr_out are int64 (rows + 1 / cols + 1 entries).  Synchronises `stream`.
extern "C" int cumf_csr_to_csc_device(int rows, int cols, long long nnz, const long long* d_rowptr, const int* d_col, const float* d_val,
                                      long long* d_colptr_out, int* d_row_out, float* d_val_out, void* stream) {
    CUMF_REQUIRE(rows >= 0 && cols >= 0 && nnz >= 0 && d_rowptr && d_colptr_out, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf keys_a, keys_b, tmp;
    const size_t cap = (size_t)std::max<long long>(nnz, 1);
    CUMF_TRY(keys_a.alloc(sizeof(unsigned long long) * cap));
    CUMF_TRY(keys_b.alloc(sizeof(unsigned long long) * cap));
    if (rows && nnz) csr_keys_kernel<<<(unsigned)(((size_t)rows * 32 + 255) / 256), 256, 0, st>>>(d_rowptr, d_col, rows, keys_a.as<unsigned long long>());
    int col_bits = 1;
    while ((1ll << col_bits) < std::max(cols, 2)) ++col_bits;
    size_t bytes = 0;
    CUMF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_a.as<unsigned long long>(), keys_b.as<unsigned long long>(), d_val, d_val_out,
                                                  nnz, 0, 32 + col_bits, st));
    CUMF_TRY(tmp.alloc(bytes));
    CUMF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys_a.as<unsigned long long>(), keys_b.as<unsigned long long>(), d_val, d_val_out,
                                                  nnz, 0, 32 + col_bits, st));
    if (nnz) split_keys_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(keys_b.as<unsigned long long>(), nnz, d_row_out);
    col_ptr_kernel<<<(cols + 1 + 255) / 256, 256, 0, st>>>(keys_b.as<unsigned long long>(), nnz, cols, d_colptr_out);
    CUMF_CUDA_TRY(cudaGetLastError());
    CUMF_CUDA_TRY(cudaStreamSynchronize(st));
    return CUMF_OK;
}

namespace cumf {
namespace {
__global__ void init_uniform_kernel(float* __restrict__ p, size_t n, unsigned long long seed, float scale) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = scale * u01(mix64(seed ^ (unsigned long long)i * 0x2545F4914F6CDD1Dull));
}
}  // namespace
}  // namespace cumf

// theta <- scale * uniform[0,1) (a pure function of seed and position: every replica gets the same values), X <- 0: the shape
// of the front ends' initialisation (main.cpp:72-78) without a host copy of the factors (20 GB of X at Hugewiki scale)
extern "C" void cumf_als_internal_rewrite_factors(cumf_als_solver* s, float** theta, float** x);
extern "C" int cumf_als_init_factors_device(cumf_als_solver* s, unsigned long long seed, float scale) {
    CUMF_REQUIRE(s, "null pointer");
    float* th = nullptr;
    float* x = nullptr;
    cumf_als_internal_rewrite_factors(s, &th, &x);
    int m = 0, n = 0, f = 0;
    cumf_als_shape(s, &m, &n, &f);
    init_uniform_kernel<<<1184, 256>>>(th, (size_t)n * f, seed, scale);
    CUMF_CUDA_TRY(cudaGetLastError());
    CUMF_CUDA_TRY(cudaMemsetAsync(x, 0, sizeof(float) * (size_t)m * f, nullptr));
    CUMF_CUDA_TRY(cudaDeviceSynchronize());
    return CUMF_OK;
}

// the solver of this shard, built on the slices where they lie (cumf_als_create_device)
extern "C" int cumf_synth_solver(cumf_synth_shard* sh, cumf_als_solver** out, int f, float lambda, long nnz_test_total, int solver, int path) {
    CUMF_REQUIRE(sh && out, "null pointer");
    return cumf_als_create_device(out, sh->h_csr_ptr.data(), sh->csr_col.as<int>(), sh->csr_val.as<float>(), sh->h_csc_ptr.data(),
                                  sh->csc_row.as<int>(), sh->csc_val.as<float>(), sh->test_row.as<int>(), sh->test_col.as<int>(),
                                  sh->test_val.as<float>(), sh->test_cnt, (int)sh->spec.m, sh->spec.n, f, (long)sh->total_nnz, nnz_test_total,
                                  lambda, sh->x0, sh->x1, sh->t0, sh->t1, sh->device, solver, path);
}
