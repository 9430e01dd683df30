// rmse.cu -- sum of squared prediction errors over a list of (row, col, value) samples.
//
// Replaces the RMSE kernel + 1000-bin atomics + cublasSasum of the reference
// (als.cu:191-219, 979-991, 1006-1019).  e_i = val_i - <thetaT[col_i], XT[row_i]>.
// The sample SET is kept bit-exact by the caller (test tail drop, als.cu:1006);
// the summation is deterministic here (fixed-order tree), where the reference's
// float atomics are not.
//
// B200 design: 8 lanes cooperate on one sample (two coalesced f-wide row reads
// as float2 per lane instead of the reference's one-thread-per-sample strided
// reads), butterfly reduce, per-CTA partial in double, second tiny kernel sums
// the partials in index order.  HBM/L2 bytes per sample: 8f + 12.
#include "common.cuh"

namespace cumf {
namespace {

constexpr int kThreads = 256;
constexpr int kLanesPerSample = 8;
constexpr int kMaxBlocks = 148 * 16;

__global__ void __launch_bounds__(kThreads)
sse_kernel(const float* __restrict__ val, const int* __restrict__ row, const int* __restrict__ col,
           const float* __restrict__ thetaT, const float* __restrict__ XT, long count, int f,
           double* __restrict__ partials) {
    // 4 samples per warp, 8 lanes each; the loop is warp-uniform so the full-mask
    // shuffles below are always executed by converged warps.
    const int lane = threadIdx.x & 31;
    const int sub = lane % kLanesPerSample;
    const int sidx = lane / kLanesPerSample;
    constexpr int kSamplesPerWarp = 32 / kLanesPerSample;
    const long warp_global = ((long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const long nwarps_total = (long)gridDim.x * kThreads / 32;
    const int f2 = f >> 1;
    double local = 0.0;
    for (long base = warp_global * kSamplesPerWarp; base < count; base += nwarps_total * kSamplesPerWarp) {
        const long i = base + sidx;
        const bool valid = i < count;
        float dot = 0.f;
        if (valid) {
            const float2* a = reinterpret_cast<const float2*>(thetaT + (size_t)col[i] * f);
            const float2* b = reinterpret_cast<const float2*>(XT + (size_t)row[i] * f);
            for (int c = sub; c < f2; c += kLanesPerSample) {
                const float2 av = __ldg(a + c), bv = __ldg(b + c);
                dot = fmaf(av.x, bv.x, dot);
                dot = fmaf(av.y, bv.y, dot);
            }
        }
#pragma unroll
        for (int off = kLanesPerSample / 2; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
        if (valid && sub == 0) {
            const float e = val[i] - dot;
            local += (double)(e * e);    // e*e in fp32 like als.cu:216, accumulated in double
        }
    }
    __shared__ double red[kThreads / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += red[w];
        partials[blockIdx.x] = s;
    }
}

__global__ void sse_final_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        *out = t;
    }
}

}  // namespace

int sse_partial_capacity() { return kMaxBlocks; }

// d_partials: sse_partial_capacity() doubles of scratch.
int launch_sse(const float* d_val, const int* d_row, const int* d_col, const float* d_thetaT, const float* d_XT,
               long count, int f, double* d_sse_out, double* d_partials, int partial_capacity, cudaStream_t st) {
    double* partials = d_partials;
    if (count <= 0) {
        CUMF_CUDA_TRY(cudaMemsetAsync(d_sse_out, 0, sizeof(double), st));
        return CUMF_OK;
    }
    long want = (count * kLanesPerSample + kThreads - 1) / kThreads;
    int blocks = (int)(want < (long)kMaxBlocks ? want : (long)kMaxBlocks);
    if (blocks > partial_capacity) blocks = partial_capacity;
    if (blocks < 1) blocks = 1;
    sse_kernel<<<blocks, kThreads, 0, st>>>(d_val, d_row, d_col, d_thetaT, d_XT, count, f, partials);
    CUMF_CUDA_TRY(cudaGetLastError());
    sse_final_kernel<<<1, 256, 0, st>>>(partials, blocks, d_sse_out);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

}  // namespace cumf
