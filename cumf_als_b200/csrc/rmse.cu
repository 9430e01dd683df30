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

#include <algorithm>

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

// Same sum over the samples of a CSR-like structure, one warp per row chunk: the chunk's own factor row stays
// in registers and only the opposing factor's rows are gathered (per sample: one f-wide row instead of two).
// Used for the train RMSE when cooRowIndex is consistent with the CSR row pointer (then the sample multiset
// {(cooRow[i], csrCol[i], csrVal[i])} of als.cu:979-980 is the matrix itself and may be walked in either
// orientation); the gather side is the smaller factor, which is L2-resident for the BASELINE shapes.
// Per sample the arithmetic is that of sse_kernel (same lane split, same FMA order, same butterfly).
constexpr int kMaxF2PerLane = 13;      // f <= 200 -> f/2 <= 100 float2 over 8 lanes

__global__ void __launch_bounds__(kThreads)
sse_chunks_kernel(const Chunk* __restrict__ chunks, int nchunks, const int* __restrict__ idx,
                  const float* __restrict__ val, const float* __restrict__ own, const float* __restrict__ other,
                  int own_is_theta, int f, double* __restrict__ partials) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % kLanesPerSample;
    const int sidx = lane / kLanesPerSample;
    constexpr int kSamplesPerWarp = 32 / kLanesPerSample;
    const int warp_global = (int)(((long)blockIdx.x * kThreads + threadIdx.x) >> 5);
    const int nwarps_total = (int)((long)gridDim.x * kThreads / 32);
    const int f2 = f >> 1;
    double local = 0.0;
    for (int c = warp_global; c < nchunks; c += nwarps_total) {
        const Chunk ck = chunks[c];
        float2 o[kMaxF2PerLane];
        const float2* op = reinterpret_cast<const float2*>(own + (size_t)ck.row * f);
#pragma unroll
        for (int k = 0; k < kMaxF2PerLane; ++k) {
            const int cc = sub + k * kLanesPerSample;
            o[k] = (cc < f2) ? __ldg(op + cc) : make_float2(0.f, 0.f);
        }
        for (int base = ck.begin; base < ck.end; base += kSamplesPerWarp) {
            const int i = base + sidx;
            const bool valid = i < ck.end;
            float dot = 0.f;
            if (valid) {
                const float2* g = reinterpret_cast<const float2*>(other + (size_t)idx[i] * f);
#pragma unroll
                for (int k = 0; k < kMaxF2PerLane; ++k) {
                    const int cc = sub + k * kLanesPerSample;
                    if (cc < f2) {
                        const float2 gv = __ldg(g + cc);
                        // sse_kernel multiplies theta by X in this order; keep the operand order (fmaf is commutative
                        // in its first two arguments, so this is for the reader, not the result)
                        const float2 av = own_is_theta ? o[k] : gv, bv = own_is_theta ? gv : o[k];
                        dot = fmaf(av.x, bv.x, dot);
                        dot = fmaf(av.y, bv.y, dot);
                    }
                }
            }
#pragma unroll
            for (int off = kLanesPerSample / 2; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
            if (valid && sub == 0) {
                const float e = val[i] - dot;
                local += (double)(e * e);
            }
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

// flag[0] |= 1 if some rating of a chunk carries a cooRow different from the chunk's row
__global__ void coo_check_kernel(const Chunk* __restrict__ chunks, int nchunks, const int* __restrict__ coo_row,
                                 int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps_total = (int)((long)gridDim.x * blockDim.x / 32);
    bool bad = false;
    for (int c = warp_global; c < nchunks; c += nwarps_total) {
        const Chunk ck = chunks[c];
        for (int i = ck.begin + lane; i < ck.end; i += 32) bad |= (coo_row[i] != ck.row);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flag, 1);
}

}  // namespace

int sse_partial_capacity() { return kMaxBlocks; }

// SSE over the ratings of a chunk list (positions relative to idx/val): own = factor of the chunk rows,
// other = factor indexed by idx.  own_is_theta only documents which of the two is theta.
int launch_sse_chunks(const Chunk* d_chunks, int nchunks, const int* d_idx, const float* d_val, const float* d_own,
                      const float* d_other, int own_is_theta, int f, double* d_sse_out, double* d_partials,
                      int partial_capacity, cudaStream_t st) {
    if (nchunks <= 0) {
        CUMF_CUDA_TRY(cudaMemsetAsync(d_sse_out, 0, sizeof(double), st));
        return CUMF_OK;
    }
    if (f > 2 * kLanesPerSample * kMaxF2PerLane || (f & 1)) {
        set_last_error("launch_sse_chunks: f must be even and <= 208");
        return CUMF_EUNSUPPORTED;
    }
    int blocks = (nchunks + kThreads / 32 - 1) / (kThreads / 32);
    blocks = std::min(std::min(blocks, kMaxBlocks), partial_capacity);
    sse_chunks_kernel<<<blocks, kThreads, 0, st>>>(d_chunks, nchunks, d_idx, d_val, d_own, d_other, own_is_theta, f, d_partials);
    CUMF_CUDA_TRY(cudaGetLastError());
    sse_final_kernel<<<1, 256, 0, st>>>(d_partials, blocks, d_sse_out);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

// sum of v[i]^2 in double (fixed-order partials): the constant term of the by-product train SSE
__global__ void __launch_bounds__(kThreads)
sumsq_kernel(const float* __restrict__ v, long n, double* __restrict__ partials) {
    double local = 0.0;
    for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long)gridDim.x * kThreads) {
        const float x = v[i];
        local += (double)x * (double)x;
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

int launch_sumsq(const float* d_v, long n, double* d_out, double* d_partials, int partial_capacity, cudaStream_t st) {
    if (n <= 0) {
        CUMF_CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(double), st));
        return CUMF_OK;
    }
    int blocks = (int)std::min<long>((n + kThreads - 1) / kThreads, (long)std::min(kMaxBlocks, partial_capacity));
    sumsq_kernel<<<blocks, kThreads, 0, st>>>(d_v, n, d_partials);
    CUMF_CUDA_TRY(cudaGetLastError());
    sse_final_kernel<<<1, 256, 0, st>>>(d_partials, blocks, d_out);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

// *d_out = sum of n doubles, fixed order
int launch_sum_doubles(const double* d_v, int n, double* d_out, cudaStream_t st) {
    sse_final_kernel<<<1, 256, 0, st>>>(d_v, n, d_out);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

// *d_flag (int, zeroed by the caller) becomes non-zero iff coo_row disagrees with the chunk rows
int launch_coo_check(const Chunk* d_chunks, int nchunks, const int* d_coo_row, int* d_flag, cudaStream_t st) {
    if (nchunks <= 0) return CUMF_OK;
    const int blocks = std::min((nchunks + 7) / 8, 148 * 8);
    coo_check_kernel<<<blocks, 256, 0, st>>>(d_chunks, nchunks, d_coo_row, d_flag);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

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
