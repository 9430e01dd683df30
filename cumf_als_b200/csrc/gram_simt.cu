// gram_simt.cu -- exact-fp32 Gram (+RHS) formation on the FFMA pipe.
//
// Role: the correctness anchor and generic-f path of the hot path
//   A_u = sum_{j in Omega_u} theta_j theta_j^T + lambda*|Omega_u|*I,   b_u = sum r_uj theta_j
// replacing get_hermitian100 / get_hermitianT10 (reference als.cu:443-659) and the
// cuSPARSE RHS pass (als.cu:750-757).  Per element it performs the same fp32 FMA
// chain in CSR order as the reference (als.h:39-143), so on rows that are not split
// its output is bit-identical to the reference's `tt`.  The tensor-core path
// (gram_tc.cu) is validated against this kernel.
//
// Design (B200): one CTA per row *chunk* (rows longer than the split threshold are
// cut up; see plan builder), cp.async double-buffered staging of the gathered
// f-wide factor rows in shared memory (coalesced 8-byte reads along f), 10x10
// register tiles over the upper triangle of the f x f output, mirrored on store.
// HBM traffic per chunk: nnz_chunk*(4f+4+4) read, f*f*4 + f*4 written.
#include "common.cuh"

namespace cumf {

namespace {

constexpr int KC = 32;   // gathered rows per shared-memory stage
constexpr int TS = 10;   // register tile edge (f is a multiple of 10: main.cpp:33-36)

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__global__ void gram_simt_kernel(const Chunk* __restrict__ chunks, int c0, const int* __restrict__ colidx,
                                 const float* __restrict__ val, const float* __restrict__ factor, int f,
                                 float lambda, int out_row_base, float* __restrict__ tt,
                                 float* __restrict__ rhs, float* __restrict__ scratchA,
                                 float* __restrict__ scratchB) {
    extern __shared__ __align__(16) float smem[];
    float* vals = smem + 2 * KC * f;

    const Chunk ch = chunks[c0 + blockIdx.x];
    const int t = threadIdx.x;
    const int nthreads = blockDim.x;
    const int lane = t & 31, warp = t >> 5, nwarps = nthreads >> 5;
    const int N = f / TS;
    const int ntiles = N * (N + 1) / 2;
    const bool has_tile = t < ntiles;
    // upper-triangular tile enumeration, row-major: (0,0),(0,1)..(0,N-1),(1,1)...
    int tx = 0, ty = 0;
    if (has_tile) {
        int rem = t, i = 0;
        while (rem >= N - i) { rem -= N - i; ++i; }
        tx = i;
        ty = i + rem;
    }
    const bool want_rhs = (val != nullptr);

    float acc[TS][TS];
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
        for (int j = 0; j < TS; ++j) acc[i][j] = 0.f;
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};

    const int total = ch.end - ch.begin;
    const int nstages = (total + KC - 1) / KC;
    const int f2 = f >> 1;

    auto issue = [&](int s) {
        const int k0 = ch.begin + s * KC;
        const int cnt = min(KC, ch.end - k0);
        float* dst = smem + (s & 1) * KC * f;
        for (int kk = warp; kk < cnt; kk += nwarps) {
            const int col = __ldg(colidx + k0 + kk);
            const float* src = factor + (size_t)col * f;
            for (int c = lane; c < f2; c += 32) cp_async_8(dst + kk * f + 2 * c, src + 2 * c);
        }
        if (want_rhs && t < cnt) vals[(s & 1) * KC + t] = __ldg(val + k0 + t);
        cp_async_commit();
    };

    if (nstages > 0) issue(0);
    for (int s = 0; s < nstages; ++s) {
        if (s + 1 < nstages) {
            issue(s + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* tl = smem + (s & 1) * KC * f;
        const float* vl = vals + (s & 1) * KC;
        const int cnt = min(KC, total - s * KC);
        if (has_tile) {
            const float* pa = tl + tx * TS;
            const float* pb = tl + ty * TS;
            for (int kk = 0; kk < cnt; ++kk) {
                float a[TS], b[TS];
#pragma unroll
                for (int q = 0; q < TS / 2; ++q) {
                    const float2 va = *reinterpret_cast<const float2*>(pa + kk * f + 2 * q);
                    const float2 vb = *reinterpret_cast<const float2*>(pb + kk * f + 2 * q);
                    a[2 * q] = va.x; a[2 * q + 1] = va.y;
                    b[2 * q] = vb.x; b[2 * q + 1] = vb.y;
                }
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TS; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        if (want_rhs) {
            for (int kk = 0; kk < cnt; ++kk) {
                const float r = vl[kk];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = t + q * nthreads;
                    if (c < f) bacc[q] = fmaf(r, tl[kk * f + c], bacc[q]);
                }
            }
        }
        __syncthreads();
    }

    const bool direct = ch.slot < 0;
    float* Aout = direct ? tt + (size_t)(ch.row - out_row_base) * f * f : scratchA + (size_t)ch.slot * f * f;
    float* bout = direct ? (rhs ? rhs + (size_t)(ch.row - out_row_base) * f : nullptr)
                         : (scratchB ? scratchB + (size_t)ch.slot * f : nullptr);
    if (has_tile) {
        if (direct && tx == ty) {
            // weighted-lambda regularisation: (end-start)*lambda on the diagonal (als.cu:546, 655);
            // the reference binary executes it as one FFMA (pinned by tests/golden/gram_*.npz)
#pragma unroll
            for (int i = 0; i < TS; ++i) acc[i][i] = fmaf((float)total, lambda, acc[i][i]);
        }
#pragma unroll
        for (int i = 0; i < TS; ++i) {
            float* dst = Aout + (size_t)(tx * TS + i) * f + ty * TS;
#pragma unroll
            for (int q = 0; q < TS / 2; ++q)
                *reinterpret_cast<float2*>(dst + 2 * q) = make_float2(acc[i][2 * q], acc[i][2 * q + 1]);
        }
        if (tx != ty) {
#pragma unroll
            for (int j = 0; j < TS; ++j) {
                float* dst = Aout + (size_t)(ty * TS + j) * f + tx * TS;
#pragma unroll
                for (int q = 0; q < TS / 2; ++q)
                    *reinterpret_cast<float2*>(dst + 2 * q) = make_float2(acc[2 * q][j], acc[2 * q + 1][j]);
            }
        }
    }
    if (want_rhs && bout) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = t + q * nthreads;
            if (c < f) bout[c] = bacc[q];
        }
    }
}

// Deterministic split-K tail: out = sum_{s in slots, ascending} partial_s (+ lambda*n_u on the diagonal).
__global__ void split_reduce_kernel(const SplitRow* __restrict__ rows, int r0, int f, float lambda, int compact,
                                    int out_row_base, float* __restrict__ tt,
                                    float* __restrict__ rhs, const float* __restrict__ scratchA,
                                    const float* __restrict__ scratchB) {
    const SplitRow sr = rows[r0 + blockIdx.x];
    const int ff = f * f;
    const size_t oidx = compact ? (size_t)(r0 + blockIdx.x) : (size_t)(sr.row - out_row_base);
    float* Aout = tt + oidx * ff;
    for (int e = threadIdx.x; e < ff; e += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < sr.count; ++k) s += scratchA[(size_t)(sr.first_slot + k) * ff + e];
        const int i = e / f, j = e - i * f;
        if (i == j) s = fmaf((float)sr.nnz, lambda, s);
        Aout[e] = s;
    }
    if (rhs && scratchB) {
        float* bout = rhs + oidx * f;
        for (int e = threadIdx.x; e < f; e += blockDim.x) {
            float s = 0.f;
            for (int k = 0; k < sr.count; ++k) s += scratchB[(size_t)(sr.first_slot + k) * f + e];
            bout[e] = s;
        }
    }
}

}  // namespace

int launch_gram_simt(const Chunk* d_chunks, int c0, int c1, const int* d_colidx, const float* d_val,
                     const float* d_factor, int f, float lambda, int out_row_base, float* d_tt, float* d_rhs,
                     float* d_scratchA, float* d_scratchB, cudaStream_t st) {
    if (c1 <= c0) return CUMF_OK;
    const int N = f / TS;
    int threads = N * (N + 1) / 2;
    threads = ((threads + 31) / 32) * 32;
    if (threads < 32) threads = 32;
    const size_t smem = (size_t)(2 * KC * f + 2 * KC) * sizeof(float);
    // per device and cheap: set before every launch (f = 200 needs 51 KB, above the 48 KB default)
    CUMF_CUDA_TRY(cudaFuncSetAttribute(gram_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    gram_simt_kernel<<<c1 - c0, threads, smem, st>>>(d_chunks, c0, d_colidx, d_rhs ? d_val : nullptr, d_factor, f,
                                                     lambda, out_row_base, d_tt, d_rhs, d_scratchA, d_scratchB);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

int launch_split_reduce(const SplitRow* d_rows, int r0, int r1, int f, float lambda, int compact,
                        int out_row_base, float* d_tt, float* d_rhs, const float* d_scratchA,
                        const float* d_scratchB, cudaStream_t st) {
    if (r1 <= r0) return CUMF_OK;
    split_reduce_kernel<<<r1 - r0, 256, 0, st>>>(d_rows, r0, f, lambda, compact, out_row_base, d_tt, d_rhs,
                                                 d_scratchA, d_scratchB);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

}  // namespace cumf
