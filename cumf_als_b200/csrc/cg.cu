// cg.cu -- batched conjugate-gradient solve of the f x f SPD systems (unfused path).
//
// Replaces updateXWithCGKernel / updateXWithCGHost (reference cg.cu:36-231, 682-686):
//   x <- CG(A, b, x0 = x), at most cgIter steps, break when rsnew < 1e-4 (absolute).
// Semantics kept exactly: warm start (cg.cu:48), alpha = rsold/pAp with no guard
// (cg.cu:128), the break test in double (CG_ERROR is a double literal, cg.cu:31,195),
// the loop bound compared as float (cg.h:30).
//
// B200 design: the reference streams A from HBM (cgIter+1) = 7 times
// (280 KB/system at f=100).  Here each thread loads its slice of one (symmetric)
// row of A into registers ONCE with coalesced column reads (A[j][i] over lanes i),
// and all SpMVs run out of the register file; p is broadcast from shared memory
// with vector loads; the three dot products per step are xor-butterfly warp
// reductions plus a fixed-order cross-warp sum (deterministic, no atomics --
// the reference's shuffle with a partial last warp at blockDim=f=100 is UB,
// SURVEY.md A.2-7).  The SpMV accumulation order equals the reference's
// (temp += A[f*i+t]*p[i], i ascending), so with identical inputs the products
// match bit for bit.  HBM bytes per system: 4f^2 + 12f.
#include "common.cuh"

namespace cumf {
namespace {

constexpr double kCgError = 1e-4;  // cg.cu:31

template <int NT>
__device__ __forceinline__ float block_sum(float v, float* red /*[NT/32]*/, int tid) {
    // butterfly: every lane ends with the same value, same association as the
    // reference's shfl_down tree for lane 0 (device_utilities.h:9-13)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    constexpr int NW = NT / 32;
    if (NW == 1) return v;
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w];
    return s;
}

// F: rank.  S: threads per row (1 for F <= 128, 2 above: 100 registers of A per thread).
template <int F, int S>
__global__ void __launch_bounds__(((F * S + 31) / 32) * 32)
cg_kernel(const float* __restrict__ A, float* __restrict__ x, const float* __restrict__ b, float cg_iter,
          const SplitRow* __restrict__ sys_rows, float lambda, double* __restrict__ sse_rows, PeerOut peers, int x_row_offset,
          SplitOut split_out) {
    constexpr int SEG = F / S;
    constexpr int NT = ((F * S + 31) / 32) * 32;
    __shared__ __align__(16) float sp[F];
    __shared__ float red[3][NT / 32 + 1];

    const int tid = threadIdx.x;
    const int i = tid / S;          // row / unknown owned by this thread
    const int h = tid % S;          // which segment of the row
    const bool active = i < F;
    const size_t sys = blockIdx.x;
    const float* As = A + sys * (size_t)F * F;

    // A[h*SEG + j][i] == A[i][h*SEG + j] (symmetric): coalesced over i (cg.cu:54-55).
    float a[SEG];
#pragma unroll
    for (int j = 0; j < SEG; ++j) a[j] = active ? __ldg(As + (size_t)(h * SEG + j) * F + i) : 0.f;

    // compact batch -> factor row; x_row_offset: row of the factor that system 0 solves (peer replicas are addressed by
    // absolute row, x by batch-relative row)
    const size_t xrow = sys_rows ? (size_t)sys_rows[sys].row : sys;
    const size_t prow = sys_rows ? xrow : sys + (size_t)x_row_offset;
    float xi = active ? x[xrow * F + i] : 0.f;
    const float bi = active ? b[sys * F + i] : 0.f;

    auto spmv = [&]() -> float {   // y_i = sum_j A[i][j] * sp[j], j ascending (one FMA chain per segment)
        float y = 0.f;
        const float* ps = sp + h * SEG;
        if constexpr (SEG % 4 == 0) {
#pragma unroll
            for (int j = 0; j < SEG; j += 4) {
                const float4 pv = *reinterpret_cast<const float4*>(ps + j);
                y = fmaf(a[j], pv.x, y); y = fmaf(a[j + 1], pv.y, y);
                y = fmaf(a[j + 2], pv.z, y); y = fmaf(a[j + 3], pv.w, y);
            }
        } else if constexpr (SEG % 2 == 0) {
#pragma unroll
            for (int j = 0; j < SEG; j += 2) {
                const float2 pv = *reinterpret_cast<const float2*>(ps + j);
                y = fmaf(a[j], pv.x, y); y = fmaf(a[j + 1], pv.y, y);
            }
        } else {   // odd segment (f = 130, 150, 170, 190 split in two): scalar broadcast reads
#pragma unroll
            for (int j = 0; j < SEG; ++j) y = fmaf(a[j], ps[j], y);
        }
        if constexpr (S == 2) y += __shfl_xor_sync(0xffffffffu, y, 1);
        return y;
    };
    const float own = (active && h == 0) ? 1.f : 0.f;   // one contribution per unknown to the dots

    if (active && h == 0) sp[i] = xi;
    __syncthreads();
    float r = bi - spmv();                               // r = b - A x      (cg.cu:58-64)
    float p = r;                                         // p = r            (cg.cu:66)
    float rsold = block_sum<NT>(own * r * r, red[0], tid);   // rsold = r'r  (cg.cu:68-73)

    for (int iter = 0; (float)iter < cg_iter; ++iter) {  // cg.cu:85
        __syncthreads();                                 // everyone is done reading sp / red[0]
        if (active && h == 0) sp[i] = p;
        __syncthreads();
        const float ap = spmv();                         // ap = A p         (cg.cu:88-93)
        const float pap = block_sum<NT>(own * p * ap, red[1], tid);   // cg.cu:119-122
        const float alpha = rsold / pap;                 // cg.cu:128 (no guard: 0/0 -> NaN on empty rows)
        xi = fmaf(alpha, p, xi);                         // cg.cu:142-143
        r = fmaf(-alpha, ap, r);                         // cg.cu:145-146
        const float rsnew = block_sum<NT>(own * r * r, red[2], tid);  // cg.cu:174-175
        if ((double)rsnew < kCgError) break;             // cg.cu:195 (uniform across the CTA)
        const float beta = rsnew / rsold;                // cg.cu:201
        rsold = rsnew;                                   // cg.cu:203
        p = fmaf(beta, p, r);                            // cg.cu:208-209
    }
    if (active && h == 0) {
        x[xrow * F + i] = xi;                            // cg.cu:230
        for (int k = 0; k < peers.n; ++k) peers.p[k][prow * F + i] = xi;      // the same row of every peer replica
        if (split_out.n > 0) split_row_store(split_out, prow, i, xi);         // and its split form (common.cuh, SplitOut)
    }
    if (sse_rows != nullptr) {
        // x^T b + x^T r + reg x^T x: with sum r_j^2 it gives the row's squared error (see gram_tc.cu); only used
        // for the compact batch of split rows, where the rating count comes with the row map
        const float reg = sys_rows ? (float)sys_rows[sys].nnz * lambda : 0.f;
        __syncthreads();
        const float srow = block_sum<NT>(own * xi * (bi + r + reg * xi), red[1], tid);
        if (tid == 0) sse_rows[sys] = (sys_rows && sys_rows[sys].nnz > 0) ? (double)srow : 0.0;
    }
}

template <int F>
int launch_one(const float* A, float* x, const float* b, int batch, float cg_iter, const SplitRow* rows,
               cudaStream_t st, float lambda, double* sse_rows, const PeerOut& peers, int x_row_offset, const SplitOut& split_out) {
    constexpr int S = (F > 128) ? 2 : 1;
    constexpr int NT = ((F * S + 31) / 32) * 32;
    cg_kernel<F, S><<<batch, NT, 0, st>>>(A, x, b, cg_iter, rows, lambda, sse_rows, peers, x_row_offset, split_out);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

}  // namespace

int launch_cg(const float* d_A, float* d_x, const float* d_b, int batch, int f, float cg_iter,
              const SplitRow* d_sys_rows, cudaStream_t st, float lambda, double* d_sse_rows, const PeerOut* peers_in, int x_row_offset,
              const SplitOut* split_in) {
    if (batch <= 0) return CUMF_OK;
    PeerOut peers{};
    if (peers_in) peers = *peers_in;
    SplitOut split_out{};
    if (split_in && f == 100) split_out = *split_in;      // the table format belongs to the f = 100 kernel
    switch (f) {
#define CUMF_CG_CASE(F) case F: return launch_one<F>(d_A, d_x, d_b, batch, cg_iter, d_sys_rows, st, lambda, d_sse_rows, peers, x_row_offset, split_out);
        CUMF_CG_CASE(10) CUMF_CG_CASE(20) CUMF_CG_CASE(30) CUMF_CG_CASE(40) CUMF_CG_CASE(50)
        CUMF_CG_CASE(60) CUMF_CG_CASE(70) CUMF_CG_CASE(80) CUMF_CG_CASE(90) CUMF_CG_CASE(100)
        CUMF_CG_CASE(110) CUMF_CG_CASE(120) CUMF_CG_CASE(130) CUMF_CG_CASE(140) CUMF_CG_CASE(150)
        CUMF_CG_CASE(160) CUMF_CG_CASE(170) CUMF_CG_CASE(180) CUMF_CG_CASE(190) CUMF_CG_CASE(200)
#undef CUMF_CG_CASE
        default:
            set_last_error("cumf_cg: f must be a multiple of 10 in [10,200]");
            return CUMF_EUNSUPPORTED;
    }
}

}  // namespace cumf
