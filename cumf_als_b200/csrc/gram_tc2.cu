// gram_tc2.cu -- host side and helper kernels of the generic-f fused kernel (gram_tc2.cuh):
//   absmax_kernel          largest finite |v| of the opposing factor / of the ratings (one power-of-two scale per half-step)
//   split_factor2_kernel   fp32 factor -> fp16 table [hi 2^c | lo 2^c] (the TMA gather source) + the unscale factors
//   fill_stage_table2      one descriptor per stage (32 / 64 ratings) with everything the MMA issuer needs
//   update()               the launches of one half-step
#include <vector>
#include <cstdio>
#include "gram_tc2.cuh"

namespace cumf {
namespace tc2 {

namespace {

__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ v, size_t n, unsigned* __restrict__ out) {
    unsigned m = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const unsigned b = __float_as_uint(__ldg(v + k)) & 0x7fffffffu;
        if (b < 0x7f800000u) m = max(m, b);         // NaN / inf rows (empty rows of the reference, README.md:113) do not set the scale
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// power of two that brings a value with biased exponent field `e` to [2^top, 2^(top+1))
__device__ __forceinline__ int scale_exp(unsigned absmax_bits, int top) {
    const int e = (int)(absmax_bits >> 23);
    if (e == 0) return 0;                            // all zero (or denormal): no scaling
    return max(-60, min(60, top - (e - 127)));
}
__device__ __forceinline__ float pow2f(int c) { return __uint_as_float((unsigned)(c + 127) << 23); }

// out[row] = [ hi 2^c (f, zero padded to nb = Geo::LO_COL) | lo 2^c (f, zero padded) | zero to tab_cols ], row == rows is the all-zero row.
//   hi = v 2^c with the low 13 mantissa bits cleared (exact in fp16: 11 significant bits, max |v| 2^c < 2^15),
//   lo = fp16(v 2^c - hi)   (sym: 2 lo -- the long-row variant forms hi^T (2 lo) only and halves G + G^T)
__global__ void __launch_bounds__(256) split_factor2_kernel(const float* __restrict__ fac, int rows, int f, int nb, int tab_cols, int sym,
                                                            const unsigned* __restrict__ absmax, uint4* __restrict__ out,
                                                            float* __restrict__ scales) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ppr = tab_cols >> 3, ppreg = nb >> 3;
    const size_t row = gid / (size_t)ppr;
    const int pc = (int)(gid - row * (size_t)ppr);
    const int c = scale_exp(__ldg(absmax), 14);
    if (gid == 0) {
        const int cr = scale_exp(__ldg(absmax + 1), 13);
        scales[0] = pow2f(-2 * c);
        scales[1] = pow2f(-(c + cr));
        scales[2] = pow2f(cr);
        scales[3] = pow2f(c);
    }
    if (row > (size_t)rows) return;
    uint4 o = make_uint4(0, 0, 0, 0);
    const int region = pc / ppreg, e0 = (pc - region * ppreg) * 8;
    if (row < (size_t)rows && region < 2 && e0 < f) {
        const float sc = pow2f(c), lmul = sym ? 2.f : 1.f;
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            if (e0 + k < f) {                       // f is even: pairs never straddle the end
                const float2 v = __ldg(reinterpret_cast<const float2*>(fac + row * (size_t)f + e0 + k));
                const float s0 = v.x * sc, s1 = v.y * sc;
                const float h0 = __uint_as_float(__float_as_uint(s0) & 0xFFFFE000u);
                const float h1 = __uint_as_float(__float_as_uint(s1) & 0xFFFFE000u);
                const __half2 hh = region == 0 ? __floats2half2_rn(h0, h1) : __floats2half2_rn((s0 - h0) * lmul, (s1 - h1) * lmul);
                w[k >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
            }
        }
        o = make_uint4(w[0], w[1], w[2], w[3]);
    }
    out[gid] = o;
}

__global__ void fill_stage_table2_kernel(const Chunk* __restrict__ chunks, const int* __restrict__ chunk_stage_base,
                                         const int* __restrict__ chunk_meta, int nchunks, StageDesc* __restrict__ table, int krows,
                                         int sub, int nbuf) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const Chunk ck = chunks[c];
    const int steps = max(1, (ck.end - ck.begin + krows - 1) / krows);
    const uint32_t sys = (uint32_t)chunk_meta[c] & 3u;
    const uint32_t tile0 = (uint32_t)chunk_meta[c] >> 2;
    StageDesc* out = table + chunk_stage_base[c];
    for (int s = 0; s < steps; ++s) {
        const int pos = ck.begin + s * krows;
        const int cnt = max(0, min(krows, ck.end - pos));
        const bool last = (s == steps - 1);
        const uint32_t tile = tile0 + (uint32_t)(s / sub);
        const uint32_t flags = (s == 0 ? FLAG_CHUNK_FIRST : 0u) | (last ? FLAG_CHUNK_LAST : 0u) |
                               ((s % sub) == 0 ? FLAG_SUB_FIRST : 0u) | ((last || (s % sub) == sub - 1) ? FLAG_SUB_LAST : 0u) |
                               ((tile % (uint32_t)nbuf) << FLAG_BUF_SHIFT) | (sys << FLAG_SYS_SHIFT) |
                               ((((tile / (uint32_t)nbuf) & 1u) ^ 1u) << FLAG_EMPTY_PARITY_SHIFT) |
                               ((uint32_t)(max(1, (cnt + KT - 1) / KT) - 1) << FLAG_GROUPS_SHIFT);
        uint32_t tile_bits = 0;
        if ((s % sub) == 0) {
            const int tile_stages = min(sub, steps - s);
            const int last_pos = ck.begin + (s + tile_stages - 1) * krows;
            const int last_cnt = max(0, min(krows, ck.end - last_pos));
            tile_bits = ((uint32_t)tile_stages << TILE_STAGES_SHIFT) |
                        ((uint32_t)(max(1, (last_cnt + KT - 1) / KT) - 1) << TILE_LAST_GROUPS_SHIFT);
        }
        out[s] = StageDesc{pos, (uint32_t)cnt | (flags << 8) | tile_bits};
    }
}

bool find_variant(int f, bool sym, Variant* out) {
    return variant_a(f, sym, out) || variant_b(f, sym, out) || variant_c(f, sym, out);
}

}  // namespace

}  // namespace tc2

bool tc2_plan_info(int f, bool sym, Tc2Info* out) {
    tc2::Variant v;
    if (!tc2::find_variant(f, sym, &v)) return false;
    out->krows = v.krows; out->sub = v.sub; out->nbuf = v.nbuf; out->nsys = v.nsys; out->tab_cols = v.tab_cols;
    out->nb = (f + 1 + 15) / 16 * 16;
    return true;
}

int tc2_fill_stage_table(const Chunk* d_chunks, const int* d_stage_base, const int* d_chunk_meta, int nchunks, void* d_table,
                         const Tc2Info& info, cudaStream_t st) {
    if (nchunks <= 0) return CUMF_OK;
    tc2::fill_stage_table2_kernel<<<(nchunks + 127) / 128, 128, 0, st>>>(d_chunks, d_stage_base, d_chunk_meta, nchunks,
                                                                         reinterpret_cast<tc2::StageDesc*>(d_table), info.krows, info.sub,
                                                                         info.nbuf);
    CUMF_CUDA_TRY(cudaGetLastError());
    return CUMF_OK;
}

int tc2_update(const Tc2Launch& a, cudaStream_t st, int* launches) {
    tc2::Variant v;
    if (!tc2::find_variant(a.f, a.sym, &v)) {
        set_last_error("tc2_update: no kernel for f = " + std::to_string(a.f));
        return CUMF_EUNSUPPORTED;
    }
    // scale of the half-step: largest finite |v| of the opposing factor and of the ratings (asynchronous, no host round trip)
    CUMF_CUDA_TRY(cudaMemsetAsync(a.d_absmax, 0, (a.scan_ratings ? 2 : 1) * sizeof(unsigned), st));
    const size_t nfac = (size_t)a.factor_rows * a.f;
    tc2::absmax_kernel<<<(unsigned)std::min<size_t>((nfac + 255) / 256, 1184), 256, 0, st>>>(a.d_factor, nfac, a.d_absmax);
    if (a.scan_ratings) {       // the ratings are data, not state: their scale is kept per plan and rating array
        tc2::absmax_kernel<<<(unsigned)std::min<size_t>(((size_t)a.val_span + 255) / 256, 1184), 256, 0, st>>>(a.d_val, (size_t)a.val_span,
                                                                                                               a.d_absmax + 1);
        *launches += 1;
    }
    const size_t pieces = (size_t)(a.factor_rows + 1) * (v.tab_cols / 8);
    tc2::split_factor2_kernel<<<(unsigned)((pieces + 255) / 256), 256, 0, st>>>(a.d_factor, a.factor_rows, a.f, v.lo_col,
                                                                               v.tab_cols, a.sym ? 1 : 0, a.d_absmax,
                                                                               reinterpret_cast<uint4*>(a.d_table), a.d_scales);
    CUMF_CUDA_TRY(cudaGetLastError());
    *launches += 2;
    tc2::Params p;
    p.chunks = a.d_chunks; p.chunk_meta = a.d_chunk_meta; p.cta_chunk_ptr = a.d_cta_ptr;
    p.stage_tab = reinterpret_cast<const tc2::StageDesc*>(a.d_stage_tab); p.cta_stage_ptr = a.d_cta_stage_ptr;
    p.colidx = a.d_colidx; p.val = a.d_val;
    p.out.n = 0;
    for (int k = 0; k < 8; ++k) p.out.p[k] = nullptr;
    p.out.p[0] = a.d_out;
    p.out.n = a.d_out ? 1 : 0;
    for (int k = 0; k < a.n_peer_out && p.out.n < 8; ++k) p.out.p[p.out.n++] = a.peer_out[k];
    p.lambda = a.lambda; p.cg_iter = a.cg_iter;
    p.scratchA = a.d_scratchA; p.scratchB = a.d_scratchB;
    p.tt = a.d_tt; p.rhs = a.d_rhs; p.tt_row_base = a.tt_row_base;
    p.scales = a.d_scales; p.sse_terms = a.d_sse_terms; p.zero_row = a.factor_rows;
    p.hi_only = a.hi_only ? 1 : 0;
    p.split_out.n = 0;
    for (int k = 0; k < 8; ++k) p.split_out.p[k] = nullptr;
    if (a.split_out && a.f == 100 && !a.d_tt) p.split_out = *a.split_out;
    p.prof = nullptr;
#ifdef CUMF_TC2_PROFILE
    // experiment builds only (tools/build_variant.sh prof -DCUMF_TC2_PROFILE): per-role cycle counters, printed per launch
    static unsigned long long* d_prof = nullptr;
    const bool prof = getenv("CUMF_TC2_PROF") != nullptr;
    if (prof) {
        if (!d_prof) cudaMalloc(&d_prof, 1024 * tc2::PROF_WORDS * sizeof(unsigned long long));
        cudaMemsetAsync(d_prof, 0, 1024 * tc2::PROF_WORDS * sizeof(unsigned long long), st);
        p.prof = d_prof;
    }
#endif
    CUMF_CUDA_TRY(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    v.fn<<<a.grid, v.threads, v.smem, st>>>(*reinterpret_cast<const CUtensorMap*>(a.tensor_map), p);
    CUMF_CUDA_TRY(cudaGetLastError());
    *launches += 1;
#ifdef CUMF_TC2_PROFILE
    if (prof) {
        std::vector<unsigned long long> h((size_t)a.grid * tc2::PROF_WORDS);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), d_prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        double s[tc2::PROF_WORDS] = {};
        for (int c = 0; c < a.grid; ++c)
            for (int k = 0; k < tc2::PROF_WORDS; ++k) s[k] += (double)h[(size_t)c * tc2::PROF_WORDS + k] / a.grid;
        printf("[tc2 prof f=%d sym=%d grid=%d] per-CTA mean kilocycles\n", a.f, (int)a.sym, a.grid);
        printf("  issuer : loop %.0f  wait landed %.0f  wait tmem-buffer %.0f  stages %.0f\n", s[3] / 1e3, s[1] / 1e3, s[2] / 1e3, s[4]);
        for (int w = 0; w < 3; ++w)
            printf("  worker%d: loop %.0f  wait slot-free %.0f  issue %.0f\n", w, s[7 + 3 * w] / 1e3, s[5 + 3 * w] / 1e3, s[6 + 3 * w] / 1e3);
        for (int g = 0; g < 3; ++g)
            printf("  solver%d: loop %.0f  wait accumulator %.0f  drain %.0f  solve+store %.0f  chunk head %.0f  systems %.0f\n", g,
                   s[17 + 6 * g] / 1e3, s[14 + 6 * g] / 1e3, s[15 + 6 * g] / 1e3, s[16 + 6 * g] / 1e3, s[19 + 6 * g] / 1e3, s[18 + 6 * g]);
        printf("  solver0 detail: before the loop %.0f  CG steps %.0f: publish+mat-vec %.0f  p.Ap sum+update %.0f  r.r sum %.0f  (cycles per step: %.0f / %.0f / %.0f)\n",
               s[32] / 1e3, s[36], s[33] / 1e3, s[34] / 1e3, s[35] / 1e3, s[33] / s[36], s[34] / s[36], s[35] / s[36]);
        fflush(stdout);
    }
#endif
    return CUMF_OK;
}

int tc2_sse_terms_per_cta() { return tc2::MAX_SYS; }

}  // namespace cumf
