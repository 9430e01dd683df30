/*
 * oracle/als_cpu.c -- CPU restatement of the cuMF ALS factor-update path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (cumf_als_b200/) never
 * links, imports or falls back to anything under oracle/.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md section 8c).  The restatement is pinned against the reference
 * itself: oracle/build_ref.sh compiles the unmodified reference sources for
 * sm_100a into oracle/_ref/, tests/golden/make_golden.py runs them on a B200
 * on small seeded inputs and commits the outputs under tests/golden/, and
 * tests/test_oracle.py checks this file against those vectors on CPU.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Arithmetic is fp32 with explicit fmaf() wherever nvcc
 * contracts `a += b*c` into FFMA (the reference is built with the default
 * -fmad=true), and summation orders follow the reference where it defines
 * one.  Build with -ffp-contract=off so the compiler adds no fusions of its
 * own.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* cg.cu:31  #define CG_ERROR 1e-4   (a double constant: the comparison at
 * cg.cu:195 promotes rsnew to double)                                      */
#define CG_ERROR 1e-4
/* als.cu:1000-ish: error_size = 1000 bins (als.cu:969, 996)                */
#define RMSE_BINS 1000

/* ------------------------------------------------------------------------
 * Gram matrix of one row.
 * als.cu:575-659 (get_hermitianT10) and als.cu:443-569 (get_hermitian100) +
 * als.h:39-143 (accumulate_in_registers): every element (i,j) is
 *     acc = 0;  for k in CSR order: acc = fma(theta[col_k][i], theta[col_k][j], acc)
 * then the diagonal gets  (end-start)*lambda  added (als.cu:546-558, 653-656).  nvcc
 * contracts that multiply-add into one FFMA (default -fmad=true); pinned by the golden
 * vectors: with a separate multiply and add the diagonals are 1 ulp off the reference.
 * Output layout: full symmetric f x f, row-major (als.h:501-627).
 * ---------------------------------------------------------------------- */
static void gram_row(const int* idx, int cnt, const float* factor, int f, float lambda, float* A) {
    for (int i = 0; i < f * f; ++i) A[i] = 0.0f;
    for (int k = 0; k < cnt; ++k) {
        const float* t = factor + (size_t)idx[k] * f;
        for (int i = 0; i < f; ++i) {
            const float ti = t[i];
            float* Ai = A + (size_t)i * f;
            /* upper triangle incl. diagonal; mirrored below (products commute,
             * so the mirrored element is bit-identical, as in the reference
             * where one register tile is stored twice: als.cu:561-565)      */
            for (int j = i; j < f; ++j) Ai[j] = fmaf(ti, t[j], Ai[j]);
        }
    }
    for (int i = 0; i < f; ++i)
        for (int j = 0; j < i; ++j) A[(size_t)i * f + j] = A[(size_t)j * f + i];
    for (int d = 0; d < f; ++d)   /* als.cu:546, 655: fused multiply-add */
        A[(size_t)d * f + d] = fmaf((float)cnt, lambda, A[(size_t)d * f + d]);
}

/* b_u = sum_k val_k * theta[col_k]   (als.cu:750-757: cusparseScsrmm2 +
 * cublasSgeam; the library's summation order is unspecified, CSR order with
 * fma is used here).                                                        */
static void rhs_row(const int* idx, const float* val, int cnt, const float* factor, int f, float* b) {
    for (int i = 0; i < f; ++i) b[i] = 0.0f;
    for (int k = 0; k < cnt; ++k) {
        const float* t = factor + (size_t)idx[k] * f;
        const float v = val[k];
        for (int i = 0; i < f; ++i) b[i] = fmaf(v, t[i], b[i]);
    }
}

/* Block-wide sum as the reference intends it (device_utilities.h:8-13, 36-48):
 * per warp of 32 consecutive threads a shuffle-down tree (offsets 16..1, lanes
 * beyond blockDim contribute 0 -- SURVEY.md A.2-7: the partial last warp is
 * formally UB in the reference, the intended value is the exact sum), then the
 * warp partials are added to the accumulator (atomicAdd order is unspecified;
 * warp order is used here).                                                  */
static float block_sum(const float* v, int n) {
    float total = 0.0f;
    for (int w = 0; w * 32 < n; ++w) {
        float lane[32];
        for (int l = 0; l < 32; ++l) lane[l] = (w * 32 + l < n) ? v[w * 32 + l] : 0.0f;
        for (int off = 16; off > 0; off >>= 1)
            for (int l = 0; l + off < 32; ++l) lane[l] = lane[l] + lane[l + off];
        total += lane[0];
    }
    return total;
}

/* A*p with the reference's access pattern and order: thread t accumulates
 * temp += A[f*i + t] * p[i] for i ascending (cg.cu:60-63, 89-92).           */
static void spmv(const float* A, const float* p, int f, float* out) {
    for (int t = 0; t < f; ++t) {
        float temp = 0.0f;
        for (int i = 0; i < f; ++i) temp = fmaf(A[(size_t)f * i + t], p[i], temp);
        out[t] = temp;
    }
}

/* One system of updateXWithCGKernel, cg.cu:36-231.                          */
static void cg_one(const float* A, float* x, const float* b, int f, float cg_iter, float* work) {
    float* p = work;
    float* r = work + f;
    float* ap = work + 2 * f;
    float* tmp = work + 3 * f;
    spmv(A, x, f, ap);                                   /* cg.cu:58-63 */
    for (int t = 0; t < f; ++t) { r[t] = b[t] - ap[t]; p[t] = r[t]; }   /* cg.cu:64-66 */
    for (int t = 0; t < f; ++t) tmp[t] = r[t] * r[t];
    float rsold = block_sum(tmp, f);                     /* cg.cu:68-73 */
    for (int iter = 0; (float)iter < cg_iter; ++iter) {  /* cg.cu:85 (cgIter is a float) */
        spmv(A, p, f, ap);                               /* cg.cu:88-93 */
        for (int t = 0; t < f; ++t) tmp[t] = p[t] * ap[t];
        const float pap = block_sum(tmp, f);             /* cg.cu:119-122 */
        const float alpha = rsold / pap;                 /* cg.cu:128, no guard on 0/0 */
        for (int t = 0; t < f; ++t) {
            x[t] = x[t] + alpha * p[t];                  /* cg.cu:142-143 (contracted to fma by nvcc) */
            r[t] = r[t] - alpha * ap[t];                 /* cg.cu:145-146 */
        }
        for (int t = 0; t < f; ++t) tmp[t] = r[t] * r[t];
        const float rsnew = block_sum(tmp, f);           /* cg.cu:174-175 */
        if ((double)rsnew < CG_ERROR) break;             /* cg.cu:195 */
        const float beta = rsnew / rsold;                /* cg.cu:201 */
        rsold = rsnew;                                   /* cg.cu:203 */
        for (int t = 0; t < f; ++t) p[t] = r[t] + beta * p[t];   /* cg.cu:208-209 */
    }
}

/* nvcc contracts x + alpha*p and r - alpha*ap into FFMA (default -fmad=true);
 * cg_one above is compiled with -ffp-contract=off, so spell the fused form
 * out in a second variant and let the caller choose.  The default used by the
 * parity tests is the fused form (what the reference binary executes).      */
static void cg_one_fused(const float* A, float* x, const float* b, int f, float cg_iter, float* work) {
    float* p = work;
    float* r = work + f;
    float* ap = work + 2 * f;
    float* tmp = work + 3 * f;
    spmv(A, x, f, ap);
    for (int t = 0; t < f; ++t) { r[t] = b[t] - ap[t]; p[t] = r[t]; }
    for (int t = 0; t < f; ++t) tmp[t] = r[t] * r[t];
    float rsold = block_sum(tmp, f);
    for (int iter = 0; (float)iter < cg_iter; ++iter) {
        spmv(A, p, f, ap);
        for (int t = 0; t < f; ++t) tmp[t] = p[t] * ap[t];
        const float pap = block_sum(tmp, f);
        const float alpha = rsold / pap;
        for (int t = 0; t < f; ++t) {
            x[t] = fmaf(alpha, p[t], x[t]);
            r[t] = fmaf(-alpha, ap[t], r[t]);
        }
        for (int t = 0; t < f; ++t) tmp[t] = r[t] * r[t];
        const float rsnew = block_sum(tmp, f);
        if ((double)rsnew < CG_ERROR) break;
        const float beta = rsnew / rsold;
        rsold = rsnew;
        for (int t = 0; t < f; ++t) p[t] = fmaf(beta, p[t], r[t]);
    }
}

/* LU without pivoting + two triangular solves: cublasSgetrfBatched(NULL
 * pivot) / cublasSgetrsBatched as called at als.cu:77, 98 (closed library;
 * the published algorithm is Doolittle right-looking LU, column-major input --
 * A is symmetric so row/column-major are the same matrix).  A is destroyed.  */
static void lu_nopivot_solve(float* A, float* b, int f) {
    for (int k = 0; k < f; ++k) {
        const float piv = A[(size_t)k * f + k];
        for (int i = k + 1; i < f; ++i) {
            const float l = A[(size_t)i * f + k] / piv;
            A[(size_t)i * f + k] = l;
            for (int j = k + 1; j < f; ++j)
                A[(size_t)i * f + j] = fmaf(-l, A[(size_t)k * f + j], A[(size_t)i * f + j]);
        }
    }
    for (int i = 0; i < f; ++i) {            /* L y = b, unit diagonal */
        float s = b[i];
        for (int j = 0; j < i; ++j) s = fmaf(-A[(size_t)i * f + j], b[j], s);
        b[i] = s;
    }
    for (int i = f - 1; i >= 0; --i) {       /* U x = y */
        float s = b[i];
        for (int j = i + 1; j < f; ++j) s = fmaf(-A[(size_t)i * f + j], b[j], s);
        b[i] = s / A[(size_t)i * f + i];
    }
}

/* ------------------------------------------------------------------------
 * Exported stage-level entry points (argument meaning = the reference's
 * internal seams).
 * ---------------------------------------------------------------------- */

/* get_hermitian*<<<batch_size,...>>>(batch_offset, tt, rowPtr, colIdx, lambda, m, F, factor)
 * als.cu:445-447, 576-578, launches als.cu:804, 816.  tt is [batch_size][f][f]. */
ORACLE_API void oracle_gram(int batch_offset, int batch_size, float* tt, const int* rowptr,
                            const int* colidx, float lambda, int m, int f, const float* factor) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int r = 0; r < batch_size; ++r) {
        const int row = r + batch_offset;
        if (row >= m) continue;               /* als.cu:449-450 */
        const int s = rowptr[row], e = rowptr[row + 1];
        gram_row(colidx + s, e - s, factor, f, lambda, tt + (size_t)r * f * f);
    }
}

/* ythetaT (rows x f, row-major) as left by als.cu:750-757 / 867-874.        */
ORACLE_API void oracle_rhs(int rows, float* out, const int* rowptr, const int* colidx,
                           const float* val, int f, const float* factor) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int row = 0; row < rows; ++row) {
        const int s = rowptr[row], e = rowptr[row + 1];
        rhs_row(colidx + s, val + s, e - s, factor, f, out + (size_t)row * f);
    }
}

/* updateXWithCGHost(A, x, b, batchSize, f, cgIter)  cg.cu:682-686, cg.h:30.
 * fused_fma != 0 spells the fma contractions nvcc applies.                  */
ORACLE_API void oracle_cg(const float* A, float* x, const float* b, int batch, int f,
                          float cg_iter, int fused_fma) {
#pragma omp parallel
    {
        float* work = (float*)malloc(sizeof(float) * 4 * (size_t)f);
#pragma omp for schedule(dynamic, 16)
        for (int s = 0; s < batch; ++s) {
            if (fused_fma) cg_one_fused(A + (size_t)s * f * f, x + (size_t)s * f, b + (size_t)s * f, f, cg_iter, work);
            else           cg_one(A + (size_t)s * f * f, x + (size_t)s * f, b + (size_t)s * f, f, cg_iter, work);
        }
        free(work);
    }
}

/* updateX / updateTheta with the LU solver, als.cu:58-122 / 124-189: solve in
 * place in the RHS, then copy into the factor (als.cu:108, 175).  A is
 * overwritten with its factors exactly like the library call.               */
ORACLE_API void oracle_lu(float* A, float* x, const float* b, int batch, int f) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int s = 0; s < batch; ++s) {
        float* xs = x + (size_t)s * f;
        memcpy(xs, b + (size_t)s * f, sizeof(float) * f);
        lu_nopivot_solve(A + (size_t)s * f * f, xs, f);
    }
}

/* RMSE kernel + reduction, als.cu:191-219, 967-1020.
 *   launched = number of threads that exist: train  ((nnz-1)/256+1)*256 -> all
 *   samples (als.cu:979); test ((nnz_test-1)/256)*256 -> the tail is dropped
 *   (als.cu:1006) while the divisor stays `count`.
 *   e = val - sum_k theta[col][k]*x[row][k], k ascending, e -= a*b contracted
 *   to fma; bins[i % 1000] += e*e (atomic order unspecified; index order used);
 *   S = sum |bins| (cublasSasum); rmse = sqrt(S / count).
 * Returns sqrt(S/count) as als.cu:991/1018 compute it: float / (long->float),
 * std::sqrt(float) -- the value the reference stores in final_rmse.         */
ORACLE_API float oracle_rmse(const float* val, const int* row, const int* col, const float* thetaT,
                             const float* XT, long count, int f, int drop_tail) {
    long launched = drop_tail ? ((count - 1) / 256) * 256 : count;
    if (launched > count) launched = count;
    float bins[RMSE_BINS];
    for (int i = 0; i < RMSE_BINS; ++i) bins[i] = 0.0f;
    for (long i = 0; i < launched; ++i) {
        const float* a = thetaT + (size_t)col[i] * f;
        const float* b = XT + (size_t)row[i] * f;
        float e = val[i];
        for (int k = 0; k < f; ++k) e = fmaf(-a[k], b[k], e);
        bins[i % RMSE_BINS] += e * e;
    }
    float s = 0.0f;
    for (int i = 0; i < RMSE_BINS; ++i) s += fabsf(bins[i]);
    return sqrtf(s / (float)count);
}

/* batch split arithmetic of doALS, als.cu:768-777 / 881-890 (integer path,
 * must be bit-exact).                                                       */
ORACLE_API void oracle_batch_range(int rows, int nbatch, int batch_id, int* batch_size, int* batch_offset) {
    if (batch_id != nbatch - 1) *batch_size = rows / nbatch;
    else *batch_size = rows - batch_id * (rows / nbatch);
    *batch_offset = batch_id * (rows / nbatch);
}

/* One half-step: for all rows, Gram + RHS + solve, written into `out` (which
 * also carries the CG warm start), als.cu:727-853 (X) / 858-961 (theta).
 * solver: 0 = CG (USE_CG, als.cu:28, CG_ITER 6 als.cu:32), 1 = LU.
 * row_begin/row_end restrict the update to a row range (bench cpu sample).   */
ORACLE_API void oracle_half_step(const int* rowptr, const int* colidx, const float* val, int rows,
                                 int row_begin, int row_end, const float* factor, float* out, int f,
                                 float lambda, int solver, float cg_iter) {
    (void)rows;
#pragma omp parallel
    {
        float* A = (float*)malloc(sizeof(float) * (size_t)f * f);
        float* b = (float*)malloc(sizeof(float) * (size_t)f);
        float* work = (float*)malloc(sizeof(float) * 4 * (size_t)f);
#pragma omp for schedule(dynamic, 1)
        for (int row = row_begin; row < row_end; ++row) {
            const int s = rowptr[row], e = rowptr[row + 1];
            gram_row(colidx + s, e - s, factor, f, lambda, A);
            rhs_row(colidx + s, val + s, e - s, factor, f, b);
            float* x = out + (size_t)row * f;
            if (solver == 0) cg_one_fused(A, x, b, f, cg_iter, work);
            else { lu_nopivot_solve(A, b, f); memcpy(x, b, sizeof(float) * f); }
        }
        free(A); free(b); free(work);
    }
}

/* doALS, als.cu:662-1035 (signature als.h:676-681) with the solver choice the
 * reference makes at compile time (als.cu:28) as a run-time argument.
 * rmse_out (optional) receives 2*ITERS floats: train, test per iteration
 * (als.cu:991, 1019).  Returns the last test RMSE (als.cu:1034).            */
ORACLE_API float oracle_doALS(const int* csrRowIndex, const int* csrColIndex, const float* csrVal,
                              const int* cscRowIndex, const int* cscColIndex, const float* cscVal,
                              const int* cooRowIndex, float* thetaT, float* XT,
                              const int* cooRowIndexTest, const int* cooColIndexTest,
                              const float* cooValTest, int m, int n, int f, long nnz, long nnz_test,
                              float lambda, int iters, int solver, float* rmse_out) {
    float final_rmse = 0.0f;
    for (int it = 0; it < iters; ++it) {
        /* update X from theta over CSR rows (als.cu:727-853) */
        oracle_half_step(csrRowIndex, csrColIndex, csrVal, m, 0, m, thetaT, XT, f, lambda, solver, 6.0f);
        /* update theta from X over CSC columns (als.cu:858-961; kernel args
         * swapped at als.cu:918-919: cscColIndex is the pointer array)       */
        oracle_half_step(cscColIndex, cscRowIndex, cscVal, n, 0, n, XT, thetaT, f, lambda, solver, 6.0f);
        /* train RMSE pairs cooRowIndex[i] with csrColIndex[i], csrVal[i] (als.cu:979-980) */
        const float tr = oracle_rmse(csrVal, cooRowIndex, csrColIndex, thetaT, XT, nnz, f, 0);
        const float te = oracle_rmse(cooValTest, cooRowIndexTest, cooColIndexTest, thetaT, XT, nnz_test, f, 1);
        if (rmse_out) { rmse_out[2 * it] = tr; rmse_out[2 * it + 1] = te; }
        final_rmse = te;
    }
    return final_rmse;
}

ORACLE_API int oracle_version(void) { return 1; }
