/*
 * include/cumf_als.h -- C ABI of the B200-native ALS factor-update path.
 *
 * This is the drop-in boundary for cuMF/cumf_als's one hot path (per-row Gram
 * formation + batched f x f solve).  Plain pointers and sizes only; no torch,
 * no C++ types.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference tree).  The library is
 * cumf_als_b200/libcumf_als_b200.so; it additionally exports the C++-mangled
 * symbol `doALS` (_Z5doALSPKiS0_PKfS0_S0_S2_S0_PfS3_S0_S0_S2_iiillfiiii) so
 * the reference's unmodified main.cpp and tensorflow/als_tf.cc link against it
 * (als.h:676-681, als_tf.cc:33-38), and the four loaders of host_utilities.h.
 *
 * Conventions
 *   - "d_" arguments are DEVICE pointers on the current CUDA device, "h_"/
 *     "...HostPtr" arguments are HOST pointers.
 *   - Factor matrices are row-major [rows][f] fp32 (thetaT: n x f, XT: m x f),
 *     exactly the reference's layout (als.cu:805, 1024-1025).
 *   - Functions returning int return CUMF_OK (0) or a negative CUMF_E* code and
 *     leave a message retrievable with cumf_last_error().  cumf_doALS mirrors
 *     the reference instead: on a CUDA error it prints and exit(EXIT_FAILURE)s
 *     (als.h:628-665).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - There is NO CPU fallback anywhere behind this header.
 */
#ifndef CUMF_ALS_H_
#define CUMF_ALS_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CUMF_OK 0
#define CUMF_EINVAL (-1)   /* bad argument (f odd / f%10 != 0 where required, null pointer, ...) */
#define CUMF_ECUDA (-2)    /* a CUDA runtime / cuBLAS call failed                                   */
#define CUMF_ENOGPU (-3)   /* no sm_100 device visible                                              */
#define CUMF_EUNSUPPORTED (-4)

/* solver selection: the reference chooses at compile time (`#define USE_CG`, als.cu:28) */
#define CUMF_SOLVER_CG 0   /* batched CG, CG_ITER = 6 (als.cu:32), break at rsnew < 1e-4 (cg.cu:31,195) */
#define CUMF_SOLVER_LU 1   /* cuBLAS getrf/getrsBatched without pivoting (als.cu:58-189): oracle mode   */

/* kernel family selection for the Gram + solve half-step */
#define CUMF_PATH_AUTO 0   /* fused tcgen05 kernel where it applies, else SIMT                       */
#define CUMF_PATH_SIMT 1   /* exact-fp32 FFMA Gram (materialised A) + register-resident CG kernel    */
#define CUMF_PATH_TC   2   /* fused TMA-gather + tcgen05 split-fp16 Gram + in-register CG            */

const char* cumf_last_error(void);
int cumf_version(void);

/* ---- b1: doALS -------------------------------------------------------------
 * Replaces  float doALS(...)  als.h:676-681 / als.cu:662-1035 (same argument
 * order and meaning).  All pointers are host memory owned by the caller;
 * thetaTHost (n*f) and XTHost (m*f) are in/out (initial guess in -- CG warm
 * starts from XTHost, main.cpp:76-78 -- final factors out).  Returns the last
 * iteration's test RMSE (als.cu:1018, 1034).  X_BATCH / THETA_BATCH are
 * accepted and advisory: results do not depend on them.
 * Environment knobs (all optional): CUMF_SOLVER=cg|lu, CUMF_PATH=auto|simt|tc,
 * CUMF_DEBUG=1 (prints the reference's -DDEBUG line set -- "update X kernel
 * run", "updateX solver run seconds", "update X run", the theta twins and
 * "Calculate RMSE." (als.cu:821, 845, 850, 928, 957, 961-963) -- so that
 * hermitiantime.sh / solvertime.sh / print-test-result.sh read our logs),
 * CUMF_QUIET=1 (no stdout), CUMF_CACHE_MB=<n> (opt-in: keep up to n MiB of
 * device buffers for the next call; default 0 = everything is freed on
 * return like als.cu:1026-1033), CUMF_GPUS=<n> (row-shard the half-steps over
 * n GPUs of this node inside the call, DEVICEID = first device).             */
float cumf_doALS(const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr, const float* csrValHostPtr,
                 const int* cscRowIndexHostPtr, const int* cscColIndexHostPtr, const float* cscValHostPtr,
                 const int* cooRowIndexHostPtr, float* thetaTHost, float* XTHost,
                 const int* cooRowIndexTestHostPtr, const int* cooColIndexTestHostPtr,
                 const float* cooValHostTestPtr, const int m, const int n, const int f, const long nnz,
                 const long nnz_test, const float lambda, const int ITERS, const int X_BATCH,
                 const int THETA_BATCH, const int DEVICEID);

/* ---- b2: the .bin loaders of the CLI ---------------------------------------
 * Replace host_utilities.h:31-40 / host_utilities.cpp:19-98 (same file format:
 * headerless little-endian int32 / float32).  Return 0, or -1 if a file cannot
 * be opened or is short (the reference prints "Unable to open file!" and
 * carries on with garbage, host_utilities.cpp:27-31; the C++-linkage symbols
 * with the reference's names keep that void signature).                      */
int cumf_load_csr_bin(const char* dataFile, const char* rowFile, const char* colFile, float* data,
                      int* row, int* col, int m, long nnz);
int cumf_load_csc_bin(const char* dataFile, const char* rowFile, const char* colFile, float* data,
                      int* row, int* col, int n, long nnz);
int cumf_load_coo_row_bin(const char* rowFile, int* row, long nnz);
int cumf_load_coo_bin(const char* dataFile, const char* rowFile, const char* colFile, float* data,
                      int* row, int* col, long nnz);

/* Sharded loading: one rank reads only ITS rows (columns) of the CLI's .bin files with seeks -- no process holds the whole
 * matrix (hugewiki.cu:2332-2340 keeps one file set per GPU batch instead).  cumf_bin_shard_extent: first entry and number
 * of entries of rows [row_begin, row_end); cumf_load_csr_shard_bin: rebased int64 pointers (row_end - row_begin + 1) and the
 * index / value slices, ready for cumf_plan_create64 / (after upload) cumf_als_create_device.  0 on success, -1 otherwise. */
int cumf_bin_shard_extent(const char* indptrFile, int rows, int row_begin, int row_end, long long* first, long long* count);
int cumf_load_bin_slice(const char* file, int elem_size, long long first, long long count, void* dst);
int cumf_load_csr_shard_bin(const char* dataFile, const char* indptrFile, const char* indicesFile, int rows,
                            int row_begin, int row_end, long long* ptr_out, int* idx_out, float* val_out);

/* Factor initialisation of the reference's front ends (they do it inline, before doALS):
 *   thetaTHost[k] = scale * rand() / RAND_MAX  (glibc rand),  XTHost[k] = 0
 * main.cpp:72-78 uses srand(0) and scale 0.2; the TensorFlow op als_tf.cc:118-125 never
 * seeds and uses 0.1.  seed < 0: do not call srand.  NULL pointers are skipped.           */
void cumf_init_factors(float* thetaTHost, float* XTHost, int m, int n, int f, double scale, long seed);

/* ---- b4: stage-level seams (device pointers) --------------------------------
 * cumf_gram replaces the launches
 *   get_hermitian100 / get_hermitianT10 <<<batch_size, ...>>>(batch_offset, tt,
 *       rowPtr, colIdx, lambda, m, F, factor)          als.cu:445-447, 576-578, 804, 816
 * and (when d_rhs != NULL) the RHS pass cusparseScsrmm2 + cublasSgeam als.cu:750-757:
 *   d_tt  [batch_size][f][f] full symmetric fp32,  A_u = sum theta theta^T + lambda*nnz_u*I
 *   d_rhs [batch_size][f]   b_u = sum val * theta   (rows batch_offset .. +batch_size)
 * d_rowptr has m+1 entries, indices are absolute (not rebased).  f must be even
 * and a multiple of 10 (main.cpp:33-36), 10 <= f <= 200.                      */
int cumf_gram(int batch_offset, int batch_size, float* d_tt, float* d_rhs, const int* d_rowptr,
              const int* d_colidx, const float* d_val, float lambda, int m, int f,
              const float* d_factor, int path, void* stream);

/* Replaces  void updateXWithCGHost(float* A, float* x, float* b, const int
 * batchSize, const int f, const float cgIter)  cg.h:30 / cg.cu:682-686.
 * x is in/out (warm start, cg.cu:48).                                         */
int cumf_cg(const float* d_A, float* d_x, const float* d_b, int batchSize, int f, float cgIter,
            void* stream);

/* Replaces updateX / updateTheta with the LU solver, als.cu:58-122 / 124-189
 * (cublasSgetrfBatched + cublasSgetrsBatched, NULL pivot).  d_A is overwritten
 * with its LU factors, d_b with the solution, which is also copied to d_x.    */
int cumf_lu(float* d_A, float* d_x, float* d_b, int batchSize, int f, void* stream);

/* Replaces the RMSE kernel + cublasSasum + sqrt, als.cu:191-219, 979-991,
 * 1006-1019.  drop_tail != 0 reproduces the test-set launch (nnz_test-1)/256
 * blocks (als.cu:1006): only the first 256*((count-1)/256) samples are summed
 * while the divisor stays `count`.  *sse_out (optional) receives the sum of
 * squared errors in double; returns via *rmse_out = sqrtf(sse/count).         */
int cumf_rmse(const float* d_val, const int* d_row, const int* d_col, const float* d_thetaT,
              const float* d_XT, long count, int f, int drop_tail, float* rmse_out, double* sse_out,
              void* stream);

/* ---- the hot path as one call: a half-step plan -------------------------------
 * A plan holds the work decomposition (row chunks, split-row partial slots,
 * workspace) for updating rows [row_begin,row_end) of one factor from a CSR-like
 * structure.  h_rowptr (rows+1 ints) is read on the host once; the plan keeps a
 * device copy of what it needs.  One plan per side (CSR for X, CSC for theta).  */
typedef struct cumf_plan cumf_plan;
int cumf_plan_create(cumf_plan** out, const int* h_rowptr, int rows, int row_begin, int row_end,
                     int f, int path);
/* int64 row pointers: a matrix with more than 2^31 ratings (hugewiki.cu:27-42: 3.1 G; the reference squeezes them
 * through `unsigned int`, hugewiki.cu:2266).  Only the ratings of ONE plan (one shard) must number < 2^31.   */
int cumf_plan_create64(cumf_plan** out, const long long* h_rowptr, int rows, int row_begin, int row_end,
                       int f, int path);
int cumf_plan_destroy(cumf_plan* plan);
/* number of kernels launched by the last cumf_update_factor on this plan */
int cumf_plan_last_launches(const cumf_plan* plan);
/* Optional hint: the number of rows of the opposing factor (what d_factor will hold).  The
 * fused path stages a pre-split fp16 copy of that factor; without the hint its first launch
 * scans d_colidx for the largest id (one extra kernel + a stream synchronisation).  The
 * reference's kernels take the same information implicitly as the extent of thetaT / XT
 * (als.cu:805, 918-919).                                                        */
int cumf_plan_set_factor_rows(cumf_plan* plan, int rows);

/* One half-step of ALS for the plan's rows: for every row u in the plan
 *   A_u, b_u  as in cumf_gram;  d_out[u] <- solve(A_u, b_u, x0 = d_out[u])
 * = the body of the iteration loop als.cu:727-853 (X) / 858-961 (theta).
 * d_colidx/d_val: device pointers to the first rating of row_begin (i.e. the
 * CSR index/value arrays offset by h_rowptr[row_begin]; the plan keeps the row
 * structure); d_factor: the opposing factor [*, f]; d_out: the factor being
 * updated, full [rows][f] array (in/out, rows addressed absolutely).          */
int cumf_update_factor(cumf_plan* plan, const int* d_colidx, const float* d_val,
                       const float* d_factor, float* d_out, float lambda, int solver, float cgIter,
                       void* stream);

/* ---- data-parallel partial Gram (the multi-GPU form of hugewiki.cu:2629-2696) ---
 * When the OPPOSING factor is sharded by rows across GPUs, GPU g forms, for every
 * row u of the side being updated, the partial system over the ratings whose
 * column id falls in g's shard.  Column ids are sorted inside a CSR row, so that
 * share is one contiguous range [h_begin[u], h_end[u]) of the row's ratings
 * (absolute positions in colidx/val; h_begin[u] == h_end[u] for "none").
 * cumf_plan_create_ranges builds the work decomposition for those ranges;
 * cumf_plan_gram materialises
 *   tt[u]  = sum_{j in range(u)} theta_j theta_j^T + lambda * |range(u)| * I
 *   rhs[u] = sum_{j in range(u)} r_uj theta_j
 * (lambda scaled by the LOCAL count, hugewiki.cu:1675-1678, so that the sum over
 * GPUs carries lambda * n_u).  The caller all-reduces tt and rhs over the GPUs
 * (hugewiki.cu:2769-2827 does it with peer copies + reduce kernels; here NCCL on
 * the same device pointers) and solves with cumf_cg / cumf_lu.  d_colidx / d_val
 * are the FULL arrays the positions index.  Asynchronous on `stream`.          */
int cumf_plan_create_ranges(cumf_plan** out, const long long* h_begin, const long long* h_end,
                            int rows, int f, int path);
int cumf_plan_gram(cumf_plan* plan, const int* d_colidx, const float* d_val, const float* d_factor,
                   float lambda, float* d_tt, float* d_rhs, void* stream);

/* ---- resident solver handle (what cumf_doALS is built from) -------------------
 * Uploads CSR/CSC/COO once (the reference re-uploads CSR every iteration,
 * als.cu:734-739) and keeps the factors on the device.  Row ranges select the
 * shard this process updates (one process per GPU: rank g owns X rows
 * [x_begin,x_end) and theta rows [t_begin,t_end)); pass 0,m / 0,n for a single
 * GPU.  Only the owned CSR/CSC slices are uploaded (64-bit offsets inside).   */
typedef struct cumf_als_solver cumf_als_solver;
int cumf_als_create(cumf_als_solver** out, const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr,
                    const float* csrValHostPtr, const int* cscRowIndexHostPtr,
                    const int* cscColIndexHostPtr, const float* cscValHostPtr,
                    const int* cooRowIndexHostPtr, const int* cooRowIndexTestHostPtr,
                    const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, int m, int n, int f,
                    long nnz, long nnz_test, float lambda, int x_begin, int x_end, int t_begin, int t_end,
                    int device, int solver, int path);
/* the same with int64 csrRowIndex / cscColIndex (whole-matrix host arrays beyond 2^31 ratings)                  */
int cumf_als_create64(cumf_als_solver** out, const long long* csrRowIndexHostPtr, const int* csrColIndexHostPtr,
                      const float* csrValHostPtr, const int* cscRowIndexHostPtr,
                      const long long* cscColIndexHostPtr, const float* cscValHostPtr,
                      const int* cooRowIndexHostPtr, const int* cooRowIndexTestHostPtr,
                      const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, int m, int n, int f,
                      long nnz, long nnz_test, float lambda, int x_begin, int x_end, int t_begin, int t_end,
                      int device, int solver, int path);
/* A shard whose rating slices are ALREADY on the device -- generated there (cumf_synth_*) or loaded there shard by
 * shard, the B200 form of hugewiki's per-GPU batch files (hugewiki.cu:2332-2340, 2508-2516): no host copy of the
 * whole matrix.  h_csr_ptr / h_csc_ptr: this shard's own row / column pointers, rebased to 0 (host, int64,
 * x_end - x_begin + 1 and t_end - t_begin + 1 entries); d_*: device arrays of the slices, borrowed (they must
 * outlive the solver); the train samples are the CSR entries; d_test_* (optional): this shard's test samples.    */
int cumf_als_create_device(cumf_als_solver** out, const long long* h_csr_ptr, const int* d_csr_col,
                           const float* d_csr_val, const long long* h_csc_ptr, const int* d_csc_row,
                           const float* d_csc_val, const int* d_test_row, const int* d_test_col,
                           const float* d_test_val, long test_cnt, int m, int n, int f, long nnz, long nnz_test,
                           float lambda, int x_begin, int x_end, int t_begin, int t_end, int device, int solver,
                           int path);
int cumf_als_destroy(cumf_als_solver* s);
/* By default cumf_als_destroy (and so cumf_doALS) cudaFree's every device buffer, like the reference
 * (als.cu:1026-1033).  With CUMF_CACHE_MB=<n> (opt-in) up to n MiB of them are kept for the next
 * solver; this call returns those, and the pinned staging arena of the plan uploads, to the driver. */
int cumf_release_cached_memory(void);
/* Train RMSE as a by-product of the theta half-step: when on (returns 1 if the solver can do it:
 * fused CG path, cooRowIndex == CSR rows; a row shard then reports the ratings of its theta rows, so every shard of a run
 * must switch it on or none), cumf_als_update_theta also accumulates, per row, x^T b + x^T r + reg x^T x
 * from the CG state, and cumf_als_sse returns  sum r^2 - that  for the train set instead of streaming over the ratings
 * (als.cu:967-991) -- valid while X is unchanged since that half-step; otherwise, and whenever the subtraction would
 * keep fewer than three digits, the streaming kernel runs.  Off by default; cumf_doALS turns it on.                    */
int cumf_als_collect_train_sse(cumf_als_solver* s, int on);
int cumf_als_set_factors(cumf_als_solver* s, const float* thetaTHost, const float* XTHost);
int cumf_als_get_factors(cumf_als_solver* s, float* thetaTHost, float* XTHost);
int cumf_als_shape(const cumf_als_solver* s, int* m, int* n, int* f);
/* device pointers of the resident full factor replicas (for NCCL all-gather by the caller) */
float* cumf_als_theta_ptr(cumf_als_solver* s);
float* cumf_als_x_ptr(cumf_als_solver* s);
int cumf_als_update_x(cumf_als_solver* s, void* stream);       /* als.cu:727-853 for the owned rows  */
int cumf_als_update_theta(cumf_als_solver* s, void* stream);   /* als.cu:858-961 for the owned rows  */
/* sum of squared errors over this shard's share of the train / test samples
 * (all of them on a single GPU); the caller all-reduces and takes
 * sqrt(sse/count) (als.cu:991, 1018).                                          */
int cumf_als_sse(cumf_als_solver* s, double* train_sse, double* test_sse, void* stream);
/* `iters` full iterations (X step then theta step), device-timed with CUDA
 * events on `stream`; *ms_out = elapsed milliseconds for the iters.            */
int cumf_als_iterate(cumf_als_solver* s, int iters, float* ms_out, void* stream);
/* per-phase device time (ms) accumulated since the last call with reset != 0:
 * out[0] = X step, out[1] = theta step, out[2] = dominant Gram kernel X side,
 * out[3] = dominant Gram kernel theta side, out[4] = kernel launches, out[5] = iterations */
int cumf_als_timers(cumf_als_solver* s, double* out6, int reset);

/* ---- multi-GPU row sharding -----------------------------------
 * Replaces the X_BATCH / THETA_BATCH model-parallel loop (als.cu:768-777, 881-890) and the
 * peer-copy exchange of hugewiki.cu:2562-2572, 2744-2745: rank g owns rating-balanced row
 * ranges of X and theta and full replicas of both factors; the solver epilogue of its half-step
 * stores every updated row into ALL replicas (peer-mapped pointers over NVLink), so the
 * exchange overlaps the kernel, and one tiny barrier kernel (epoch flags in peer memory) ends
 * the half-step.  No NCCL call on this path.  At f = 100 with long X rows the theta half-step's epilogues also
 * store each solved row's fp16 split form into every rank's X-side gather table, so no rank re-splits all of theta
 * before its X half-step (CUMF_FUSED_SPLIT=0 turns that off; taking cumf_als_theta_ptr does too).
 *
 * (a) one process per GPU (torchrun / MPI): create the solver with this rank's ranges, exchange
 *     the blobs of cumf_als_ipc_export (CUDA IPC handles of the two replicas, the flag
 *     words and the allocation holding the gather table; cumf_als_ipc_blob_bytes() each) with any transport, hand all of them, in rank
 *     order, to cumf_als_ipc_import on every rank, synchronise the ranks once on the host, then
 *     cumf_als_iterate (which ends each half-step with cumf_als_peer_barrier) or
 *     update_x / peer_barrier / update_theta / peer_barrier by hand.                          */
int cumf_als_ipc_blob_bytes(void);
int cumf_als_ipc_export(cumf_als_solver* s, void* blob);
int cumf_als_ipc_import(cumf_als_solver* s, const void* blobs, int nranks, int my_rank);
int cumf_als_peer_barrier(cumf_als_solver* s, void* stream);
/* (b) one process, n devices [first_device, first_device + n): the group creates one shard per
 *     device on its own host thread (plans, allocations and uploads in parallel), connects them
 *     with cudaDeviceEnablePeerAccess and drives them.  cumf_doALS does this itself when
 *     CUMF_GPUS=n is set, so the reference's main.cpp goes multi-GPU unmodified.               */
typedef struct cumf_als_group cumf_als_group;
int cumf_group_create(cumf_als_group** out, const int* csrRowIndexHostPtr, const int* csrColIndexHostPtr,
                      const float* csrValHostPtr, const int* cscRowIndexHostPtr,
                      const int* cscColIndexHostPtr, const float* cscValHostPtr,
                      const int* cooRowIndexHostPtr, const int* cooRowIndexTestHostPtr,
                      const int* cooColIndexTestHostPtr, const float* cooValHostTestPtr, int m, int n, int f,
                      long nnz, long nnz_test, float lambda, int first_device, int n_devices, int solver,
                      int path);
int cumf_group_destroy(cumf_als_group* g);
int cumf_group_size(const cumf_als_group* g);
cumf_als_solver* cumf_group_shard(cumf_als_group* g, int k);
int cumf_group_set_factors(cumf_als_group* g, const float* thetaTHost, const float* XTHost);
int cumf_group_get_factors(cumf_als_group* g, float* thetaTHost, float* XTHost);   /* one replica */
int cumf_group_iterate(cumf_als_group* g, int iters, float* ms_out);   /* ms_out: slowest shard */
int cumf_group_collect_train_sse(cumf_als_group* g, int on);
int cumf_group_sse(cumf_als_group* g, double* train_sse, double* test_sse);   /* summed over shards */


/* ---- synthetic matrices generated shard by shard ON the device (bench / test tooling) ----------
 * BASELINE.json's Hugewiki-scale configuration (m ~ 50 M, n ~ 40 K, 3.1 G ratings over 8 GPUs, hugewiki.cu:27-42)
 * never exists as one host copy: every GPU derives its CSR slice (rows [x0,x1)) and CSC slice (columns [t0,t1)) of
 * one global matrix -- a pure function of (seed, row, column) -- in device memory, without communication
 * (cumf_als_b200/csrc/synth.cu).  cumf_synth_solver builds the shard's solver on those slices where they lie
 * (cumf_als_create_device); cumf_group_create_synth does it for n devices and connects them (row sharding, above).   */
typedef struct cumf_synth_shard cumf_synth_shard;
int cumf_synth_create(cumf_synth_shard** out, long long m, int n, float avg_deg, unsigned long long seed,
                      int x0, int x1, int t0, int t1, long test_cnt, int device);
int cumf_synth_destroy(cumf_synth_shard* sh);
long long cumf_synth_slice(const cumf_synth_shard* sh, int what /* 0 CSR, 1 CSC */, long long* ptr_out);
long long cumf_synth_total_nnz(const cumf_synth_shard* sh);
int cumf_synth_download(const cumf_synth_shard* sh, int what, int* idx_out, float* val_out);
int cumf_synth_download_test(const cumf_synth_shard* sh, int* row_out, int* col_out, float* val_out);
int cumf_synth_solver(cumf_synth_shard* sh, cumf_als_solver** out, int f, float lambda, long nnz_test_total,
                      int solver, int path);
/* CSR -> CSC on the device: one radix sort of (column, row) keys, rows ascending inside every column (what scipy's
 * tocsc gives prepare_netflix_data.py:98-110).  Device pointers; int64 pointer arrays.  Synchronises `stream`.        */
int cumf_csr_to_csc_device(int rows, int cols, long long nnz, const long long* d_rowptr, const int* d_col,
                           const float* d_val, long long* d_colptr_out, int* d_row_out, float* d_val_out, void* stream);
/* theta <- scale * uniform[0,1) (same values on every replica), X <- 0, on the device (main.cpp:72-78's shape) */
int cumf_als_init_factors_device(cumf_als_solver* s, unsigned long long seed, float scale);
int cumf_group_create_synth(cumf_als_group** out, long long m, int n, float avg_deg, unsigned long long seed,
                            long test_per_shard, int f, float lambda, int first_device, int n_devices,
                            int solver, int path);
long cumf_group_nnz(const cumf_als_group* g);
long cumf_group_nnz_test(const cumf_als_group* g);

#ifdef __cplusplus
}
#endif
#endif /* CUMF_ALS_H_ */
