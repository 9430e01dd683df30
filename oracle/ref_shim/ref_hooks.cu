// oracle/ref_shim/ref_hooks.cu -- TEST INFRASTRUCTURE, not product code.
//
// Stage-level entry points into the REAL reference kernels so that golden vectors
// can be produced at the seams of SURVEY.md 8b-b4.  The reference's als.cu is
// textually included from where it lies (REF_ALS_SOURCE, set by build_ref.sh to
// /root/reference/als.cu or to the throw-away LU copy); nothing is copied into
// the repository.  Each launcher repeats the launch configuration of the cited
// call site and nothing else.
#include REF_ALS_SOURCE

extern "C" {

// als.cu:788-817 (X side) -- `tt` is [batch_size][f][f], all pointers device memory.
void ref_get_hermitian(int batch_offset, int batch_size, float* tt, const int* rowIndex, const int* colIndex,
                       float lambda, int m, int f, const float* factor) {
    int block_dim = f / T10 * (f / T10 + 1) / 2;          // als.cu:766-767
    if (block_dim < f / 2) block_dim = f / 2;
    if (f == 100)
        get_hermitian100<<<batch_size, 64, SCAN_BATCH * f / 2 * sizeof(float2)>>>(
            batch_offset, (float2*)tt, rowIndex, colIndex, lambda, m, f, (float2*)factor);   // als.cu:804-805
    else
        get_hermitianT10<<<batch_size, block_dim, SCAN_BATCH * f / 2 * sizeof(float2)>>>(
            batch_offset, tt, rowIndex, colIndex, lambda, m, f, factor);                     // als.cu:816-817
    cudaDeviceSynchronize();
    cudaCheckError();
}

// als.cu:745-757: ythetaT (rows x f row-major) = (R * theta) transposed, via the legacy csrmm2 call.
void ref_rhs(int rows, int cols, int f, int nnz, const float* val, const int* rowIndex, const int* colIndex,
             const float* factor, float* ythetaT) {
    cublasHandle_t handle;
    cublascall(cublasCreate(&handle));
    cusparseHandle_t cushandle = 0;
    cusparsecall(cusparseCreate(&cushandle));
    cusparseMatDescr_t descr;
    cusparsecall(cusparseCreateMatDescr(&descr));
    cusparseSetMatType(descr, CUSPARSE_MATRIX_TYPE_GENERAL);
    cusparseSetMatIndexBase(descr, CUSPARSE_INDEX_BASE_ZERO);
    float* ytheta = 0;
    cudacall(cudaMalloc((void**)&ytheta, (size_t)f * rows * sizeof(float)));
    const float alpha = 1.0f, beta = 0.0f;
    cusparsecall(cusparseScsrmm2(cushandle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_TRANSPOSE, rows, f,
                                 cols, nnz, &alpha, descr, val, rowIndex, colIndex, factor, f, &beta, ytheta, rows));
    cublascall(cublasSgeam(handle, CUBLAS_OP_T, CUBLAS_OP_N, f, rows, &alpha, (const float*)ytheta, rows, &beta,
                           ythetaT, f, ythetaT, f));
    cudaDeviceSynchronize();
    cudacall(cudaFree(ytheta));
    cublasDestroy(handle);
    cusparseDestroy(cushandle);
}

// cg.cu:682-686
void ref_cg(float* A, float* x, float* b, int batchSize, int f, float cgIter) {
    updateXWithCGHost(A, x, b, batchSize, f, cgIter);
}

// als.cu:58-122 with its host pointer scratch (als.cu:834-841)
void ref_lu(int batch_size, int batch_offset, float* ythetaT, float* tt, float* XT, int m, int n, int f, int nnz) {
    cublasHandle_t handle;
    cublascall(cublasCreate(&handle));
    float** devPtrTTHost = 0;
    cudacall(cudaMallocHost((void**)&devPtrTTHost, batch_size * sizeof(*devPtrTTHost)));
    float** devPtrYthetaTHost = 0;
    cudacall(cudaMallocHost((void**)&devPtrYthetaTHost, batch_size * sizeof(*devPtrYthetaTHost)));
    updateX(batch_size, batch_offset, ythetaT, tt, XT, handle, m, n, f, nnz, devPtrTTHost, devPtrYthetaTHost);
    cudacall(cudaFreeHost(devPtrTTHost));
    cudacall(cudaFreeHost(devPtrYthetaTHost));
    cublasDestroy(handle);
}

// als.cu:979-991 (train: grid (count-1)/256+1) and als.cu:1006-1018 (test: grid (count-1)/256)
float ref_rmse(const float* val, const int* row, const int* col, const float* thetaT, const float* XT, int count,
               int f, int test_grid) {
    cublasHandle_t handle;
    cublascall(cublasCreate(&handle));
    float* errors = 0;
    int error_size = 1000;
    cudacall(cudaMalloc((void**)&errors, error_size * sizeof(errors[0])));
    cudacall(cudaMemset(errors, 0, error_size * sizeof(float)));
    if (test_grid)
        RMSE<<<(count - 1) / 256, 256>>>(val, row, col, thetaT, XT, errors, count, error_size, f);
    else
        RMSE<<<(count - 1) / 256 + 1, 256>>>(val, row, col, thetaT, XT, errors, count, error_size, f);
    cudaDeviceSynchronize();
    cudaCheckError();
    float s = 0;
    cublascall(cublasSasum(handle, error_size, errors, 1, &s));
    cudaDeviceSynchronize();
    cudacall(cudaFree(errors));
    cublasDestroy(handle);
    return sqrt(s / count);
}

}  // extern "C"
