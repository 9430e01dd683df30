// oracle/ref_shim/csrmm2_shim.cu -- TEST INFRASTRUCTURE, not product code.
//
// Re-creates the legacy cuSPARSE entry point the reference calls for its RHS
// (als.cu:750-752 and als.cu:867-869: transA=N, transB=T, C(m x n, col-major,
// ldc=m) = A(m x k CSR) * B^T with B stored n x k col-major, ldb=n) using the
// generic cusparseSpMM API that replaced it.  Only the argument combination
// the reference uses is supported; anything else returns NOT_SUPPORTED.
#include <cuda_runtime.h>
#include <cusparse.h>

extern "C++" cusparseStatus_t cusparseScsrmm2(cusparseHandle_t handle, cusparseOperation_t transA,
                                 cusparseOperation_t transB, int m, int n, int k, int nnz,
                                 const float* alpha, const cusparseMatDescr_t /*descrA*/,
                                 const float* csrValA, const int* csrRowPtrA, const int* csrColIndA,
                                 const float* B, int ldb, const float* beta, float* C, int ldc) {
    if (transA != CUSPARSE_OPERATION_NON_TRANSPOSE || transB != CUSPARSE_OPERATION_TRANSPOSE)
        return CUSPARSE_STATUS_NOT_SUPPORTED;
    cusparseSpMatDescr_t matA = nullptr;
    cusparseDnMatDescr_t matB = nullptr, matC = nullptr;
    cusparseStatus_t st;
    st = cusparseCreateCsr(&matA, m, k, nnz, (void*)csrRowPtrA, (void*)csrColIndA, (void*)csrValA,
                           CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    // B as stored: n rows x k cols, column-major, leading dimension ldb; op(B) = B^T (k x n).
    st = cusparseCreateDnMat(&matB, n, k, ldb, (void*)B, CUDA_R_32F, CUSPARSE_ORDER_COL);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    st = cusparseCreateDnMat(&matC, m, n, ldc, (void*)C, CUDA_R_32F, CUSPARSE_ORDER_COL);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    size_t bufSize = 0;
    st = cusparseSpMM_bufferSize(handle, transA, transB, alpha, matA, matB, beta, matC, CUDA_R_32F,
                                 CUSPARSE_SPMM_ALG_DEFAULT, &bufSize);
    if (st != CUSPARSE_STATUS_SUCCESS) return st;
    void* buf = nullptr;
    if (bufSize > 0 && cudaMalloc(&buf, bufSize) != cudaSuccess) return CUSPARSE_STATUS_ALLOC_FAILED;
    st = cusparseSpMM(handle, transA, transB, alpha, matA, matB, beta, matC, CUDA_R_32F,
                      CUSPARSE_SPMM_ALG_DEFAULT, buf);
    cudaDeviceSynchronize();
    if (buf) cudaFree(buf);
    cusparseDestroySpMat(matA);
    cusparseDestroyDnMat(matB);
    cusparseDestroyDnMat(matC);
    return st;
}
