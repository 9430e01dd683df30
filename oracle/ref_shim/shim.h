// oracle/ref_shim/shim.h -- TEST INFRASTRUCTURE, not product code.
//
// Force-included (`nvcc -include`) when oracle/build_ref.sh compiles the
// UNMODIFIED reference sources where they lie under /root/reference, so the
// real reference can run on sm_100a as the parity oracle.  It only papers over
// two APIs that CUDA 12 removed:
//   * __shfl_down / __shfl   (device_utilities.h:11 in the reference)
//   * cusparseScsrmm2        (als.cu:750, als.cu:867 in the reference)
// Nothing here changes the reference's arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <cusparse.h>

#ifdef __CUDACC__
#define __shfl_down(v, o) __shfl_down_sync(0xffffffffu, (v), (o))
#define __shfl(v, l) __shfl_sync(0xffffffffu, (v), (l))
#endif

// Legacy prototype; defined in csrmm2_shim.cu on top of cusparseSpMM.
cusparseStatus_t cusparseScsrmm2(cusparseHandle_t handle, cusparseOperation_t transA,
                                 cusparseOperation_t transB, int m, int n, int k, int nnz,
                                 const float* alpha, const cusparseMatDescr_t descrA,
                                 const float* csrValA, const int* csrRowPtrA, const int* csrColIndA,
                                 const float* B, int ldb, const float* beta, float* C, int ldc);
