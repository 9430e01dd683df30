// tools/gather4_probe.cu -- bring-up probe (GPU box): does TMA tile::gather4 fetch four arbitrary
// 400-byte factor rows with one instruction, and which box shape does the tensor map need?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/gather4_probe tools/gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int W = 100;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int r0, int r1, int r2, int r3) {
    __shared__ __align__(128) float tile[4 * W];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = threadIdx.x; i < 4 * W; i += blockDim.x) tile[i] = -1.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4 * W * 4));
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            ::"r"(smem_u32(tile)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
    }
    // bounded wait so a wrong descriptor cannot hang the box
    uint32_t done = 0;
    for (int spin = 0; spin < 2000000 && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)));
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * W; i += blockDim.x) out[i] = tile[i];
    if (threadIdx.x == 0) out[4 * W] = (float)done;
}

int main() {
    const int rows = 64;
    std::vector<float> h((size_t)rows * W);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < W; ++c) h[(size_t)r * W + c] = r * 1000.f + c;
    float *d, *dout;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&dout, (4 * W + 1) * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        printf("no cuTensorMapEncodeTiled entry point\n");
        return 1;
    }
    EncodeFn encode = (EncodeFn)fn;
    const int want[4] = {5, 2, 9, 40};
    for (int boxrows : {1, 4}) {
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)rows};
        cuuint64_t gstride[1] = {(cuuint64_t)W * 4};
        cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)boxrows};
        cuuint32_t estride[2] = {1, 1};
        CUresult rc = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box {%d,%d}: encode rc=%d\n", W, boxrows, (int)rc);
        if (rc != CUDA_SUCCESS) continue;
        cudaMemset(dout, 0, (4 * W + 1) * 4);
        probe<<<1, 128>>>(tmap, dout, want[0], want[1], want[2], want[3]);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> o(4 * W + 1);
        cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 0;
        for (int k = 0; k < 4; ++k) {
            bool row_ok = true;
            for (int c = 0; c < W; ++c) row_ok &= (o[k * W + c] == want[k] * 1000.f + c);
            ok += row_ok;
            printf("   slot %d: first vals %.0f %.0f ... last %.0f  (want row %d) %s\n", k, o[k * W], o[k * W + 1], o[k * W + W - 1],
                   want[k], row_ok ? "OK" : "MISMATCH");
        }
        printf("   barrier completed=%d, cuda=%s, rows ok=%d/4\n", (int)o[4 * W], cudaGetErrorString(e), ok);
    }
    return 0;
}
