// tools/mn_major_probe.cu -- bring-up probe (GPU box) for the "direct" staging of the fused kernel:
//   (1) does TMA tile::gather4 with CU_TENSOR_MAP_SWIZZLE_128B land four 128-byte pieces of four arbitrary
//       rows of an fp16 table where the UMMA MN-major SWIZZLE_128B canonical layout expects them?
//   (2) which (leading, stride) byte offsets does a tcgen05.mma kind::f16 shared-memory descriptor need to
//       read a 16(k) x 256(mn) stage laid out that way, for both operands MN-major?
// One CTA: 16 gathered rows -> one 8 KB stage -> D[128 x 256] = A^T B with A = columns [0,128), B = columns
// [0,256) of the gathered rows (exactly the k-step of the fused kernel), read back from TMEM and compared on
// the host.  Every wait is bounded, so a wrong descriptor cannot hang the box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/mn_major_probe tools/mn_major_probe.cu
//   tools/mn_major_probe <arrangement 0|1> <lbo bytes> <sbo bytes>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int COLS = 256, KT = 16, STAGE_BYTES = 8192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool bounded_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < 4000000 && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

struct Idx16 { int v[KT]; };

// kg / ch: byte strides between 8-row k-groups and between 64-element chunks inside a stage
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmap, Idx16 idx, int kg, int ch, uint64_t desc_tmpl, uint32_t idesc,
      unsigned char* out_smem, float* out_d, int* flags, int a_off_bytes, int b_off_bytes) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* stage = smem_raw;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + STAGE_BYTES + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        flags[0] = (int)(smem_u32(stage) & 1023u);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < STAGE_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(stage)[i] = 0x7e007e00u;   // fp16 NaN
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // 16 gathers: lane L -> chunk c = L >> 2, row quad q = L & 3 (rows 4q .. 4q+3)
    if (tid == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(STAGE_BYTES) : "memory");
    __syncthreads();
    if (tid < 16) {
        const int c = tid >> 2, q = tid & 3;
        unsigned char* dst = stage + (q >> 1) * kg + c * ch + (q & 1) * 512;
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            ::"r"(smem_u32(dst)), "l"(&tmap), "r"(smem_u32(&bars[0])), "r"(c * 64), "r"(idx.v[4 * q]), "r"(idx.v[4 * q + 1]),
              "r"(idx.v[4 * q + 2]), "r"(idx.v[4 * q + 3]) : "memory");
    }
    const bool landed = bounded_wait(&bars[0], 0);
    if (tid == 0) flags[1] = landed ? 1 : 0;
    // the ratings of the fused kernel: generic-proxy writes into MN positions 112 / 113 of every gathered row
    if (tid < KT) {
        const int k = tid;
        unsigned char* p = stage + (k >> 3) * kg + 1 * ch + (k & 7) * 128 + ((6 ^ (k & 7)) << 4);
        *reinterpret_cast<__half2*>(p) = __floats2half2_rn((float)(k + 1), (float)(-(k + 1)));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (int i = tid; i < STAGE_BYTES / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(out_smem)[i] = reinterpret_cast<uint32_t*>(stage)[i];
    __syncthreads();
    if (!landed) return;   // (TMEM stays allocated; the process exits anyway)

    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // operand start addresses may sit inside a swizzle atom (multiples of 16 bytes = 8 MN elements): round 2 asks whether
        // the hardware applies the 128B XOR pattern to the final address, as TMA does
        const uint64_t da = desc_tmpl | (uint64_t)(((smem_u32(stage) + (uint32_t)a_off_bytes) & 0x3FFFFu) >> 4);
        const uint64_t db = desc_tmpl | (uint64_t)(((smem_u32(stage) + (uint32_t)b_off_bytes) & 0x3FFFFu) >> 4);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
    }
    const bool mma_done = bounded_wait(&bars[1], 0);
    if (tid == 0) flags[2] = mma_done ? 1 : 0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mma_done) {
        for (int cc = 0; cc < COLS; cc += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) out_d[(size_t)(warp * 32 + lane) * COLS + cc + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

int main(int argc, char** argv) {
    const int arr = argc > 1 ? atoi(argv[1]) : 0;
    const int kg = arr == 0 ? 4096 : 1024, ch = arr == 0 ? 1024 : 2048;
    const int lbo = argc > 2 ? atoi(argv[2]) : ch;     // MN-major SW128: leading = next 64-element chunk
    const int sbo = argc > 3 ? atoi(argv[3]) : kg;     //                 stride  = next 8-row k-group
    // round 2: operand windows that start inside a swizzle atom.  a_off / b_off = first MN element of the A / B operand
    // (multiples of 8), n = UMMA N; A is 128 elements from a_off, B n elements from b_off (windows must end <= 256)
    const int a_off = argc > 4 ? atoi(argv[4]) : 0;
    const int b_off = argc > 5 ? atoi(argv[5]) : 0;
    const int nn = argc > 6 ? atoi(argv[6]) : 256;
    const int rows = 64;
    std::vector<__half> h((size_t)rows * COLS);
    auto tval = [](int r, int c) { return (float)(((r * 7 + c * 3) % 17) - 8); };
    for (int r = 0; r < rows; ++r) for (int c = 0; c < COLS; ++c) h[(size_t)r * COLS + c] = __float2half(tval(r, c));
    __half* d;
    unsigned char* d_smem;
    float* d_out;
    int* d_flags;
    cudaMalloc(&d, h.size() * 2);
    cudaMalloc(&d_smem, STAGE_BYTES);
    cudaMalloc(&d_out, 128 * COLS * 4);
    cudaMalloc(&d_flags, 16);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0, 128 * COLS * 4);
    cudaMemset(d_flags, 0xff, 16);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        printf("no cuTensorMapEncodeTiled entry point\n");
        return 1;
    }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    cuuint64_t gdim[2] = {(cuuint64_t)COLS, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)COLS * 2};
    cuuint32_t box[2] = {64u, 1u};
    cuuint32_t estride[2] = {1, 1};
    CUresult rc = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, gdim, gstride, box, estride,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("arrangement %d (k-group stride %d, chunk stride %d), LBO %d, SBO %d: encode rc=%d\n", arr, kg, ch, lbo, sbo, (int)rc);
    if (rc != CUDA_SUCCESS) return 1;
    Idx16 idx;
    const int want[KT] = {5, 2, 9, 40, 63, 0, 17, 17, 33, 8, 21, 50, 1, 62, 30, 11};
    for (int k = 0; k < KT; ++k) idx.v[k] = want[k];
    // MN-major SWIZZLE_128B descriptor: version 1 (bit 46), layout type 2 (bits 61..63)
    const uint64_t desc_tmpl = ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    // c F32, a/b F16, both MN-major (bits 15, 16), N = 256, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(nn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto elem_bytes = [&](int e) { return (e >> 6) * ch + ((e & 63) >> 3) * 16; };
    printf("  operand windows: A = MN [%d, %d), B = MN [%d, %d) (N = %d)\n", a_off, a_off + 128, b_off, b_off + nn, nn);
    const int smem = STAGE_BYTES + 128;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(tmap, idx, kg, ch, desc_tmpl, idesc, d_smem, d_out, d_flags, elem_bytes(a_off), elem_bytes(b_off));
    cudaError_t e = cudaDeviceSynchronize();
    int flags[4];
    cudaMemcpy(flags, d_flags, 16, cudaMemcpyDeviceToHost);
    printf("  cuda=%s  smem base & 1023 = %d  tma landed=%d  mma done=%d\n", cudaGetErrorString(e), flags[0], flags[1], flags[2]);
    if (e != cudaSuccess) return 2;
    // expected gathered matrix (with the rating columns 112/113 overwritten)
    std::vector<float> g((size_t)KT * COLS);
    for (int k = 0; k < KT; ++k) {
        for (int c = 0; c < COLS; ++c) g[(size_t)k * COLS + c] = tval(want[k], c);
        g[(size_t)k * COLS + 112] = (float)(k + 1);
        g[(size_t)k * COLS + 113] = (float)(-(k + 1));
    }
    std::vector<unsigned char> raw(STAGE_BYTES);
    cudaMemcpy(raw.data(), d_smem, STAGE_BYTES, cudaMemcpyDeviceToHost);
    int bad_smem = 0;
    for (int k = 0; k < KT; ++k)
        for (int mn = 0; mn < COLS; ++mn) {
            const int c = mn >> 6, el = mn & 63;
            const int off = (k >> 3) * kg + c * ch + (k & 7) * 128 + (((el >> 3) ^ (k & 7)) << 4) + (el & 7) * 2;
            __half v;
            memcpy(&v, &raw[off], 2);
            if (__half2float(v) != g[(size_t)k * COLS + mn]) {
                if (bad_smem < 6) printf("  smem mismatch k=%d mn=%d off=%d got %g want %g\n", k, mn, off, __half2float(v), g[(size_t)k * COLS + mn]);
                ++bad_smem;
            }
        }
    printf("  smem layout (TMA swizzle == UMMA SW128 MN-major expectation): %d mismatches of %d\n", bad_smem, KT * COLS);
    if (bad_smem) {
        // where did row 0 / chunk 0 land?  print the first 128-byte line per row of k-group 0
        for (int k = 0; k < 8; ++k) {
            printf("   line %d:", k);
            for (int j = 0; j < 64; j += 8) { __half v; memcpy(&v, &raw[k * 128 + j * 2], 2); printf(" %g", __half2float(v)); }
            printf("   (row %d starts %g %g)\n", want[k], tval(want[k], 0), tval(want[k], 8));
        }
    }
    if (flags[2] == 1) {
        std::vector<float> o((size_t)128 * COLS);
        cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 128 && a_off + i < COLS; ++i)
            for (int j = 0; j < nn && b_off + j < COLS; ++j) {
                float acc = 0.f;
                for (int k = 0; k < KT; ++k) acc += g[(size_t)k * COLS + a_off + i] * g[(size_t)k * COLS + b_off + j];
                if (o[(size_t)i * COLS + j] != acc) {
                    if (bad < 6) printf("  D mismatch i=%d j=%d got %g want %g\n", i, j, o[(size_t)i * COLS + j], acc);
                    ++bad;
                }
            }
        printf("  MMA result: %d mismatches of %d  -> %s\n", bad, 128 * nn, bad == 0 ? "DESCRIPTOR OK" : "descriptor wrong");
    }
    return 0;
}
