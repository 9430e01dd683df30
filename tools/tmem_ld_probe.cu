// tools/tmem_ld_probe.cu -- bring-up probe (GPU box) for the 2-D blocked solve:
// which (lane, column) of a TMEM tile does every thread receive from tcgen05.ld.16x256b?
// One CTA of 4 warps fills a 128-lane x 64-column tile with  value = lane * 1000 + column  through
// tcgen05.st.32x32b (lane = TMEM lane of the thread, the layout the fused kernel's epilogue reads today), reads it
// back with .16x256b.x1 (8 columns per instruction) and .x4 (32 columns) at both 16-lane halves of the warp's
// quadrant, and prints the decoded map plus a check of the expected mma-C-fragment pattern
//   register 2q + e of thread t  <-  lane 8 q' + t / 4 (+ 16 for the upper half), column 8 j + 2 (t % 4) + e.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_ld_probe tools/tmem_ld_probe.cu && tools/tmem_ld_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int COLS = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(float* out_x1, float* out_x4) {
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_slot;
    const uint32_t quad = base + ((uint32_t)(warp * 32) << 16);          // this warp's 32 lanes
    // fill: thread = lane, 16 consecutive columns per instruction
    for (int cc = 0; cc < COLS; cc += 16) {
        uint32_t v[16];
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint((float)((warp * 32 + lane) * 1000 + cc + j));
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
            ::"r"(quad + cc), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
              "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read back: half h = lanes 16 h .. 16 h + 15 of the quadrant
    for (int h = 0; h < 2; ++h) {
        const uint32_t taddr = quad + ((uint32_t)(16 * h) << 16);
        uint32_t a[4];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int r = 0; r < 4; ++r) out_x1[((size_t)tid * 2 + h) * 4 + r] = __uint_as_float(a[r]);
        uint32_t b[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
              "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
            : "r"(taddr + 8) : "memory");                               // columns 8 .. 39
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int r = 0; r < 16; ++r) out_x4[((size_t)tid * 2 + h) * 16 + r] = __uint_as_float(b[r]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(COLS) : "memory");
    }
}

int main() {
    float *d1, *d4;
    cudaMalloc(&d1, 128 * 2 * 4 * sizeof(float));
    cudaMalloc(&d4, 128 * 2 * 16 * sizeof(float));
    probe<<<1, 128>>>(d1, d4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cuda: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> x1(128 * 2 * 4), x4(128 * 2 * 16);
    cudaMemcpy(x1.data(), d1, x1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(x4.data(), d4, x4.size() * 4, cudaMemcpyDeviceToHost);
    int bad1 = 0, bad4 = 0;
    for (int tid = 0; tid < 128; ++tid)
        for (int h = 0; h < 2; ++h) {
            const int warp = tid >> 5, t = tid & 31;
            for (int r = 0; r < 4; ++r) {
                const int v = (int)x1[((size_t)tid * 2 + h) * 4 + r];
                const int want_lane = warp * 32 + 16 * h + (t >> 2) + 8 * (r >> 1), want_col = 2 * (t & 3) + (r & 1);
                if (v != want_lane * 1000 + want_col) ++bad1;
                if (tid < 8 && h == 0) printf("  x1 thread %2d reg %d <- lane %3d col %2d\n", tid, r, v / 1000, v % 1000);
            }
            for (int r = 0; r < 16; ++r) {
                const int v = (int)x4[((size_t)tid * 2 + h) * 16 + r];
                const int j = r >> 2, rr = r & 3;
                const int want_lane = warp * 32 + 16 * h + (t >> 2) + 8 * (rr >> 1), want_col = 8 + 8 * j + 2 * (t & 3) + (rr & 1);
                if (v != want_lane * 1000 + want_col) ++bad4;
                if (tid == 5 && h == 1) printf("  x4 thread %2d (upper half) reg %2d <- lane %3d col %2d\n", tid, r, v / 1000, v % 1000);
            }
        }
    printf("16x256b.x1: %d of %d registers off the mma-C-fragment pattern; .x4: %d of %d\n", bad1, 128 * 2 * 4, bad4, 128 * 2 * 16);
    return 0;
}
