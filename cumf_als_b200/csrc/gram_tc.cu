// gram_tc.cu -- the fused B200 half-step kernel:
//   TMA tile::gather4 row gather -> split-fp16 operands -> tcgen05.mma (accumulators in TMEM) ->
//   epilogue straight into an in-register conjugate-gradient solve.
// A_u is never written to HBM; per rating the kernel moves one f-wide factor row, one int32 column
// index and one fp32 value (B_gram_fused).
//
// What it replaces in the reference: get_hermitian100 (als.cu:443-569) + the cuSPARSE RHS
// pass (als.cu:750-757) + updateXWithCGKernel (cg.cu:36-231) for f = 100, i.e. the body of
// the iteration loop als.cu:727-853 / 858-961.  The disabled fused kernel
// alsUpdateFeature100 (cg.cu:726-1189, "register pressure and low occupancy", als.cu:809)
// is the precedent; here the Gram tile lives in TMEM instead of 100 registers x 55 threads.
//
// Numerics.  The reference accumulates theta_j theta_j^T in fp32.  Tensor cores have no fp32
// input mode, so every gathered fp32 value v is split as  v = hi + lo,
//   hi = v with the low 13 mantissa bits cleared (exactly representable in fp16),
//   lo' = fp16( (v - hi) * 2048 )                      (scaled so it stays a normal fp16)
// and  A = hi^T hi + (hi^T lo' + lo'^T hi) / 2048  is accumulated in fp32 in TMEM with
// kind::f16 MMAs (the dropped lo^T lo term is < 2^-20 relative).  The ratings get the same
// split and ride along as two extra operand rows, so  b = sum r_uj theta_j  comes out of the
// same MMAs.  The tensor core truncates when it accumulates (measured ~6e-8 relative bias per
// 16 ratings): accumulation chains are cut every 256 ratings and the tiles summed with
// round-to-nearest in registers.  The CG is the fp32 register-resident solve of cg.cu with
// identical semantics (cg.cu:47-230).
//
// Staging, two variants that feed the tensor core bit-identical operands in the same order:
//   direct (default): split_factor_kernel rewrites the opposing factor once per half-step as an fp16 table
//     [hi | r slots | lo'] (512 B per row); tile::gather4 with CU_TENSOR_MAP_SWIZZLE_128B drops the gathered rows
//     straight into the UMMA MN-major SWIZZLE_128B operand layout (tools/mn_major_probe.cu pins the layout against
//     the hardware).  No staging ring, no conversion pass: per 16 ratings the shared-memory pipe carries 8 KB of TMA
//     writes + the operand reads.  Stages hold 64 (or 32) ratings = 4 (2) MMA k-groups per barrier round trip.
//   fp32 ring (CUMF_TC_DIRECT=0): gather the fp32 rows, convert them to (hi | lo' | r) fp16 in the UMMA K-major
//     no-swizzle layout with the stage-worker warps (the round-1 kernel; 16 ratings per stage).
//
// Persistent CTAs, one per SM, each owning a contiguous, cost-balanced range of row chunks (so its ratings are one
// stream); warp roles (Cfg<> below has the shapes and setmaxnreg budgets):
//   MMA issuer (warp 3): walks its tiles (one table word per 256-rating TMEM tile), waits for the stage barrier and
//     issues tcgen05.mma cta_group::1 kind::f16, M = 128 -- long rows (kSym): one MMA per k-group, N = 240,
//     [hi|r]^T [hi|r|0|lo'], the epilogue symmetrises; short rows: that with N = 256 plus N = 128 lo'^T [hi|r];
//     tcgen05.commit frees operand stages / publishes accumulator tiles; owns the TMEM allocation (2 x 256 columns)
//   stage workers: read their stages' descriptors from the precomputed stage table, prefetch column ids / ratings,
//     arm the stage mbarrier, one elected lane issues the gathers (UTMALDG.2D.GATHER4, 4 rows x 128 B each), drop
//     the split ratings into the landed rows, hand the stage to the issuer, refill the slot when its MMAs retire
//   solver warpgroups (2 for long rows, 3 for short rows with direct staging): tcgen05.ld of row i of [A | b] into
//     registers, + lambda*n_u, 6-step CG with named-barrier reductions, x written back; chunks of split rows store
//     their partial [A|b] instead (reduced + solved by the unfused kernels)
// Waits park the warp (try_wait with a suspend-time hint -> NANOSLEEP.SYNCS) and carry a watchdog.
#include "common.cuh"

#include <cuda.h>          // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_fp16.h>

#include <algorithm>

namespace cumf {
namespace {

// ---- compile-time geometry (f = 100) ---------------------------------------------------
constexpr int F = 100;                    // rank handled by this kernel
constexpr int FP = 112;                   // rank padded to a multiple of 16 (UMMA N granularity at M=128)
constexpr int KT = 16;                    // gathered rows per MMA k-step (fp16 UMMA K)
constexpr int S1 = 16;                    // fp32 staging ring depth: 16 x 6.5 KB of gathered rows in flight per SM
constexpr int S2 = 8;                     // fp16 operand ring depth: one slot per staging warp (a parity wait
                                          // needs a waiter that observes every phase of its barrier)
constexpr int ROW_BYTES = F * 4;          // one factor row
constexpr int GROUP_ROWS = 4;             // rows fetched by one TMA tile::gather4 instruction
constexpr int GROUP_BYTES = 1664;         // 4 x 400 B padded to a multiple of 128 B (TMA destination alignment)
constexpr int GROUP_FLOATS = GROUP_BYTES / 4;
constexpr int STAGE_F32_BYTES = (KT / GROUP_ROWS) * GROUP_BYTES;   // 6656
// operand stage (256 rows x 16 k): [0,112) hi, 112 r_hi, 113 r_lo', [114,128) zero, [128,240) lo', [240,256) zero
constexpr int OP_ROWS = 256;
constexpr int R_ROW = FP;                 // 112: the two rating rows sit right behind hi
constexpr int LO_ROW = FP + 16;           // 128: first lo' row
constexpr int OP_GROUP_BYTES = 256;       // 8 rows x (2 K-core-matrices x 16 B): SBO
constexpr int OP_KCORE_BYTES = 128;       // one 8x16B core matrix: LBO
constexpr int OP_STAGE_BYTES = OP_ROWS / 8 * OP_GROUP_BYTES;   // 8192
constexpr int MMA_WARP = 3;
// Shape of the CTA (20 or 16 warps; the roles are fixed per warp, the setmaxnreg budgets per warpgroup).  The register
// pool is what the launch allocated (threads x launch registers), the budgets must fit into it.
//   fp32 staging, and direct staging for long rows (kSym):  warps 0-2 idle | 3 MMA issuer | 4-11 stage workers (direct
//     64-rating stages: five of them, one 32 KB ring slot each) | 12-19 two solver warpgroups;
//     640 threads x 96 registers = 61440 >= 128 (P + 2 S + 2 E).
//   direct staging for short rows (!kSym, "wide"): the launch is bound by the per-row drain + CG (ncu: the solver
//     warpgroups are 85 % busy, the issuer waits for free TMEM tiles), and the direct workers only issue gathers, so
//     three of them share the issuer's warpgroup and a THIRD solver warpgroup takes their place:
//     warps 0-2 stage workers (two 32 KB or four 16 KB ring slots each) | 3 MMA issuer | 4-15 three solver warpgroups;
//     512 threads x 128 registers = 65536 = 128 (56 + 3 x 152).
// (An earlier 3-warpgroup shape that kept columns [64,100) of every row in shared memory to fit the 640-thread CTA's
// register pool was slower; here every row stays in registers.)
// kRows: ratings per stage -- 16 (fp32 staging: one MMA k-group per stage), 32 or 64 (direct staging: 2 / 4 k-groups per
// barrier round trip, flag word and commit of the MMA-issuing warp).
template <bool kSym, bool kDirect, int kRows> struct Cfg {
    static_assert(kDirect ? (kRows == 32 || kRows == 64) : kRows == 16, "stage size");
    static constexpr bool kWide = kDirect && !kSym;
    static constexpr int kGroups = kRows / 16;                     // MMA k-groups per stage
    static constexpr int kSubSteps = 256 / kRows;                  // stages per TMEM tile: 256 ratings in every variant
    static constexpr int kStageBytes = kDirect ? kRows * 512 : 8192;   // operand bytes of one stage (SPLIT_ROW_BYTES / OP_STAGE_BYTES)
    static constexpr int kWG = kWide ? 3 : 2;                      // solver warpgroups
    static constexpr int kFirstWorker = kWide ? 0 : 4;
    static constexpr int kWorkers = kWide ? 3 : (kRows == 64 ? 5 : 8);       // stage worker warps (the other warps of 4..11 idle)
    static constexpr int kSlotsPerWorker = kWide ? (kRows == 64 ? 2 : 4) : 1;   // direct ring: slots owned by one worker
    static constexpr int kSlots = kWorkers * kSlotsPerWorker;      // direct ring depth
    static constexpr int kFirstEpiWarp = kWide ? 4 : 12;
    static constexpr int kThreads = (kFirstEpiWarp + 4 * kWG) * 32;          // 640 / 512
    static constexpr int kRegsLaunch = kWide ? 128 : 96;
    static constexpr int kRegsProd = kWide ? 56 : (kDirect ? 80 : 48);       // warpgroup 0 (issuer; + the workers if kWide)
    static constexpr int kRegsStage = kDirect ? 48 : 64;                     // warpgroups 1, 2 (workers; absent if kWide)
    static constexpr int kRegsEpi = 152;
    static_assert(128 * (kRegsProd + (kWide ? 0 : 2 * kRegsStage) + kWG * kRegsEpi) <= kThreads * kRegsLaunch, "setmaxnreg budgets exceed the CTA register pool");
    static_assert(kThreads * kRegsLaunch <= 65536, "launch registers");
    static_assert(!kDirect || kSlots <= 16, "ring barriers");
    static_assert(kGroups >= 1 && kGroups <= 4, "k-groups per stage (2-bit fields in the stage flags)");
};
constexpr int MAX_WG = 3;

constexpr int SUB_STEPS = 16;             // k-steps (x16 ratings) accumulated in TMEM before the tile is drained
// "direct" staging (kDirect): the opposing factor is pre-split once per half-step into an fp16 table
//   row j = [ hi_j (100) | 0 (12) | r slots (2) | 0 (14) | lo'_j (100) | 0 (28) ]      256 halfs = 512 B
// (+ one all-zero row behind the last), and TMA tile::gather4 with CU_TENSOR_MAP_SWIZZLE_128B drops the 16 gathered
// rows of a k-step straight into the UMMA **MN-major** SWIZZLE_128B canonical layout: four 64-element chunks per
// row, 8 rows x 128 B per swizzle atom.  No fp32 staging ring, no conversion pass: per k-step the shared-memory pipe
// carries 8 KB of TMA writes + the UMMA operand reads instead of 6.4 + 6.4 + 6.5 KB of staging traffic on top of them.
// A direct stage holds 32 or 64 ratings = two / four 16-row MMA k-groups: one barrier round trip, one flag word and one
// commit of the MMA-issuing warp (whose instruction stream bounds long rows) per stage instead of per 16 ratings.
constexpr int DKT_MAX = 64;               // largest direct stage
constexpr int SPLIT_COLS = 256;           // fp16 elements per row of the pre-split table
constexpr int SPLIT_ROW_BYTES = SPLIT_COLS * 2;
constexpr int SPLIT_CHUNK = 64;           // elements per 128-byte swizzle line
constexpr int DGROUP_BYTES = KT * SPLIT_ROW_BYTES;    // 8192: one MMA k-group (16 rows)
// inside a k-group: 8-row swizzle atoms of one 64-element chunk are 1 KB, the four chunks of 8 rows 4 KB
// (the CUTLASS tile_to_shape order of Layout_MN_SW128_Atom; tools/mn_major_probe.cu checks it against the hardware)
constexpr int D_CHUNK_STRIDE = 1024;      // -> descriptor leading byte offset
constexpr int D_KG_STRIDE = 4096;         // -> descriptor stride byte offset
static_assert(SUB_STEPS * KT == 256, "every staging cuts the accumulation chains at the same ratings");
constexpr int TMEM_COLS = 512;
// accumulator tile, one MMA per k-step:  D[0:128, 0:240] (+)= [hi | r]^T [hi | r | 0 | lo']
//   lanes 0..99 (features i):  [0,112) P[i][:] = hi_i . hi_j | 112 hi_i . r_hi | 113 hi_i . r_lo' | [128,240) S[i][:] = hi_i . lo'_j
//   lane 112 (the r_hi row):   [0,112) r_hi . hi_j                                                | [128,240) r_hi . lo'_j
// The symmetric half of the cross term (lo'^T hi = S^T) is not computed by the tensor core: the epilogue forms
//   G = P/2 + S/2048  (row i per thread, lane 112 carries the rating row)   and   [A|b] = G + G^T
// through a shared-memory transpose once per chunk.  That halves the MMA work of the cross terms and the
// operand bytes the tensor core reads per k-step.
constexpr int ACC_COLS = 256;
constexpr int N1_SYM = 240;               // kSym: B = operand rows [0,240), A = rows [0,128)
// !kSym (short rows, where the solver warpgroups are the bottleneck and the transpose would cost more than the
// MMA it saves): the tensor core also forms lo'^T [hi | r] onto columns [128,256), so S = hi^T lo' + lo'^T hi,
// column 240 = lo'^T r_hi, and row i of [A|b] = P + S/2048 needs no exchange between threads.
constexpr int N1 = 256;                   // B = all 256 operand rows, A = hi
constexpr int N2 = 128;                   // B = rows [0,128) (hi, r), A = lo', D columns [128,256)
constexpr int BCOL_LO = LO_ROW + R_ROW;   // 240: TMEM column of  lo'^T r_hi  (!kSym)
constexpr int SCOL = LO_ROW;              // 128: first TMEM column of S
constexpr int BCOL_HI = R_ROW;            // 112: TMEM column of  hi^T r_hi  (next column: hi^T r_lo')
constexpr int TR_ROWS = 50;               // rows of G exchanged per transpose pass (2 passes cover f = 100)
static_assert(2 * TR_ROWS == F, "two transpose passes");
constexpr float kLoScale = 2048.0f;
constexpr float kLoInv = 1.0f / 2048.0f;
// cg.cu:31,195: `rsnew < 1e-4` compares in double.  For a float rsnew that is exactly
// rsnew < nextafterf(1e-4f, +inf): the double 1e-4 lies strictly between two adjacent floats.
constexpr float kCgErrorF = 1.00000005e-4f;

// stage flags: chunk = one row (or one piece of a split row); sub = one TMEM accumulation tile
constexpr uint32_t FLAG_CHUNK_FIRST = 1u, FLAG_CHUNK_LAST = 2u, FLAG_SUB_FIRST = 4u, FLAG_SUB_LAST = 8u;
// ... followed by what the MMA issuer would otherwise have to count at run time (the CTA partition is fixed per
// plan, so the tile sequence of every CTA is known when the table is built):
//   bit 4  TMEM buffer of the stage's tile (tile index within the CTA & 1)
//   bits 5-6  solver warpgroup that drains it (chunk index within the CTA mod #warpgroups) -> acc_full[wg][bit4]
//   bit 7  parity to wait for on acc_empty[bit4] before the tile's first MMA
//   bits 8-9  (direct stages) number of 16-rating MMA k-groups the stage fetches and issues, minus one
constexpr int FLAG_BUF_SHIFT = 4, FLAG_WG_SHIFT = 5, FLAG_EMPTY_PARITY_SHIFT = 7, FLAG_GROUPS_SHIFT = 8;
// StageDesc::info of the first stage of a tile additionally carries (bits above the flags):
constexpr int TILE_STAGES_SHIFT = 20;             // bits 20-24: stages in the tile (1 .. 16)
constexpr int TILE_LAST_GROUPS_SHIFT = 25;        // bits 25-26: k-groups of the tile's last stage, minus one (earlier stages are full)

// One k-step of work: 16 (or fewer) consecutive ratings of one chunk.  Precomputed per plan
// (stage table), so every stage worker warp is autonomous.
struct StageDesc {
    int pos;             // offset of the stage's first rating (relative to the plan's first rating)
    uint32_t info;       // cnt (bits 0..7) | flags << 8
};
static_assert(sizeof(StageDesc) == 8, "StageDesc is loaded as one 8-byte word");

// fp32 staging: fp32 gather ring (S1 x 6.5 KB) followed by the converted fp16 operand ring (S2 x 8 KB)
constexpr int CONV_RING_BYTES = S1 * STAGE_F32_BYTES + S2 * OP_STAGE_BYTES;     // 172032
constexpr int NBAR = 16;                  // stage barriers / metadata slots (the largest ring)
static_assert(NBAR >= S1 && NBAR >= S2, "barrier / metadata arrays cover every ring");
// dynamic shared memory, used in place (SWIZZLE_128B atoms need 1024-byte alignment)
template <int kRingBytes, int kScratchFloats, int kStageBytes> struct __align__(1024) SmemT {
    unsigned char ring[kRingBytes];   // kDirect: kSlots stages, TMA destination == UMMA operand; else the two rings above
    float stage_vals[NBAR][DKT_MAX]; // the ratings of the stage in flight in each gather slot (zero beyond cnt)
    uint32_t meta_op[NBAR];      // stage flags forwarded to the MMA warp
    __align__(16) int stage_idx[NBAR][DKT_MAX];   // direct staging: the column ids of the stage being fetched into each slot
    // per solver warpgroup, kSym only: TR_ROWS rows of G (+ the rating row) for the transpose (2 x 5100 floats)
    float solver_scratch[kScratchFloats];
    float sp[MAX_WG][2][128];    // CG direction vector per solver warpgroup, double buffered
    float red[MAX_WG][3][4];     // cross-warp partial sums
    unsigned long long full_f32[NBAR], full_op[NBAR], empty_op[NBAR];
    // acc_full[w][buf]: tile in TMEM buffer `buf` complete, for solver warpgroup w.  One barrier per
    // (consumer, buffer): a parity wait is only sound if its waiter observes every phase, and the
    // warpgroups take turns irregularly on the buffers (tiles per chunk vary).
    unsigned long long acc_full[MAX_WG][2], acc_empty[2];
    uint32_t tmem_base;
    __device__ __forceinline__ unsigned char* f32_stage(int slot) { return ring + slot * STAGE_F32_BYTES; }
    __device__ __forceinline__ unsigned char* op_stage(int slot) { return ring + S1 * STAGE_F32_BYTES + slot * OP_STAGE_BYTES; }
    __device__ __forceinline__ unsigned char* dstage(int slot) { return ring + slot * kStageBytes; }
};
template <bool kSym, bool kDirect, int kRows> struct SmemFor {
    using C = Cfg<kSym, kDirect, kRows>;
    static constexpr int kRing = kDirect ? C::kSlots * C::kStageBytes : CONV_RING_BYTES;
    static constexpr int kScratch = kSym ? 2 * (TR_ROWS + 1) * F : 4;
    using type = SmemT<kRing, kScratch, C::kStageBytes>;
    static_assert(sizeof(type) <= 232448, "Smem exceeds the 227 KB a CTA can opt into");
};

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Waiting warps must not spin: a polling loop is ready to issue most of the time and takes issue slots from the warps of
// its scheduler that have work (the MMA issuer shares its SM sub-partition with two solver warps and two stage workers
// that mostly wait; in the ncu source view a fifth of all issued instructions were wait-loop polls).  try_wait with a
// suspend-time hint parks the warp in hardware until the phase completes or the hint (in ns) expires.
// Watchdog: a protocol error (or a fault in another role's warp) must not leave the GPU waiting for ever -- after
// ~4 s (2^33 SM cycles; a healthy wait is microseconds) the CTA traps and the launch fails loudly.
constexpr uint32_t kWaitHintNs = 100000u;
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    long long t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity), "r"(kWaitHintNs) : "memory");
        if (!done) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > (1ll << 33)) __trap();
        }
    } while (!done);
}
// TMA tile::gather4: four rows (row coordinates r0..r3, column coordinate 0) of the 2-D tensor described
// by `tmap` (box {F, 1}) land back to back at smem_dst; complete_tx(4 * ROW_BYTES) on `bar`.
__device__ __forceinline__ void tma_gather4(void* smem_dst, const CUtensorMap* tmap, int r0, int r1, int r2, int r3,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
// same, for the pre-split fp16 table: four 64-element (128-byte) pieces starting at column `col`
__device__ __forceinline__ void tma_gather4_col(void* smem_dst, const CUtensorMap* tmap, int col, int r0, int r1, int r2, int r3,
                                                unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 (fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_smem_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_smem_addr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) leading (K-direction core-matrix) byte offset >> 4,
//   [32,46) stride (8-row group) byte offset >> 4, [46,48) version = 1, [61,64) layout = 0.
__host__ __device__ constexpr uint64_t smem_desc_template(bool swap_lbo_sbo) {
    return ((uint64_t)((swap_lbo_sbo ? OP_GROUP_BYTES : OP_KCORE_BYTES) >> 4) << 16) |
           ((uint64_t)((swap_lbo_sbo ? OP_KCORE_BYTES : OP_GROUP_BYTES) >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint64_t tmpl) {
    return tmpl | (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format F16 (0),
// K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool mn_major = false) {
    return (1u << 4) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major SWIZZLE_128B descriptor of a direct stage (cute::UMMA::make_umma_desc<Major::MN>, LayoutType::B128:
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): leading offset = distance between 64-element chunks,
// stride offset = distance between 8-row k-groups, layout type 2 at [61,64).
__host__ __device__ constexpr uint64_t smem_desc_template_direct(int lbo_bytes, int sbo_bytes) {
    return ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// ---- solver warpgroup helpers ---------------------------------------------------------------
__device__ __forceinline__ float wg_sum(float v, float* red4, int warp_in_wg, int lane, int bar_id) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red4[warp_in_wg] = v;
    named_bar_sync(bar_id, 128);
    return (red4[0] + red4[1]) + (red4[2] + red4[3]);
}

template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// drain one TMEM accumulator tile into / onto this thread's copy of row i (all 100 columns in registers):
//   kSym:  G = P/2 + S/2048 and its rating column (symmetrised later);   !kSym:  [A|b] = P + S/2048 directly.
template <bool kFirst, bool kSym>
__device__ __forceinline__ void drain_tile(uint32_t taddr, float (&a)[F], float& b) {
    constexpr float kP = kSym ? 0.5f : 1.0f;
#pragma unroll
    for (int cc = 0; cc < 96; cc += 16) {
        uint32_t p[16], s[16];
        tmem_ld16(taddr + cc, p);
        tmem_ld16(taddr + SCOL + cc, s);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float v = fmaf(__uint_as_float(s[j]), kLoInv, kP * __uint_as_float(p[j]));
            a[cc + j] = kFirst ? v : a[cc + j] + v;
        }
    }
    uint32_t p[4], s[4], bh[4], bl[4];
    tmem_ld4(taddr + 96, p);
    tmem_ld4(taddr + SCOL + 96, s);
    tmem_ld4(taddr + BCOL_HI, bh);      // hi^T r_hi, hi^T r_lo'
    if (!kSym) tmem_ld4(taddr + BCOL_LO, bl);      // lo'^T r_hi
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float v = fmaf(__uint_as_float(s[j]), kLoInv, kP * __uint_as_float(p[j]));
        a[96 + j] = kFirst ? v : a[96 + j] + v;
    }
    const float tb = kSym ? fmaf(__uint_as_float(bh[1]), kLoInv, 0.5f * __uint_as_float(bh[0]))
                          : fmaf(__uint_as_float(bh[1]) + __uint_as_float(bl[0]), kLoInv, __uint_as_float(bh[0]));
    b = kFirst ? tb : b + tb;
}

template <int kRows> __device__ __forceinline__ int chunk_steps(const Chunk& ck) { return max(1, (ck.end - ck.begin + kRows - 1) / kRows); }

// stage table of a plan: one StageDesc per stage (kRows ratings), in chunk order (built once per plan)
// chunk_meta[c] = (index of the chunk's first tile within its CTA) << 2 | (index of the chunk within its CTA mod #solver warpgroups)
template <int kRows, int kSub>
__global__ void fill_stage_table_kernel(const Chunk* __restrict__ chunks, const int* __restrict__ chunk_stage_base,
                                        const int* __restrict__ chunk_meta, int nchunks, StageDesc* __restrict__ table) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const Chunk ck = chunks[c];
    const int steps = chunk_steps<kRows>(ck);
    const uint32_t wg = (uint32_t)chunk_meta[c] & 3u;
    const uint32_t tile0 = (uint32_t)chunk_meta[c] >> 2;
    StageDesc* out = table + chunk_stage_base[c];
    for (int s = 0; s < steps; ++s) {
        const int pos = ck.begin + s * kRows;
        const int cnt = max(0, min(kRows, ck.end - pos));
        const bool last = (s == steps - 1);
        const uint32_t tile = tile0 + (uint32_t)(s / kSub);
        const uint32_t flags = (s == 0 ? FLAG_CHUNK_FIRST : 0u) | (last ? FLAG_CHUNK_LAST : 0u) |
                               ((s % kSub) == 0 ? FLAG_SUB_FIRST : 0u) |
                               ((last || (s % kSub) == kSub - 1) ? FLAG_SUB_LAST : 0u) |
                               ((tile & 1u) << FLAG_BUF_SHIFT) | (wg << FLAG_WG_SHIFT) |
                               ((((tile >> 1) & 1u) ^ 1u) << FLAG_EMPTY_PARITY_SHIFT) |
                               ((uint32_t)(max(1, (cnt + KT - 1) / KT) - 1) << FLAG_GROUPS_SHIFT);
        uint32_t tile_bits = 0;
        if ((s % kSub) == 0) {
            // first stage of a TMEM tile: what the direct-staging MMA issuer needs for the whole tile, so that it reads one
            // word per tile instead of one per stage -- number of stages, and whether the tile's last stage has a second k-group
            const int tile_stages = min(kSub, steps - s);
            const int last_pos = ck.begin + (s + tile_stages - 1) * kRows;
            const int last_cnt = max(0, min(kRows, ck.end - last_pos));
            tile_bits = ((uint32_t)tile_stages << TILE_STAGES_SHIFT) |
                        ((uint32_t)(max(1, (last_cnt + KT - 1) / KT) - 1) << TILE_LAST_GROUPS_SHIFT);
        }
        out[s] = StageDesc{pos, (uint32_t)cnt | (flags << 8) | tile_bits};
    }
}

// ---- direct staging: pre-split of the opposing factor --------------------------------------------
// out[row] (32 x 16 bytes) = [ hi (13 pieces, elements >= 100 zero) | 0 x 3 | lo' (13 pieces) | 0 x 3 ]; row == rows is
// the all-zero padding row.  Same arithmetic as the in-kernel conversion of the fp32 staging path, so both stagings feed
// the tensor core identical operands.
__global__ void __launch_bounds__(256) split_factor_kernel(const float* __restrict__ fac, int rows, uint4* __restrict__ out) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t row = gid >> 5;
    const int g = (int)(gid & 31);
    if (row > (size_t)rows) return;
    uint4 o = make_uint4(0, 0, 0, 0);
    const int gg = g & 15, c0 = gg * 8;
    if (row < (size_t)rows && c0 < F) {
        const float4* src = reinterpret_cast<const float4*>(fac + row * F + c0);
        const float4 a = __ldg(src);
        const float4 b = (c0 + 4 < F) ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const float h0 = __uint_as_float(__float_as_uint(v[k]) & 0xFFFFE000u);
            const float h1 = __uint_as_float(__float_as_uint(v[k + 1]) & 0xFFFFE000u);
            const __half2 hh = (g < 16) ? __floats2half2_rn(h0, h1)                                               // exact
                                        : __floats2half2_rn((v[k] - h0) * kLoScale, (v[k + 1] - h1) * kLoScale);
            w[k >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        o = make_uint4(w[0], w[1], w[2], w[3]);
    }
    out[gid] = o;
}

// largest column id a plan touches (+1 = rows of the opposing factor that can be gathered)
__global__ void max_index_kernel(const int* __restrict__ idx, long long n, int* __restrict__ out) {
    int m = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = max(m, __ldg(idx + i));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(out, m);
}

// kDirect: `factor_map` describes the pre-split fp16 table (box {64, 1}, SWIZZLE_128B), `zero_row` is the index of its
// all-zero row (the padding of ragged k-groups).
template <bool kSym, bool kDirect, int kRows>
__global__ void __launch_bounds__(Cfg<kSym, kDirect, kRows>::kThreads, 1)
als_fused_f100_kernel(const Chunk* __restrict__ chunks, const int* __restrict__ cta_chunk_ptr,
                      const StageDesc* __restrict__ stage_tab, const int* __restrict__ cta_stage_ptr,
                      const int* __restrict__ colidx, const float* __restrict__ val,
                      const __grid_constant__ CUtensorMap factor_map, float* __restrict__ out, float lambda, float cg_iter,
                      float* __restrict__ scratchA, float* __restrict__ scratchB, uint64_t desc_tmpl,
                      double* __restrict__ sse_terms, int zero_row, PeerOut peers) {
    // dynamic shared memory is used in place (no pointer arithmetic through integers, so the
    // compiler keeps the shared address space and emits LDS/STS)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    using C = Cfg<kSym, kDirect, kRows>;
    using Smem = typename SmemFor<kSym, kDirect, kRows>::type;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    constexpr int NUM_THREADS = C::kThreads;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int c_begin = cta_chunk_ptr[blockIdx.x], c_end = cta_chunk_ptr[blockIdx.x + 1];
    const int n_chunks = c_end - c_begin;
    const int s_begin = cta_stage_ptr[blockIdx.x];
    const int total_stages = cta_stage_ptr[blockIdx.x + 1] - s_begin;

    // ---- one-time setup --------------------------------------------------------------------
    // TMA destinations need 128 B, UMMA descriptors 16 B, SWIZZLE_128B atoms 1024 B
    if ((smem_u32(smem_raw) & (kDirect ? 1023u : 127u)) != 0u) __trap();
    if constexpr (!kDirect) {   // zero the operand ring: padded feature rows and the spare rows stay zero for ever
        uint4* p = reinterpret_cast<uint4*>(sm.op_stage(0));
        for (int i = tid; i < S2 * OP_STAGE_BYTES / 16; i += NUM_THREADS) p[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int s = 0; s < NBAR; ++s) { mbar_init(&sm.full_f32[s], 1); mbar_init(&sm.full_op[s], 1); mbar_init(&sm.empty_op[s], 1); }
        for (int b = 0; b < 2; ++b) {
            for (int g = 0; g < MAX_WG; ++g) mbar_init(&sm.acc_full[g][b], 1);
            mbar_init(&sm.acc_empty[b], 4);
        }
        fence_mbar_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&sm.tmem_base, TMEM_COLS);
    fence_proxy_async();      // the zero fill above must be visible to the tensor-core (async) proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    // register budget per warpgroup (Cfg: the increases below block until the decreases freed enough)
    if (warp < C::kFirstEpiWarp) {
        if (warp < 4) reg_dec<C::kRegsProd>(); else reg_dec<C::kRegsStage>();
    }
    if (warp == MMA_WARP) {
        if (n_chunks > 0) {
            // ================================ MMA issuer ========================================
            // The whole warp runs the loop (uniform control flow: waits, flag reads, bookkeeping stay off the
            // divergent path); one elected lane issues the tcgen05 instructions.  The loop is unrolled over
            // the 8 operand slots so every shared-memory descriptor is base + compile-time constant.
            constexpr uint32_t idesc1 = make_idesc(128, kSym ? N1_SYM : N1, kDirect);
            constexpr uint32_t idesc2 = make_idesc(128, N2, kDirect);
            constexpr int STAGE_BYTES = C::kStageBytes;
            // operand rows 128.. (lo' | 0): 16 eight-row groups further (K-major), or two 64-element chunks further (direct)
            constexpr uint32_t LO_OFF16 = (uint32_t)((kDirect ? 2 * D_CHUNK_STRIDE : (LO_ROW / 8) * OP_GROUP_BYTES) >> 4);
            const uint32_t op_base0 = kDirect ? smem_u32(sm.dstage(0)) : smem_u32(sm.op_stage(0));
            const uint64_t dbase = make_smem_desc(op_base0, desc_tmpl);     // descriptor of slot 0, row 0
            // this warp's own copy of the TMEM base: the kernel-wide value is spilled around the role branches, and a local
            // memory reload inside the elected region sits on the critical path of every pass
            const uint32_t tmem_base = *reinterpret_cast<const volatile uint32_t*>(&sm.tmem_base);
            const uint32_t empty_bar0 = smem_u32(&sm.empty_op[0]);
            const uint32_t acc_full_bar0 = smem_u32(&sm.acc_full[0][0]);    // [wg][buf], 8 bytes each
            // one k-group: D[0:128, 0:240] (+)= [hi | r]^T [hi | r | 0 | lo']   (+ D[:, 128:256] += lo'^T [hi | r] if !kSym)
            // Everything the step needs beyond the slot number comes from the stage's flag word (see FLAG_*): no
            // run-time tile / chunk counters in this warp, whose instruction stream bounds the k-step rate on long rows.
            [[maybe_unused]] auto issue_step = [&](int slot, uint32_t m) {      // fp32 staging: one k-group per stage
                // start-address field is (byte address >> 4); slots and row groups are 16-byte multiples and the
                // whole ring lies below the field's 256 KB wrap, so plain addition is exact
                const uint64_t d_hi = dbase + (uint64_t)((slot * STAGE_BYTES) >> 4);       // rows 0.. : hi | r | 0 | lo'
                const uint32_t d_tmem = tmem_base + ((m >> FLAG_BUF_SHIFT) & 1u) * (uint32_t)ACC_COLS;
                umma_f16(d_tmem, d_hi, d_hi, idesc1, (m & FLAG_SUB_FIRST) ? 0u : 1u);
                if (!kSym) umma_f16(d_tmem + SCOL, d_hi + (uint64_t)LO_OFF16, d_hi, idesc2, 1u);      // rows 128.. : lo' | 0
                umma_commit_addr(empty_bar0 + (uint32_t)slot * 8u);          // operand stage reusable once the MMAs retire
                if (m & FLAG_SUB_LAST) umma_commit_addr(acc_full_bar0 + ((m >> FLAG_BUF_SHIFT) & 7u) * 8u);   // acc_full[wg][buf]: index 2 wg + buf
            };
            [[maybe_unused]] auto wait_tile_free = [&](uint32_t m) {
                if (m & FLAG_SUB_FIRST) mbar_wait(&sm.acc_empty[(m >> FLAG_BUF_SHIFT) & 1u], (m >> FLAG_EMPTY_PARITY_SHIFT) & 1u);
            };
            if constexpr (kDirect) {
                // Tile by tile: one table word per TMEM tile (its first stage's info) tells how many stages it has, which
                // buffer it accumulates in, which warpgroup drains it and the acc_empty parity to wait for; inside a tile the
                // per-stage work is a barrier wait and the MMAs, with descriptors that depend on the ring position only.
                const uint32_t nslot = (uint32_t)C::kSlots;
                uint32_t slot = 0, ph = 0;
                int S = s_begin;
                const int s_end = s_begin + total_stages;
                uint32_t info_next = stage_tab[S].info;
                while (S < s_end) {
                    const uint32_t info = info_next;
                    const uint32_t stages = (info >> TILE_STAGES_SHIFT) & 31u;
                    const uint32_t m = info >> 8;
                    S += (int)stages;
                    if (S < s_end) info_next = stage_tab[S].info;        // next tile's word: in flight during this tile
                    const uint32_t buf = (m >> FLAG_BUF_SHIFT) & 1u;
                    const uint32_t d_tmem = tmem_base + buf * (uint32_t)ACC_COLS;
                    const uint32_t full_bar = acc_full_bar0 + ((m >> FLAG_BUF_SHIFT) & 7u) * 8u;      // acc_full[wg][buf]
                    const uint32_t last_groups = ((info >> TILE_LAST_GROUPS_SHIFT) & 3u) + 1u;
                    mbar_wait(&sm.acc_empty[buf], (m >> FLAG_EMPTY_PARITY_SHIFT) & 1u);
                    // (issuing two stages per elected region was measured: no gain on long rows, 9 % slower on short rows,
                    // where the first stage's MMAs then wait for the second stage's data)
                    for (uint32_t st = 0; st < stages; ++st) {
                        mbar_wait(&sm.full_op[slot], ph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t d_hi = dbase + (uint64_t)(slot * (uint32_t)(STAGE_BYTES >> 4));
                            // every stage of a tile but the last is full; k-group g = ratings 16 g .. 16 g + 15, 8 KB apart
                            const uint32_t groups = (st + 1u < stages) ? (uint32_t)C::kGroups : last_groups;
#pragma unroll
                            for (int g = 0; g < C::kGroups; ++g) {
                                if ((uint32_t)g < groups) {
                                    const uint64_t d_g = d_hi + (uint64_t)((g * DGROUP_BYTES) >> 4);
                                    umma_f16(d_tmem, d_g, d_g, idesc1, (g == 0 && st == 0u) ? 0u : 1u);
                                    if (!kSym) umma_f16(d_tmem + SCOL, d_g + (uint64_t)LO_OFF16, d_g, idesc2, 1u);
                                }
                            }
                            umma_commit_addr(empty_bar0 + slot * 8u);            // operand stage reusable once the MMAs retire
                            if (st + 1u == stages) umma_commit_addr(full_bar);
                        }
                        __syncwarp();
                        if (++slot == nslot) { slot = 0; ph ^= 1u; }
                    }
                }
            } else {
                // 8 operand slots, the loop is unrolled over them so every shared-memory descriptor is base + constant
                for (int n0 = 0; n0 < total_stages; n0 += S2) {
                    const uint32_t ph = ((uint32_t)n0 / S2) & 1u;
                    // two stages per pass: their barrier waits and flag reads overlap, one elected region issues both
#pragma unroll
                    for (int slot = 0; slot < S2; slot += 2) {
                        if (n0 + slot < total_stages) {
                            const bool two = (n0 + slot + 1 < total_stages);
                            mbar_wait(&sm.full_op[slot], ph);
                            if (two) mbar_wait(&sm.full_op[slot + 1], ph);
                            const uint32_t m0 = sm.meta_op[slot];
                            const uint32_t m1 = two ? sm.meta_op[slot + 1] : 0u;
                            wait_tile_free(m0);
                            wait_tile_free(m1);
                            tc_fence_after();
                            if (elect_one()) {
                                issue_step(slot, m0);
                                if (two) issue_step(slot + 1, m1);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
    } else if (warp >= C::kFirstWorker && warp < C::kFirstWorker + C::kWorkers) {
        if (n_chunks > 0) {
            // ============ autonomous stage workers: warp w owns stages w, w+8, w+16, ... ===========
            // own-stage t (global stage n = w + 8t) lives in fp32 slot w + 8(t&1) and operand slot w.
            // kDirect: one ring; own-stage t lives in slot w + 8(t&1), which is also the MMA operand.
            constexpr int STAGE_WARPS = C::kWorkers;
            const int sw = warp - C::kFirstWorker;
            [[maybe_unused]] unsigned char* obase = sm.op_stage(sw);
            const int own = (total_stages > sw) ? (total_stages - sw + STAGE_WARPS - 1) / STAGE_WARPS : 0;
            auto load_desc = [&](int t) -> StageDesc {
                return (t < own) ? stage_tab[s_begin + sw + STAGE_WARPS * t] : StageDesc{0, 0u};
            };
            // lane k < 16 fetches column index and rating k of a stage (coalesced 64-byte reads)
            [[maybe_unused]] auto load_idx = [&](const StageDesc& d) -> int {
                return (lane < (int)(d.info & 0xffu)) ? __ldg(colidx + d.pos + lane) : 0;
            };
            [[maybe_unused]] auto load_val = [&](const StageDesc& d) -> float {
                return (lane < (int)(d.info & 0xffu)) ? __ldg(val + d.pos + lane) : 0.f;
            };
            if constexpr (kDirect) {
                // Worker w owns ring slots w + W j (j < R): own-stage t (global stage n = w + W t) lives in slot n mod (W R) =
                // w + W (t mod R) and is use number t / R of that slot -- every barrier has one producer and one consumer.
                constexpr int R = C::kSlotsPerWorker;
                // arm the slot's mbarrier and launch the gathers of one stage (warp-collective).  Rows past cnt inside a
                // fetched k-group carry the index of the table's all-zero row; a stage of <= 16 ratings fetches one k-group.
                constexpr int LPR = kRows / 32;       // ratings per lane of a stage: lane handles ratings lane, lane + 32
                struct Rat { int idx[LPR]; float val[LPR]; };
                auto load_rat = [&](const StageDesc& d) -> Rat {
                    Rat r;
                    const int cnt = (int)(d.info & 0xffu);
#pragma unroll
                    for (int e = 0; e < LPR; ++e) {
                        const int k = lane + 32 * e;
                        r.idx[e] = (k < cnt) ? __ldg(colidx + d.pos + k) : zero_row;
                        r.val[e] = (k < cnt) ? __ldg(val + d.pos + k) : 0.f;
                    }
                    return r;
                };
                auto issue_direct = [&](int slot, uint32_t flags, const Rat& rat) {
                    const uint32_t groups = ((flags >> FLAG_GROUPS_SHIFT) & 3u) + 1u;
                    unsigned char* sbase = sm.dstage(slot);
#pragma unroll
                    for (int e = 0; e < LPR; ++e) {
                        sm.stage_vals[slot][lane + 32 * e] = rat.val[e];
                        sm.stage_idx[slot][lane + 32 * e] = rat.idx[e];
                    }
                    if (lane == 0) sm.meta_op[slot] = flags;      // read back by this warp when the stage has landed
                    __syncwarp();
                    // One lane issues all gathers from an unrolled loop: the four row coordinates of a quad are loaded once
                    // (LDS.128) and reused by its four 64-element chunks, so successive UTMALDG differ only in destination and
                    // column.  (Issuing from 32 divergent lanes costs an ELECT + 6 R2UR.BROADCAST round per instruction: ~60
                    // cycles each, half of a worker's time in the ncu source view.)  elect.sync, not `lane == 0`: ptxas then
                    // keeps the operands in uniform registers.
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&sm.full_f32[slot], groups * (uint32_t)DGROUP_BYTES);
#pragma unroll
                        for (int g = 0; g < C::kGroups; ++g) {
                            if ((uint32_t)g < groups) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const int4 ix = *reinterpret_cast<const int4*>(&sm.stage_idx[slot][g * KT + q * GROUP_ROWS]);
#pragma unroll
                                    for (int c = 0; c < 4; ++c)
                                        tma_gather4_col(sbase + g * DGROUP_BYTES + (q >> 1) * D_KG_STRIDE + c * D_CHUNK_STRIDE + (q & 1) * 512,
                                                        &factor_map, c * SPLIT_CHUNK, ix.x, ix.y, ix.z, ix.w, &sm.full_f32[slot]);
                                }
                            }
                        }
                    }
                    __syncwarp();
                };
                // prologue: own-stages 0 .. R-1 into their (free) slots
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    if (j < own) {
                        const StageDesc d = load_desc(j);
                        const Rat rat = load_rat(d);
                        issue_direct(sw + STAGE_WARPS * j, d.info >> 8, rat);
                    }
                }
                StageDesc dn = load_desc(R);              // descriptor of the stage the next refill fetches
                int slot_j = 0;                           // t mod R
                uint32_t par = 0;                         // (t / R) & 1
                for (int t = 0; t < own; ++t) {
                    // software prefetch: descriptor of own-stage t+R+1, indices/ratings of t+R (consumed after the waits below)
                    const StageDesc dcur = dn;
                    dn = load_desc(t + R + 1);
                    const Rat nrat = load_rat(dcur);
                    const int slot = sw + STAGE_WARPS * slot_j;
                    unsigned char* sbase = sm.dstage(slot);
                    mbar_wait(&sm.full_f32[slot], par);               // the rows have landed
                    const uint32_t groups = ((sm.meta_op[slot] >> FLAG_GROUPS_SHIFT) & 3u) + 1u;      // this warp's own write at issue time
#pragma unroll
                    for (int e = 0; e < LPR; ++e) {
                        const uint32_t kk = (uint32_t)lane + 32u * e;       // rating (= gathered row) of the stage
                        if ((kk >> 4) < groups) {
                            // the ratings ride along as operand columns 112 (r_hi) and 113 (r_lo') of gathered row kk: chunk 1,
                            // element 48 -> 16-byte piece 6 of the row's 128-byte line, XOR-swizzled with the line number
                            const float r0 = sm.stage_vals[slot][kk];
                            const float h0 = __uint_as_float(__float_as_uint(r0) & 0xFFFFE000u);
                            const uint32_t k = kk & 15u;
                            unsigned char* ob = sbase + (kk >> 4) * DGROUP_BYTES + (k >> 3) * D_KG_STRIDE + D_CHUNK_STRIDE +
                                                (k & 7u) * 128u + ((6u ^ (k & 7u)) << 4);
                            *reinterpret_cast<__half2*>(ob) = __floats2half2_rn(h0, (r0 - h0) * kLoScale);
                        }
                    }
                    fence_proxy_async();                  // the generic-proxy rating writes ordered before the tensor core's reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.full_op[slot]);
                    if (t + R < own) {
                        // own-stage t+R reuses the slot: the MMAs of stage t must have retired (tcgen05.commit -> empty_op)
                        mbar_wait(&sm.empty_op[slot], par);
                        issue_direct(slot, dcur.info >> 8, nrat);
                    }
                    if (++slot_j == R) { slot_j = 0; par ^= 1u; }
                }
            } else {
            // arm the slot's mbarrier and launch the gathers of one stage (warp-collective)
            auto issue = [&](int fs, const StageDesc& d, int my_idx, float my_val) {
                const int cnt = (int)(d.info & 0xffu);
                if (lane < KT) sm.stage_vals[fs][lane] = my_val;
                const int src = (lane & 3) * GROUP_ROWS;
                const int i0 = __shfl_sync(0xffffffffu, my_idx, src + 0);
                const int i1 = __shfl_sync(0xffffffffu, my_idx, src + 1);
                const int i2 = __shfl_sync(0xffffffffu, my_idx, src + 2);
                const int i3 = __shfl_sync(0xffffffffu, my_idx, src + 3);
                // rows past cnt carry index 0 (load_idx): fetched from row 0 and ignored by the conversion
                if (lane == 0)
                    mbar_arrive_expect_tx(&sm.full_f32[fs], (uint32_t)((cnt + GROUP_ROWS - 1) / GROUP_ROWS) * GROUP_ROWS * ROW_BYTES);
                __syncwarp();
                if (lane < KT / GROUP_ROWS && lane * GROUP_ROWS < cnt)
                    tma_gather4(sm.f32_stage(fs) + lane * GROUP_BYTES, &factor_map, i0, i1, i2, i3, &sm.full_f32[fs]);
            };

            // prologue: descriptors of own-stages 0..3, gathers of 0 and 1 in flight, indices of 2 ready
            StageDesc d0 = load_desc(0), d1 = load_desc(1), d2 = load_desc(2), d3 = load_desc(3);
            {
                const int ia = load_idx(d0); const float va = load_val(d0);
                const int ib = load_idx(d1); const float vb = load_val(d1);
                if (own > 0) issue(sw, d0, ia, va);
                if (own > 1) issue(sw + STAGE_WARPS, d1, ib, vb);
            }
            int idx2 = load_idx(d2);
            float val2 = load_val(d2);

            for (int t = 0; t < own; ++t) {
                // software prefetch: descriptor of t+4, indices/ratings of t+3 (consumed next iteration)
                const StageDesc d4 = load_desc(t + 4);
                const int idx3 = load_idx(d3);
                const float val3 = load_val(d3);

                const int fs = sw + STAGE_WARPS * (t & 1);
                const uint32_t cnt = d0.info & 0xffu;
                const uint32_t flags = d0.info >> 8;
                const unsigned char* fbase = sm.f32_stage(fs);
                mbar_wait(&sm.full_f32[fs], ((uint32_t)t >> 1) & 1u);
                mbar_wait(&sm.empty_op[sw], ((uint32_t)t & 1u) ^ 1u);
                if (cnt < KT) {
                    // ragged last stage of a chunk: rows beyond cnt were not fetched -> zero them once, so the
                    // conversion below has a single unpredicated path
                    float* base = reinterpret_cast<float*>(const_cast<unsigned char*>(fbase));
                    for (int k = (int)cnt; k < KT; ++k)
                        for (int c = lane; c < F; c += 32) base[(k >> 2) * GROUP_FLOATS + (k & 3) * F + c] = 0.f;
                    __syncwarp();
                }
                // features 0..95: three passes, lane = feature within the pass, all 16 gathered rows
#pragma unroll 1
                for (int j = 0; j < 3; ++j) {
                    const int c = lane + 32 * j;
                    const float* src = reinterpret_cast<const float*>(fbase) + c;
                    float v[KT];
#pragma unroll
                    for (int k = 0; k < KT; ++k) v[k] = src[(k >> 2) * GROUP_FLOATS + (k & 3) * F];
                    uint32_t hi2[8], lo2[8];
#pragma unroll
                    for (int k = 0; k < KT; k += 2) {
                        const float h0 = __uint_as_float(__float_as_uint(v[k]) & 0xFFFFE000u);
                        const float h1 = __uint_as_float(__float_as_uint(v[k + 1]) & 0xFFFFE000u);
                        const __half2 hh = __floats2half2_rn(h0, h1);                       // exact
                        const __half2 ll = __floats2half2_rn((v[k] - h0) * kLoScale, (v[k + 1] - h1) * kLoScale);
                        hi2[k >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                        lo2[k >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    unsigned char* ob = obase + (c >> 3) * OP_GROUP_BYTES + (c & 7) * 16;
                    *reinterpret_cast<uint4*>(ob) = make_uint4(hi2[0], hi2[1], hi2[2], hi2[3]);                                   // k 0..7
                    *reinterpret_cast<uint4*>(ob + OP_KCORE_BYTES) = make_uint4(hi2[4], hi2[5], hi2[6], hi2[7]);                  // k 8..15
                    *reinterpret_cast<uint4*>(ob + (LO_ROW / 8) * OP_GROUP_BYTES) = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
                    *reinterpret_cast<uint4*>(ob + (LO_ROW / 8) * OP_GROUP_BYTES + OP_KCORE_BYTES) = make_uint4(lo2[4], lo2[5], lo2[6], lo2[7]);
                }
                {
                    // features 96..99: 4 x 16 elements spread over the 32 lanes (two consecutive rows each)
                    const int c = 96 + (lane & 3), k = (lane >> 2) * 2;
                    const float* src = reinterpret_cast<const float*>(fbase) + c;
                    const float v0 = src[(k >> 2) * GROUP_FLOATS + (k & 3) * F];
                    const float v1 = src[((k + 1) >> 2) * GROUP_FLOATS + ((k + 1) & 3) * F];
                    const float h0 = __uint_as_float(__float_as_uint(v0) & 0xFFFFE000u);
                    const float h1 = __uint_as_float(__float_as_uint(v1) & 0xFFFFE000u);
                    const __half2 hh = __floats2half2_rn(h0, h1);
                    const __half2 ll = __floats2half2_rn((v0 - h0) * kLoScale, (v1 - h1) * kLoScale);
                    unsigned char* ob = obase + (c >> 3) * OP_GROUP_BYTES + (k >> 3) * OP_KCORE_BYTES + (c & 7) * 16 + (k & 7) * 2;
                    *reinterpret_cast<uint32_t*>(ob) = *reinterpret_cast<const uint32_t*>(&hh);
                    *reinterpret_cast<uint32_t*>(ob + (LO_ROW / 8) * OP_GROUP_BYTES) = *reinterpret_cast<const uint32_t*>(&ll);
                }
                if (lane < KT) {
                    // the ratings ride along as operand rows 112 (r_hi) and 113 (r_lo'): b = sum r theta from the
                    // same MMAs.  Lane k converts rating k and drops its two halves into place.
                    const float r0 = sm.stage_vals[fs][lane];
                    const float h0 = __uint_as_float(__float_as_uint(r0) & 0xFFFFE000u);
                    unsigned char* ob = obase + (R_ROW / 8) * OP_GROUP_BYTES + (lane >> 3) * OP_KCORE_BYTES + (lane & 7) * 2;
                    *reinterpret_cast<__half*>(ob) = __float2half_rn(h0);
                    *reinterpret_cast<__half*>(ob + 16) = __float2half_rn((r0 - h0) * kLoScale);
                }
                if (lane == 0) sm.meta_op[sw] = flags;
                fence_proxy_async();                  // generic-proxy accesses of both rings ordered before async-proxy ones
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.full_op[sw]);
                // this warp is the only user of fp32 slot fs: refill it with own-stage t+2 right away
                if (t + 2 < own) issue(fs, d2, idx2, val2);
                d0 = d1; d1 = d2; d2 = d3; d3 = d4;
                idx2 = idx3; val2 = val3;
            }
            }   // !kDirect
        }
    } else if (warp >= C::kFirstEpiWarp) {
        reg_inc<C::kRegsEpi>();
        if (n_chunks > 0) {
            // ========================= epilogue + solver warpgroups =============================
            const int wg = (warp - C::kFirstEpiWarp) >> 2;  // this warpgroup takes chunks with (index % kWG) == wg
            const int quad = warp & 3;                    // TMEM lane quadrant this warp may read (warp id % 4)
            const int i = quad * 32 + lane;               // row of A / unknown owned by this thread
            const bool active = i < F;
            const int bar_id = 1 + wg;
            // tiles appear in chunk-list order: all warpgroups walk the list, each drains only its chunks
            int q = 0;
            uint32_t seen0 = 0, seen1 = 0;                // tiles this warpgroup has taken from TMEM buffer 0 / 1
            uint32_t spb = 0;                             // which copy of sp the next broadcast uses
            double sse_acc = 0.0;                         // sum over this warpgroup's rows of x^T b + x^T r + reg x^T x
            Chunk ck_next = chunks[c_begin];
            for (int c = c_begin; c < c_end; ++c) {
                const Chunk ck = ck_next;
                if (c + 1 < c_end) ck_next = chunks[c + 1];        // hide the descriptor load behind this chunk
                const int tiles = (chunk_steps<kRows>(ck) + C::kSubSteps - 1) / C::kSubSteps;
                if (((c - c_begin) % C::kWG) != wg) { q += tiles; continue; }
                float a[F];
                float bi = 0.f;
                // warm start x_u (cg.cu:47): requested before the tile waits so its latency is hidden behind them
                float* xrow = out + (size_t)ck.row * F;
                float xi = (active && ck.slot < 0) ? xrow[i] : 0.f;
                for (int tile = 0; tile < tiles; ++tile, ++q) {
                    const int buf = q & 1;
                    if (buf == 0) { mbar_wait(&sm.acc_full[wg][0], seen0 & 1u); ++seen0; }
                    else          { mbar_wait(&sm.acc_full[wg][1], seen1 & 1u); ++seen1; }
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * ACC_COLS);
                    // tile = (hi^T hi)/2 + (hi^T lo')/2048 over <= SUB_STEPS k-steps; the tiles of one chunk
                    // are summed here in fp32 (round-to-nearest), which bounds the length of the tensor core's
                    // own (truncating) accumulation chain
                    if (tile == 0) drain_tile<true, kSym>(taddr, a, bi);
                    else drain_tile<false, kSym>(taddr, a, bi);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.acc_empty[buf]);       // accumulator drained
                }

                // ---- [A | b] = G + G^T: rows of G go through shared memory, TR_ROWS at a time.  Thread i adds
                // G[j][i] to its G[i][j]; where row j was symmetrised in an earlier pass (i < base) the buffer
                // already holds the finished A[j][i], which is taken as is (A is symmetric bit for bit).
                if constexpr (kSym) {
                    float* tb = sm.solver_scratch + wg * (TR_ROWS + 1) * F;
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        constexpr int kRowFloat4 = F / 4;
                        const int base = pass * TR_ROWS;
                        if (i >= base && i < base + TR_ROWS) {
                            float4* dst = reinterpret_cast<float4*>(tb + (i - base) * F);
#pragma unroll
                            for (int j = 0; j < kRowFloat4; ++j) dst[j] = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
                        }
                        if (pass == 0 && i == R_ROW) {          // lane 112: r_hi . (hi_j/2 + lo'_j/2048)
                            float4* dst = reinterpret_cast<float4*>(tb + TR_ROWS * F);
#pragma unroll
                            for (int j = 0; j < kRowFloat4; ++j) dst[j] = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
                        }
                        named_bar_sync(bar_id, 128);
                        if (active) {
                            if (pass == 0) bi += tb[TR_ROWS * F + i];
#pragma unroll
                            for (int jj = 0; jj < TR_ROWS; ++jj) {
                                const float v = tb[jj * F + i];
                                a[pass * TR_ROWS + jj] = (i < base) ? v : a[pass * TR_ROWS + jj] + v;
                            }
                        }
                        named_bar_sync(bar_id, 128);             // all reads done before the buffer is rewritten
                    }
                }

                if (ck.slot >= 0) {
                    // chunk of a row split across CTAs: store the partial [A | b]; reduced and solved later
                    if (active) {
                        float4* dst = reinterpret_cast<float4*>(scratchA + (size_t)ck.slot * F * F + (size_t)i * F);
#pragma unroll
                        for (int j = 0; j < F; j += 4) dst[j >> 2] = make_float4(a[j], a[j + 1], a[j + 2], a[j + 3]);
                        scratchB[(size_t)ck.slot * F + i] = bi;
                    }
                    continue;
                }
                // weighted-lambda regularisation (als.cu:546): A + (end-start)*lambda*I, applied inside the
                // mat-vec as  (A p)_i + reg * p_i  (same real-number result, no 100-way register select)
                const float reg = (float)(ck.end - ck.begin) * lambda;

                // ---- CG (cg.cu:47-230), A row i in registers, p broadcast from shared memory ----
                auto spmv = [&](const float* sp, float self) -> float {   // four independent FMA chains, summed pairwise
                    float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
#pragma unroll
                    for (int j = 0; j < F; j += 4) {
                        const float4 pv = *reinterpret_cast<const float4*>(sp + j);
                        y0 = fmaf(a[j], pv.x, y0); y1 = fmaf(a[j + 1], pv.y, y1);
                        y2 = fmaf(a[j + 2], pv.z, y2); y3 = fmaf(a[j + 3], pv.w, y3);
                    }
                    return fmaf(reg, self, (y0 + y1) + (y2 + y3));
                };
                const float own = active ? 1.f : 0.f;
                // sp is double buffered: a copy is rewritten only two broadcasts later, i.e. after at least one
                // warpgroup barrier that every reader of the old contents has passed
                float* sp = sm.sp[wg][spb]; spb ^= 1u;
                if (active) sp[i] = xi;
                named_bar_sync(bar_id, 128);
                float r = active ? bi - spmv(sp, xi) : 0.f;        // r = b - A x
                float p = r;
                float rsold = wg_sum(own * r * r, sm.red[wg][0], quad, lane, bar_id);
                for (int it = 0; (float)it < cg_iter; ++it) {
                    sp = sm.sp[wg][spb]; spb ^= 1u;
                    if (active) sp[i] = p;
                    named_bar_sync(bar_id, 128);
                    const float ap = active ? spmv(sp, p) : 0.f;
                    const float pap = wg_sum(own * p * ap, sm.red[wg][1], quad, lane, bar_id);
                    const float alpha = rsold / pap;               // cg.cu:128 (no guard)
                    xi = fmaf(alpha, p, xi);
                    r = fmaf(-alpha, ap, r);
                    const float rsnew = wg_sum(own * r * r, sm.red[wg][2], quad, lane, bar_id);
                    if (rsnew < kCgErrorF) break;                  // cg.cu:195
                    const float beta = rsnew / rsold;
                    rsold = rsnew;
                    p = fmaf(beta, p, r);
                }
                if (active) {
                    xrow[i] = xi;
                    for (int k = 0; k < peers.n; ++k) peers.p[k][(size_t)ck.row * F + i] = xi;     // every peer replica gets the row
                }
                if (sse_terms != nullptr) {
                    // Squared error of the row's ratings under the x just computed, without touching them again:
                    //   sum_j (r_j - x.theta_j)^2 = sum r_j^2 - 2 x^T b + x^T G x,   G = A - reg I,   A x = b - r
                    //                             = sum r_j^2 - (x^T b + x^T r + reg x^T x)
                    // with r the CG residual carried above.  The caller holds sum r_j^2 (a constant of the data).
                    const float srow = wg_sum(own * xi * (bi + r + reg * xi), sm.red[wg][1], quad, lane, bar_id);
                    if (ck.end > ck.begin) sse_acc += (double)srow;     // empty rows: no ratings, no error (x is NaN there)
                }
            }
            if (sse_terms != nullptr && i == 0) sse_terms[blockIdx.x * MAX_WG + wg] = sse_acc;
        }
    }

    // ---- teardown ----------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------
struct TcWork {
    DevBuf cta_ptr;                     // first chunk of every CTA (+1)
    DevBuf cta_stage_ptr;               // first stage (k-step) of every CTA (+1)
    DevBuf stage_tab;                   // StageDesc per k-step, chunk order
    DevBuf chunk_stage_base, chunk_meta; // inputs of the table fill; kept until destroy (cudaFree would synchronise the
                                        // device, i.e. wait for rating uploads still in flight on another stream)
    int grid = 0;
    int nchunks = 0;
    bool sym = false;                   // long chunks: the tensor core forms half of the cross term, the epilogue transposes
    CUtensorMap factor_map;             // 2-D view [rows][F] fp32 of the opposing factor, box {F, 1}
    const float* mapped_factor = nullptr;
    // direct staging (CUMF_TC_DIRECT=1): fp16 pre-split copy of the opposing factor, refreshed before every launch
    bool direct = false;
    int stage_rows = KT;                // ratings per stage: 16 (fp32 staging), 32 or 64 (direct)
    long long idx_span = 0;             // ratings the plan's chunks cover (largest chunk end)
    const int* scanned_colidx = nullptr;
    int hint_rows = 0;                  // rows of the opposing factor as told by the caller (0: scan the indices)
    int factor_rows = 0;                // largest column id + 1 found in scanned_colidx[0, idx_span)
    DevBuf split_tab;                   // [factor_rows + 1][256] fp16, last row zero
    const void* table_filled = nullptr; // == split_tab.p once a full split pass has written padding and zero row of this allocation
    DevBuf max_idx;
    // with a row-count hint the index scan runs asynchronously (no stream synchronisation on the hot path): its result
    // lands in pinned host memory and is checked by the next launch of this plan
    // generic-f kernel (gram_tc2.cuh): impl == 2
    int impl = 1, f = F;
    Tc2Info info2{};
    DevBuf absmax, scales;              // per-launch scale state of the generic-f kernel
    const float* scaled_val = nullptr;  // rating array whose largest |r| the plan already holds (CUMF_TC_RESCAN_RATINGS=1: never)
    int* h_max_idx = nullptr;           // pinned
    cudaEvent_t max_idx_ready = nullptr;
    bool max_idx_pending = false;
    CUtensorMap split_map;              // 2-D view [factor_rows + 1][256] fp16, box {64, 1}, SWIZZLE_128B
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn tensor_map_encoder() {
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            set_last_error("cuTensorMapEncodeTiled is not available from this driver");
            return nullptr;
        }
        encode = (EncodeFn)fn;
    }
    return encode;
}

// pre-split fp16 table [rows][256]: tile::gather4 fetches four {64, 1} boxes (one 128-byte swizzle line per row);
// coordinates beyond `rows` are out of bounds and read as zero
static int encode_split_map(CUtensorMap* map, const void* d_table, long long rows, int cols) {
    EncodeFn encode = tensor_map_encoder();
    if (!encode) return CUMF_ECUDA;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)SPLIT_CHUNK, 1u};
    const cuuint32_t estride[2] = {1u, 1u};
    const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d_table), gdim, gstride, box, estride,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled (split table) failed with CUresult " + std::to_string((int)rc));
        return CUMF_ECUDA;
    }
    return CUMF_OK;
}

static int encode_factor_map(CUtensorMap* map, const float* d_factor) {
    EncodeFn encode = tensor_map_encoder();
    if (!encode) return CUMF_ECUDA;
    // The row count only bounds the coordinates the hardware accepts; the gathered indices are CSR column ids
    // of rows that exist, so a generous bound is safe.
    const cuuint64_t gdim[2] = {(cuuint64_t)F, (cuuint64_t)1 << 31};
    const cuuint64_t gstride[1] = {(cuuint64_t)ROW_BYTES};
    const cuuint32_t box[2] = {(cuuint32_t)F, 1u};      // tile::gather4 fetches four such boxes (tools/gather4_probe.cu)
    const cuuint32_t estride[2] = {1u, 1u};
    const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(d_factor), gdim, gstride, box, estride,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)rc));
        return CUMF_ECUDA;
    }
    return CUMF_OK;
}

void tc_plan_destroy(TcWork* w, bool cache);

int tc_plan_grid(const TcWork* w) { return w ? w->grid : 0; }
unsigned short* tc_plan_split_table_f100(TcWork* w) {
    return (w && w->impl == 1 && w->direct && w->f == 100) ? w->split_tab.as<unsigned short>() : nullptr;
}

int tc_plan_impl(const TcWork* w) { return w ? w->impl : 0; }
// rows of the opposing factor, when the caller knows them: saves the index scan (and its stream synchronisation) of
// the first direct-staging launch
// (also invalidates the cached scan: the next launch validates the index buffer it is given against `rows`)
static int encode_split_map(CUtensorMap* map, const void* d_table, long long rows, int cols);
// Everything the launches need is allocated HERE, not by the first launch: a device allocation synchronises the device, and in
// a multi-GPU run the first half-step of one rank must not wait for a barrier kernel that is spinning for its peers.
void tc_plan_set_factor_rows(TcWork* w, int rows) {
    if (!w || rows <= 0) return;
    w->hint_rows = rows;
    w->scanned_colidx = nullptr;
    w->scaled_val = nullptr;
    if (!w->direct) return;
    if (rows != w->factor_rows || !w->split_tab.p) {
        const int cols = w->impl == 2 ? w->info2.tab_cols : SPLIT_COLS;
        w->split_tab.release();
        w->table_filled = nullptr;
        if (w->split_tab.alloc((size_t)(rows + 1) * cols * 2) == CUMF_OK &&
            encode_split_map(&w->split_map, w->split_tab.p, (long long)rows + 1, cols) == CUMF_OK)
            w->factor_rows = rows;
        else
            w->split_tab.release();          // the launch path reports the failure
    }
    if (!w->max_idx.p) w->max_idx.alloc(sizeof(int));
    if (!w->h_max_idx) w->h_max_idx = DevBuf::pinned_int();
    if (!w->max_idx_ready) cudaEventCreateWithFlags(&w->max_idx_ready, cudaEventDisableTiming);
    if (w->impl == 2) {
        if (!w->absmax.p) w->absmax.alloc(4 * sizeof(unsigned));
        if (!w->scales.p) w->scales.alloc(4 * sizeof(float));
    }
}
int tc_sse_terms_per_cta() { return MAX_WG; }

bool tc_path_supports(int f) {
    const char* off = getenv("CUMF_DISABLE_TC");
    if (off && *off == '1') return false;
    return f >= 10 && f <= 200 && f % 10 == 0;
}

// which fused kernel serves rank f: gram_tc2.cuh (generic f, one accumulator) unless CUMF_TC_IMPL=1 asks for the round-1
// f = 100 kernel of this file
// f = 100 without CUMF_TC_IMPL: measured per side on the Netflix shape (profiles/README.md, round 2) -- long rows (the "sym"
// variant, X side) are 10 % faster on this file's kernel (one N = 240 MMA per k-group instead of two N = 112 ones keeps the
// single issuing warp ahead of the L2-bound gather), short rows (theta side) 3 % faster on the generic kernel.
static int tc_impl_for(int f, bool long_rows) {
    const char* e = getenv("CUMF_TC_IMPL");
    const char* h = getenv("CUMF_TT_FP16");
    if (h && *h == '1') return 2;                  // the reduced-precision mode lives in the generic kernel
    if (f == F && e && *e == '1') return 1;
    if (f == F && !(e && *e)) return long_rows ? 1 : 2;
    return 2;
}

int tc_plan_create(TcWork** out, const std::vector<Chunk>& chunks, const Chunk* d_chunks, const std::vector<SplitRow>&, int, int f) {
    if (!tc_path_supports(f)) {
        set_last_error("the fused tcgen05 kernels handle f = 10, 20, ..., 200");
        return CUMF_EUNSUPPORTED;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const char* g = getenv("CUMF_TC_CTAS");
    int grid = (g && *g) ? atoi(g) : sms;
    if (grid < 1) grid = 1;
    const int n = (int)chunks.size();
    if (grid > n) grid = std::max(1, n);
    // long chunks -> the symmetric variant (the tensor core forms half of the cross term, the epilogue transposes): pays from
    // ~1024 ratings per chunk on average; measured on Netflix (round 1): X side (5.6 k ratings/chunk) 9.6 -> 8.6 ms, theta side
    // (206 ratings/chunk) 14.8 -> 16.2 ms
    long long total_ratings = 0;
    for (int c = 0; c < n; ++c) total_ratings += chunks[c].end - chunks[c].begin;
    bool sym = false;
    {
        const char* m = getenv("CUMF_TC_SYM");
        sym = (m && *m) ? (*m == '1') : (n > 0 && total_ratings >= 1024LL * n);
    }
    const int impl = tc_impl_for(f, sym);
    if (impl == 2 && f > F) sym = false;              // the generic kernel's symmetric variant exists for f <= 100
    Tc2Info info2{};
    if (impl == 2 && !tc2_plan_info(f, sym, &info2)) {
        set_last_error("no generic-f kernel variant for f = " + std::to_string(f));
        return CUMF_EUNSUPPORTED;
    }
    // staging variant of the f = 100 kernel: "direct" (pre-split fp16 table gathered straight into the UMMA operand) or the
    // fp32 ring + in-kernel conversion (16-rating stages)
    const char* denv = getenv("CUMF_TC_DIRECT");
    const bool direct = impl == 2 || !(denv && *denv == '0');        // default; CUMF_TC_DIRECT=0 selects the fp32 staging ring
    const char* renv = getenv("CUMF_TC_STAGE_ROWS");
    const int kt = impl == 2 ? info2.krows : (direct ? ((renv && atoi(renv) == 32) ? 32 : 64) : KT);      // default 64 (measured: 32 is 3-5 % slower)
    const int sub = 256 / kt;
    // contiguous, cost-balanced partition of the (row-ordered) chunk list: cost = MMA k-steps
    // plus a per-chunk epilogue/solve term, so every CTA streams one contiguous rating range.
    const char* rc_env = getenv("CUMF_TC_ROW_COST");
    const long long per_chunk = (rc_env && *rc_env) ? atoll(rc_env) : 96;
    std::vector<long long> prefix(n + 1, 0);
    for (int c = 0; c < n; ++c) {
        const long long nnz = chunks[c].end - chunks[c].begin;
        prefix[c + 1] = prefix[c] + ((nnz + KT - 1) / KT) * KT + per_chunk;     // MMA k-groups of 16 ratings in both stagings
    }
    std::vector<int> ptr(grid + 1, 0);
    for (int b = 1; b < grid; ++b) {
        const long long target = prefix[n] * b / grid;
        int c = (int)(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
        c = std::min(std::max(c, ptr[b - 1]), n);
        ptr[b] = c;
    }
    ptr[grid] = n;
    // stage table: k-steps per chunk -> prefix -> one descriptor per k-step, filled on the device
    std::vector<int> stage_base(n + 1, 0);
    for (int c = 0; c < n; ++c) {
        const long long nnz = chunks[c].end - chunks[c].begin;
        const long long steps = std::max<long long>(1, (nnz + kt - 1) / kt);
        if (stage_base[c] + steps > 0x7fffffffLL) {
            set_last_error("tc_plan_create: too many k-steps for one shard");
            return CUMF_EINVAL;
        }
        stage_base[c + 1] = stage_base[c] + (int)steps;
    }
    std::vector<int> sptr(grid + 1, 0);
    for (int b = 0; b <= grid; ++b) sptr[b] = stage_base[ptr[b]];

    TcWork* w = new TcWork();
    w->grid = grid;
    w->nchunks = n;
    w->direct = direct;
    w->stage_rows = kt;
    w->impl = impl;
    w->f = f;
    w->info2 = info2;
    w->sym = sym;
    for (int c = 0; c < n; ++c) w->idx_span = std::max<long long>(w->idx_span, chunks[c].end);
    // what the MMA issuer would otherwise count: first tile of every chunk within its CTA, chunk parity within its CTA
    std::vector<int> chunk_meta(std::max(n, 1), 0);
    const int n_wg = impl == 2 ? info2.nsys : ((direct && !w->sym) ? 3 : 2);       // systems in flight of the variant that will run
    static_assert(Cfg<false, true, 32>::kWG == 3 && Cfg<false, true, 64>::kWG == 3 && Cfg<true, true, 32>::kWG == 2 &&
                  Cfg<true, true, 64>::kWG == 2 && Cfg<true, false, 16>::kWG == 2 && Cfg<false, false, 16>::kWG == 2, "solver warpgroups");
    for (int b = 0; b < grid; ++b) {
        long long tile = 0;
        for (int c = ptr[b]; c < ptr[b + 1]; ++c) {
            chunk_meta[c] = (int)(((tile & 0x1fffffffLL) << 2) | ((c - ptr[b]) % n_wg));   // only tile & 3 is consumed
            tile += (stage_base[c + 1] - stage_base[c] + sub - 1) / sub;
        }
    }
    DevBuf& d_base = w->chunk_stage_base;
    int rc = w->cta_ptr.alloc(sizeof(int) * (grid + 1));
    if (rc == CUMF_OK) rc = w->cta_stage_ptr.alloc(sizeof(int) * (grid + 1));
    if (rc == CUMF_OK) rc = w->stage_tab.alloc(sizeof(StageDesc) * (size_t)std::max(1, stage_base[n]));
    if (rc == CUMF_OK) rc = d_base.alloc(sizeof(int) * (n + 1));
    if (rc == CUMF_OK) rc = w->chunk_meta.alloc(sizeof(int) * std::max(n, 1));
    StagingSection sec;      // the device's pinned staging arena, until the fill kernel below has read it
    if (rc == CUMF_OK &&
        (upload_via_kernel(w->cta_ptr.p, ptr.data(), sizeof(int) * (grid + 1), 0) != CUMF_OK ||
         upload_via_kernel(w->chunk_meta.p, chunk_meta.data(), sizeof(int) * std::max(n, 1), 0) != CUMF_OK ||
         upload_via_kernel(w->cta_stage_ptr.p, sptr.data(), sizeof(int) * (grid + 1), 0) != CUMF_OK ||
         upload_via_kernel(d_base.p, stage_base.data(), sizeof(int) * (n + 1), 0) != CUMF_OK)) {
        set_last_error("tc_plan_create: upload failed");
        rc = CUMF_ECUDA;
    }
    if (rc == CUMF_OK && n > 0) {
        if (impl == 2) {
            rc = tc2_fill_stage_table(d_chunks, d_base.as<int>(), w->chunk_meta.as<int>(), n, w->stage_tab.p, info2, 0);
        } else {
            auto fill = kt == 64 ? fill_stage_table_kernel<64, 4> : (kt == 32 ? fill_stage_table_kernel<32, 8> : fill_stage_table_kernel<KT, SUB_STEPS>);
            fill<<<(n + 127) / 128, 128>>>(d_chunks, d_base.as<int>(), w->chunk_meta.as<int>(), n, w->stage_tab.as<StageDesc>());
        }
    }
    {   // (the stream, not the device: rating uploads may be running on another stream)
        const int rc2 = sec.finish(0);
        if (rc == CUMF_OK) rc = rc2;
    }
    if (rc != CUMF_OK) { tc_plan_destroy(w, false); return rc; }
    *out = w;
    return CUMF_OK;
}

void tc_plan_destroy(TcWork* w, bool cache) {
    if (!w) return;
    auto rel = [cache](DevBuf& b) { if (cache) b.release_to_cache(); else b.release(); };
    rel(w->cta_ptr);
    rel(w->cta_stage_ptr);
    rel(w->stage_tab);
    rel(w->chunk_stage_base);
    rel(w->chunk_meta);
    rel(w->split_tab);
    rel(w->max_idx);
    rel(w->absmax);
    rel(w->scales);
    if (w->max_idx_ready) cudaEventDestroy(w->max_idx_ready);
    delete w;
}

int tc_update_factor(TcWork* w, const Chunk* d_chunks, int nchunks, const int* d_colidx, const float* d_val,
                     const float* d_factor, float* d_out, int f, float lambda, float cg_iter, float* d_scratchA,
                     float* d_scratchB, cudaStream_t st, int* launches, double* d_sse_terms, const TcExtra* extra) {
    if (!w || f != w->f) {
        set_last_error("tc_update_factor: bad plan");
        return CUMF_EINVAL;
    }
    if (nchunks == 0) return CUMF_OK;
    if (w->impl != 2 && extra && extra->d_tt) {
        set_last_error("tc_update_factor: direct stores need the generic-f kernel (CUMF_TC_IMPL=2)");
        return CUMF_EUNSUPPORTED;
    }
    // kernel variant: (long rows -> symmetric single-MMA mode) x (staging, stage size)
    using KernelFn = void (*)(const Chunk*, const int*, const StageDesc*, const int*, const int*, const float*, const CUtensorMap, float*,
                              float, float, float*, float*, uint64_t, double*, int, PeerOut);
    PeerOut peers{};
    if (extra) for (int k = 0; k < extra->n_peer_out && k < 8; ++k) peers.p[peers.n++] = extra->peer_out[k];
    struct Variant { KernelFn fn; int threads; size_t smem; };
    auto variant = [&]() -> Variant {
#define CUMF_VARIANT(SYM, DIRECT, ROWS) \
        Variant{als_fused_f100_kernel<SYM, DIRECT, ROWS>, Cfg<SYM, DIRECT, ROWS>::kThreads, sizeof(typename SmemFor<SYM, DIRECT, ROWS>::type)}
        if (!w->direct) return w->sym ? CUMF_VARIANT(true, false, 16) : CUMF_VARIANT(false, false, 16);
        if (w->stage_rows == 64) return w->sym ? CUMF_VARIANT(true, true, 64) : CUMF_VARIANT(false, true, 64);
        return w->sym ? CUMF_VARIANT(true, true, 32) : CUMF_VARIANT(false, true, 32);
#undef CUMF_VARIANT
    };
    const Variant v = variant();
    CUMF_CUDA_TRY(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    CUMF_REQUIRE((reinterpret_cast<uintptr_t>(d_factor) & 15u) == 0, "the factor matrix must be 16-byte aligned for TMA");
    if (w->direct) {
        // a hinted plan's asynchronous index validation (below): an id beyond the table would be gathered as zeros
        if (w->max_idx_pending && cudaEventQuery(w->max_idx_ready) == cudaSuccess) {
            w->max_idx_pending = false;
            if (*w->h_max_idx >= w->factor_rows) {
                set_last_error("column id " + std::to_string(*w->h_max_idx) + " exceeds the opposing factor's " +
                               std::to_string(w->factor_rows) + " rows (cumf_plan_set_factor_rows)");
                w->scanned_colidx = nullptr;
                return CUMF_EINVAL;
            }
        }
        // rows of the opposing factor the plan can gather = largest column id + 1 (one scan per plan and index array)
        if (w->scanned_colidx != d_colidx) {
            int rows = w->hint_rows;
            if (rows > 0) {
                // trust the hint for this launch, validate it behind the launch without synchronising
                if (!w->max_idx.p) CUMF_TRY(w->max_idx.alloc(sizeof(int)));
                if (!w->h_max_idx) w->h_max_idx = DevBuf::pinned_int();
                CUMF_REQUIRE(w->h_max_idx != nullptr, "no pinned host memory for the index validation");
                if (!w->max_idx_ready) CUMF_CUDA_TRY(cudaEventCreateWithFlags(&w->max_idx_ready, cudaEventDisableTiming));
                CUMF_CUDA_TRY(cudaMemsetAsync(w->max_idx.p, 0, sizeof(int), st));
                max_index_kernel<<<592, 256, 0, st>>>(d_colidx, w->idx_span, w->max_idx.as<int>());
                CUMF_CUDA_TRY(cudaMemcpyAsync(w->h_max_idx, w->max_idx.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                CUMF_CUDA_TRY(cudaEventRecord(w->max_idx_ready, st));
                w->max_idx_pending = true;
                *launches += 1;
            }
            if (rows <= 0) {
                if (!w->max_idx.p) CUMF_TRY(w->max_idx.alloc(sizeof(int)));
                CUMF_CUDA_TRY(cudaMemsetAsync(w->max_idx.p, 0, sizeof(int), st));
                max_index_kernel<<<592, 256, 0, st>>>(d_colidx, w->idx_span, w->max_idx.as<int>());
                int h_max = 0;
                CUMF_CUDA_TRY(cudaMemcpyAsync(&h_max, w->max_idx.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                CUMF_CUDA_TRY(cudaStreamSynchronize(st));
                *launches += 1;
                rows = h_max + 1;
            }
            if (rows != w->factor_rows || !w->split_tab.p) {
                const int cols = w->impl == 2 ? w->info2.tab_cols : SPLIT_COLS;
                w->split_tab.release();
                w->table_filled = nullptr;
                CUMF_TRY(w->split_tab.alloc((size_t)(rows + 1) * cols * 2));
                CUMF_TRY(encode_split_map(&w->split_map, w->split_tab.p, (long long)rows + 1, cols));
                w->factor_rows = rows;
            }
            w->scanned_colidx = d_colidx;
        }
        if (w->impl == 2) {
            if (!w->absmax.p) CUMF_TRY(w->absmax.alloc(4 * sizeof(unsigned)));
            if (!w->scales.p) CUMF_TRY(w->scales.alloc(4 * sizeof(float)));
            Tc2Launch a;
            a.f = f; a.sym = w->sym; a.grid = w->grid;
            a.d_chunks = d_chunks; a.d_chunk_meta = w->chunk_meta.as<int>(); a.d_cta_ptr = w->cta_ptr.as<int>();
            a.d_stage_tab = w->stage_tab.p; a.d_cta_stage_ptr = w->cta_stage_ptr.as<int>();
            a.d_colidx = d_colidx; a.d_val = d_val; a.val_span = w->idx_span;
            {
                const char* rs = getenv("CUMF_TC_RESCAN_RATINGS");
                a.scan_ratings = (w->scaled_val != d_val) || (rs && *rs == '1');
                w->scaled_val = d_val;
            }
            a.d_factor = d_factor; a.factor_rows = w->factor_rows;
            a.d_table = w->split_tab.p; a.tensor_map = &w->split_map;
            a.d_absmax = w->absmax.as<unsigned>(); a.d_scales = w->scales.as<float>();
            {   // the reference's CUMF_TT_FP16 / CUMF_XX_FP16 build flags (als.cu:30-31, 335-441: fp16 storage of the gathered
                // factor) as a run-time switch: only the 11-bit hi halves are gathered and multiplied (half the bytes, a third
                // of the MMAs; A is then accurate to ~5e-4 relative -- its own tolerance, tests/test_gpu_generic_f.py)
                const char* h = getenv("CUMF_TT_FP16");
                a.hi_only = h && *h == '1';
            }
            a.d_out = d_out; a.lambda = lambda; a.cg_iter = cg_iter;
            a.d_scratchA = d_scratchA; a.d_scratchB = d_scratchB; a.d_sse_terms = d_sse_terms;
            if (extra) {
                a.d_tt = extra->d_tt; a.d_rhs = extra->d_rhs; a.tt_row_base = extra->tt_row_base;
                a.peer_out = extra->peer_out; a.n_peer_out = extra->n_peer_out;
                a.split_out = extra->split_out;
            }
            return tc2_update(a, st, launches);
        }
        // the split pass over the whole opposing factor -- unless the half-step that produced that factor already wrote the
        // split form of every row into this table (TcExtra::table_current; the table's padding and zero row come from the
        // first full pass)
        if (!(extra && extra->table_current && w->table_filled == w->split_tab.p)) {
            const size_t pieces = (size_t)(w->factor_rows + 1) * 32;
            split_factor_kernel<<<(unsigned)((pieces + 255) / 256), 256, 0, st>>>(d_factor, w->factor_rows, w->split_tab.as<uint4>());
            CUMF_CUDA_TRY(cudaGetLastError());
            *launches += 1;
            w->table_filled = w->split_tab.p;
        }
        const uint64_t desc_tmpl = smem_desc_template_direct(D_CHUNK_STRIDE, D_KG_STRIDE);
        v.fn<<<w->grid, v.threads, v.smem, st>>>(d_chunks, w->cta_ptr.as<int>(), w->stage_tab.as<StageDesc>(), w->cta_stage_ptr.as<int>(),
                                               d_colidx, d_val, w->split_map, d_out, lambda, cg_iter, d_scratchA, d_scratchB,
                                               desc_tmpl, d_sse_terms, w->factor_rows, peers);
    } else {
        if (w->mapped_factor != d_factor) {
            CUMF_TRY(encode_factor_map(&w->factor_map, d_factor));
            w->mapped_factor = d_factor;
        }
        const char* swap = getenv("CUMF_TC_SWAP_LBO_SBO");   // bring-up knob: swap the two descriptor strides
        const uint64_t desc_tmpl = smem_desc_template(swap && *swap == '1');
        v.fn<<<w->grid, v.threads, v.smem, st>>>(d_chunks, w->cta_ptr.as<int>(), w->stage_tab.as<StageDesc>(), w->cta_stage_ptr.as<int>(),
                                               d_colidx, d_val, w->factor_map, d_out, lambda, cg_iter, d_scratchA, d_scratchB,
                                               desc_tmpl, d_sse_terms, 0, peers);
    }
    CUMF_CUDA_TRY(cudaGetLastError());
    *launches += 1;
    return CUMF_OK;
}

}  // namespace cumf
