// gram_tc2.cuh -- the generic-f fused half-step kernel (round 2): f = 10 ... 200 on tcgen05.
//
// Replaces get_hermitianT10 (als.cu:575-659, dispatch als.cu:788, 906) + the cuSPARSE RHS pass (als.cu:750-757) +
// updateXWithCGKernel (cg.cu:36-231) for every rank the reference accepts (f % 10 == 0, main.cpp:33-36), and is a second
// implementation of the f = 100 kernel of gram_tc.cu.  Same pipeline -- TMA tile::gather4 of pre-split fp16 factor rows
// straight into the UMMA MN-major SWIZZLE_128B operand layout -> tcgen05.mma kind::f16 into TMEM -> tcgen05.ld ->
// in-register CG -- with three changes:
//
// 1. ONE accumulator for the split product.  gram_tc.cu keeps P = hi^T hi and S = hi^T lo' + lo'^T hi in separate TMEM
//    columns because lo' carries a 2^11 scale (fp16 has a 5-bit exponent).  Here the whole factor is scaled by one power
//    of two, 2^c with max|v| * 2^c in [2^14, 2^15) (absmax_kernel + split_factor2_kernel, once per half-step), so that
//    hi * 2^c AND lo * 2^c are both normal fp16 numbers for every |v| >= max|v| * 2^-17 (below that lo degrades gradually
//    to an absolute error of max|v| * 2^-39), and
//        D  =  (hi^T hi + hi^T lo + lo^T hi) * 2^2c        accumulates in the SAME columns (fp32 in TMEM).
//    The ratings ride along as feature number f (their own scale 2^cr), so column f of D is b.  Per k-group of 16 ratings:
//    three MMAs of N = NB = ceil16(f + 1) columns instead of N = 256 + 128; a TMEM tile is NB columns instead of 256 (four
//    buffers instead of two at f <= 120); the drain is one tcgen05.ld per 16 columns with no arithmetic; the table row
//    shrinks from 512 to 4 NB bytes (448 B at f = 100).  The power-of-two unscale is folded into the CG mat-vec (exact).
//    Long rows ("sym" variant): the table stores 2 lo, two MMAs form G = hi^T hi + hi^T (2 lo) and the epilogue computes
//    [A|b] = (G + G^T) / 2 through shared memory (same idea as gram_tc.cu's kSym).
// 2. f > 127: the accumulator has two 128-lane row blocks (features 0..127 and 128..f, 2 x NB columns, one TMEM buffer), six
//    MMAs per k-group, and TWO warpgroups per system: warpgroup rb drains row block rb, thread i holds row 128 rb + i of
//    [A | b] in registers (up to 201 values, setmaxnreg 224), the CG's block sums and the p broadcast span 256 threads.
// 3. A leaner solver: packed fp32 FMAs (fma.rn.f32x2 -> FFMA2: 50 instead of 100 issue slots per 100-column mat-vec), each
//    solver warpgroup walks only its own chunks (first TMEM tile of every chunk precomputed in chunk_meta), poll-counted
//    watchdog instead of clock reads in the wait loops.
//
// Operand layout in shared memory (per k-group of 16 gathered rows; NC = 2 CR chunks of 64 features):
//    atom(chunk c, rows 8 kg .. 8 kg + 7) at  kg * (NC * 1024) + c * 1024        (1 KB SWIZZLE_128B atoms)
//    chunks [0, CR): hi region (features, rating slot at feature index f), chunks [CR, 2 CR): lo region (same order)
//    descriptors: LBO = 1024 (next chunk), SBO = NC * 1024 (next 8 rows), layout type 2, both operands MN-major
//    B operand = NB features from the start of a region; A operand = 128 features from chunk 2 rb of a region (lanes that
//    fall beyond feature f read neighbouring data and are never looked at).
// Table row in HBM: [hi (f) | 0 .. NB) | lo (f) | 0 .. NB)], 2 NB halfs; the last chunk of the lo region is partly out of
// bounds (zero-filled by TMA, not fetched).
#pragma once
#include "common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <type_traits>

namespace cumf {
namespace tc2 {

constexpr int KT = 16;                    // gathered rows per MMA k-group (fp16 UMMA K)
constexpr int CHUNK = 64;                 // features per 128-byte swizzle line
constexpr int ATOM_BYTES = 1024;          // 8 rows x 128 B
constexpr int TMEM_COLS = 512;
constexpr int MMA_WARP = 3;
constexpr int MAX_SYS = 3;                // systems in flight per CTA
constexpr int MAX_BUF = 4;                // TMEM accumulator buffers
constexpr int NBAR = 16;
constexpr int MAX_ROWS = 64;              // ratings per stage, at most
#ifndef CUMF_TC2_PROD_REGS
// setmaxnreg of warps 0-3 / 4-15 of the 512-thread shape at f > 90 (128 * (prod + 3 * epi) <= 65536; smaller f: 56 / 152).  56 / 152 until round 2's last A/B:
// with 160 registers the solver keeps six or seven LDS.128 of the mat-vec in flight instead of five (theta side 9.2 -> 8.8 ms),
// and the stage workers / the issuer fit 32 without spills
#define CUMF_TC2_PROD_REGS 32
#define CUMF_TC2_EPI_REGS 160
#endif
constexpr int ROP_KG_BYTES = 512;         // rating operand of one k-group: 16 (N) x 16 (K) fp16, 8 x 16-byte core matrices
constexpr int SYM = 1, WIDE = 0;

// cg.cu:31,195: `rsnew < 1e-4` compares in double; for a float rsnew that is rsnew < nextafterf(1e-4f, +inf)
constexpr float kCgErrorF = 1.00000005e-4f;

// stage flags (StageDesc::info = cnt | flags << 8 | tile bits)
constexpr uint32_t FLAG_CHUNK_FIRST = 1u, FLAG_CHUNK_LAST = 2u, FLAG_SUB_FIRST = 4u, FLAG_SUB_LAST = 8u;
constexpr int FLAG_BUF_SHIFT = 4;             // bits 4-5: TMEM buffer of the stage's tile
constexpr int FLAG_SYS_SHIFT = 6;             // bits 6-7: system (solver warpgroup set) that drains it
constexpr int FLAG_EMPTY_PARITY_SHIFT = 8;    // bit 8: parity to wait for on acc_empty[buf] before the tile's first MMA
constexpr int FLAG_GROUPS_SHIFT = 9;          // bits 9-10: MMA k-groups of the stage, minus one
constexpr int TILE_STAGES_SHIFT = 20;         // first stage of a tile: stages in the tile (1 .. 16), bits 20-24
constexpr int TILE_LAST_GROUPS_SHIFT = 25;    // bits 25-26: k-groups of the tile's last stage, minus one

struct StageDesc {
    int pos;             // offset of the stage's first rating (relative to the plan's first rating)
    uint32_t info;
};
static_assert(sizeof(StageDesc) == 8, "StageDesc is loaded as one 8-byte word");

template <int F_> struct Geo {
    static constexpr int F = F_;
    static constexpr int FA = (F + 3) / 4 * 4;                     // coefficient registers per thread (zero padded)
    static constexpr int NB = (F + 1 + 15) / 16 * 16;              // MMA N: features + the rating column, multiple of 16
    static constexpr int CR = (NB + CHUNK - 1) / CHUNK;            // chunks per region (hi / lo)
    static constexpr int NC = 2 * CR;                              // chunks per gathered row in shared memory
    static constexpr int ROW_BYTES = NC * 128;                     // one gathered row in shared memory
    // Table row in HBM (halfs): [hi | lo], each region padded so that every 64-half box the gathers fetch is ONE aligned
    // 128-byte line (2 NB <= 64: both regions share the one line of a 128-byte row).  Round 2 first packed the regions to
    // 2 NB halfs (448-byte rows at f = 100): half of the boxes then straddled two lines, and the gather rate of the short-row
    // launch -- requests, not bytes -- dropped by a third (-DCUMF_TC2_PACKED_TABLE rebuilds that layout for A/B runs).
#ifdef CUMF_TC2_PACKED_TABLE
    static constexpr int LO_COL = NB;
    static constexpr int TAB_COLS = (2 * NB < CHUNK) ? CHUNK : 2 * NB;
#else
    static constexpr int LO_COL = (2 * NB <= CHUNK) ? NB : CR * CHUNK;
    static constexpr int TAB_COLS = (2 * NB <= CHUNK) ? CHUNK : 2 * CR * CHUNK;
#endif
    static constexpr int RB = (F + 1 > 128) ? 2 : 1;               // 128-lane row blocks of the accumulator
    static constexpr int TILE_COLS = RB * NB;
    static constexpr int NBUF = (TMEM_COLS / TILE_COLS >= 4) ? 4 : (TMEM_COLS / TILE_COLS >= 2 ? 2 : 1);
#ifdef CUMF_TC2_KROWS32
    static constexpr int KROWS = 32;                               // experiment: 32-rating stages everywhere (twice the slots)
#else
    static constexpr int KROWS = (RB == 2) ? 32 : 64;              // ratings per stage
#endif
    static constexpr int KGROUPS = KROWS / KT;
    static constexpr int SUB = 256 / KROWS;                        // stages per TMEM tile: chains are cut every 256 ratings
    static constexpr int KG_BYTES = KT * ROW_BYTES;
    static constexpr int STAGE_BYTES = KROWS * ROW_BYTES;
    static constexpr int SBO = NC * ATOM_BYTES;
    static_assert(F % 2 == 0 && F >= 2 && F <= 254, "rank");
    static_assert(RB * 128 >= F + 1 || RB == 2, "lanes");
    static_assert(2 * RB <= CR + 1 || RB == 1, "row block 1 of the hi region must start inside the region");
};

// CTA shapes.  Warp 3 issues the MMAs; stage workers are the other warps below kFirstEpiWarp, in order, the first kWorkers of
// them (worker w owns ring slots w + kWorkers j); solver warpgroups follow.
//   long rows (SYM):       20 warps  | 0-2 idle | 3 issuer | 4-8 five workers x 1 slot | 12-19 two solver warpgroups (2 systems)
//   short rows, f <= 100:  16 warps  | 0-2 three workers x 2 slots | 3 issuer | 4-15 three solver warpgroups (3 systems)
//                          (CUMF_TC2_W6S2: 0-2, 4-6 six workers x 1 slot | 3 issuer | 8-15 two solver warpgroups)
//   f = 110, 120:          12 warps  | 0-2 workers x 2 slots | 3 issuer | 4-11 two solver warpgroups (2 systems, 224 registers)
//   f >= 130 (RB = 2):     12 warps  | 0-2 workers x 2 slots | 3 issuer | 4-11 two warpgroups = ONE system
template <int F, int MODE> struct Cfg {
    using G = Geo<F>;
    static constexpr bool kSym = (MODE == SYM);
    static_assert(!kSym || (G::RB == 1 && F <= 100), "the symmetric variant holds F + transposition temporaries in 152 registers");
#if defined(CUMF_TC2_W6S2) || defined(CUMF_TC2_W6S3)
    static constexpr bool kSixWorkers = !kSym && G::RB == 1 && F <= 100;     // experiment: more gather-issuing warps, one system less
#else
    static constexpr bool kSixWorkers = false;
#endif
    static constexpr int kSysWG = G::RB;                                          // warpgroups per system
#ifdef CUMF_TC2_W6S3
    static constexpr int kWG = kSym ? 2 : (G::RB == 2 ? 2 : (F <= 100 ? 3 : 2));  // experiment: six workers AND three systems (640 threads, 128-register solvers)
#else
    static constexpr int kWG = kSym ? 2 : (G::RB == 2 ? 2 : ((F <= 100 && !kSixWorkers) ? 3 : 2));  // solver warpgroups
#endif
    static constexpr int kSys = kWG / kSysWG;                                     // systems in flight
    static constexpr int kFirstWorker = kSym ? 4 : 0;
    static constexpr int kWorkers = kSym ? 5 : (kSixWorkers ? 6 : 3);
#ifdef CUMF_TC2_KROWS32
    static constexpr int kSlotsPerWorker = kSym ? 2 : (G::RB == 2 ? 2 : 4);
#else
    static constexpr int kSlotsPerWorker = (kSym || kSixWorkers) ? 1 : 2;
#endif
    static constexpr int kSlots = kWorkers * kSlotsPerWorker;
    static constexpr int kFirstEpiWarp = kSym ? 12 : (kSixWorkers ? 8 : 4);
    static constexpr int kThreads = (kFirstEpiWarp + 4 * kWG) * 32;               // 640 / 512 / 384
    static constexpr int kRegsLaunch = (kSym || kThreads == 640) ? 96 : (kThreads == 512 ? 128 : 168);
    static constexpr int kRegsProd = kSym ? 80 : (kThreads == 640 ? 48 : (kThreads == 512 ? (F > 90 ? CUMF_TC2_PROD_REGS : 56) : 56));
    static constexpr int kRegsStage = (!kSym && kThreads == 640) ? 40 : 48;       // warpgroups of workers only (warps 4 .. kFirstEpiWarp)
    static constexpr int kRegsEpi = kSym ? 152 : (kThreads == 640 ? 128 : (kThreads == 512 ? (F > 90 ? CUMF_TC2_EPI_REGS : 152) : 224));
    static constexpr int kRing = kSlots * G::STAGE_BYTES;
    // -DCUMF_TC2_RATING_OPERAND (experiment, off): the ratings of a stage as a SECOND B operand (K-major, no swizzle, N = 16: 512
    // bytes per k-group) that the stage's worker writes when it issues the gathers; two small MMAs per k-group add hi^T r and
    // lo^T r into columns F, F + 1 of the tile, and nobody patches the landed rows.  Correct for every f (tests green), but
    // slower: each extra MMA re-reads the 4 KB A tile from shared memory, and shared-memory bandwidth -- MMA operand fetches
    // (7.5 KB per 128x112x16 MMA = 134 B per tensor-pipe cycle, above the 128 B/cycle the SM delivers) plus the TMA writes --
    // is what bounds this kernel (profiles/README.md, round 2 "what bounds the short-row launch").
#ifdef CUMF_TC2_RATING_OPERAND
    static constexpr bool kRatingOperand = !kSym;
#else
    static constexpr bool kRatingOperand = false;      // measured: theta side 10.3-10.6 ms with the operand, 9.7 ms with the patch
#endif
    static constexpr int kRopBytes = kRatingOperand ? kSlots * G::KGROUPS * ROP_KG_BYTES : 16;
    static constexpr int kTrRows = F / 2;                                         // kSym: rows of G exchanged per pass (2 passes)
    static constexpr int kScratch = kSym ? kSys * (kTrRows + 1) * F : 4;
    static constexpr int kSpN = 128 * G::RB;
    static constexpr int kStageWGs = (kFirstEpiWarp - 4) / 4;                     // warpgroups 1 .. that hold only workers / idle warps
    static_assert(128 * (kRegsProd + kStageWGs * kRegsStage + kWG * kRegsEpi) <= kThreads * kRegsLaunch, "setmaxnreg budgets exceed the CTA register pool");
    static_assert(kThreads * kRegsLaunch <= 65536, "launch registers");
    static_assert(kSlots <= NBAR, "ring barriers");
    // worker index of a warp (-1: not a worker): warps below kFirstEpiWarp except the issuer, counted from kFirstWorker
    __host__ __device__ static constexpr int worker_of(int warp) {
        return (warp == MMA_WARP || warp < kFirstWorker || warp >= kFirstEpiWarp) ? -1
               : ((warp - kFirstWorker - ((warp > MMA_WARP && kFirstWorker <= MMA_WARP) ? 1 : 0)) < kWorkers
                      ? (warp - kFirstWorker - ((warp > MMA_WARP && kFirstWorker <= MMA_WARP) ? 1 : 0)) : -1);
    }
};

template <int F, int MODE> struct __align__(1024) Smem {
    using C = Cfg<F, MODE>;
    unsigned char ring[C::kRing];                 // kSlots stages, TMA destination == UMMA operand
    unsigned char ring_guard[2 * ATOM_BYTES];     // A-operand windows of the last stage may read one atom past the ring
    __align__(128) unsigned char rop[C::kRopBytes];   // rating operands [slot][k-group][512] (kRatingOperand)
    float stage_vals[NBAR][MAX_ROWS];
    __align__(16) int stage_idx[NBAR][MAX_ROWS];
    float solver_scratch[C::kScratch];
    __align__(16) float sp[C::kSys][2][C::kSpN];  // CG direction vector per system, double buffered
    float red[C::kSys][3][8];                     // cross-warp partial sums
    unsigned long long full_tma[NBAR], empty_op[NBAR];
#ifdef CUMF_TC2_EXPLICIT_HANDOFF
    unsigned long long vals_ready[NBAR];          // sanitizer build: see the issuer's wait
#endif
    unsigned long long acc_full[MAX_SYS][MAX_BUF], acc_empty[MAX_BUF];
    uint32_t tmem_base;
    __device__ __forceinline__ unsigned char* dstage(int slot) { return ring + slot * C::G::STAGE_BYTES; }
};

template <int F, int MODE> struct SmemCheck {
    static_assert(sizeof(Smem<F, MODE>) <= 232448, "Smem exceeds the 227 KB a CTA can opt into");
    static constexpr bool ok = true;
};

struct OutPtrs {
    float* p[8];      // replicas of the factor being updated that receive every solved row (p[0] = the local one)
    int n;
};

struct Params {
    const Chunk* chunks;
    const int* chunk_meta;        // (first TMEM tile of the chunk within its CTA) << 2 | system
    const int* cta_chunk_ptr;
    const StageDesc* stage_tab;
    const int* cta_stage_ptr;
    const int* colidx;
    const float* val;
    OutPtrs out;
    float lambda, cg_iter;
    float* scratchA;              // split-row partials [slot][f*f], [slot][f]
    float* scratchB;
    float* tt;                    // store mode (tt != nullptr): unsplit rows are written as A + lambda n I, b instead of solved
    float* rhs;
    int tt_row_base;
    const float* scales;          // [0] 2^-2c (A), [1] 2^-(c+cr) (b), [2] 2^cr (ratings)
    double* sse_terms;
    int zero_row;
    int hi_only;                  // reduced-precision mode: gather and multiply only the fp16 hi halves
    SplitOut split_out;           // F == 100: tables that receive every solved row's split form (common.cuh)
    unsigned long long* prof;     // -DCUMF_TC2_PROFILE builds: [cta][PROF_WORDS] cycle counters per role (tools/theta_probe.py)
};
constexpr int PROF_WORDS = 40;
#ifdef CUMF_TC2_PROFILE
#define CUMF_PROF(stmt) stmt
#else
#define CUMF_PROF(stmt)
#endif

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
// Waits park the warp (try_wait with a suspend-time hint -> NANOSLEEP.SYNCS).  Watchdog: a protocol error must fail the
// launch instead of hanging the GPU -- after 2^18 unsuccessful polls (each parks for up to 100 us: 0.3 .. 26 s; a healthy wait
// is microseconds) the CTA traps.  A poll counter, not a clock read: the wait loops are part of every role's hot path.
#ifndef CUMF_TC2_WAIT_NS
#define CUMF_TC2_WAIT_NS 100000
#endif
constexpr uint32_t kWaitHintNs = CUMF_TC2_WAIT_NS;
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done, polls = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity), "r"(kWaitHintNs) : "memory");
        if (!done && ++polls > (1u << 18)) __trap();
    } while (!done);
}
// the issuer's wait for a landed stage: plain polling (no suspend hint) in -DCUMF_TC2_ISSUER_SPIN builds
__device__ __forceinline__ void mbar_wait_hot(unsigned long long* bar, uint32_t parity) {
#ifdef CUMF_TC2_ISSUER_SPIN
    const uint32_t addr = smem_u32(bar);
    uint32_t done, polls = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && ++polls > (1u << 28)) __trap();
    } while (!done);
#else
    mbar_wait(bar, parity);
#endif
}
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// two fp32 FMAs in one issue slot (sm_100: fma.rn.f32x2 -> FFMA2); each half rounds like fmaf
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
#ifdef CUMF_NO_FFMA2
    d0 = fmaf(a0, b0, d0); d1 = fmaf(a1, b1, d1);
    return;
#endif
    uint64_t a, b, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}

// MN-major SWIZZLE_128B descriptor (cute::UMMA::make_umma_desc<Major::MN>, LayoutType::B128): leading offset = distance
// between 64-feature chunks, stride offset = distance between 8-row k-groups, version 1 at [46,48), layout type 2 at [61,64)
__host__ __device__ constexpr uint64_t desc_template(int lbo_bytes, int sbo_bytes) {
    return ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// c_format F32 (bit 4), a/b F16, both MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// A MN-major (the gathered rows), B K-major (the rating operand)
__host__ __device__ constexpr uint32_t make_idesc_rating(int M, int N) {
    return (1u << 4) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// K-major, no-swizzle descriptor (layout type 0; the layout of round 1's fp32-ring operands, gram_tc.cu smem_desc_template):
// element (n, k) of a 16-deep k-group at (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2; leading offset = distance of
// the two K core matrices (128), stride offset = distance of 8-row groups (256)
__host__ __device__ constexpr uint64_t desc_template_rating() {
    return ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr int rating_operand_offset(int n, int k) {
    return (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
}

// sum over the 4 * kSysWG warps of a system; every thread gets the same value, fixed order (deterministic)
template <int NW>
__device__ __forceinline__ float sys_sum(float v, float* red, int warp_in_sys, int lane, int bar_id) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp_in_sys] = v;
    named_bar_sync(bar_id, NW * 32);
    if constexpr (NW == 4) {
        const float4 q = *reinterpret_cast<const float4*>(red);
        return (q.x + q.y) + (q.z + q.w);
    } else {
        const float4 q0 = *reinterpret_cast<const float4*>(red);
        const float4 q1 = *reinterpret_cast<const float4*>(red + 4);
        return ((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w));
    }
}

template <int KROWS> __device__ __forceinline__ int chunk_steps(const Chunk& ck) { return max(1, (ck.end - ck.begin + KROWS - 1) / KROWS); }

// drain this thread's TMEM lane of one tile: columns [0, F) -> a[], column F -> b (raw accumulator values, scale 2^2c / 2^(c+cr))
// kTwoB: column F + 1 holds the rest of b (rating operand: hi and lo halves of the ratings land in two columns)
template <int F, bool kFirst, bool kTwoB>
__device__ __forceinline__ void drain_tile(uint32_t taddr, float (&a)[Geo<F>::FA], float& b) {
    constexpr int NB = Geo<F>::NB;
    static_assert(!kTwoB || (F % 16 != 15 && F + 1 < NB), "column F + 1 must lie in the 16-column block of column F");
    if constexpr (kFirst) {
        // the first tile of a chunk lands straight in a[] (16-register blocks, no copies); all loads are in flight before the
        // one wait (-DCUMF_TC2_DRAIN_SERIAL: a wait per block, the round-2 form)
        uint32_t v[NB / 16][16];
#pragma unroll
        for (int cc = 0; cc < NB; cc += 16) {
            tmem_ld16(taddr + cc, v[cc / 16]);
#ifdef CUMF_TC2_DRAIN_SERIAL
            tmem_ld_wait();
#endif
        }
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < NB; cc += 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float x = __uint_as_float(v[cc / 16][j]);
                if (cc + j < F) a[cc + j] = x;
                else if (cc + j == F) b = x;
                else if (kTwoB && cc + j == F + 1) b += x;
            }
        }
    } else {
        // later tiles are added through 8-register temporaries, two loads in flight (the 16-register blocks of a[] leave no
        // second aligned block of 16 free in a 152-register budget)
#pragma unroll
        for (int cc = 0; cc < NB; cc += 16) {
            uint32_t v[8], w[8];
            tmem_ld8(taddr + cc, v);
            tmem_ld8(taddr + cc + 8, w);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = __uint_as_float(v[j]), y = __uint_as_float(w[j]);
                if (cc + j < F) a[cc + j] += x;
                else if (cc + j == F || (kTwoB && cc + j == F + 1)) b += x;
                if (cc + 8 + j < F) a[cc + 8 + j] += y;
                else if (cc + 8 + j == F || (kTwoB && cc + 8 + j == F + 1)) b += y;
            }
        }
    }
}

template <int F, int MODE>
__global__ void __launch_bounds__(Cfg<F, MODE>::kThreads, 1)
als_fused2_kernel(const __grid_constant__ CUtensorMap factor_map, const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    using C = Cfg<F, MODE>;
    using G = Geo<F>;
    using SmemT = Smem<F, MODE>;
    SmemT& sm = *reinterpret_cast<SmemT*>(smem_raw);
    constexpr bool kSym = C::kSym;
    constexpr int RB = G::RB, NB = G::NB, CR = G::CR, NC = G::NC, FA = G::FA, NBUF = G::NBUF;
    constexpr int KROWS = G::KROWS, KGROUPS = G::KGROUPS;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int c_begin = P.cta_chunk_ptr[blockIdx.x], c_end = P.cta_chunk_ptr[blockIdx.x + 1];
    const int n_chunks = c_end - c_begin;
    const int s_begin = P.cta_stage_ptr[blockIdx.x];
    const int total_stages = P.cta_stage_ptr[blockIdx.x + 1] - s_begin;

    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();      // SWIZZLE_128B atoms need 1024-byte alignment
    if (tid == 0) {
        for (int s = 0; s < NBAR; ++s) { mbar_init(&sm.full_tma[s], 1); mbar_init(&sm.empty_op[s], 1); }
#ifdef CUMF_TC2_EXPLICIT_HANDOFF
        for (int s = 0; s < NBAR; ++s) mbar_init(&sm.vals_ready[s], 1);
#endif
        for (int b = 0; b < MAX_BUF; ++b) {
            for (int g = 0; g < MAX_SYS; ++g) mbar_init(&sm.acc_full[g][b], 1);
            mbar_init(&sm.acc_empty[b], 4 * RB);
        }
        fence_mbar_init();
    }
    // padding entries of the CG direction vector stay zero for ever (the mat-vec reads FA >= F of them)
    for (int k = tid; k < C::kSys * 2 * C::kSpN; k += C::kThreads) (&sm.sp[0][0][0])[k] = 0.f;
    // rating operands: only rows F % 16 and F % 16 + 1 are ever written, the other fourteen stay zero
    if constexpr (C::kRatingOperand)
        for (int k = tid; k < C::kRopBytes / 16; k += C::kThreads) reinterpret_cast<uint4*>(sm.rop)[k] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == MMA_WARP) tmem_alloc(&sm.tmem_base, TMEM_COLS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < C::kFirstEpiWarp) {
        if (warp < 4) reg_dec<C::kRegsProd>(); else reg_dec<C::kRegsStage>();
    }
    if (warp == MMA_WARP) {
        if (n_chunks > 0) {
            // ================================ MMA issuer ========================================
            constexpr uint32_t idesc = make_idesc(128, NB);
            constexpr uint64_t tmpl = desc_template(ATOM_BYTES, G::SBO);
            const uint64_t dbase = tmpl | (uint64_t)((smem_u32(sm.ring) & 0x3FFFFu) >> 4);
            constexpr uint32_t idesc_r = make_idesc_rating(128, 16);
            const uint64_t rbase = desc_template_rating() | (uint64_t)((smem_u32(sm.rop) & 0x3FFFFu) >> 4);
            const uint32_t tmem_base = *reinterpret_cast<const volatile uint32_t*>(&sm.tmem_base);
            const uint32_t empty_bar0 = smem_u32(&sm.empty_op[0]);
            const uint32_t acc_full_bar0 = smem_u32(&sm.acc_full[0][0]);      // [sys][buf], 8 bytes each
            const float rscale = __ldg(P.scales + 2);
            uint32_t slot = 0, ph = 0;
            int S = s_begin;
            const int s_end = s_begin + total_stages;
            uint32_t info_next = P.stage_tab[S].info;
            CUMF_PROF(long long pf_full = 0; long long pf_acc = 0; long long pf_n = 0; const long long pf_t0 = clock64();)
            while (S < s_end) {
                const uint32_t info = info_next;
                const uint32_t stages = (info >> TILE_STAGES_SHIFT) & 31u;
                const uint32_t m = info >> 8;
                S += (int)stages;
                if (S < s_end) info_next = P.stage_tab[S].info;
                const uint32_t buf = (m >> FLAG_BUF_SHIFT) & 3u;
                const uint32_t sys = (m >> FLAG_SYS_SHIFT) & 3u;
                const uint32_t d_tmem = tmem_base + buf * (uint32_t)G::TILE_COLS;
                const uint32_t full_bar = acc_full_bar0 + (sys * MAX_BUF + buf) * 8u;
                const uint32_t last_groups = ((info >> TILE_LAST_GROUPS_SHIFT) & 3u) + 1u;
                CUMF_PROF(long long pf_a = clock64();)
                mbar_wait(&sm.acc_empty[buf], (m >> FLAG_EMPTY_PARITY_SHIFT) & 1u);
                CUMF_PROF(pf_acc += clock64() - pf_a;)
                for (uint32_t st = 0; st < stages; ++st) {
                    const uint32_t groups = (st + 1u < stages) ? (uint32_t)KGROUPS : last_groups;
                    CUMF_PROF(pf_a = clock64();)
                    mbar_wait_hot(&sm.full_tma[slot], ph);      // the gathered rows have landed (TMA complete_tx)
                    CUMF_PROF(pf_full += clock64() - pf_a; ++pf_n;)
#ifdef CUMF_TC2_EXPLICIT_HANDOFF
                    // compute-sanitizer racecheck build only.  The stage worker's stage_vals[] reach this warp through full_tma[slot]:
                    // the worker's arrive.expect_tx (release) is one of the two things that complete the phase this warp acquired
                    // above -- the other is the TMA's complete_tx, and racecheck does not follow transaction counts, so it reports
                    // the handoff as a hazard.  This redundant thread-to-thread barrier makes the same ordering visible to the tool.
                    mbar_wait(&sm.vals_ready[slot], ph);
#endif
                    if constexpr (!C::kRatingOperand) {
                        // The rating rides along as feature F of gathered row kk: hi part in the hi region, lo part in the lo region.
                        // This warp drops them in itself -- the slot's worker is free to refill other slots, and a landed stage
                        // never waits for a worker that is busy issuing gathers (round 2: 40 % of the workers' time was spent
                        // blocked between these two duties while landed stages sat unprocessed).
                        unsigned char* sbase = sm.dstage((int)slot);
                        constexpr int LPR = KROWS / 32;
#pragma unroll
                        for (int e = 0; e < LPR; ++e) {
                            const uint32_t kk = (uint32_t)lane + 32u * e;
                            if ((kk >> 4) < groups) {
                                const float r0 = sm.stage_vals[slot][kk] * rscale;
                                const float h0 = __uint_as_float(__float_as_uint(r0) & 0xFFFFE000u);
                                const float l0 = kSym ? 2.f * (r0 - h0) : (r0 - h0);
                                const uint32_t k = kk & 15u;
                                constexpr uint32_t cF = F / CHUNK, eF = F % CHUNK;
                                unsigned char* row = sbase + (kk >> 4) * G::KG_BYTES + (k >> 3) * G::SBO + (k & 7u) * 128u +
                                                     (((eF >> 3) ^ (k & 7u)) << 4) + (eF & 7u) * 2u;
                                *reinterpret_cast<__half*>(row + cF * ATOM_BYTES) = __float2half_rn(h0);
                                *reinterpret_cast<__half*>(row + (CR + cF) * ATOM_BYTES) = __float2half_rn(l0);
                            }
                        }
                        fence_proxy_async();                    // generic-proxy writes ordered before the tensor core's reads
                        __syncwarp();
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t d_stage = dbase + (uint64_t)(slot * (uint32_t)(G::STAGE_BYTES >> 4));
#pragma unroll
                        for (int g = 0; g < KGROUPS; ++g) {
                            if ((uint32_t)g < groups) {
                                const uint64_t d_g = d_stage + (uint64_t)((g * G::KG_BYTES) >> 4);
                                const uint64_t b_hi = d_g, b_lo = d_g + (uint64_t)((CR * ATOM_BYTES) >> 4);
#pragma unroll
                                for (int rb = 0; rb < RB; ++rb) {
                                    const uint64_t a_hi = d_g + (uint64_t)((2 * rb * ATOM_BYTES) >> 4);
                                    const uint64_t a_lo = d_g + (uint64_t)(((CR + 2 * rb) * ATOM_BYTES) >> 4);
                                    const uint32_t dt = d_tmem + (uint32_t)(rb * NB);
#ifdef CUMF_TC2_EXP_NOMMA
                                    if (g != 0 || st != 0u) continue;      // experiment: one MMA per tile (defines the accumulator)
#endif
                                    umma_f16(dt, a_hi, b_hi, idesc, (g == 0 && st == 0u) ? 0u : 1u);    // hi^T hi
#ifdef CUMF_TC2_EXP_NOMMA
                                    continue;
#endif
                                    if (!P.hi_only) {
#ifdef CUMF_TC2_MMA_ORDER_B
                                        if (!kSym) umma_f16(dt, a_lo, b_hi, idesc, 1u);                 // experiment: consecutive MMAs share B
                                        umma_f16(dt, a_hi, b_lo, idesc, 1u);
#else
                                        umma_f16(dt, a_hi, b_lo, idesc, 1u);                            // hi^T lo   (kSym: hi^T 2 lo)
                                        if (!kSym) umma_f16(dt, a_lo, b_hi, idesc, 1u);                 // lo^T hi
#endif
                                    }
                                    if constexpr (C::kRatingOperand) {
                                        // columns F, F + 1 += hi^T (r_hi, r_lo) [+ lo^T (r_hi, r_lo)]; the other 14 rows of the
                                        // operand are zero, so the neighbouring Gram columns get + 0
                                        const uint64_t b_r = rbase + (uint64_t)((slot * (uint32_t)(KGROUPS * ROP_KG_BYTES) + g * ROP_KG_BYTES) >> 4);
                                        umma_f16(dt + (uint32_t)(F / 16 * 16), a_hi, b_r, idesc_r, 1u);
                                        if (!P.hi_only) umma_f16(dt + (uint32_t)(F / 16 * 16), a_lo, b_r, idesc_r, 1u);
                                    }
                                }
                            }
                        }
                        umma_commit_addr(empty_bar0 + slot * 8u);
                        if (st + 1u == stages) umma_commit_addr(full_bar);
                    }
                    __syncwarp();
                    if (++slot == (uint32_t)C::kSlots) { slot = 0; ph ^= 1u; }
                }
            }
            CUMF_PROF(if (lane == 0 && P.prof) { unsigned long long* q = P.prof + (size_t)blockIdx.x * PROF_WORDS;
                                                 q[1] = pf_full; q[2] = pf_acc; q[3] = clock64() - pf_t0; q[4] = pf_n; })
        }
    } else if (C::worker_of(warp) >= 0) {
        if (n_chunks > 0) {
            // ============ stage workers: worker w owns stages w, w + W, ... and ring slots w + W j ===========
            constexpr int W = C::kWorkers, R = C::kSlotsPerWorker;
            constexpr int LPR = KROWS / 32;       // ratings per lane of a stage
            const int sw = C::worker_of(warp);
            const int own = (total_stages > sw) ? (total_stages - sw + W - 1) / W : 0;
            auto load_desc = [&](int t) -> StageDesc {
                return (t < own) ? P.stage_tab[s_begin + sw + W * t] : StageDesc{0, 0u};
            };
            const float rscale = __ldg(P.scales + 2);
            struct Rat { int idx[LPR]; float val[LPR]; };
            auto load_rat = [&](const StageDesc& d) -> Rat {
                Rat r;
                const int cnt = (int)(d.info & 0xffu);
#pragma unroll
                for (int e = 0; e < LPR; ++e) {
                    const int k = lane + 32 * e;
                    r.idx[e] = (k < cnt) ? __ldg(P.colidx + d.pos + k) : P.zero_row;
                    r.val[e] = (k < cnt) ? __ldg(P.val + d.pos + k) : 0.f;
                }
                return r;
            };
            auto issue = [&](int slot, uint32_t flags, const Rat& rat) {
                const uint32_t groups = ((flags >> FLAG_GROUPS_SHIFT) & 3u) + 1u;
                unsigned char* sbase = sm.dstage(slot);
#pragma unroll
                for (int e = 0; e < LPR; ++e) {
                    if constexpr (!C::kRatingOperand) sm.stage_vals[slot][lane + 32 * e] = rat.val[e];
                    sm.stage_idx[slot][lane + 32 * e] = rat.idx[e];
                }
                if constexpr (C::kRatingOperand) {
                    // r * 2^cr = hi + lo, both fp16 (hi: the top 11 bits, exact), as rows F % 16 and F % 16 + 1 of the operand
                    unsigned char* rop = sm.rop + slot * (KGROUPS * ROP_KG_BYTES);
#pragma unroll
                    for (int e = 0; e < LPR; ++e) {
                        const int kk = lane + 32 * e;
                        const float r0 = rat.val[e] * rscale;
                        const float h0 = __uint_as_float(__float_as_uint(r0) & 0xFFFFE000u);
                        unsigned char* dst = rop + (kk >> 4) * ROP_KG_BYTES + rating_operand_offset(F % 16, kk & 15);
                        *reinterpret_cast<__half*>(dst) = __float2half_rn(h0);
                        *reinterpret_cast<__half*>(dst + 16) = __float2half_rn(r0 - h0);
                    }
                    fence_proxy_async();                        // generic-proxy writes ordered before the tensor core's reads
                }
                __syncwarp();
                if (elect_one()) {
#ifdef CUMF_TC2_EXPLICIT_HANDOFF
                    mbar_arrive(&sm.vals_ready[slot]);
#endif
#if defined(CUMF_TC2_EXP_NOGATHER)
                    mbar_arrive(&sm.full_tma[slot]);                 // experiment: nothing is gathered
#elif defined(CUMF_TC2_EXP_GATHER1)
                    mbar_arrive_expect_tx(&sm.full_tma[slot], groups * (uint32_t)(G::KG_BYTES / NC));
#else
                    mbar_arrive_expect_tx(&sm.full_tma[slot], groups * (uint32_t)(P.hi_only ? G::KG_BYTES / 2 : G::KG_BYTES));
#endif
#pragma unroll
                    for (int g = 0; g < KGROUPS; ++g) {
#ifdef CUMF_TC2_EXP_NOGATHER
                        if (g >= 0) break;
#endif
                        if ((uint32_t)g < groups) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int4 ix = *reinterpret_cast<const int4*>(&sm.stage_idx[slot][g * KT + q * 4]);
#pragma unroll
                                for (int c = 0; c < NC; ++c)
#ifdef CUMF_TC2_EXP_GATHER1
                                    if (c == 0)
#endif
                                    if (c < CR || !P.hi_only)
                                        tma_gather4_col(sbase + g * G::KG_BYTES + (q >> 1) * G::SBO + c * ATOM_BYTES + (q & 1) * 512, &factor_map,
                                                    (c < CR ? c * CHUNK : G::LO_COL + (c - CR) * CHUNK), ix.x, ix.y, ix.z, ix.w, &sm.full_tma[slot]);
                            }
                        }
                    }
                }
                __syncwarp();
            };
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (j < own) {
                    const StageDesc d = load_desc(j);
                    const Rat rat = load_rat(d);
                    issue(sw + W * j, d.info >> 8, rat);
                }
            }
            // refills: own-stage t goes into slot sw + W (t mod R) once the MMAs of own-stage t - R (use number t / R - 1 of that
            // slot) have retired (tcgen05.commit -> empty_op); descriptor and ratings of the next refill are prefetched
            // (descriptors two refills ahead, ratings one: the rating loads never wait for the descriptor they depend on)
            CUMF_PROF(long long pf_wait = 0; long long pf_issue = 0; const long long pf_t0 = clock64();)
            StageDesc dcur = load_desc(R);
            StageDesc dnext = load_desc(R + 1);
            Rat rat = load_rat(dcur);
            int slot_j = 0;
            uint32_t par = 0;
            for (int t = R; t < own; ++t) {
                const StageDesc dnext2 = load_desc(t + 2);
                const Rat nrat = load_rat(dnext);
                const int slot = sw + W * slot_j;
                CUMF_PROF(const long long pf_a = clock64();)
                mbar_wait(&sm.empty_op[slot], par);
                CUMF_PROF(const long long pf_b = clock64();)
                issue(slot, dcur.info >> 8, rat);
                CUMF_PROF(pf_wait += pf_b - pf_a; pf_issue += clock64() - pf_b;)
                dcur = dnext;
                dnext = dnext2;
                rat = nrat;
                if (++slot_j == R) { slot_j = 0; par ^= 1u; }
            }
            CUMF_PROF(if (lane == 0 && P.prof) { unsigned long long* q = P.prof + (size_t)blockIdx.x * PROF_WORDS + 5 + 3 * sw;
                                                 q[0] = pf_wait; q[1] = pf_issue; q[2] = clock64() - pf_t0; })
        }
    } else if (warp >= C::kFirstEpiWarp) {
        reg_inc<C::kRegsEpi>();
        if (n_chunks > 0) {
            // ========================= drain + solver warpgroups =============================
            constexpr int NW = 4 * RB;                     // warps per system
            const int wgi = (warp - C::kFirstEpiWarp) >> 2;
            const int sys = wgi / RB, rb = wgi % RB;
            const int quad = warp & 3;                     // TMEM lane quadrant this warp may read (warp id % 4)
            const int li = quad * 32 + lane;               // TMEM lane
            const int i = rb * 128 + li;                   // row of [A | b] / unknown owned by this thread
            const int warp_in_sys = rb * 4 + quad;
            const bool active = i < F;
            const int bar_id = 1 + sys;
            const uint32_t tmem_base = sm.tmem_base;
            const float sA = __ldg(P.scales + 0) * (kSym ? 0.5f : 1.f);
            const float sB = __ldg(P.scales + 1) * (kSym ? 0.5f : 1.f);
            const bool store = (P.tt != nullptr);
            uint32_t seen = 0;                             // bit b: phase parity of acc_full[sys][b] this system waits for next
            uint32_t spb = 0;
            double sse_acc = 0.0;
            CUMF_PROF(long long pf_wait = 0; long long pf_drain = 0; long long pf_solve = 0; long long pf_n = 0; long long pf_head = 0;
                      long long pf_pre = 0; long long pf_mv = 0; long long pf_red1 = 0; long long pf_red2 = 0; long long pf_iters = 0;
                      const long long pf_t0 = clock64();)
            // the head of the next chunk is loaded a chunk ahead: its latency hides behind this chunk's solve
            int4 ck_raw = make_int4(0, 0, 0, -1);
            int meta_raw = 0;
            if (c_begin + sys < c_end) {
                ck_raw = __ldg(reinterpret_cast<const int4*>(P.chunks + c_begin + sys));
                meta_raw = __ldg(P.chunk_meta + c_begin + sys);
            }
            for (int c = c_begin + sys; c < c_end; c += C::kSys) {
                CUMF_PROF(const long long pf_h = clock64();)
                Chunk ck;
                ck.row = ck_raw.x; ck.begin = ck_raw.y; ck.end = ck_raw.z; ck.slot = ck_raw.w;
                const int tile0 = meta_raw >> 2;
                if (c + C::kSys < c_end) {
                    ck_raw = __ldg(reinterpret_cast<const int4*>(P.chunks + c + C::kSys));
                    meta_raw = __ldg(P.chunk_meta + c + C::kSys);
                }
                const int tiles = (chunk_steps<KROWS>(ck) + G::SUB - 1) / G::SUB;
                float a[FA];
                float bi = 0.f;
                float xi = (active && ck.slot < 0 && !store) ? P.out.p[0][(size_t)ck.row * F + i] : 0.f;    // warm start (cg.cu:47)
#pragma unroll
                for (int j = F; j < FA; ++j) a[j] = 0.f;
                // one accumulation chain per 256 ratings: the tiles of a chunk are summed here in fp32 (round to nearest), which
                // bounds the length of the tensor core's own (truncating) chain.  The first tile lands straight in a[].
                auto take_tile = [&](int tile, auto first) {
                    const uint32_t buf = (uint32_t)(tile0 + tile) & (uint32_t)(NBUF - 1);
                    CUMF_PROF(const long long pf_a = clock64();)
                    mbar_wait(&sm.acc_full[sys][buf], (seen >> buf) & 1u);
                    CUMF_PROF(const long long pf_b = clock64(); pf_wait += pf_b - pf_a;)
                    seen ^= 1u << buf;
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)G::TILE_COLS + (uint32_t)(rb * NB);
                    drain_tile<F, decltype(first)::value, C::kRatingOperand>(taddr, a, bi);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.acc_empty[buf]);
                    CUMF_PROF(pf_drain += clock64() - pf_b;)
                };
                CUMF_PROF(pf_head += clock64() - pf_h;)
                take_tile(0, std::true_type{});
#pragma unroll 1
                for (int tile = 1; tile < tiles; ++tile) take_tile(tile, std::false_type{});
                CUMF_PROF(const long long pf_s = clock64(); ++pf_n;)
                if constexpr (kSym) {
                    // [A | b] = (G + G^T) / 2 (the 1/2 lives in sA / sB): rows of G go through shared memory, kTrRows at a time.
                    // Thread i adds G[j][i] to its G[i][j]; where row j was symmetrised in an earlier pass (i < base) the buffer
                    // already holds the finished sum, which is taken as is (so A is symmetric bit for bit).  Lane F carries the
                    // rating row G[F][:].
                    constexpr int TR = C::kTrRows;
                    float* tb = sm.solver_scratch + sys * (TR + 1) * F;
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        const int base = pass * TR;
                        if (i >= base && i < base + TR) {
                            float2* dst = reinterpret_cast<float2*>(tb + (i - base) * F);
#pragma unroll
                            for (int j = 0; j < F / 2; ++j) dst[j] = make_float2(a[2 * j], a[2 * j + 1]);
                        }
                        if (pass == 0 && i == F) {
                            float2* dst = reinterpret_cast<float2*>(tb + TR * F);
#pragma unroll
                            for (int j = 0; j < F / 2; ++j) dst[j] = make_float2(a[2 * j], a[2 * j + 1]);
                        }
                        named_bar_sync(bar_id, 128);
                        if (active) {
                            if (pass == 0) bi += tb[TR * F + i];
#pragma unroll
                            for (int jj = 0; jj < TR; ++jj) {
                                const float v = tb[jj * F + i];
                                a[pass * TR + jj] = (i < base) ? v : a[pass * TR + jj] + v;
                            }
                        }
                        named_bar_sync(bar_id, 128);
                    }
                }
                if (ck.slot >= 0) {
                    // chunk of a row split across CTAs (or a partial-Gram plan): store the partial [A | b], unscaled
                    if (active) {
                        float2* dst = reinterpret_cast<float2*>(P.scratchA + (size_t)ck.slot * F * F + (size_t)i * F);
#pragma unroll
                        for (int j = 0; j < F; j += 2) dst[j >> 1] = make_float2(a[j] * sA, a[j + 1] * sA);
                        P.scratchB[(size_t)ck.slot * F + i] = bi * sB;
                    }
                    continue;
                }
                const float reg = (float)(ck.end - ck.begin) * P.lambda;      // weighted-lambda regularisation (als.cu:546)
                if (store) {
                    if (active) {
                        const size_t o = (size_t)(ck.row - P.tt_row_base);
                        float2* dst = reinterpret_cast<float2*>(P.tt + o * F * F + (size_t)i * F);
#pragma unroll
                        for (int j = 0; j < F; j += 2) {
                            float v0 = a[j] * sA, v1 = a[j + 1] * sA;
                            if (j == i) v0 = fmaf((float)(ck.end - ck.begin), P.lambda, v0);
                            if (j + 1 == i) v1 = fmaf((float)(ck.end - ck.begin), P.lambda, v1);
                            dst[j >> 1] = make_float2(v0, v1);
                        }
                        if (P.rhs) P.rhs[o * F + i] = bi * sB;
                    }
                    continue;
                }
#ifdef CUMF_TC2_EXP_NOSOLVE
                if (P.lambda > -1.f) continue;                       // experiment: drain only
#endif
                bi *= sB;
                // ---- CG (cg.cu:47-230): row i of A in registers (still scaled by 2^2c: the power-of-two unscale sA is applied
                // to the finished dot product, which is exact), p broadcast from shared memory, packed FMAs ----
                auto spmv = [&](const float* sp, float self) -> float {
#if defined(CUMF_TC2_LDS_BATCH)
                    // experiment: CUMF_TC2_LDS_BATCH loads of p in flight before the first FMA needs one
                    constexpr int NQ = FA / 4, B = CUMF_TC2_LDS_BATCH;
                    float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
                    float4 q[B];
#pragma unroll
                    for (int k = 0; k < B; ++k) if (k < NQ) q[k] = *reinterpret_cast<const float4*>(sp + 4 * k);
#pragma unroll
                    for (int k = 0; k < NQ; ++k) {
                        const float4 pv = q[k % B];
                        if (k + B < NQ) q[k % B] = *reinterpret_cast<const float4*>(sp + 4 * (k + B));
                        ffma2(y0, y1, a[4 * k], a[4 * k + 1], pv.x, pv.y);
                        ffma2(y2, y3, a[4 * k + 2], a[4 * k + 3], pv.z, pv.w);
                    }
                    return fmaf(reg, self, sA * ((y0 + y1) + (y2 + y3)));
#elif defined(CUMF_TC2_SPMV4)
                    float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
#pragma unroll
                    for (int j = 0; j < FA; j += 4) {
                        const float4 pv = *reinterpret_cast<const float4*>(sp + j);
                        ffma2(y0, y1, a[j], a[j + 1], pv.x, pv.y);
                        ffma2(y2, y3, a[j + 2], a[j + 3], pv.z, pv.w);
                    }
                    return fmaf(reg, self, sA * ((y0 + y1) + (y2 + y3)));
#else
                    // four independent FFMA2 chains (eight partial sums): the chain, not the issue rate, sets the mat-vec time
                    float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < FA; j += 4) {
                        const float4 pv = *reinterpret_cast<const float4*>(sp + j);
                        const int h = (j >> 2) & 1;
                        ffma2(y[4 * h], y[4 * h + 1], a[j], a[j + 1], pv.x, pv.y);
                        ffma2(y[4 * h + 2], y[4 * h + 3], a[j + 2], a[j + 3], pv.z, pv.w);
                    }
                    return fmaf(reg, self, sA * (((y[0] + y[1]) + (y[2] + y[3])) + ((y[4] + y[5]) + (y[6] + y[7]))));
#endif
                };
                const float own = active ? 1.f : 0.f;
                float* sp = sm.sp[sys][spb]; spb ^= 1u;
                if (active) sp[i] = xi;
                named_bar_sync(bar_id, NW * 32);
                float r = active ? bi - spmv(sp, xi) : 0.f;        // r = b - A x
                float p = r;
                float rsold = sys_sum<NW>(own * r * r, sm.red[sys][0], warp_in_sys, lane, bar_id);
                CUMF_PROF(pf_pre += clock64() - pf_s;)
                for (int it = 0; (float)it < P.cg_iter; ++it) {
                    CUMF_PROF(const long long pf_i0 = clock64(); ++pf_iters;)
                    sp = sm.sp[sys][spb]; spb ^= 1u;
                    if (active) sp[i] = p;
                    named_bar_sync(bar_id, NW * 32);
                    const float ap = active ? spmv(sp, p) : 0.f;
                    CUMF_PROF(const long long pf_i1 = clock64() + (long long)(ap != ap); pf_mv += pf_i1 - pf_i0;)
                    const float pap = sys_sum<NW>(own * p * ap, sm.red[sys][1], warp_in_sys, lane, bar_id);
                    const float alpha = rsold / pap;               // cg.cu:128 (no guard)
                    xi = fmaf(alpha, p, xi);
                    r = fmaf(-alpha, ap, r);
                    CUMF_PROF(const long long pf_i2 = clock64() + (long long)(r != r && alpha == 0.5f); pf_red1 += pf_i2 - pf_i1;)
                    const float rsnew = sys_sum<NW>(own * r * r, sm.red[sys][2], warp_in_sys, lane, bar_id);
                    CUMF_PROF(pf_red2 += clock64() + (long long)(rsnew != rsnew) - pf_i2;)
                    if (rsnew < kCgErrorF) break;                  // cg.cu:195
                    const float beta = rsnew / rsold;
                    rsold = rsnew;
                    p = fmaf(beta, p, r);
                }
                if (active) {
#pragma unroll 1
                    for (int k = 0; k < P.out.n; ++k) P.out.p[k][(size_t)ck.row * F + i] = xi;
                    if constexpr (F == 100) {
                        if (P.split_out.n > 0) split_row_store(P.split_out, (size_t)ck.row, i, xi);
                    }
                }
                if (P.sse_terms != nullptr) {
                    // sum_j (r_j - x.theta_j)^2 = sum r_j^2 - (x^T b + x^T r + reg x^T x) with r the CG residual (see gram_tc.cu)
                    const float srow = sys_sum<NW>(own * xi * (bi + r + reg * xi), sm.red[sys][1], warp_in_sys, lane, bar_id);
                    if (ck.end > ck.begin) sse_acc += (double)srow;
                }
                CUMF_PROF(pf_solve += clock64() - pf_s;)
            }
            CUMF_PROF(if (i == 0 && P.prof) { unsigned long long* q = P.prof + (size_t)blockIdx.x * PROF_WORDS + 14 + 6 * sys;
                                              q[0] = pf_wait; q[1] = pf_drain; q[2] = pf_solve; q[3] = clock64() - pf_t0; q[4] = pf_n; q[5] = pf_head;
                                              if (sys == 0) { unsigned long long* e = P.prof + (size_t)blockIdx.x * PROF_WORDS + 32;
                                                              e[0] = pf_pre; e[1] = pf_mv; e[2] = pf_red1; e[3] = pf_red2; e[4] = pf_iters; } })
            if (P.sse_terms != nullptr && i == 0) P.sse_terms[blockIdx.x * MAX_SYS + sys] = sse_acc;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(sm.tmem_base, TMEM_COLS);
    }
}

// host-side launcher of one instantiation; the three translation units gram_tc2_{a,b,c}.cu hold the instantiations
struct Variant {
    void (*fn)(const CUtensorMap, const Params);
    int threads;
    size_t smem;
    int krows, sub, nbuf, nsys, tab_cols, lo_col;
};
template <int F, int MODE> Variant make_variant() {
    using C = Cfg<F, MODE>;
    static_assert(SmemCheck<F, MODE>::ok, "shared memory");
    return Variant{als_fused2_kernel<F, MODE>, C::kThreads, sizeof(Smem<F, MODE>), Geo<F>::KROWS, Geo<F>::SUB, Geo<F>::NBUF, C::kSys,
                   Geo<F>::TAB_COLS, Geo<F>::LO_COL};
}
bool variant_a(int f, bool sym, Variant* out);   // f = 10 .. 50
bool variant_b(int f, bool sym, Variant* out);   // f = 60 .. 100
bool variant_c(int f, bool sym, Variant* out);   // f = 110 .. 200

}  // namespace tc2
}  // namespace cumf
