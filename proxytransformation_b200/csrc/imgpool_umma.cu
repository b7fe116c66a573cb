// S9 image proxies, pass B (scores -> softmax -> attention-weighted feature sums) on tcgen05 tensor cores, TMA-fed:
// get_img_proxy (:335-342) + AttentionPool2d.forward (:154-177) in single-query form (algebra in imgpool.cu), shipped
// geometry C = 512 channels, 15 x 15 = 225 positions, 8 heads of 32, 16-bit features (bf16).
//
// The feature map of a view is [512 channels][225 tokens] with a 450-byte row pitch: the rows are not 16-byte aligned, which
// TMA boxes and UMMA shared-memory descriptors need.  225 = 1 (mod 8): the rows of one residue CLASS s = channel mod 8
// (channels s + 8 r, r = 0..63) are 3600 bytes apart and start 2 s bytes after a 16-byte boundary.  In the coordinate
// u = token + s every class is a regular, 16-byte aligned matrix X_s[r][u] at byte 3600 r + 448 s + 2 u of the view (columns
// u < s and u >= s + 225 belong to the neighbouring channels).  So one 4-D tensor map (u, r, s, view) lets TMA deliver
// SWIZZLE_128B tiles [64 class rows][64 u] straight from the raw NCHW tensor: no register staging, no realignment.  The price
// is that the token axis of class s is shifted by s:
//   scores  D1_s[u][n]    = sum_r X_s[r][u] w_eff[n][s + 8 r]      one accumulator per class; the two classes of a pair share one
//                           tcgen05.mma M128 N32 K16 (A = both tiles read MN-major, rows 0-63 / 64-127; B = the w_eff rows
//                           n = (hi|lo, head) of both classes, K-major); the score of token t is sum_s D1_s[t + s]: the softmax
//                           warps add the eight classes through a 16 KB shared-memory exchange (publish by row, gather shifted);
//   sums    D2_s[r][n]   += sum_u X_s[r][u] P[u - s][n]            (tcgen05.mma M128 N32 K16, A = both tiles read K-major,
//                           B = the probabilities as an MN-major, unswizzled operand [token][8 heads x 2 B] per (hi|lo) plane:
//                           a shift of s tokens is a 16 s-byte shift of the descriptor start address, so one copy of the
//                           probabilities (plus a one-row-shifted copy for the odd class of a pair) serves all classes).
// The fp32 operands (w_eff, probabilities) enter as bf16 hi + lo halves in separate N columns, so every product is exact and
// the accumulation is fp32 in TMEM (same numerics as the 3xBF16 GEMMs: ~2^-17 relative).
//
// One persistent CTA per SM; a view is streamed as 4 OVERLAPPING windows of 64 u-columns that advance by 56 (flash-attention
// structure, one query per head).  The score of token t needs u = t .. t + 7 and its weighted sum touches the same columns, so
// window g = [56 g, 56 g + 64) is self-contained for the token SET g = [56 g, 56 g + 55] (last set: up to token 224): scores and
// sums of a set read the SAME tiles and no slot outlives its window (the 8 re-read columns per window are L2 hits).
//   warps 0-3    MMA issuers, one per class pair p (ring slots 4 g + p, accumulators of classes 2p, 2p+1); the warp stays
//                converged and lane 0 issues: scores of window g+1 are issued before the sums of set g; a ring slot goes back to
//                the producer (tcgen05.commit) when the sums that read it have completed.  Warp 0 also fetches the w_eff planes
//   warps 4-11   softmax (a row's work is split over 4 threads: class parity x pair group, 2 heads each): tcgen05.ld of the class
//                score tiles (lane = u), class exchange, + position term, exp
//                relative to a per-view reference maximum (established by window 0, raised FA-style by rescaling the
//                accumulators in TMEM only when a later window exceeds it by more than TAU — never on ordinary data),
//                bf16 hi/lo probability rows -> shared memory (16-byte stores), running sum in registers; end of view: final
//                probabilities -> global
//   warps 12-15  epilogue: mean-token score s0 = w_eff . xbar (CUDA cores, from L2), then per view Y = (D2 f + p0 xbar) / L
//                from TMEM -> bf16 hi/lo planes for the value-side GEMM (same output format as the mma.sync kernel)
//   warp 16      TMA producer of the feature tiles: one 16 KB box (64 u x 64 rows x 2 classes) per ring slot, 10 slots
// Channel order of w_eff columns and of the weighted sums: position 64 s + r  <->  channel s + 8 r (absorbed into the folded
// GEMM weights on the host, pt_img_pool_params variant 1).
#include "common.cuh"
#include "gemm_tc.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

namespace pt {

namespace ipu {
constexpr int C = 512, HW = 225, HEADS = 8;
constexpr int TP = 228;                      // cterm row pitch (floats)
constexpr int YA = 768;                      // output row: 512 weighted sums + 256 probabilities
constexpr int NWIN = 4, WSTEP = 56;          // overlapping u-windows per view: [56 w, 56 w + 64), u = token + class in [0, 232)
constexpr int UCOLS = 232;                   // valid u range of the tensor map: 225 tokens + 7 class shifts
constexpr int TILE_BYTES = 64 * 128;         // [64 class rows][64 u] bf16, SWIZZLE_128B
constexpr int SLOT_BYTES = 2 * TILE_BYTES;   // a class pair: rows 0-63 class 2p, rows 64-127 class 2p+1 (one TMA box)
constexpr int RING = 10;
constexpr int WCLASS_BYTES = 16 * 128;       // w_eff rows (hi|lo, head) x 64 class channels
constexpr int W_BYTES = 8 * WCLASS_BYTES;
// Probabilities of a token set as the MN-major, unswizzled B operand of the sums (N = 32: both classes of a pair in one MMA):
// four planes of [P_ROWS][8 heads] bf16 = 16 bytes per row: hi, lo, and a copy of each stored ONE ROW LATER, so that the same
// start address reads the copies shifted by one more token (class 2p+1 next to class 2p; the four 8-column blocks of the operand
// have to be equidistant).  Row r holds token 56 w - 7 + r (zero if that token is not in the set); the rows above 63 stay zero
// (the shifted 16-row k-steps of a class reach 7 rows further).  Buffer = window index mod PBUF.
constexpr int P_ROWS = 73, P_PLANE = P_ROWS * 16, P_BYTES = 4 * P_PLANE, PBUF = 2;
// class exchange: [class][head pair][64 rows] float2 = the hi + lo score sums of a window, published by the thread that read them
// from TMEM and gathered (shifted by 7 - class rows) by the thread that owns the token
constexpr int XCH_BYTES = 8 * 4 * 64 * 8;
constexpr int OFF_RING = 0;
constexpr int OFF_W = OFF_RING + RING * SLOT_BYTES;
constexpr int OFF_P = OFF_W + 2 * W_BYTES;
constexpr int OFF_XCH = OFF_P + PBUF * P_BYTES;
constexpr int OFF_MISC = OFF_XCH + XCH_BYTES;         // floats: smax[32] sred[32] ered[32] s0[16] stat_l[16] stat_m[16] fcs[8]
constexpr int OFF_BAR = OFF_MISC + 1024;
// mbarriers: full[RING] empty[RING] wfull[2] wempty[2] d1_full[2] p_full[PBUF] p_empty[PBUF] d2_full[2] d2_empty[2] s0_full[2] l_full[2]
constexpr int NBAR = 2 * RING + 6 + 2 * PBUF + 8;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;             // no slack: the dynamic shared memory window itself is 1024-byte aligned (checked)
// warp roles: 0-3 = MMA issuers (one per class pair; warp 0 also TMA of w_eff + TMEM allocation), 4-11 softmax, 12-15 epilogue
// (TMEM lane quarter = warp mod 4), 16 = TMA producer
constexpr int ISSUE_WARPS = 4, SOFTMAX_WARP0 = 4, SOFTMAX_WARPS = 8, EPI_WARP0 = 12, PRODUCER_WARP = 16;
constexpr int THREADS = 32 * (PRODUCER_WARP + 1);
// TMEM: D1 (scores) 2 window buffers x 4 class pairs x 32 columns at 0 ; D2 (sums) 2 view buffers x 4 class pairs x 32 columns at
// 256.  Both are M = 128 accumulators (row = lane): rows 0-63 belong to class 2p and are valid in columns 0-15 (hi | lo heads) of
// the pair's block, rows 64-127 to class 2p+1, valid in columns 16-31; the other two quadrants hold cross products nobody reads.
constexpr int TMEM_COLS = 512;
constexpr int D1_BUF_COLS = 128, D2_COL = 256, D2_BUF_COLS = 128;
constexpr float TAU = 16.0f;                 // the reference maximum is raised when a score exceeds it by more than this
constexpr float TAU_FP16 = 10.0f;            // ... half probability operand: exp(TAU) has to fit a half
constexpr float LOG2E = 1.4426950408889634f;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(OFF_W % 1024 == 0 && SLOT_BYTES % 1024 == 0 && OFF_P % 16 == 0 && OFF_XCH % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
}  // namespace ipu

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t iu_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void iu_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(iu_smem(bar)), "r"(count));
}
__device__ __forceinline__ void iu_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(iu_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void iu_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(iu_smem(bar)) : "memory");
}
__device__ __forceinline__ void iu_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(iu_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void iu_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(iu_smem(dst)),
                 "l"(map), "r"(iu_smem(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void iu_tma_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(iu_smem(dst)),
                 "l"(map), "r"(iu_smem(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void iu_tma_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(iu_smem(dst)),
                 "l"(map), "r"(iu_smem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void iu_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void iu_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void iu_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(iu_smem(bar)) : "memory");
}
__device__ __forceinline__ void iu_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Issued by lane 0 of a CONVERGED warp (the whole warp runs the issue loop, so that addresses and descriptors stay warp-uniform;
// a divergent `if (lane == 0)` region makes the compiler wrap every tcgen05.mma in an elect / R2UR.BROADCAST loop).
__device__ __forceinline__ void iu_mma_l0(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        ".reg .u32 l;\n"
        "mov.u32 l, %%laneid;\n"
        "setp.eq.u32 e, l, 0;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void iu_commit_l0(uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .pred e;\n"
        ".reg .u32 l;\n"
        "mov.u32 l, %%laneid;\n"
        "setp.eq.u32 e, l, 0;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(iu_smem(bar)) : "memory");
}
__device__ __forceinline__ uint64_t iu_mk64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
// Shared-memory operand descriptors (sm_100 version 1, SWIZZLE_128B, 8-row groups 1024 bytes apart).
//   K-major: rows = M/N index, 128-byte rows hold 64 K elements; a K = 16 step advances the start address by 32 bytes.
//   MN-major: rows = K index, 128-byte rows hold 64 M elements; a K = 16 step advances by two 8-row groups (2048 bytes);
//             LBO = distance between 64-element blocks along M (one block here).
__device__ __forceinline__ uint64_t iu_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MN-major operand WITHOUT swizzle (canonical layout ((8,1,m),(8,k)):((1,8,SBO),(8,LBO)) in elements): 16-byte rows of 8
// MN elements, 8 consecutive K rows form a 128-byte core matrix, LBO = distance between K blocks of 8 rows, SBO = distance
// between MN blocks of 8 elements.  The start address only needs 16-byte alignment: a K shift of one row is 16 bytes.
__device__ __forceinline__ uint64_t iu_desc_mn_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16; bit 15 = A is MN-major, bit 16 = B is MN-major; N >> 3 at bit 17,
// M >> 4 at bit 24.
// (A / B formats at bits 7 / 10: 1 = bf16 here; IU_AB_FP16 clears both: features, w_eff planes and probabilities are IEEE half —
// the hardware rejects an f16 A next to a bf16 B ("illegal instruction"))
constexpr uint32_t IU_AB_FP16 = ~((1u << 7) | (1u << 10));
__host__ __device__ constexpr uint32_t iu_idesc(int m, int n, bool a_mn_major, bool b_mn_major = false) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major ? (1u << 15) : 0u) | (b_mn_major ? (1u << 16) : 0u) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void iu_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load without the wait: issue several, then iu_tmem_wait_ld(), then iu_tmem_use16 on each register set (ties the
// registers to a point after the wait so that no consumer is scheduled above it)
__device__ __forceinline__ void iu_tmem_ld16_async(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void iu_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void iu_tmem_use16(uint32_t (&v)[16]) {
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                      "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}
__device__ __forceinline__ void iu_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::"r"(v[0]),
        "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void iu_bar_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void iu_bar_sync256(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ bool iu_bar_or(int id, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.b32 q, %2, 0;\n"
        "bar.red.or.pred p, %1, 256, q;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(r)
        : "r"(id), "r"((uint32_t)pred)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ void iu_split(float x, unsigned short& hi, unsigned short& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
}
__device__ __forceinline__ void iu_split_h(float x, unsigned short& hi, unsigned short& lo) {     // IEEE half hi + lo
    const __half h = __float2half_rn(x);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(x - __half2float(h)));
}
__device__ __forceinline__ float iu_bf(uint32_t packed, int k) {      // element k (0/1) of a packed bf16 pair
    return __uint_as_float(k ? (packed & 0xffff0000u) : (packed << 16));
}

struct UmmaPoolMaps {
    CUtensorMap w;         // w_eff planes as (512 columns, BV*16 rows (view, hi|lo, head)), box (64, 16)
    CUtensorMap x;         // features as (u 232, class row 64, class 8, view BV) with strides (3600, 448, 230400) bytes, box (64, 64, 2, 1)
};

struct UmmaPoolArgs {
    const __nv_bfloat16* wpl;    // (BV, 2, 8, 512) bf16 hi / lo planes of w_eff, columns 64 s + r
    const float* cterm;          // (BV, 8, TP)
    const float* xbar;           // (BV, 512) natural channel order
    __nv_bfloat16* ya_hi;        // (BV, 8, 768) bf16 hi plane (lo plane ya_plane elements later)
    long long ya_plane;
    int BV;
    float scale;
    float* dbg;                  // optional (PT_POOL_DEBUG bit 64): [BV][8][256] scaled scores, attention tokens 0..225
    float wscale_inv;            // 1 / scale of the half w_eff planes (fp16 instantiation: 1/16)
    int debug;                   // PT_UMMA_DEBUG bring-up switches (garbage results): 1 no score MMAs, 2 no sum MMAs, 16 per-role trace (trace
                                 // build), timing probes: 64 no raise vote, 128 no final probability stores, 256 no s0, 512 no Y stores
};

// Per-role cycle trace of CTA 0 (PT_UMMA_DEBUG bit 16): SM clocks spent in each wait / work section, summed over its views;
// read with pt_debug_umma_trace.  Slots: 0 producer wait empty | 1 issuer wait full, 2 wait p_full, 3 wait wfull, 4 wait d2_empty,
// 5 issuer total | 6 softmax wait d1_full, 7 class exchange (tmem loads + shuffles + barrier), 8 wait p_empty, 9 softmax total,
// 10 bar_or / raise | 11 epilogue s0, 12 wait l_full, 13 wait d2_full, 14 epilogue store, 15 epilogue total
#ifdef PT_UMMA_TRACE
#define PT_UMMA_TRACE_ON 1
#else
#define PT_UMMA_TRACE_ON 0      // build with PT_NVCC_DEFINES=-DPT_UMMA_TRACE (tools/umma_trace.py): the timers cost issue slots in every CTA
#endif
__device__ unsigned long long g_umma_trace[24];   // 16..23: softmax sub-sections (tmem load | shuffles + stores | barrier | gather | exp + split | stores + fences + arrive | end of view)
// (accumulated in registers, written once at the end of the role: a global read-modify-write per section would itself cost an
// L2 round trip)
#define UT(acc, stmt) do { const long long ut0_ = tracing ? clock64() : 0; stmt; if (tracing) acc += clock64() - ut0_; } while (0)

// FP16: features, w_eff planes and the probability operand are IEEE half (compile-time: the bf16 instantiation is the round's
// tuned kernel unchanged; run-time switches in the softmax chain cost 30 % of the kernel).
template <bool FP16>
__global__ void __launch_bounds__(ipu::THREADS, 1) img_pool_umma_kernel(const __grid_constant__ UmmaPoolMaps maps, const UmmaPoolArgs a) {
    using namespace ipu;
    extern __shared__ __align__(1024) uint8_t smem[];
    if (threadIdx.x == 0 && (iu_smem(smem) & 1023u) != 0u) __trap();      // SWIZZLE_128B tiles need a 1024-byte aligned base
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty = full + RING;
    uint64_t* wfull = empty + RING;
    uint64_t* wempty = wfull + 2;
    uint64_t* d1_full = wempty + 2;
    uint64_t* p_full = d1_full + 2;
    uint64_t* p_empty = p_full + PBUF;
    uint64_t* d2_full = p_empty + PBUF;
    uint64_t* d2_empty = d2_full + 2;
    uint64_t* s0_full = d2_empty + 2;
    uint64_t* l_full = s0_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    float* smax = reinterpret_cast<float*>(smem + OFF_MISC);     // [4 warps][8]
    float* sred = smax + 32;                                     // [4][8]
    float* ered = sred + 32;                                     // [4][8]
    float* sm_s0 = ered + 32;                                    // [2][8]
    float* stat_l = sm_s0 + 16;                                  // [2][8]
    float* stat_m = stat_l + 16;                                 // [2][8]
    float* fcs = stat_m + 16;                                    // [8]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) { iu_mbar_init(full + i, 1); iu_mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) {
            iu_mbar_init(wfull + i, 1); iu_mbar_init(wempty + i, ISSUE_WARPS); iu_mbar_init(d1_full + i, ISSUE_WARPS);
            iu_mbar_init(d2_full + i, ISSUE_WARPS); iu_mbar_init(d2_empty + i, 4); iu_mbar_init(s0_full + i, 1); iu_mbar_init(l_full + i, 4);
        }
        for (int i = 0; i < PBUF; ++i) { iu_mbar_init(p_full + i, SOFTMAX_WARPS); iu_mbar_init(p_empty + i, ISSUE_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
    }
    // the zero margins of the probability buffers are never written again
    for (int i = threadIdx.x; i < PBUF * P_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem + OFF_P)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(iu_smem(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    iu_fence_before();
    __syncthreads();
    iu_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int nviews = (int)blockIdx.x < a.BV ? (a.BV - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const bool tracing = PT_UMMA_TRACE_ON && (a.debug & 16) && blockIdx.x == 0 && (lane == 0);

    if (warp == PRODUCER_WARP) {
        // ===== TMA producer: window w, class pair p of the view -> ring slot (4 g + p) mod RING =====
        if (lane == 0) {
            unsigned it = 0;
            long long tw0 = 0;
            for (int vi = 0; vi < nviews; ++vi) {
                const int bv = blockIdx.x + vi * gridDim.x;
                for (int w = 0; w < NWIN; ++w)
                    for (int p = 0; p < 4; ++p, ++it) {
                        const unsigned slot = it % RING;
                        UT(tw0, iu_wait(empty + slot, ((it / RING) & 1) ^ 1));
                        iu_expect_tx(full + slot, SLOT_BYTES);
                        iu_tma_4d(smem + OFF_RING + slot * SLOT_BYTES, &maps.x, WSTEP * w, 0, 2 * p, bv, full + slot);
                    }
            }
            if (tracing) g_umma_trace[0] += (unsigned long long)tw0;
        }
    } else if (warp < ISSUE_WARPS) {
        // ===== MMA issuers: warp p owns class pair p (ring slots 4 g + p, D1 / D2 columns of classes 2p, 2p+1); the warp stays
        // converged, lane 0 issues.  Warp 0 also fetches the per-view w_eff planes one view ahead. =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");     // the softmax / epilogue warpgroups take the registers
        const int p = warp;
        constexpr uint32_t AMASK = FP16 ? IU_AB_FP16 : 0xffffffffu;
        constexpr uint32_t IDESC1 = iu_idesc(128, 32, true, false) & AMASK, IDESC2 = iu_idesc(128, 32, false, true) & AMASK;
        // descriptor words (addresses in 16-byte units): SWIZZLE_128B tiles (SBO 1024, version 1, layout 2) and the unswizzled
        // MN-major probability planes (LBO 128 between K blocks, SBO P_PLANE between the hi / lo planes, version 1)
        constexpr uint32_t HI_SW = (1024u >> 4) | (1u << 14) | (2u << 29), HI_P = ((uint32_t)P_PLANE >> 4) | (1u << 14);
        constexpr uint32_t LBO_TILE = ((uint32_t)TILE_BYTES >> 4) << 16, LBO_P = (128u >> 4) << 16;
        const uint32_t ring = iu_smem(smem + OFF_RING) >> 4, wbase = iu_smem(smem + OFF_W) >> 4, pbase = iu_smem(smem + OFF_P) >> 4;
        unsigned g = 0;
        long long tw1 = 0, tw2 = 0, tw3 = 0, tw4 = 0;
        // sums of token set gp (window wp of view vp), issued after the scores of the next window
        auto sums = [&](unsigned gp, unsigned vp, int wp) {
            const unsigned pbi = gp & (PBUF - 1);
            UT(tw2, iu_wait(p_full + pbi, (gp / PBUF) & 1));                        // probabilities of set gp are in shared memory
            if (wp == 0) UT(tw4, iu_wait(d2_empty + (vp & 1), ((vp >> 1) & 1) ^ 1));  // the epilogue has drained this accumulator buffer
            iu_fence_after();
            const unsigned slot = (4 * gp + p) % RING;
            const uint32_t sa = ring + slot * (SLOT_BYTES >> 4), pb = pbase + pbi * (P_BYTES >> 4) + 7 - 2 * p;
            const uint32_t d2 = tmem + D2_COL + (vp & 1) * D2_BUF_COLS + 32 * p;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (!(a.debug & 2))
                    iu_mma_l0(d2, iu_mk64(sa + 2 * j, HI_SW), iu_mk64((pb + 16 * j) | LBO_P, HI_P), IDESC2, (wp > 0 || j > 0) ? 1u : 0u);
            iu_commit_l0(empty + slot);                                              // slot back to the producer once read
            iu_commit_l0(p_empty + pbi);
            if (wp == NWIN - 1) iu_commit_l0(d2_full + (vp & 1));
        };
        auto load_w = [&](int vi) {            // w_eff planes of this CTA's view vi -> buffer vi & 1 (free once the scores of view vi - 2 are done)
            if (vi >= nviews || p != 0) return;
            const int bv = blockIdx.x + vi * gridDim.x, wb = vi & 1;
            iu_wait(wempty + wb, ((vi >> 1) & 1) ^ 1);
            if (lane == 0) {
                iu_expect_tx(wfull + wb, W_BYTES);
#pragma unroll
                for (int s = 0; s < 8; ++s) iu_tma_2d(smem + OFF_W + wb * W_BYTES + s * WCLASS_BYTES, &maps.w, 64 * s, 16 * bv, wfull + wb);
            }
            __syncwarp();
        };
        const long long ti0 = tracing ? clock64() : 0;
        load_w(0);
        for (int vi = 0; vi < nviews; ++vi) {
            const int wb = vi & 1;
            load_w(vi + 1);
            UT(tw3, iu_wait(wfull + wb, (vi >> 1) & 1));
            iu_fence_after();
            const uint32_t sw = wbase + wb * (W_BYTES >> 4) + 2 * p * (WCLASS_BYTES >> 4);
            for (int w = 0; w < NWIN; ++w, ++g) {
                const uint32_t d1 = tmem + (g & 1) * D1_BUF_COLS + 32 * p;
                const unsigned it1 = 4 * g + p, slot = it1 % RING;
                UT(tw1, iu_wait(full + slot, (it1 / RING) & 1));
                iu_fence_after();
                const uint32_t sa = ring + slot * (SLOT_BYTES >> 4);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (!(a.debug & 1)) iu_mma_l0(d1, iu_mk64((sa + 128 * j) | LBO_TILE, HI_SW), iu_mk64(sw + 2 * j, HI_SW), IDESC1, j != 0 ? 1u : 0u);
                iu_commit_l0(d1_full + (g & 1));
                if (w == NWIN - 1) iu_commit_l0(wempty + wb);
                if (g > 0) sums(g - 1, w == 0 ? vi - 1 : vi, w == 0 ? NWIN - 1 : w - 1);
            }
        }
        if (g > 0) sums(g - 1, nviews - 1, NWIN - 1);
        if (tracing && p == 0) {
            g_umma_trace[5] += (unsigned long long)(clock64() - ti0);
            g_umma_trace[1] += (unsigned long long)tw1; g_umma_trace[2] += (unsigned long long)tw2;
            g_umma_trace[3] += (unsigned long long)tw3; g_umma_trace[4] += (unsigned long long)tw4;
        }
    } else if (warp >= SOFTMAX_WARP0 && warp < EPI_WARP0) {
        // ===== softmax: row r of a window is shared by 4 threads (half = class parity, pg = pair group): thread (half, pg, r) reads
        // row r of the class-(2p + half) score tiles of pairs p = 2 pg, 2 pg + 1 (TMEM lane 64 half + r) and keeps heads
        // 4 half + 2 pg, + 1 of token 56 w - 7 + r =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
        const int ws = warp - SOFTMAX_WARP0, q = warp & 3, half = q >> 1, qq = q & 1, pg = ws >> 2;
        const int r = 32 * qq + lane, hb = 4 * half + 2 * pg, src = 2 * half + pg, wpair = 4 * pg + 2 * half;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        float2* xch = reinterpret_cast<float2*>(smem + OFF_XCH);
        float mref[2], lsum[2], pr[NWIN][2];
        unsigned g = 0;
        const bool tr_s = tracing && warp == SOFTMAX_WARP0;
        const long long ts0 = tr_s ? clock64() : 0;
        long long ta6 = 0, ta7 = 0, ta8 = 0, ta10 = 0, tb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define TB(i) do { if (tr_s) { const long long n_ = clock64(); tb[i] += n_ - tbp; tbp = n_; } } while (0)
        long long tbp = 0;
        // token of this thread in window w of a view (-1: not in the set), and its position terms, fetched one VIEW ahead (a load
        // issued one window ahead is still in flight when the window's fence.proxy.async drains the memory pipe)
        auto token_of = [&](int w) { return (r >= 7 && r <= (w == NWIN - 1 ? 63 : 62)) ? WSTEP * w - 7 + r : -1; };
        float ctn[NWIN][2];
        auto load_ct = [&](int vi2) {
            const int bv2 = blockIdx.x + vi2 * gridDim.x;
#pragma unroll
            for (int w2 = 0; w2 < NWIN; ++w2) {
                const int t = token_of(w2);
#pragma unroll
                for (int k = 0; k < 2; ++k) ctn[w2][k] = (t >= 0 && vi2 < nviews) ? __ldg(a.cterm + ((size_t)bv2 * HEADS + hb + k) * TP + 1 + t) : 0.f;
            }
        };
        load_ct(0);
        for (int vi = 0; vi < nviews; ++vi) {
            const int bv = blockIdx.x + vi * gridDim.x;
            float ctv[NWIN][2];
#pragma unroll
            for (int w = 0; w < NWIN; ++w) { ctv[w][0] = ctn[w][0]; ctv[w][1] = ctn[w][1]; }
            load_ct(vi + 1);
#pragma unroll
            for (int k = 0; k < 2; ++k) { lsum[k] = 0.f; mref[k] = 0.f; }
#pragma unroll
            for (int w = 0; w < NWIN; ++w, ++g) {
                const int t = token_of(w);
                const bool valid = t >= 0;
                float ct[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) ct[k] = ctv[w][k];
                { const long long c0_ = tr_s ? clock64() : 0; iu_wait(d1_full + (g & 1), (g >> 1) & 1); if (tr_s) ta6 += clock64() - c0_; }
                iu_fence_after();
                if (a.debug & 4096) {                                   // timing probe: no softmax work at all, the window is handed on at once
                    const unsigned pbi0 = g & (PBUF - 1);
                    iu_wait(p_empty + pbi0, ((g / PBUF) & 1) ^ 1);
                    iu_fence_before();
                    __syncwarp();
                    if (lane == 0) iu_arrive(p_full + pbi0);
                    continue;
                }
                const long long cx0 = tr_s ? clock64() : 0;
                tbp = cx0;
                // class exchange: token t needs row r - (7 - s) of class s.  Every thread publishes the hi + lo sums of its two classes
                // (row r, all 8 heads) in shared memory as [class][head pair][row] float2 (consecutive rows = consecutive 8-byte
                // slots: conflict-free both ways); after the barrier it gathers the eight shifted rows of its own head pair.
                uint32_t v[2][16];
#pragma unroll
                for (int i = 0; i < 2; ++i) iu_tmem_ld16_async(trow + (g & 1) * D1_BUF_COLS + 32 * (2 * pg + i) + 16 * half, v[i]);
                iu_tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 2; ++i) iu_tmem_use16(v[i]);
                TB(0);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int s = 2 * (2 * pg + i) + half;
                    float2* dst = xch + (s * 4) * 64 + r;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dst[j * 64] = make_float2(__uint_as_float(v[i][2 * j]) + __uint_as_float(v[i][8 + 2 * j]),
                                                  __uint_as_float(v[i][2 * j + 1]) + __uint_as_float(v[i][8 + 2 * j + 1]));
                }
                TB(1);
                iu_bar_sync256(1);
                TB(2);
                float tot[2] = {0.f, 0.f};
                {
                    float2 o[8];
#pragma unroll
                    for (int s = 0; s < 8; ++s) o[s] = xch[(s * 4 + src) * 64 + max(r - (7 - s), 0)];   // (rows below 7 are never tokens of a set)
#pragma unroll
                    for (int s = 0; s < 8; ++s) { tot[0] += o[s].x; tot[1] += o[s].y; }
                }
                if (tr_s) ta7 += clock64() - cx0;
                TB(3);
                float sc[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) sc[k] = valid ? a.scale * ((FP16 ? tot[k] * a.wscale_inv : tot[k]) + ct[k]) : -INFINITY;
                if (a.dbg != nullptr && valid) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) a.dbg[((size_t)bv * HEADS + hb + k) * 256 + 1 + t] = sc[k];
                }
                const long long cr0 = tr_s ? clock64() : 0;
                bool raise = w == 0;
                if (w > 0) {
                    bool ex = false;
#pragma unroll
                    for (int k = 0; k < 2; ++k) ex = ex || (sc[k] > mref[k] + (FP16 ? TAU_FP16 : TAU));
                    raise = (a.debug & 64) ? false : iu_bar_or(1, ex);   // (also orders the exchange reads before the next window's writes)
                }
                if (raise) {
                    // window 0 establishes the reference maximum; a later window raises it only in the (rare) case above
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float x = warp_max(sc[k]);
                        if (lane == 0) smax[ws * 2 + k] = x;
                    }
                    iu_bar_sync256(1);
                    float fc[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float wm = fmaxf(smax[wpair * 2 + k], smax[(wpair + 1) * 2 + k]);
                        const float mn = w == 0 ? wm : fmaxf(mref[k], wm);
                        fc[k] = w == 0 ? 1.f : exp2f((mref[k] - mn) * LOG2E);
                        mref[k] = mn;
                    }
                    if (w > 0) {
                        // everything accumulated so far is relative to the old reference: rescale the weighted sums in TMEM
                        // (the sums of set g-1 must have completed; those of set g are not issued before this warp's
                        // arrival on p_full), the running sums and the probabilities kept for the final output
                        if (qq == 0 && lane == 0) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) fcs[hb + k] = fc[k];
                        }
                        iu_bar_sync256(1);
                        float fc8[8];
#pragma unroll
                        for (int h = 0; h < 8; ++h) fc8[h] = fcs[h];
                        iu_wait(p_empty + ((g - 1) & (PBUF - 1)), ((g - 1) / PBUF) & 1);
                        iu_fence_after();
#pragma unroll 1
                        for (int c16 = 4 * pg; c16 < 4 * pg + 4; ++c16) {
                            uint32_t y[16];
                            const uint32_t ta = trow + D2_COL + (vi & 1) * D2_BUF_COLS + 16 * c16;
                            iu_tmem_ld16(ta, y);
#pragma unroll
                            for (int n = 0; n < 16; ++n) y[n] = __float_as_uint(__uint_as_float(y[n]) * fc8[n & 7]);
                            iu_tmem_st16(ta, y);
                        }
                        iu_fence_before();
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            lsum[k] *= fc[k];
#pragma unroll
                            for (int w2 = 0; w2 < NWIN; ++w2)
                                if (w2 < w) pr[w2][k] *= fc[k];
                        }
                    }
                    iu_bar_sync256(1);                              // smax / fcs are rewritten by the next raise
                }
                if (tr_s) ta10 += clock64() - cr0;
                if (tr_s) tbp = clock64();
                unsigned short ph[2], pl[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float p = valid ? exp2f((sc[k] - mref[k]) * LOG2E) : 0.f;
                    lsum[k] += p;
                    pr[w][k] = p;
                    if (FP16) iu_split_h(p, ph[k], pl[k]); else iu_split(p, ph[k], pl[k]);
                }
                const unsigned pbi = g & (PBUF - 1);
                TB(4);
                { const long long c0_ = tr_s ? clock64() : 0; iu_wait(p_empty + pbi, ((g / PBUF) & 1) ^ 1); if (tr_s) ta8 += clock64() - c0_; }
                if (tr_s) tbp = clock64();   // the sums that read this buffer last have completed
                {
                    // row r of the hi / lo planes and row r + 1 of their copies: this thread's 2 heads = 4 bytes each
                    const uint32_t pt = iu_smem(smem + OFF_P) + pbi * P_BYTES + r * 16 + 2 * hb;
                    const uint32_t h01 = (uint32_t)ph[0] | ((uint32_t)ph[1] << 16), l01 = (uint32_t)pl[0] | ((uint32_t)pl[1] << 16);
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(pt), "r"(h01) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(pt + P_PLANE), "r"(l01) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(pt + 2 * P_PLANE + 16), "r"(h01) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(pt + 3 * P_PLANE + 16), "r"(l01) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                iu_fence_before();
                __syncwarp();
                if (lane == 0) iu_arrive(p_full + pbi);
                TB(5);
            }
            // ---- end of the view: total of the running sums, the mean token, final probabilities
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float x = warp_sum(lsum[k]);
                if (lane == 0) sred[ws * 2 + k] = x;
            }
            iu_bar_sync256(1);
            iu_wait(s0_full + (vi & 1), (vi >> 1) & 1);
            float fin[2], p0n[2], lt[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                lt[k] = sred[wpair * 2 + k] + sred[(wpair + 1) * 2 + k];
                const float s0 = sm_s0[(vi & 1) * 8 + hb + k];
                const float mf = fmaxf(mref[k], s0);
                const float f = exp2f((mref[k] - mf) * LOG2E), p0 = exp2f((s0 - mf) * LOG2E);
                const float inv = 1.0f / (lt[k] * f + p0);
                fin[k] = f * inv;
                p0n[k] = p0 * inv;
            }
            if (qq == 0 && lane == 0) {
#pragma unroll
                for (int k = 0; k < 2; ++k) { stat_l[(vi & 1) * 8 + hb + k] = lt[k]; stat_m[(vi & 1) * 8 + hb + k] = mref[k]; }
                iu_arrive(l_full + (vi & 1));
            }
#pragma unroll
            for (int w = 0; w < NWIN; ++w) {
                const int t = token_of(w);
                if (t >= 0 && !(a.debug & 128)) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        unsigned short hi, lo;
                        iu_split(pr[w][k] * fin[k], hi, lo);
                        __nv_bfloat16* dst = a.ya_hi + ((size_t)bv * HEADS + hb + k) * YA + C + 1 + t;
                        dst[0] = __ushort_as_bfloat16(hi);
                        dst[a.ya_plane] = __ushort_as_bfloat16(lo);
                    }
                }
            }
            if (r < 256 - (HW + 1)) {                                   // zero padding of the probability block (the value GEMM reads 256 columns)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    __nv_bfloat16* dst = a.ya_hi + ((size_t)bv * HEADS + hb + k) * YA + C + HW + 1 + r;
                    dst[0] = __ushort_as_bfloat16((unsigned short)0);
                    dst[a.ya_plane] = __ushort_as_bfloat16((unsigned short)0);
                }
            }
            if (r == 0) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    unsigned short hi, lo;
                    iu_split(p0n[k], hi, lo);
                    __nv_bfloat16* dst = a.ya_hi + ((size_t)bv * HEADS + hb + k) * YA + C;
                    dst[0] = __ushort_as_bfloat16(hi);
                    dst[a.ya_plane] = __ushort_as_bfloat16(lo);
                }
            }
            iu_bar_sync256(1);                                      // sred is rewritten by the next view
            TB(6);
        }
        if (tr_s) {
            g_umma_trace[9] += (unsigned long long)(clock64() - ts0);
            g_umma_trace[6] += (unsigned long long)ta6; g_umma_trace[7] += (unsigned long long)ta7;
            g_umma_trace[8] += (unsigned long long)ta8; g_umma_trace[10] += (unsigned long long)ta10;
            for (int i = 0; i < 8; ++i) g_umma_trace[16 + i] += (unsigned long long)tb[i];
        }
    } else if (warp >= EPI_WARP0 && warp < PRODUCER_WARP) {
        // ===== epilogue: mean-token score, then Y = (D2 f + p0 xbar) / L -> bf16 hi/lo planes =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
        const int q = warp & 3, et = threadIdx.x - 32 * EPI_WARP0, half = q >> 1, row = 32 * (q & 1) + lane;   // class row of class 2p + half
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        const bool tr_e = tracing && warp == EPI_WARP0;
        const long long te0 = tr_e ? clock64() : 0;
        long long ta11 = 0, ta12 = 0, ta13 = 0, ta14 = 0;
        for (int vi = 0; vi < nviews; ++vi) {
            const int bv = blockIdx.x + vi * gridDim.x;
            const long long ce0 = tr_e ? clock64() : 0;
            if (!(a.debug & 256)) {   // s0[h] = scale (w_eff[h] . xbar + cterm[h][0]); this thread: columns 4 et .. 4 et + 3 (class et >> 4, rows 4 (et & 15)..)
                const int c0 = 4 * et, cls = c0 >> 6, r0 = c0 & 63;
                float xb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) xb[k] = __ldg(a.xbar + (size_t)bv * C + cls + 8 * (r0 + k));
                float dot[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const uint2 hi = __ldg(reinterpret_cast<const uint2*>(a.wpl + (((size_t)bv * 2 + 0) * HEADS + h) * C + c0));
                    const uint2 lo = __ldg(reinterpret_cast<const uint2*>(a.wpl + (((size_t)bv * 2 + 1) * HEADS + h) * C + c0));
                    float d;
                    if (FP16) {
                        const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&hi.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
                        const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&lo.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&lo.y));
                        d = (h0.x + l0.x) * xb[0];
                        d = fmaf(h0.y + l0.y, xb[1], d);
                        d = fmaf(h1.x + l1.x, xb[2], d);
                        d = fmaf(h1.y + l1.y, xb[3], d);
                        d *= a.wscale_inv;
                    } else {
                        d = (iu_bf(hi.x, 0) + iu_bf(lo.x, 0)) * xb[0];
                        d = fmaf(iu_bf(hi.x, 1) + iu_bf(lo.x, 1), xb[1], d);
                        d = fmaf(iu_bf(hi.y, 0) + iu_bf(lo.y, 0), xb[2], d);
                        d = fmaf(iu_bf(hi.y, 1) + iu_bf(lo.y, 1), xb[3], d);
                    }
                    dot[h] = d;
                }
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const float x = warp_sum(dot[h]);
                    if (lane == 0) ered[q * 8 + h] = x;
                }
                iu_bar_sync(2);
                if (et == 0) {
#pragma unroll
                    for (int h = 0; h < 8; ++h) {
                        const float s0 = a.scale * (((ered[h] + ered[8 + h]) + (ered[16 + h] + ered[24 + h])) + __ldg(a.cterm + ((size_t)bv * HEADS + h) * TP));
                        sm_s0[(vi & 1) * 8 + h] = s0;
                        if (a.dbg != nullptr) a.dbg[((size_t)bv * HEADS + h) * 256] = s0;
                    }
                    iu_arrive(s0_full + (vi & 1));
                }
            } else if (et == 0) {
                for (int h = 0; h < 8; ++h) sm_s0[(vi & 1) * 8 + h] = 0.f;
                iu_arrive(s0_full + (vi & 1));
            }
            const long long ce1 = tr_e ? clock64() : 0;
            iu_wait(l_full + (vi & 1), (vi >> 1) & 1);
            const long long ce2 = tr_e ? clock64() : 0;
            float fin[8], p0n[8];
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const float m = stat_m[(vi & 1) * 8 + h], s0 = sm_s0[(vi & 1) * 8 + h];
                const float mf = fmaxf(m, s0);
                const float f = exp2f((m - mf) * LOG2E), p0 = exp2f((s0 - mf) * LOG2E);
                const float inv = 1.0f / (stat_l[(vi & 1) * 8 + h] * f + p0);
                fin[h] = f * inv;
                p0n[h] = p0 * inv;
            }
            iu_wait(d2_full + (vi & 1), (vi >> 1) & 1);
            const long long ce3 = tr_e ? clock64() : 0;
            iu_fence_after();
#pragma unroll 1
            for (int p = 0; p < 4; ++p) {
                uint32_t y[16];
                iu_tmem_ld16(trow + D2_COL + (vi & 1) * D2_BUF_COLS + 32 * p + 16 * half, y);
                const int s = 2 * p + half;
                const int cp = 64 * s + row;                                // output column 64 s + r of channel s + 8 r
                const float xb = __ldg(a.xbar + (size_t)bv * C + s + 8 * row);
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const float val = (__uint_as_float(y[h]) + __uint_as_float(y[8 + h])) * fin[h] + p0n[h] * xb;
                    unsigned short hi, lo;
                    iu_split(val, hi, lo);
                    __nv_bfloat16* dst = a.ya_hi + ((size_t)bv * HEADS + h) * YA + cp;
                    if (a.debug & 512) continue;
                    dst[0] = __ushort_as_bfloat16(hi);
                    dst[a.ya_plane] = __ushort_as_bfloat16(lo);
                }
            }
            iu_fence_before();
            __syncwarp();
            if (lane == 0) iu_arrive(d2_empty + (vi & 1));
            if (tr_e) {
                const long long ce4 = clock64();
                ta11 += ce1 - ce0; ta12 += ce2 - ce1; ta13 += ce3 - ce2; ta14 += ce4 - ce3;
            }
            iu_bar_sync(2);                                         // ered / sm_s0 of view vi + 2 reuse this parity: keep the warps together
        }
        if (tr_e) {
            g_umma_trace[15] += (unsigned long long)(clock64() - te0);
            g_umma_trace[11] += (unsigned long long)ta11; g_umma_trace[12] += (unsigned long long)ta12;
            g_umma_trace[13] += (unsigned long long)ta13; g_umma_trace[14] += (unsigned long long)ta14;
        }
    }
    iu_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------- host
}  // namespace pt
extern "C" int pt_debug_umma_trace(unsigned long long* out16, int reset) {   // 24 slots
    if (out16 && cudaMemcpyFromSymbol(out16, pt::g_umma_trace, sizeof(pt::g_umma_trace)) != cudaSuccess) return PT_ERR_CUDA;
    if (reset) {
        unsigned long long z[24] = {};
        if (cudaMemcpyToSymbol(pt::g_umma_trace, z, sizeof(z)) != cudaSuccess) return PT_ERR_CUDA;
    }
    return PT_OK;
}
namespace pt {
bool img_pool_umma_supported(int img_dtype) { return img_dtype == PT_DTYPE_BF16 || img_dtype == PT_DTYPE_F16; }

int launch_img_pool_umma(const void* img_feat, bool fp16, float wscale, const __nv_bfloat16* wpl, const float* cterm, const float* xbar, __nv_bfloat16* ya_hi,
                         long long ya_plane, int BV, float* dbg, cudaStream_t s) {
    using namespace ipu;
    PT_REQUIRE(((uintptr_t)img_feat & 15) == 0 && ((uintptr_t)wpl & 15) == 0, "pt_img_attnpool: img_feat / workspace must be 16-byte aligned");
    UmmaPoolMaps maps;
    int rc;
    {
        const unsigned long long dims[2] = {(unsigned long long)C, (unsigned long long)BV * 16};
        const unsigned long long strides[1] = {(unsigned long long)C * 2};
        const unsigned box[2] = {64u, 16u};
        if ((rc = encode_tensor_map_16bit(&maps.w, wpl, 2, dims, strides, box, fp16))) return rc;
    }
    {   // class-aligned view of the raw (BV, 512, 225) features: element (u, r, s, v) at byte 2 u + 3600 r + 448 s + 230400 v
        const unsigned long long dims[4] = {(unsigned long long)UCOLS, 64ull, 8ull, (unsigned long long)BV};
        const unsigned long long strides[3] = {3600ull, 448ull, (unsigned long long)C * HW * 2};
        const unsigned box[4] = {64u, 64u, 2u, 1u};
        if ((rc = encode_tensor_map_16bit(&maps.x, img_feat, 4, dims, strides, box, fp16))) return rc;
    }
    static bool attr_set[PT_MAX_DEVICES] = {};
    if (first_use_on_current_device(attr_set)) {
        PT_CUDA_OK(cudaFuncSetAttribute(img_pool_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        PT_CUDA_OK(cudaFuncSetAttribute(img_pool_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    UmmaPoolArgs a;
    a.wpl = wpl; a.cterm = cterm; a.xbar = xbar; a.ya_hi = ya_hi; a.ya_plane = ya_plane; a.BV = BV;
    a.scale = (float)(1.0 / sqrt(32.0));
    a.dbg = dbg;
    a.wscale_inv = 1.0f / wscale;
    const char* dbe = getenv("PT_UMMA_DEBUG");
    a.debug = dbe ? atoi(dbe) : 0;
    int grid = BV < sms ? BV : sms;
    if (const char* ge = getenv("PT_POOL_GRID")) { const int gv = atoi(ge); if (gv >= 1 && gv < grid) grid = gv; }   // probes: fewer persistent CTAs
    {
        ProfScope prof_(PROF_IMG_POOL, s);
        if (fp16) img_pool_umma_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>(maps, a);
        else img_pool_umma_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>(maps, a);
    }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

}  // namespace pt
