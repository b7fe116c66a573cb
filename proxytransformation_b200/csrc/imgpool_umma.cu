// S9 image proxies, pass B (scores -> softmax -> attention-weighted feature sums) on tcgen05 tensor cores:
// get_img_proxy (:335-342) + AttentionPool2d.forward (:154-177) in single-query form (algebra in imgpool.cu), shipped
// geometry C = 512 channels, 15 x 15 = 225 positions, 8 heads of 32, 16-bit features (bf16).
//
// The feature map of a view is [512 channels][225 tokens] with a 450-byte row pitch.  UMMA shared-memory descriptors (and TMA
// tensor maps) need 16-byte aligned rows, so the rows have to be REALIGNED on their way into shared memory.  225 = 1 (mod 8):
// the rows of one residue CLASS s = channel mod 8 (channels s + 8r, r = 0..63) are 3600 bytes apart and all start 2 s bytes
// after a 16-byte boundary, so a class is realigned by ONE shift of s elements.  TMA cannot do it (box origins must be 16-byte
// aligned in global memory — measured: the first box with an odd origin raises "illegal instruction"), so loader warps do:
// coalesced 16-byte loads of the aligned chunks, a funnel shift by s elements across neighbouring chunks (shuffles), and
// 16-byte stores into the SWIZZLE_128B tile layout the tensor core reads.
// One 8 KB tile [64 class rows][64 tokens] serves both contractions without any copy:
//   scores  D1[token][n]   += X^T W^T : A = the tile read MN-major (M = 64 tokens, K = 16 class rows per MMA),
//                                       B = w_eff rows n = (hi|lo, head) of this class, K-major           (tcgen05.mma M64 N16 K16)
//   sums    D2[channel][n] += X  P^T : A = two class tiles read K-major (M = 128 channel rows, K = 16 tokens per MMA),
//                                       B = probabilities n = (hi|lo, head), K-major over the tokens      (tcgen05.mma M128 N16 K16)
// The fp32 operands (w_eff, probabilities) enter as bf16 hi + lo halves in separate N columns, so every product is exact and
// the accumulation is fp32 in TMEM (same numerics as the 3xBF16 GEMMs: ~2^-17 relative).
//
// One persistent CTA per SM, a view is streamed ONCE as 4 token windows of 64 (flash-attention structure, one query per head):
//   warp 0       TMA producer of the per-view w_eff planes (8 boxes of 2 KB, SWIZZLE_128B)
//   warp 1       MMA issuer (one elected thread): scores of window g+1 interleaved with the sums of window g; every ring slot is
//                handed back to the loaders by tcgen05.commit when the sums that read it have completed
//   warps 4-7    softmax: tcgen05.ld the 64 x 16 score tile (lane = token), + position term, exp relative to a per-view
//                reference maximum (established by window 0, raised FA-style by rescaling the accumulators in TMEM only when
//                a later window exceeds it by more than TAU — never on ordinary data), bf16 hi/lo probability tile -> shared
//                memory (the B operand of the sums), running sum in registers; end of view: final probabilities -> global
//   warps 8-11   epilogue: mean-token score s0 = w_eff . xbar (CUDA cores, from L2), then per view Y = (D2 f + p0 xbar) / L
//                from TMEM -> bf16 hi/lo planes for the value-side GEMM (same output format as the mma.sync kernel)
//   warps 12-19  loaders: warp l owns rows 8 l .. 8 l + 7 of every class tile; per window it issues its 20 loads (16 x 4 rows x
//                8 chunks + the 9th chunk of its 64 rows) before it touches the ring, so 80 KB of loads are in flight per SM
//                while earlier windows are consumed; 10-slot ring of class-pair slots (160 KB)
// Channel order of w_eff columns and of the weighted sums: position 64 s + r  <->  channel s + 8 r (absorbed into the folded
// GEMM weights on the host, pt_img_pool_params variant 1).
#include "common.cuh"
#include "gemm_tc.cuh"

#include <cuda.h>
#include <math.h>
#include <stdlib.h>

namespace pt {

namespace ipu {
constexpr int C = 512, HW = 225, HEADS = 8;
constexpr int TP = 228;                      // cterm row pitch (floats)
constexpr int YA = 768;                      // output row: 512 weighted sums + 256 probabilities
constexpr int NWIN = 4, WTOK = 64;           // token windows per view
constexpr int TILE_BYTES = 64 * 128;         // [64 class rows][64 tokens] bf16
constexpr int SLOT_BYTES = 2 * TILE_BYTES;   // a class pair: rows 0-63 class 2p, rows 64-127 class 2p+1
constexpr int RING = 10;
constexpr int WCLASS_BYTES = 16 * 128;       // w_eff rows (hi|lo, head) x 64 class channels
constexpr int W_BYTES = 8 * WCLASS_BYTES;
constexpr int P_BYTES = 16 * 128, PBUF = 4;  // probability tile [16 rows (hi|lo, head)][64 tokens]; buffer = window index
constexpr int OFF_RING = 0;
constexpr int OFF_W = OFF_RING + RING * SLOT_BYTES;
constexpr int OFF_P = OFF_W + 2 * W_BYTES;
constexpr int OFF_MISC = OFF_P + PBUF * P_BYTES;      // floats: smax[32] sred[32] ered[32] s0[16] stat_l[16] stat_m[16]
constexpr int OFF_BAR = OFF_MISC + 1024;
// mbarriers: full[RING] empty[RING] wfull[2] wempty[2] d1_full[2] p_full[PBUF] p_empty[PBUF] d2_full[2] d2_empty[2] s0_full[2] l_full[2]
constexpr int NBAR = 2 * RING + 6 + 2 * PBUF + 8;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;      // + slack for the 1024-byte alignment of the swizzled tiles
// warp roles: 0 = MMA issuer + w_eff TMA + TMEM allocation, 1-4 softmax, 5-8 epilogue (TMEM lane quarter = warp mod 4), 9-16 loaders
constexpr int SOFTMAX_WARP0 = 1, EPI_WARP0 = 5, LOADER_WARP0 = 9, LOADER_WARPS = 8;
constexpr int THREADS = 32 * (LOADER_WARP0 + LOADER_WARPS);
constexpr int TMEM_COLS = 256;               // D1: 2 x 16 columns at 0 ; D2: 2 x 64 columns at 64
constexpr int D2_COL = 64;
constexpr float TAU = 16.0f;                 // the reference maximum is raised when a score exceeds it by more than this
constexpr float LOG2E = 1.4426950408889634f;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(OFF_W % 1024 == 0 && OFF_P % 1024 == 0 && SLOT_BYTES % 1024 == 0, "swizzled tiles need 1024-byte alignment");
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
// kind::f16 instruction descriptor: D = f32, A = B = bf16; bit 15 = A is MN-major; N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t iu_idesc(int m, int n, bool a_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major ? (1u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void iu_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
__device__ __forceinline__ bool iu_bar_or(int id, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.b32 q, %2, 0;\n"
        "bar.red.or.pred p, %1, 128, q;\n"
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
__device__ __forceinline__ float iu_bf(uint32_t packed, int k) {      // element k (0/1) of a packed bf16 pair
    return __uint_as_float(k ? (packed & 0xffff0000u) : (packed << 16));
}

struct UmmaPoolMaps {
    CUtensorMap w;         // w_eff planes as (512 columns, BV*16 rows (view, hi|lo, head)), box (64, 16)
};

__device__ __forceinline__ uint4 iu_ldg_stream(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// One 16-byte output chunk of a class-S row: elements [8 j + S, 8 j + S + 8) of the row in u = token + S coordinates, i.e. the
// tail of aligned chunk j (this lane's `a`) and the head of chunk j + 1 (the next lane's `a`; for j == 7 the 9th chunk of the
// row, which lane `tail_lane` holds in `t`), stored at chunk j of tile row `row` (SWIZZLE_128B: chunk index XOR row mod 8).
template <int S>
__device__ __forceinline__ void iu_shift_store(uint32_t tile, int row, int j, const uint4& a, const uint4& t, int tail_lane) {
    constexpr int Q = S >> 1, ODD = S & 1, NB = Q + ODD;
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, tw[4] = {t.x, t.y, t.z, t.w};
    uint32_t A[9] = {a.x, a.y, a.z, a.w, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const uint32_t nxt = __shfl_down_sync(FULL, aw[i], 1);
        const uint32_t tl = __shfl_sync(FULL, tw[i], tail_lane);
        A[4 + i] = j == 7 ? tl : nxt;
    }
    uint4 o;
    if (ODD) {
        o.x = __funnelshift_r(A[Q], A[Q + 1], 16); o.y = __funnelshift_r(A[Q + 1], A[Q + 2], 16);
        o.z = __funnelshift_r(A[Q + 2], A[Q + 3], 16); o.w = __funnelshift_r(A[Q + 3], A[Q + 4], 16);
    } else {
        o.x = A[Q]; o.y = A[Q + 1]; o.z = A[Q + 2]; o.w = A[Q + 3];
    }
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tile + row * 128 + ((j ^ (row & 7)) << 4)), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
}

// Class pair P of one window: both tiles of ring slot `tile0`, this warp's 8 rows of each (two groups of 4 rows x 8 chunks).
template <int P>
__device__ __forceinline__ void iu_store_pair(uint32_t tile0, int row0, int rr, int j, const uint4 (&am)[2][2], const uint4& at) {
    iu_shift_store<2 * P>(tile0, row0 + rr, j, am[0][0], at, rr);
    iu_shift_store<2 * P>(tile0, row0 + 4 + rr, j, am[0][1], at, 4 + rr);
    iu_shift_store<2 * P + 1>(tile0 + ipu::TILE_BYTES, row0 + rr, j, am[1][0], at, 8 + rr);
    iu_shift_store<2 * P + 1>(tile0 + ipu::TILE_BYTES, row0 + 4 + rr, j, am[1][1], at, 12 + rr);
}

struct UmmaPoolArgs {
    const uint8_t* img;          // (BV, 512, 225) 16-bit features
    const __nv_bfloat16* wpl;    // (BV, 2, 8, 512) bf16 hi / lo planes of w_eff, columns 64 s + r
    const float* cterm;          // (BV, 8, TP)
    const float* xbar;           // (BV, 512) natural channel order
    __nv_bfloat16* ya_hi;        // (BV, 8, 768) bf16 hi plane (lo plane ya_plane elements later)
    long long ya_plane;
    int BV;
    float scale;
    float* dbg;                  // optional (PT_POOL_DEBUG bit 64): [BV][8][256] scaled scores, attention tokens 0..225
    int debug;                   // PT_UMMA_DEBUG bring-up switches (garbage results): 1 no score MMAs, 2 no sum MMAs, 4 score MMAs with a
                                 // K-major A descriptor, 8 score MMAs with M = 128
};

__global__ void __launch_bounds__(ipu::THREADS, 1) img_pool_umma_kernel(const __grid_constant__ UmmaPoolMaps maps, const UmmaPoolArgs a) {
    using namespace ipu;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) { iu_mbar_init(full + i, LOADER_WARPS); iu_mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) {
            iu_mbar_init(wfull + i, 1); iu_mbar_init(wempty + i, 1); iu_mbar_init(d1_full + i, 1);
            iu_mbar_init(d2_full + i, 1); iu_mbar_init(d2_empty + i, 4); iu_mbar_init(s0_full + i, 1); iu_mbar_init(l_full + i, 1);
        }
        for (int i = 0; i < PBUF; ++i) { iu_mbar_init(p_full + i, 4); iu_mbar_init(p_empty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(iu_smem(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    iu_fence_before();
    __syncthreads();
    iu_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int nviews = (int)blockIdx.x < a.BV ? (a.BV - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp >= LOADER_WARP0) {
        // ===== loaders: global -> registers -> (shift by the class) -> swizzled tiles =====
        const int row0 = 8 * (warp - LOADER_WARP0), rr = lane >> 3, j = lane & 7;
        unsigned it = 0;
        for (int vi = 0; vi < nviews; ++vi) {
            const int bv = blockIdx.x + vi * gridDim.x;
            const uint8_t* view = a.img + (size_t)bv * (C * HW * 2);
#pragma unroll 1
            for (int w = 0; w < NWIN; ++w) {
                uint4 am[4][2][2], at[4];
                const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
                const int chunk = 8 * w + j;                               // aligned 16-byte chunk of the row; 29 chunks per row
#pragma unroll
                for (int p = 0; p < 4; ++p) {
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf)
                            am[p][e][hf] = chunk <= 28 ? iu_ldg_stream(view + 448 * (2 * p + e) + 3600 * (row0 + 4 * hf + rr) + 16 * chunk) : zero;
                    // 9th chunk of this warp's 16 rows of the pair: lanes 0-7 class 2p, lanes 8-15 class 2p + 1
                    at[p] = (lane < 16 && 8 * w + 8 <= 28) ? iu_ldg_stream(view + 448 * (2 * p + (lane >> 3)) + 3600 * (row0 + (lane & 7)) + 16 * (8 * w + 8)) : zero;
                }
#pragma unroll
                for (int p = 0; p < 4; ++p, ++it) {
                    const unsigned slot = it % RING;
                    iu_wait(empty + slot, ((it / RING) & 1) ^ 1);
                    const uint32_t tile0 = iu_smem(smem + OFF_RING) + slot * SLOT_BYTES;
                    if (p == 0) iu_store_pair<0>(tile0, row0, rr, j, am[0], at[0]);
                    else if (p == 1) iu_store_pair<1>(tile0, row0, rr, j, am[1], at[1]);
                    else if (p == 2) iu_store_pair<2>(tile0, row0, rr, j, am[2], at[2]);
                    else iu_store_pair<3>(tile0, row0, rr, j, am[3], at[3]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) iu_arrive(full + slot);
                }
            }
        }
    } else if (warp == 0) {
        // ===== MMA issuer (+ TMA of the per-view w_eff planes, one view ahead) =====
        if (lane == 0) {
            constexpr uint32_t IDESC1 = iu_idesc(64, 16, true), IDESC2 = iu_idesc(128, 16, false);
            const uint32_t ring = iu_smem(smem + OFF_RING), wbase = iu_smem(smem + OFF_W), pbase = iu_smem(smem + OFF_P);
            unsigned it1 = 0, it2 = 0, g = 0;
            // sums of window gp (class pair p), interleaved below with the scores of window gp + 1
            auto sums = [&](unsigned gp, int p) {
                const unsigned vp = gp >> 2, wp = gp & 3;
                if (p == 0) {
                    iu_wait(p_full + wp, (gp >> 2) & 1);                            // probability tile of window gp is in shared memory
                    if (wp == 0) iu_wait(d2_empty + (vp & 1), ((vp >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator buffer
                    iu_fence_after();
                }
                const unsigned slot = it2 % RING;
                const uint32_t sa = ring + slot * SLOT_BYTES, sp = pbase + wp * P_BYTES;
                const uint32_t d2 = tmem + D2_COL + (vp & 1) * 64 + 16 * p;
                const int nk = wp == 3 ? 3 : 4;                                      // tokens 240..255 do not exist
                for (int j = 0; j < nk; ++j)
                    if (!(a.debug & 2)) iu_mma(d2, iu_desc(sa + 32 * j, 0), iu_desc(sp + 32 * j, 0), IDESC2, (wp | j) != 0 ? 1u : 0u);
                iu_commit(empty + slot);                                             // slot back to the producer once read
                ++it2;
                if (p == 3) {
                    iu_commit(p_empty + wp);
                    if (wp == 3) iu_commit(d2_full + (vp & 1));
                }
            };
            auto load_w = [&](int vi) {            // w_eff planes of this CTA's view vi -> buffer vi & 1 (free once the scores of view vi - 2 are done)
                if (vi >= nviews) return;
                const int bv = blockIdx.x + vi * gridDim.x, wb = vi & 1;
                iu_wait(wempty + wb, ((vi >> 1) & 1) ^ 1);
                iu_expect_tx(wfull + wb, W_BYTES);
#pragma unroll
                for (int s = 0; s < 8; ++s) iu_tma_2d(smem + OFF_W + wb * W_BYTES + s * WCLASS_BYTES, &maps.w, 64 * s, 16 * bv, wfull + wb);
            };
            load_w(0);
            for (int vi = 0; vi < nviews; ++vi) {
                const int wb = vi & 1;
                load_w(vi + 1);
                iu_wait(wfull + wb, (vi >> 1) & 1);
                iu_fence_after();
                for (int w = 0; w < NWIN; ++w, ++g) {
                    const uint32_t d1 = tmem + (g & 1) * 16;
                    for (int p = 0; p < 4; ++p) {
                        const unsigned slot = it1 % RING;
                        iu_wait(full + slot, (it1 / RING) & 1);
                        iu_fence_after();
                        const uint32_t sa = ring + slot * SLOT_BYTES;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const uint32_t sw = wbase + wb * W_BYTES + (2 * p + e) * WCLASS_BYTES;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (a.debug & 1) continue;
                                const uint32_t idesc = (a.debug & 4) ? iu_idesc(64, 16, false) : (a.debug & 8) ? iu_idesc(128, 16, true) : IDESC1;
                                iu_mma(d1, iu_desc(sa + e * TILE_BYTES + 2048 * j, TILE_BYTES), iu_desc(sw + 32 * j, 0), idesc,
                                       (p | e | j) != 0 ? 1u : 0u);
                            }
                        }
                        ++it1;
                        if (p == 3) {
                            iu_commit(d1_full + (g & 1));
                            if (w == 3) iu_commit(wempty + wb);
                        }
                        if (g > 0) sums(g - 1, p);
                    }
                }
            }
            if (g > 0)
                for (int p = 0; p < 4; ++p) sums(g - 1, p);
        }
    } else if (warp >= SOFTMAX_WARP0 && warp < EPI_WARP0) {
        // ===== softmax: lane < 16 of warp q owns token 16 q + lane of every window (TMEM lanes of an M = 64 accumulator) =====
        const int q = warp & 3;
        const bool act = lane < 16;
        const int tl = 16 * q + (lane & 15);
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        float mref[8], lsum[8], pr[NWIN][8];
        unsigned g = 0;
        for (int vi = 0; vi < nviews; ++vi) {
            const int bv = blockIdx.x + vi * gridDim.x;
#pragma unroll
            for (int h = 0; h < 8; ++h) { lsum[h] = 0.f; mref[h] = 0.f; }
#pragma unroll
            for (int w = 0; w < NWIN; ++w, ++g) {
                const int t = WTOK * w + tl;
                const bool valid = act && t < HW;
                float ct[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) ct[h] = valid ? __ldg(a.cterm + ((size_t)bv * HEADS + h) * TP + 1 + t) : 0.f;
                iu_wait(d1_full + (g & 1), (g >> 1) & 1);
                iu_fence_after();
                uint32_t v[16];
                iu_tmem_ld16(trow + (g & 1) * 16, v);
                float s[8];
#pragma unroll
                for (int h = 0; h < 8; ++h)
                    s[h] = valid ? a.scale * ((__uint_as_float(v[h]) + __uint_as_float(v[8 + h])) + ct[h]) : -INFINITY;
                if (a.dbg != nullptr && valid) {
#pragma unroll
                    for (int h = 0; h < 8; ++h) a.dbg[((size_t)bv * HEADS + h) * 256 + 1 + t] = s[h];
                }
                bool raise = w == 0;
                if (w > 0) {
                    bool ex = false;
#pragma unroll
                    for (int h = 0; h < 8; ++h) ex = ex || (s[h] > mref[h] + TAU);
                    raise = iu_bar_or(1, ex);
                }
                if (raise) {
                    // window 0 establishes the reference maximum; a later window raises it only in the (rare) case above
#pragma unroll
                    for (int h = 0; h < 8; ++h) {
                        const float x = warp_max(s[h]);
                        if (lane == 0) smax[q * 8 + h] = x;
                    }
                    iu_bar_sync(1);
                    float fc[8];
#pragma unroll
                    for (int h = 0; h < 8; ++h) {
                        const float wm = fmaxf(fmaxf(smax[h], smax[8 + h]), fmaxf(smax[16 + h], smax[24 + h]));
                        const float mn = w == 0 ? wm : fmaxf(mref[h], wm);
                        fc[h] = w == 0 ? 1.f : exp2f((mref[h] - mn) * LOG2E);
                        mref[h] = mn;
                    }
                    if (w > 0) {
                        // everything accumulated so far is relative to the old reference: rescale the weighted sums in TMEM
                        // (the sums of window g-1 must have completed; those of window g are not issued before this warp's
                        // arrival on p_full), the running sums and the probabilities kept for the final output
                        iu_wait(p_empty + ((g - 1) & 3), ((g - 1) >> 2) & 1);
                        iu_fence_after();
#pragma unroll 1
                        for (int p = 0; p < 4; ++p) {
                            uint32_t y[16];
                            const uint32_t ta = trow + D2_COL + (vi & 1) * 64 + 16 * p;
                            iu_tmem_ld16(ta, y);
#pragma unroll
                            for (int n = 0; n < 16; ++n) y[n] = __float_as_uint(__uint_as_float(y[n]) * fc[n & 7]);
                            iu_tmem_st16(ta, y);
                        }
                        iu_fence_before();
#pragma unroll
                        for (int h = 0; h < 8; ++h) {
                            lsum[h] *= fc[h];
#pragma unroll
                            for (int w2 = 0; w2 < NWIN; ++w2)
                                if (w2 < w) pr[w2][h] *= fc[h];
                        }
                    }
                }
                unsigned short ph[8], pl[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const float p = valid ? exp2f((s[h] - mref[h]) * LOG2E) : 0.f;
                    lsum[h] += p;
                    pr[w][h] = p;
                    iu_split(p, ph[h], pl[h]);
                }
                iu_wait(p_empty + w, ((g >> 2) & 1) ^ 1);           // the sums of the previous view's window w have read this buffer
                if (act) {
                    // K-major SWIZZLE_128B tile: row n (128 bytes = 64 tokens), 16-byte chunk index XOR (n mod 8)
                    const uint32_t pt = iu_smem(smem + OFF_P) + w * P_BYTES;
#pragma unroll
                    for (int h = 0; h < 8; ++h) {
                        const uint32_t off = pt + h * 128 + (((tl >> 3) ^ h) << 4) + ((tl & 7) << 1);
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(off), "h"(ph[h]) : "memory");
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(off + 1024), "h"(pl[h]) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                iu_fence_before();
                __syncwarp();
                if (lane == 0) iu_arrive(p_full + w);
            }
            // ---- end of the view: total of the running sums, the mean token, final probabilities
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const float x = warp_sum(lsum[h]);
                if (lane == 0) sred[q * 8 + h] = x;
            }
            iu_bar_sync(1);
            iu_wait(s0_full + (vi & 1), (vi >> 1) & 1);
            float fin[8], p0n[8], lt[8];
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                lt[h] = (sred[h] + sred[8 + h]) + (sred[16 + h] + sred[24 + h]);
                const float s0 = sm_s0[(vi & 1) * 8 + h];
                const float mf = fmaxf(mref[h], s0);
                const float f = exp2f((mref[h] - mf) * LOG2E), p0 = exp2f((s0 - mf) * LOG2E);
                const float inv = 1.0f / (lt[h] * f + p0);
                fin[h] = f * inv;
                p0n[h] = p0 * inv;
            }
            if (warp == SOFTMAX_WARP0 && lane == 0) {
#pragma unroll
                for (int h = 0; h < 8; ++h) { stat_l[(vi & 1) * 8 + h] = lt[h]; stat_m[(vi & 1) * 8 + h] = mref[h]; }
                iu_arrive(l_full + (vi & 1));
            }
            if (act) {
#pragma unroll
                for (int w = 0; w < NWIN; ++w) {
                    const int t = WTOK * w + tl;
                    if (1 + t < 256) {
#pragma unroll
                        for (int h = 0; h < 8; ++h) {
                            unsigned short hi, lo;
                            iu_split(t < HW ? pr[w][h] * fin[h] : 0.f, hi, lo);
                            __nv_bfloat16* dst = a.ya_hi + ((size_t)bv * HEADS + h) * YA + C + 1 + t;
                            dst[0] = __ushort_as_bfloat16(hi);
                            dst[a.ya_plane] = __ushort_as_bfloat16(lo);
                        }
                    }
                }
                if (tl == 0) {
#pragma unroll
                    for (int h = 0; h < 8; ++h) {
                        unsigned short hi, lo;
                        iu_split(p0n[h], hi, lo);
                        __nv_bfloat16* dst = a.ya_hi + ((size_t)bv * HEADS + h) * YA + C;
                        dst[0] = __ushort_as_bfloat16(hi);
                        dst[a.ya_plane] = __ushort_as_bfloat16(lo);
                    }
                }
            }
        }
    } else if (warp >= EPI_WARP0 && warp < LOADER_WARP0) {
        // ===== epilogue: mean-token score, then Y = (D2 f + p0 xbar) / L -> bf16 hi/lo planes =====
        const int q = warp & 3, et = threadIdx.x - 32 * EPI_WARP0, row = 32 * q + lane;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        for (int vi = 0; vi < nviews; ++vi) {
            const int bv = blockIdx.x + vi * gridDim.x;
            {   // s0[h] = scale (w_eff[h] . xbar + cterm[h][0]); this thread: columns 4 et .. 4 et + 3 (class et >> 4, rows 4 (et & 15)..)
                const int c0 = 4 * et, cls = c0 >> 6, r0 = c0 & 63;
                float xb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) xb[k] = __ldg(a.xbar + (size_t)bv * C + cls + 8 * (r0 + k));
                float dot[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const uint2 hi = __ldg(reinterpret_cast<const uint2*>(a.wpl + (((size_t)bv * 2 + 0) * HEADS + h) * C + c0));
                    const uint2 lo = __ldg(reinterpret_cast<const uint2*>(a.wpl + (((size_t)bv * 2 + 1) * HEADS + h) * C + c0));
                    float d = (iu_bf(hi.x, 0) + iu_bf(lo.x, 0)) * xb[0];
                    d = fmaf(iu_bf(hi.x, 1) + iu_bf(lo.x, 1), xb[1], d);
                    d = fmaf(iu_bf(hi.y, 0) + iu_bf(lo.y, 0), xb[2], d);
                    d = fmaf(iu_bf(hi.y, 1) + iu_bf(lo.y, 1), xb[3], d);
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
            }
            iu_wait(l_full + (vi & 1), (vi >> 1) & 1);
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
            iu_fence_after();
#pragma unroll 1
            for (int p = 0; p < 4; ++p) {
                uint32_t y[16];
                iu_tmem_ld16(trow + D2_COL + (vi & 1) * 64 + 16 * p, y);
                const int cp = 128 * p + row;                           // output column 64 s + r of channel s + 8 r
                const float xb = __ldg(a.xbar + (size_t)bv * C + (cp >> 6) + 8 * (cp & 63));
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const float val = (__uint_as_float(y[h]) + __uint_as_float(y[8 + h])) * fin[h] + p0n[h] * xb;
                    unsigned short hi, lo;
                    iu_split(val, hi, lo);
                    __nv_bfloat16* dst = a.ya_hi + ((size_t)bv * HEADS + h) * YA + cp;
                    dst[0] = __ushort_as_bfloat16(hi);
                    dst[a.ya_plane] = __ushort_as_bfloat16(lo);
                }
            }
            iu_fence_before();
            __syncwarp();
            if (lane == 0) iu_arrive(d2_empty + (vi & 1));
        }
    }
    iu_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------- host
bool img_pool_umma_supported(int img_dtype) { return img_dtype == PT_DTYPE_BF16; }

int launch_img_pool_umma(const void* img_feat, const __nv_bfloat16* wpl, const float* cterm, const float* xbar, __nv_bfloat16* ya_hi,
                         long long ya_plane, int BV, float* dbg, cudaStream_t s) {
    using namespace ipu;
    PT_REQUIRE(((uintptr_t)img_feat & 15) == 0 && ((uintptr_t)wpl & 15) == 0, "pt_img_attnpool: img_feat / workspace must be 16-byte aligned");
    UmmaPoolMaps maps;
    int rc;
    {
        const unsigned long long dims[2] = {(unsigned long long)C, (unsigned long long)BV * 16};
        const unsigned long long strides[1] = {(unsigned long long)C * 2};
        const unsigned box[2] = {64u, 16u};
        if ((rc = encode_tensor_map_16bit(&maps.w, wpl, 2, dims, strides, box, false))) return rc;
    }
    static bool attr_set[PT_MAX_DEVICES] = {};
    if (first_use_on_current_device(attr_set))
        PT_CUDA_OK(cudaFuncSetAttribute(img_pool_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    UmmaPoolArgs a;
    a.img = (const uint8_t*)img_feat; a.wpl = wpl; a.cterm = cterm; a.xbar = xbar; a.ya_hi = ya_hi; a.ya_plane = ya_plane; a.BV = BV;
    a.scale = (float)(1.0 / sqrt(32.0));
    a.dbg = dbg;
    const char* dbe = getenv("PT_UMMA_DEBUG");
    a.debug = dbe ? atoi(dbe) : 0;
    const int grid = BV < sms ? BV : sms;
    { ProfScope prof_(PROF_IMG_POOL, s); img_pool_umma_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(maps, a); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

}  // namespace pt
