// Tensor-core GEMM for the dense layers of the preshape path (ProxyBlock S7, image-pool projections S9):
//     C[M,N] = act(A[M,K] W[N,K]^T + bias) + residual
// with fp32-ACCURATE results from bf16 tensor cores via 3xBF16 operand splitting:
//     a = a_hi + a_lo,  w = w_hi + w_lo   (hi = bf16(x), lo = bf16(x - hi))
//     a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo        (the dropped lo*lo term is ~2^-18 relative)
// accumulated in fp32 in TMEM.  The coordinate tolerance of the path (1e-4 on metres) rules out plain
// bf16/tf32 operands.
//
// Structure (sm_100a): PERSISTENT kernel, one CTA per SM looping over 128 x BN output tiles (BN = 32/64/128/256 chosen
// per problem).  warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B, mbarrier ring of {A_hi, A_lo, W_hi, W_lo}
// k-slabs of 64), warp 1 = MMA issuer (one elected thread, tcgen05.mma cta_group::1 kind::f16 M128 N=BN K16, TWO
// accumulator buffers in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1), warps 2-9 = epilogue
// (tcgen05.ld 32x32b -> smem transpose -> bias/GELU/residual -> coalesced fp32 C and/or bf16 hi/lo planes for the next GEMM).
// Batched form (blockIdx-free: the batch index is part of the tile index): per-batch A column offset (head slices),
// W row offset, C/bias element offsets.
#include "common.cuh"
#include "gemm_tc.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

namespace pt {

constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_A_TILE_BYTES = TC_BM * TC_BK * 2;            // 16 KiB, one bf16 A plane tile
constexpr int TC_EPI_WARPS = 8;                                // two per TMEM lane quarter, alternating 32-column chunks
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_EPI_SMEM = TC_EPI_WARPS * 32 * 32 * 4;         // per-warp 32x32 fp32 transpose buffer, 16-byte chunks XOR-swizzled by row
constexpr int TC_SMEM_BUDGET = 192 * 1024;

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO); LBO is unused for swizzled
// K-major layouts; descriptor version 1 (sm_100); layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t tc_idesc(int bn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void tc_tmem_ld32(uint32_t (&v)[32], uint32_t taddr) {       // asynchronous: pair with tc_tmem_wait_use32
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// wait for the outstanding TMEM loads; the "+r" list keeps every use of v behind the wait
__device__ __forceinline__ void tc_tmem_wait_use32(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// GELU(x) = x Phi(x) with erfc(|z|), z = x / sqrt(2), from the rational-exponential form of Abramowitz & Stegun 7.1.26
// (|error| <= 1.5e-7 on erfc): 16 instructions (2 MUFU) instead of the ~30 of erff — with 4 k-slabs per tile the fc1 epilogue was bound by
// its instruction count.  The negative side uses erfc directly (no 1 - erf cancellation); |GELU error| <= 0.75e-7 |x|, below the
// 2^-17 relative precision of the hi/lo planes the result is written to.
__device__ __forceinline__ float tc_gelu(float x) {
    const float az = fabsf(x) * 0.70710678118654752440f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752044448170368f * x * x));       // exp(-z^2) = 2^(-x^2 log2(e) / 2)
    const float erfc_az = p * t * e;
    return fmaf(-0.5f * fabsf(x), erfc_az, fmaxf(x, 0.f));       // x >= 0: x - (x/2) erfc(z);  x < 0: (x/2) erfc(|z|)
}

struct TcKernelArgs {
    int M, N, K, batch;
    int a_rows;          // rows of one A plane (lo plane starts at row a_rows of the tensor map)
    int a_koff_z;        // A column offset per batch index
    int w_rows;          // rows of one W plane
    int w_row_z;         // W row offset per batch index
    const float* bias; long long bias_off_z;
    const float* residual;                 // same layout as C (ldc, c_off_z); may alias C
    float* C; int ldc; long long c_off_z;
    __nv_bfloat16* Cs; long long cs_plane; int ldcs; long long cs_off_z;     // optional bf16 hi/lo planes of the result
    int cs_fp16; float cs_scale;                                             // ... as IEEE half planes of cs_scale * result
    __nv_bfloat16* Ct; int ct_col0; long long ct_ld, ct_plane;               // optional transposed planes for columns >= ct_col0
    int ct_seg, ct_seg_pad;                                                  // row segments (scenes) padded to ct_seg_pad columns
    int act;
    int tiles_m, tiles_n, stages;
    int dbg;             // -DPT_GEMM_DBG builds only (PT_GEMM_DEBUG): 1 no MMAs, 2 epilogue only hands the accumulator back, 4 epilogue without global memory, 8 no W loads
};
#ifdef PT_GEMM_DBG
#define TC_ZFAST (!(g.dbg & 16))
#else
#define TC_ZFAST true
#endif
#ifdef PT_GEMM_DBG
#define TC_DBG(bit) (g.dbg & (bit))
__device__ unsigned long long g_gemm_tl[160][8];      // per-CTA timeline (globaltimer ns), tools/gemm_timeline.py
__device__ __forceinline__ void tc_mark(int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (blockIdx.x < 160) g_gemm_tl[blockIdx.x][slot] = t;
}
#define TC_MARK(slot) tc_mark(slot)
#else
#define TC_MARK(slot) ((void)0)
#define TC_DBG(bit) false
#endif

template <int BN, int ACT, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmapA,
                                                                const __grid_constant__ CUtensorMap tmapW,
                                                                const TcKernelArgs g) {
    constexpr int W_TILE_BYTES = BN * TC_BK * 2;
    constexpr int STAGE_BYTES = 2 * TC_A_TILE_BYTES + 2 * W_TILE_BYTES;
    constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    constexpr uint32_t IDESC = tc_idesc(BN);
    constexpr int MAX_STAGES = 8;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stages = g.stages;
    uint64_t* full = (uint64_t*)(tiles + (size_t)stages * STAGE_BYTES);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tmem_full = empty + MAX_STAGES;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;         // [2]
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
    float* epi_stage = (float*)(tiles + (size_t)stages * STAGE_BYTES + 256);     // [TC_EPI_WARPS][32][32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = g.K / TC_BK;
    const int tiles_per_batch = g.tiles_m * g.tiles_n;
    const int total_tiles = tiles_per_batch * g.batch;

    if (threadIdx.x == 0) {
        TC_MARK(0);
        for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapW) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TC_MARK(1);

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int z = TC_ZFAST ? t % g.batch : t / tiles_per_batch, r = TC_ZFAST ? t / g.batch : t - z * tiles_per_batch;
                const int m0 = (r / g.tiles_n) * TC_BM, n0 = (r % g.tiles_n) * BN;
                const int ak = z * g.a_koff_z, wr = z * g.w_row_z + n0;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % stages, ph = (it / stages) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    uint8_t* st = tiles + (size_t)s * STAGE_BYTES;
                    mbar_expect_tx(full + s, TC_DBG(8) ? 2 * TC_A_TILE_BYTES : STAGE_BYTES);
                    tma_load_2d(st, &tmapA, ak + kb * TC_BK, m0, full + s);
                    tma_load_2d(st + TC_A_TILE_BYTES, &tmapA, ak + kb * TC_BK, g.a_rows + m0, full + s);
                    if (TC_DBG(8)) continue;
                    tma_load_2d(st + 2 * TC_A_TILE_BYTES, &tmapW, kb * TC_BK, wr, full + s);
                    tma_load_2d(st + 2 * TC_A_TILE_BYTES + W_TILE_BYTES, &tmapW, kb * TC_BK, g.w_rows + wr, full + s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int it = 0, tl = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
                const int as = tl & 1, aph = (tl >> 1) & 1;
                mbar_wait(tmem_empty + as, aph ^ 1);          // the epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % stages, ph = (it / stages) & 1;
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    if (it == 0) TC_MARK(2);
                    const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    const uint64_t da_hi = umma_desc_sw128(sa), da_lo = umma_desc_sw128(sa + TC_A_TILE_BYTES);
                    const uint64_t db_hi = umma_desc_sw128(sa + 2 * TC_A_TILE_BYTES), db_lo = umma_desc_sw128(sa + 2 * TC_A_TILE_BYTES + W_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        if (TC_DBG(1)) break;
                        const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);       // 32 B per K=16 step inside the 128 B swizzle atom
                        umma_f16(acc, da_hi + adv, db_hi + adv, IDESC, (kb | k) != 0 ? 1u : 0u);
                        umma_f16(acc, da_lo + adv, db_hi + adv, IDESC, 1u);
                        umma_f16(acc, da_hi + adv, db_lo + adv, IDESC, 1u);
                    }
                    umma_commit(empty + s);           // frees the smem stage once these MMAs have read it
                }
                umma_commit(tmem_full + as);          // accumulator complete
                if (tl == 0) TC_MARK(3);
            }
        }
    } else {
        // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); the two warps of a quarter take alternate
        // 32-column chunks.  Each chunk goes TMEM -> registers (lane = row) -> per-warp smem transpose -> registers
        // (lane = 4 consecutive columns of one of 4 rows) so that bias / residual / stores are fully coalesced
        // 512-byte warp accesses (the row-per-lane form wrote 32 partial lines per instruction).  The chunk loop is
        // software-pipelined — with 4 k-slabs per tile (K = 256) a tile's MMAs take ~3.7 us and a serial chunk chain
        // (TMEM load -> transpose -> residual load -> store, ~1500 cycles of latency each) made the epilogue the bound:
        // the residual rows of a chunk are requested before its TMEM load is waited for, the TMEM load of the next
        // chunk is issued as soon as the registers are free (after the transpose stores), and the accumulator buffer
        // goes back to the MMA warp right after the last TMEM load instead of after the last store.
        const int q = warp & 3, half = (warp - 2) >> 2;
        float* st = epi_stage + (size_t)(warp - 2) * 32 * 32;
        const int tr = lane >> 3, tc4 = (lane & 7) * 4;          // transposed role: row tr + 4*it, columns tc4..tc4+3
        constexpr int NCHUNK = BN / 32;
        int tl = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
            const int as = tl & 1, aph = (tl >> 1) & 1;
            const int z = TC_ZFAST ? t % g.batch : t / tiles_per_batch, r = TC_ZFAST ? t / g.batch : t - z * tiles_per_batch;
            const int m0 = (r / g.tiles_n) * TC_BM, n0 = (r % g.tiles_n) * BN;
            mbar_wait(tmem_full + as, aph);
            tc_fence_after();
            if (threadIdx.x == 64) { if (tl == 0) TC_MARK(4); TC_MARK(6); }
            const float* bias = g.bias ? g.bias + z * g.bias_off_z : nullptr;
            const int row0 = m0 + q * 32;
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            uint32_t v[32];
            if (half < NCHUNK && !TC_DBG(2)) tc_tmem_ld32(v, tacc + (uint32_t)(half * 32));
#pragma unroll 1
            for (int cc = half; cc < NCHUNK; cc += 2) {
                if (TC_DBG(2)) break;
                const bool last = cc + 2 >= NCHUNK;
                const bool transposed = g.Ct != nullptr && n0 + cc * 32 >= g.ct_col0;
                const int n = n0 + cc * 32 + tc4;
                const size_t coff0 = (size_t)z * g.c_off_z + (size_t)(row0 + tr) * g.ldc + n;      // row of iteration 0; + 4 * ldc per iteration
                float4 rr[8];
                const bool have_res = !transposed && g.residual != nullptr && n < g.N;
                if (have_res) {
                    const float* rp = g.residual + coff0;
                    const size_t rstep = (size_t)4 * g.ldc;
                    const int nit_r = (g.M - row0 - tr + 3) >> 2;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        rr[it] = it < nit_r ? *reinterpret_cast<const float4*>(rp) : make_float4(0.f, 0.f, 0.f, 0.f);
                        rp += rstep;
                    }
                }
                tc_tmem_wait_use32(v);
                if (last) {                                    // every TMEM read of this warp is done: hand the accumulator back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty + as);
                }
                if (transposed) {
                    // transposed columns: lane = row already, so a fixed register is a 64-byte run of one output row of C^T
                    // (the launcher checks that the transposed column range is a whole number of chunks)
                    const int row = row0 + lane;
                    if (row < g.M) {
                        const long long tcol = g.ct_seg > 0 ? (long long)(row / g.ct_seg) * g.ct_seg_pad + row % g.ct_seg : row;
                        __nv_bfloat16* dhi = g.Ct + (size_t)(n0 + cc * 32 - g.ct_col0) * g.ct_ld + tcol;
                        __nv_bfloat16* dlo = dhi + g.ct_plane;
                        const float4* b4p = reinterpret_cast<const float4*>(bias + n0 + cc * 32);
#pragma unroll
                        for (int j4 = 0; j4 < 32; j4 += 4) {
                            const float4 bq = bias != nullptr ? __ldg(b4p + (j4 >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float y = __uint_as_float(v[j4 + e]) + bb[e];
                                const __nv_bfloat16 hh = __float2bfloat16_rn(y);
                                *dhi = hh;
                                *dlo = __float2bfloat16_rn(y - __bfloat162float(hh));
                                dhi += g.ct_ld; dlo += g.ct_ld;
                            }
                        }
                    }
                    if (!last) tc_tmem_ld32(v, tacc + (uint32_t)((cc + 2) * 32));
                    continue;
                }
                __syncwarp();                              // previous chunk's transposed reads are done
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(st + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                if (!last) tc_tmem_ld32(v, tacc + (uint32_t)((cc + 2) * 32));
                if (TC_DBG(4)) continue;
                if (n < g.N) {                                 // N % 4 == 0: a float4 is either fully valid or fully out
                    // every step below is a straight run over the 8 row groups with one uniform branch around it: the per-element
                    // form (pointer tests, the output-format switch and 64-bit index products inside the row loop) cost ~40
                    // instructions per element, and the epilogue is bound by its instruction count.
                    const int nit = (g.M - row0 - tr + 3) >> 2;      // row groups of this lane that are inside M (may be <= 0 or > 8)
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
                    float y[8][4];
                    const float* stl = st + tr * 32;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {                 // row tr + 4 it: (row & 7) = (tr + 4 it) & 7
                        const float4 a4 = *reinterpret_cast<const float4*>(stl + it * 128 + (((lane & 7) ^ ((tr + 4 * it) & 7)) << 2));
                        y[it][0] = a4.x + b4.x; y[it][1] = a4.y + b4.y; y[it][2] = a4.z + b4.z; y[it][3] = a4.w + b4.w;
                        if (ACT == 1) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) y[it][e] = tc_gelu(y[it][e]);
                        }
                    }
                    if (have_res) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) { y[it][0] += rr[it].x; y[it][1] += rr[it].y; y[it][2] += rr[it].z; y[it][3] += rr[it].w; }
                    }
                    if (g.C != nullptr) {
                        float* cp = g.C + coff0;
                        const size_t cstep = (size_t)4 * g.ldc;
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            if (it < nit) *reinterpret_cast<float4*>(cp) = make_float4(y[it][0], y[it][1], y[it][2], y[it][3]);
                            cp += cstep;
                        }
                    }
                    if (SPLIT) {
                        __nv_bfloat16* sh = g.Cs + (size_t)z * g.cs_off_z + (size_t)(row0 + tr) * g.ldcs + n;
                        __nv_bfloat16* sl = sh + g.cs_plane;
                        const size_t sstep = (size_t)4 * g.ldcs;
                        if (g.cs_fp16) {
                            const float sc = g.cs_scale;
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const __half2 h01 = __floats2half2_rn(y[it][0] * sc, y[it][1] * sc), h23 = __floats2half2_rn(y[it][2] * sc, y[it][3] * sc);
                                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                                const __half2 l01 = __floats2half2_rn(y[it][0] * sc - f01.x, y[it][1] * sc - f01.y);
                                const __half2 l23 = __floats2half2_rn(y[it][2] * sc - f23.x, y[it][3] * sc - f23.y);
                                if (it < nit) {
                                    *reinterpret_cast<uint2*>(sh) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                                    *reinterpret_cast<uint2*>(sl) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                                }
                                sh += sstep; sl += sstep;
                            }
                        } else {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const __nv_bfloat162 h01 = __floats2bfloat162_rn(y[it][0], y[it][1]), h23 = __floats2bfloat162_rn(y[it][2], y[it][3]);
                                const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
                                // bf16 -> fp32 is a 16-bit shift: low half = element 0, high half = element 1
                                const __nv_bfloat162 l01 = __floats2bfloat162_rn(y[it][0] - __uint_as_float(u01 << 16), y[it][1] - __uint_as_float(u01 & 0xffff0000u));
                                const __nv_bfloat162 l23 = __floats2bfloat162_rn(y[it][2] - __uint_as_float(u23 << 16), y[it][3] - __uint_as_float(u23 & 0xffff0000u));
                                if (it < nit) {
                                    *reinterpret_cast<uint2*>(sh) = make_uint2(u01, u23);
                                    *reinterpret_cast<uint2*>(sl) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                                }
                                sh += sstep; sl += sstep;
                            }
                        }
                    }
                }
            }
            if (half >= NCHUNK || TC_DBG(2)) {                 // BN = 32: the second warp of a quarter has no chunk
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty + as);
            }
            if (threadIdx.x == 64 && tl == 0) TC_MARK(5);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TC_MARK(7);
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
__global__ void split_rows_bf16_kernel(const float* __restrict__ x, long long count, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ lo) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 3 < count) {
        const float4 v = *reinterpret_cast<const float4*>(x + i4);
        const float f[4] = {v.x, v.y, v.z, v.w};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { h[t] = __float2bfloat16_rn(f[t]); l[t] = __float2bfloat16_rn(f[t] - __bfloat162float(h[t])); }
        *reinterpret_cast<uint2*>(hi + i4) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lo + i4) = *reinterpret_cast<const uint2*>(l);
    } else {
        for (long long i = i4; i < count; ++i) {
            const __nv_bfloat16 h = __float2bfloat16_rn(x[i]);
            hi[i] = h;
            lo[i] = __float2bfloat16_rn(x[i] - __bfloat162float(h));
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 tensor map over [rows][cols] with row pitch ld (elements), box = 64 columns x box_rows rows, SWIZZLE_128B.
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    PT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld box_rows=%d)", (int)r, rows, cols, ld, box_rows);
    return PT_OK;
}

int encode_tensor_map_16bit(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                            const unsigned long long* strides_bytes, const unsigned* box, bool fp16) {
    EncodeTiledFn fn = encode_fn();
    PT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
    PT_REQUIRE(rank >= 1 && rank <= 5, "tensor map rank %d", rank);
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstride[i] = strides_bytes[i];
    CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                    const_cast<void*>(base), gdim, gstride, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rank %d, inner dim %llu, box %u)", (int)r, rank, dims[0], box[0]);
    return PT_OK;
}

static thread_local int t_gemm_prof_tag = PROF_GEMM_TC;
GemmProfTagScope::GemmProfTagScope(int tag) : saved(t_gemm_prof_tag) { t_gemm_prof_tag = tag; }
GemmProfTagScope::~GemmProfTagScope() { t_gemm_prof_tag = saved; }

bool gemm_tc_supported(int M, int N, int K) { return M >= 1 && N >= 4 && N % 4 == 0 && K % TC_BK == 0 && K >= TC_BK; }

size_t gemm_tc_ws_bytes(int M, int N, int K) {
    (void)N;
    return align_up((size_t)2 * M * K * sizeof(__nv_bfloat16), 256) + 256;
}

int split_rows_bf16(const float* x, long long count, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t s) {
    { ProfScope prof_(PROF_SPLIT, s); split_rows_bf16_kernel<<<(unsigned)((count / 4 + 255) / 256 + 1), 256, 0, s>>>(x, count, hi, lo); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

static int num_sms() {
    static int n[PT_MAX_DEVICES] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PT_MAX_DEVICES) return 148;
    if (n[dev] == 0 && (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0)) n[dev] = 148;
    return n[dev];
}

template <int BN, int ACT, bool SPLIT>
static int launch_variant(const CUtensorMap& mapA, const CUtensorMap& mapW, TcKernelArgs& k, cudaStream_t s) {
    constexpr int STAGE_BYTES = 2 * TC_A_TILE_BYTES + 2 * BN * TC_BK * 2;
    int stages = TC_SMEM_BUDGET / STAGE_BYTES;
    if (stages > 6) stages = 6;
    const int smem = stages * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + TC_EPI_SMEM;
    static bool attr_set[PT_MAX_DEVICES] = {};
    if (first_use_on_current_device(attr_set))
        PT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, ACT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET + 2048 + TC_EPI_SMEM));
    k.stages = stages;
    k.tiles_n = ceil_div(k.N, BN);
    const int total = k.tiles_m * k.tiles_n * k.batch;
    int grid = total < num_sms() ? total : num_sms();
#ifdef PT_GEMM_DBG
    k.dbg = getenv("PT_GEMM_DEBUG") ? atoi(getenv("PT_GEMM_DEBUG")) : 0;
    if (getenv("PT_GEMM_GRID") && atoi(getenv("PT_GEMM_GRID")) > 0 && atoi(getenv("PT_GEMM_GRID")) < grid) grid = atoi(getenv("PT_GEMM_GRID"));
    if (getenv("PT_GEMM_STAGES") && atoi(getenv("PT_GEMM_STAGES")) > 0 && atoi(getenv("PT_GEMM_STAGES")) < stages) k.stages = atoi(getenv("PT_GEMM_STAGES"));
#else
    k.dbg = 0;
#endif
    { ProfScope prof_(t_gemm_prof_tag, s); gemm_tc_kernel<BN, ACT, SPLIT><<<grid, TC_THREADS, smem, s>>>(mapA, mapW, k); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

template <int BN>
static int launch_bn(const CUtensorMap& mapA, const CUtensorMap& mapW, TcKernelArgs& k, cudaStream_t s) {
    const bool split = k.Cs != nullptr;
    if (k.act == 1) return split ? launch_variant<BN, 1, true>(mapA, mapW, k, s) : launch_variant<BN, 1, false>(mapA, mapW, k, s);
    return split ? launch_variant<BN, 0, true>(mapA, mapW, k, s) : launch_variant<BN, 0, false>(mapA, mapW, k, s);
}

int launch_gemm_tc_ex(const GemmTc& p, cudaStream_t s) {
    PT_REQUIRE(gemm_tc_supported(p.M, p.N, p.K) && p.batch >= 1, "gemm_tc: M=%d N=%d K=%d batch=%d unsupported", p.M, p.N, p.K, p.batch);
    PT_REQUIRE(p.a_split && p.w_split && (p.C || p.c_split), "gemm_tc: null operand");
    PT_REQUIRE(!p.ct_split || p.ct_seg == 0 || (p.ct_seg > 0 && p.ct_seg_pad >= p.ct_seg && p.ct_ld >= (long long)ceil_div(p.M, p.ct_seg) * p.ct_seg_pad),
               "gemm_tc: transposed output segments: ct_seg=%d ct_seg_pad=%d ct_ld=%lld", p.ct_seg, p.ct_seg_pad, p.ct_ld);
    PT_REQUIRE(!p.ct_split || (p.ct_col0 % 32 == 0 && p.ct_col0 >= 0 && (p.N - p.ct_col0) % 32 == 0 && p.ct_ld >= p.M && p.batch == 1 && !p.act &&
                               (!p.bias || ((uintptr_t)p.bias & 15) == 0)),
               "gemm_tc: transposed output needs ct_col0 %% 32 == 0, (N - ct_col0) %% 32 == 0, ct_ld >= M, batch 1, no activation, 16-byte aligned bias");
    PT_REQUIRE(((uintptr_t)p.a_split & 15) == 0 && ((uintptr_t)p.w_split & 15) == 0 && (p.lda % 8) == 0 && (p.ldw % 8) == 0,
               "gemm_tc: operand planes must be 16-byte aligned with pitches that are multiples of 8");
    PT_REQUIRE(!p.C || (((uintptr_t)p.C & 15) == 0 && p.ldc % 4 == 0 && p.c_off_z % 4 == 0), "gemm_tc: C alignment");
    PT_REQUIRE(!p.residual || ((uintptr_t)p.residual & 15) == 0, "gemm_tc: residual alignment");
    PT_REQUIRE(!p.c_split || (((uintptr_t)p.c_split & 7) == 0 && p.ldcs % 4 == 0 && p.cs_off_z % 4 == 0 && p.cs_plane % 4 == 0), "gemm_tc: split output alignment");
    CUtensorMap mapA, mapW;
    int rc;
    if ((rc = make_map(&mapA, p.a_split, 2LL * p.a_rows, p.a_cols, p.lda, TC_BM))) return rc;
    int bn = p.bn;
    if (bn == 0) {      // widest tile that still leaves ~a wave of tiles; 256 halves the A re-reads of the big layers.  With GELU the
                        // epilogue is the longer phase of a K = 256 tile: narrower tiles leave a shorter exposed epilogue at the end
                        // (fc1 of the C2 block: 37.5 us at 128 against 41.3 us at 256)
        const long long tiles128 = (long long)ceil_div(p.M, TC_BM) * ceil_div(p.N, 128) * p.batch;
        bn = p.N <= 32 ? 32 : p.N <= 64 ? 64 : (p.N % 256 == 0 && tiles128 >= 2LL * num_sms() && !p.act) ? 256 : 128;
#ifdef PT_GEMM_DBG
        if (p.act && getenv("PT_GEMM_ACT_BN")) bn = atoi(getenv("PT_GEMM_ACT_BN"));
        if (!p.act && p.N == p.K * 3 && getenv("PT_GEMM_QKV_BN")) bn = atoi(getenv("PT_GEMM_QKV_BN"));
#endif
    }
    PT_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 256, "gemm_tc: bn=%d", bn);
    if ((rc = make_map(&mapW, p.w_split, 2LL * p.w_rows, p.K, p.ldw, bn))) return rc;
    TcKernelArgs k;
    k.M = p.M; k.N = p.N; k.K = p.K; k.batch = p.batch;
    k.a_rows = p.a_rows; k.a_koff_z = p.a_koff_z; k.w_rows = p.w_rows; k.w_row_z = p.w_row_z;
    k.bias = p.bias; k.bias_off_z = p.bias_off_z; k.residual = p.residual;
    k.C = p.C; k.ldc = p.ldc; k.c_off_z = p.c_off_z;
    k.Cs = (__nv_bfloat16*)p.c_split; k.cs_plane = p.cs_plane; k.ldcs = p.ldcs; k.cs_off_z = p.cs_off_z;
    k.cs_fp16 = p.cs_fp16 ? 1 : 0; k.cs_scale = p.cs_scale;
    k.act = p.act;
    k.Ct = (__nv_bfloat16*)p.ct_split; k.ct_col0 = p.ct_col0; k.ct_ld = p.ct_ld; k.ct_plane = p.ct_plane;
    k.ct_seg = p.ct_seg; k.ct_seg_pad = p.ct_seg_pad;
    k.tiles_m = ceil_div(p.M, TC_BM);
    switch (bn) {
        case 32: return launch_bn<32>(mapA, mapW, k, s);
        case 64: return launch_bn<64>(mapA, mapW, k, s);
        case 128: return launch_bn<128>(mapA, mapW, k, s);
        default: return launch_bn<256>(mapA, mapW, k, s);
    }
}

// Plain form used by the ProxyBlock stage: A fp32 (split on the fly into ws), W pre-split [2][N][K], dense C.
int launch_gemm_tc(const float* A, const void* w_split, const float* bias, const float* residual, int act, int M, int N,
                   int K, float* C, void* ws, size_t ws_bytes, cudaStream_t s) {
    PT_REQUIRE(gemm_tc_supported(M, N, K), "gemm_tc: M=%d N=%d K=%d unsupported", M, N, K);
    PT_REQUIRE(ws != nullptr, "gemm_tc: workspace required");
    if (ws_bytes < gemm_tc_ws_bytes(M, N, K)) { set_error("gemm_tc: workspace %zu < %zu", ws_bytes, gemm_tc_ws_bytes(M, N, K)); return PT_ERR_WORKSPACE; }
    PT_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)ws & 15) == 0, "gemm_tc: pointers must be 16-byte aligned");
    __nv_bfloat16* a_split = (__nv_bfloat16*)ws;
    const long long count = (long long)M * K;
    int rc;
    if ((rc = split_rows_bf16(A, count, a_split, a_split + count, s))) return rc;
    GemmTc p;
    p.M = M; p.N = N; p.K = K;
    p.a_split = a_split; p.a_rows = M; p.a_cols = K; p.lda = K;
    p.w_split = w_split; p.w_rows = N; p.ldw = K;
    p.bias = bias; p.residual = residual; p.act = act;
    p.C = C; p.ldc = N;
    return launch_gemm_tc_ex(p, s);
}

}  // namespace pt

#ifdef PT_GEMM_DBG
extern "C" int pt_debug_gemm_timeline(unsigned long long* out) {      // [160][8] globaltimer ns of the last GEMM launch
    return cudaMemcpyFromSymbol(out, pt::g_gemm_tl, sizeof(pt::g_gemm_tl)) == cudaSuccess ? 0 : 1;
}
#endif
