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

__device__ __forceinline__ float tc_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

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
};

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

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int z = t / tiles_per_batch, r = t - z * tiles_per_batch;
                const int m0 = (r / g.tiles_n) * TC_BM, n0 = (r % g.tiles_n) * BN;
                const int ak = z * g.a_koff_z, wr = z * g.w_row_z + n0;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % stages, ph = (it / stages) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    uint8_t* st = tiles + (size_t)s * STAGE_BYTES;
                    mbar_expect_tx(full + s, STAGE_BYTES);
                    tma_load_2d(st, &tmapA, ak + kb * TC_BK, m0, full + s);
                    tma_load_2d(st + TC_A_TILE_BYTES, &tmapA, ak + kb * TC_BK, g.a_rows + m0, full + s);
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
                    const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    const uint64_t da_hi = umma_desc_sw128(sa), da_lo = umma_desc_sw128(sa + TC_A_TILE_BYTES);
                    const uint64_t db_hi = umma_desc_sw128(sa + 2 * TC_A_TILE_BYTES), db_lo = umma_desc_sw128(sa + 2 * TC_A_TILE_BYTES + W_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);       // 32 B per K=16 step inside the 128 B swizzle atom
                        umma_f16(acc, da_hi + adv, db_hi + adv, IDESC, (kb | k) != 0 ? 1u : 0u);
                        umma_f16(acc, da_lo + adv, db_hi + adv, IDESC, 1u);
                        umma_f16(acc, da_hi + adv, db_lo + adv, IDESC, 1u);
                    }
                    umma_commit(empty + s);           // frees the smem stage once these MMAs have read it
                }
                umma_commit(tmem_full + as);          // accumulator complete
            }
        }
    } else {
        // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); the two warps of a quarter take alternate
        // 32-column chunks.  Each chunk goes TMEM -> registers (lane = row) -> per-warp smem transpose -> registers
        // (lane = 4 consecutive columns of one of 4 rows) so that bias / residual / stores are fully coalesced
        // 512-byte warp accesses (the row-per-lane form wrote 32 partial lines per instruction).
        const int q = warp & 3, half = (warp - 2) >> 2;
        float* st = epi_stage + (size_t)(warp - 2) * 32 * 32;
        const int tr = lane >> 3, tc4 = (lane & 7) * 4;          // transposed role: row tr + 4*it, columns tc4..tc4+3
        int tl = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
            const int as = tl & 1, aph = (tl >> 1) & 1;
            const int z = t / tiles_per_batch, r = t - z * tiles_per_batch;
            const int m0 = (r / g.tiles_n) * TC_BM, n0 = (r % g.tiles_n) * BN;
            mbar_wait(tmem_full + as, aph);
            tc_fence_after();
            const float* bias = g.bias ? g.bias + z * g.bias_off_z : nullptr;
            const int row0 = m0 + q * 32;
#pragma unroll 1
            for (int cc = half; cc < BN / 32; cc += 2) {
                if (g.Ct != nullptr && n0 + cc * 32 >= g.ct_col0) {
                    // transposed columns: lane = row already, so a fixed register is a 64-byte run of one output row of C^T
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + cc * 32);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const int row = row0 + lane;
                    if (row < g.M) {
                        const long long tcol = g.ct_seg > 0 ? (long long)(row / g.ct_seg) * g.ct_seg_pad + row % g.ct_seg : row;
                        __nv_bfloat16* dst = g.Ct + (size_t)(n0 + cc * 32 - g.ct_col0) * g.ct_ld + tcol;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (n0 + cc * 32 + j < g.N) {
                                const float y = __uint_as_float(v[j]) + (bias != nullptr ? __ldg(bias + n0 + cc * 32 + j) : 0.f);
                                const __nv_bfloat16 hh = __float2bfloat16_rn(y);
                                dst[(size_t)j * g.ct_ld] = hh;
                                dst[(size_t)j * g.ct_ld + g.ct_plane] = __float2bfloat16_rn(y - __bfloat162float(hh));
                            }
                        }
                    }
                    continue;
                }
                {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + cc * 32);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    __syncwarp();                              // previous chunk's transposed reads are done
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(st + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    __syncwarp();
                }
                const int n = n0 + cc * 32 + tc4;
                if (n < g.N) {                                 // N % 4 == 0: a float4 is either fully valid or fully out
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rl = it * 4 + tr, row = row0 + rl;
                        if (row >= g.M) break;
                        const float4 a4 = *reinterpret_cast<const float4*>(st + rl * 32 + ((((lane & 7)) ^ (rl & 7)) << 2));
                        float y[4] = {a4.x + b4.x, a4.y + b4.y, a4.z + b4.z, a4.w + b4.w};
                        if (ACT == 1) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) y[e] = tc_gelu(y[e]);
                        }
                        const size_t coff = (size_t)z * g.c_off_z + (size_t)row * g.ldc + n;
                        if (g.residual != nullptr) {
                            const float4 rr = *reinterpret_cast<const float4*>(g.residual + coff);
                            y[0] += rr.x; y[1] += rr.y; y[2] += rr.z; y[3] += rr.w;
                        }
                        if (g.C != nullptr) *reinterpret_cast<float4*>(g.C + coff) = make_float4(y[0], y[1], y[2], y[3]);
                        if (SPLIT) {
                            __nv_bfloat16* shi = g.Cs + (size_t)z * g.cs_off_z + (size_t)row * g.ldcs + n;
                            if (g.cs_fp16) {
                                __half h[4], l[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) { const float v = y[e] * g.cs_scale; h[e] = __float2half_rn(v); l[e] = __float2half_rn(v - __half2float(h[e])); }
                                *reinterpret_cast<uint2*>(shi) = *reinterpret_cast<const uint2*>(h);
                                *reinterpret_cast<uint2*>(shi + g.cs_plane) = *reinterpret_cast<const uint2*>(l);
                            } else {
                                __nv_bfloat16 h[4], l[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) { h[e] = __float2bfloat16_rn(y[e]); l[e] = __float2bfloat16_rn(y[e] - __bfloat162float(h[e])); }
                                *reinterpret_cast<uint2*>(shi) = *reinterpret_cast<const uint2*>(h);
                                *reinterpret_cast<uint2*>(shi + g.cs_plane) = *reinterpret_cast<const uint2*>(l);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + as);
        }
    }
    tc_fence_before();
    __syncthreads();
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
    const int grid = total < num_sms() ? total : num_sms();
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
    PT_REQUIRE(!p.ct_split || (p.ct_col0 % 32 == 0 && p.ct_col0 >= 0 && p.ct_ld >= p.M && p.batch == 1 && !p.act),
               "gemm_tc: transposed output needs ct_col0 %% 32 == 0, ct_ld >= M, batch 1, no activation");
    PT_REQUIRE(((uintptr_t)p.a_split & 15) == 0 && ((uintptr_t)p.w_split & 15) == 0 && (p.lda % 8) == 0 && (p.ldw % 8) == 0,
               "gemm_tc: operand planes must be 16-byte aligned with pitches that are multiples of 8");
    PT_REQUIRE(!p.C || (((uintptr_t)p.C & 15) == 0 && p.ldc % 4 == 0 && p.c_off_z % 4 == 0), "gemm_tc: C alignment");
    PT_REQUIRE(!p.residual || ((uintptr_t)p.residual & 15) == 0, "gemm_tc: residual alignment");
    PT_REQUIRE(!p.c_split || (((uintptr_t)p.c_split & 7) == 0 && p.ldcs % 4 == 0 && p.cs_off_z % 4 == 0 && p.cs_plane % 4 == 0), "gemm_tc: split output alignment");
    CUtensorMap mapA, mapW;
    int rc;
    if ((rc = make_map(&mapA, p.a_split, 2LL * p.a_rows, p.a_cols, p.lda, TC_BM))) return rc;
    int bn = p.bn;
    if (bn == 0) {      // widest tile that still leaves ~a wave of tiles; 256 halves the A re-reads of the big layers
        const long long tiles128 = (long long)ceil_div(p.M, TC_BM) * ceil_div(p.N, 128) * p.batch;
        bn = p.N <= 32 ? 32 : p.N <= 64 ? 64 : (p.N % 256 == 0 && tiles128 >= 2LL * num_sms()) ? 256 : 128;
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
