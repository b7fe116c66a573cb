// Tensor-core GEMM for the ProxyBlock dense layers (S7): C[M,N] = act(A[M,K] W[N,K]^T + bias) + residual with
// fp32-ACCURATE results from bf16 tensor cores via 3xBF16 operand splitting:
//     a = a_hi + a_lo,  w = w_hi + w_lo   (hi = bf16(x), lo = bf16(x - hi))
//     a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo        (the dropped lo*lo term is ~2^-18 relative)
// accumulated in fp32 in TMEM.  The coordinate tolerance of the path (1e-4 on metres) rules out plain
// bf16/tf32 operands; three tcgen05.mma per k-step still run ~10x faster than the fp32 CUDA-core GEMM.
//
// Structure (sm_100a): one CTA per 128x128 output tile; warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B,
// 3-stage mbarrier ring of {A_hi, A_lo, W_hi, W_lo} 128x64 bf16 tiles), warp 1 = MMA issuer (one elected thread,
// tcgen05.mma cta_group::1 kind::f16 M128 N128 K16, accumulator in 128 TMEM columns), warps 2-5 = epilogue
// (tcgen05.ld 32x32b -> bias/GELU/residual -> global).  W is split once per weight load by the host module, A is split
// by a small pre-pass into the caller's workspace.
#include "common.cuh"

#include <cuda.h>
#include <math.h>

namespace pt {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 2;              // 16 KiB, one bf16 operand tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;             // A_hi, A_lo, W_hi, W_lo
constexpr int TC_THREADS = 192;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TC_TMEM_COLS = 128;

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ float tc_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// tmapA: 2-D bf16 [2M rows][K], hi plane rows [0,M), lo plane rows [M,2M); tmapW: [2N rows][K] likewise.
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmapA,
                                                                const __grid_constant__ CUtensorMap tmapW,
                                                                const float* __restrict__ bias, const float* residual, int act,
                                                                int M, int N, int K, float* C) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(tiles + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
    const int nkb = K / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapW) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % TC_STAGES, it = kb / TC_STAGES;
                mbar_wait(empty + s, (it & 1) ^ 1);
                uint8_t* st = tiles + s * TC_STAGE_BYTES;
                mbar_expect_tx(full + s, TC_STAGE_BYTES);
                tma_load_2d(st, &tmapA, kb * TC_BK, m0, full + s);
                tma_load_2d(st + TC_TILE_BYTES, &tmapA, kb * TC_BK, M + m0, full + s);
                tma_load_2d(st + 2 * TC_TILE_BYTES, &tmapW, kb * TC_BK, n0, full + s);
                tma_load_2d(st + 3 * TC_TILE_BYTES, &tmapW, kb * TC_BK, N + n0, full + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % TC_STAGES, it = kb / TC_STAGES;
                mbar_wait(full + s, it & 1);
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + s * TC_STAGE_BYTES);
                const uint64_t da_hi = umma_desc_sw128(sa), da_lo = umma_desc_sw128(sa + TC_TILE_BYTES);
                const uint64_t db_hi = umma_desc_sw128(sa + 2 * TC_TILE_BYTES), db_lo = umma_desc_sw128(sa + 3 * TC_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);       // 32 B per K=16 step inside the 128 B swizzle atom
                    umma_f16(tmem_base, da_hi + adv, db_hi + adv, TC_IDESC, (kb | k) != 0 ? 1u : 0u);
                    umma_f16(tmem_base, da_lo + adv, db_hi + adv, TC_IDESC, 1u);
                    umma_f16(tmem_base, da_hi + adv, db_lo + adv, TC_IDESC, 1u);
                }
                umma_commit(empty + s);           // frees the smem stage once these MMAs have read it
            }
            umma_commit(tmem_full);               // accumulator complete
        }
    } else {
        // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int cc = 0; cc < TC_BN / 32; ++cc) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < M) {
                const int n = n0 + cc * 32;
                float* crow = C + (size_t)row * N + n;
                const float* rrow = residual ? residual + (size_t)row * N + n : nullptr;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float y[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        y[t] = __uint_as_float(v[j + t]);
                        if (bias) y[t] += __ldg(bias + n + j + t);
                        if (act == 1) y[t] = tc_gelu(y[t]);
                    }
                    if (rrow) {
                        const float4 r = *reinterpret_cast<const float4*>(rrow + j);
                        y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
                    }
                    *reinterpret_cast<float4*>(crow + j) = make_float4(y[0], y[1], y[2], y[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
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

static int make_map(CUtensorMap* map, const void* base, long long rows, int K, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    PT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return PT_OK;
}

bool gemm_tc_supported(int M, int N, int K) { return M >= 1 && N % TC_BN == 0 && K % TC_BK == 0 && K >= TC_BK; }

size_t gemm_tc_ws_bytes(int M, int N, int K) {
    (void)N;
    return align_up((size_t)2 * M * K * sizeof(__nv_bfloat16), 256) + 256;
}

int launch_gemm_tc(const float* A, const void* w_split, const float* bias, const float* residual, int act, int M, int N,
                   int K, float* C, void* ws, size_t ws_bytes, cudaStream_t s) {
    PT_REQUIRE(gemm_tc_supported(M, N, K), "gemm_tc: M=%d N=%d K=%d unsupported", M, N, K);
    PT_REQUIRE(ws != nullptr, "gemm_tc: workspace required");
    if (ws_bytes < gemm_tc_ws_bytes(M, N, K)) { set_error("gemm_tc: workspace %zu < %zu", ws_bytes, gemm_tc_ws_bytes(M, N, K)); return PT_ERR_WORKSPACE; }
    PT_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)C & 15) == 0 && ((uintptr_t)ws & 15) == 0 && ((uintptr_t)w_split & 15) == 0 &&
                   (residual == nullptr || ((uintptr_t)residual & 15) == 0),
               "gemm_tc: pointers must be 16-byte aligned");
    __nv_bfloat16* a_split = (__nv_bfloat16*)ws;
    const long long count = (long long)M * K;
    { ProfScope prof_(PROF_SPLIT, s); split_rows_bf16_kernel<<<(unsigned)((count / 4 + 255) / 256 + 1), 256, 0, s>>>(A, count, a_split, a_split + count); }
    PT_LAUNCH_CHECK();
    CUtensorMap mapA, mapW;
    int rc;
    if ((rc = make_map(&mapA, a_split, 2LL * M, K, TC_BM))) return rc;
    if ((rc = make_map(&mapW, w_split, 2LL * N, K, TC_BN))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        PT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_set = true;
    }
    { ProfScope prof_(PROF_GEMM_TC, s); gemm_tc_kernel<<<dim3(N / TC_BN, ceil_div(M, TC_BM)), TC_THREADS, TC_SMEM_BYTES, s>>>(mapA, mapW, bias, residual, act, M, N, K, C); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

}  // namespace pt
