// tcgen05 / TMEM form of the two-stage proxy attention core of ProxyAttention.forward (:225-252), one CTA per (scene, head):
//   stage 1 (proxy as query, :232-238):  Pv = softmax_n((Pt*scale) K^T) V          (l x hd), unmasked
//   stage 2 (proxy as key,   :241-250):  O  = softmax_l(mask((Q*scale) Pt^T)) Pv   (n x hd)
// for heads of 32 channels, n <= 1024 point proxies (any n: the benchmark's 256, the shipped config's 691) and l <= 256 text / image
// proxies.  The clusters are streamed: stage 1 walks the keys in tiles of 256 (128 in the 8-warp form, see the kernel) with an online
// softmax (running maximum / sum per proxy row, accumulator rescaled in TMEM), stage 2 walks the cluster rows in tiles of 128; only
// one K / V^T tile and one Q tile are resident at a time.  (Other head sizes / l > 256: the mma.sync kernel in attn_mma.cu.)
//
// Every contraction is a tcgen05.mma (cta_group::1, kind::f16, M = 128) with fp32 accumulation in TMEM and 3xBF16 operand
// splitting (hi*hi + lo*hi + hi*lo), so scores and outputs keep ~2^-17 relative accuracy:
//   * Q, K, Pt arrive as bf16 hi/lo planes straight from the projection GEMMs and are staged as K-major SWIZZLE_128B
//     tiles whose 128-byte rows are [hi 32 | lo 32]: the three products are descriptor offsets into the same tile;
//   * the score tile S (128 rows x <= 256 keys, fp32) lives in TMEM; four threads per row (four warps per lane quarter,
//     interleaved 32-column chunks) reads it with tcgen05.ld, applies scale / mask / softmax and writes the probabilities
//     back IN PLACE as packed bf16 hi and lo halves (tcgen05.st), where the second MMA reads them as its TMEM A operand —
//     the probabilities never touch shared memory;
//   * V^T (from a transposed projection GEMM) and Pv^T (written by the stage-1 epilogue) are the K-major B operands of
//     the value contractions; the 1/rowsum normalisation is applied to the 32-column accumulator in the epilogue.
// Cost is O(n*l*hd), never n^2.
#include "common.cuh"

#include <math.h>

namespace pt {

namespace at {
constexpr int HD = 32, MAXR = 256;                    // MAXR: proxies (resident); NWQ warps per TMEM lane quarter take every NWQ-th 32-column chunk
constexpr int MAXN = 1024;                            // clusters (streamed in key tiles of MAXR and row tiles of 128)
constexpr int ROW_BYTES = 128;                        // [hi 32 | lo 32] bf16
constexpr int TILE_BYTES = MAXR * ROW_BYTES;          // 32768: Q / K / Pt operand tiles (256 rows)
constexpr int VT_TILE = HD * ROW_BYTES;               // 4096: one 64-key k-tile of V^T / Pv^T (32 rows)
constexpr int VT_PLANE = 4 * VT_TILE;                 // 16384: 256 keys
// K (stage 1) and Q (stage 2) share a tile, and so do V^T (stage 1) and Pv^T (written by the stage-1 epilogue, read by stage 2):
// 104 KB per CTA, so two CTAs fit on an SM
constexpr int OFF_K = 0, OFF_Q = OFF_K, OFF_P = OFF_K + TILE_BYTES;
constexpr int OFF_VT = OFF_P + TILE_BYTES;            // hi plane, then lo plane
constexpr int OFF_PV = OFF_VT;
constexpr int OFF_MISC = OFF_VT + 2 * VT_PLANE;       // rowmax [4][128], rowsum [4][128], keyflag [256], runm [256], runl [256] floats
constexpr int OFF_BAR = OFF_MISC + (8 * 128 + 3 * 256) * 4;
constexpr int SMEM_BYTES = OFF_BAR + 16 + 1024;       // + alignment slack
}  // namespace at

__device__ __forceinline__ uint32_t at_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t at_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t at_idesc(int n) {     // D=f32, A=B=bf16, K-major, M=128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void at_mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void at_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void at_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(at_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void at_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(at_smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void at_ld32(uint32_t (&v)[32], uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void at_ld16(uint32_t (&v)[16], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void at_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                 "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void at_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void at_ld8(uint32_t (&v)[8], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void at_ldn(uint32_t (&v)[8], uint32_t taddr) { at_ld8(v, taddr); }
__device__ __forceinline__ void at_ldn(uint32_t (&v)[16], uint32_t taddr) { at_ld16(v, taddr); }
__device__ __forceinline__ void at_stn(uint32_t taddr, const uint32_t (&r)[8]) { at_st8(taddr, r); }
__device__ __forceinline__ void at_stn(uint32_t taddr, const uint32_t (&r)[16]) { at_st16(taddr, r); }
__device__ __forceinline__ float at_ex2(float x) {            // 2^x, ~2 ulp; x <= 0 here
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// (x0, x1) -> packed bf16x2 hi word (x0 in the low half) and lo word, with one packed conversion each
__device__ __forceinline__ void at_split2_fast(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void at_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

struct AtArgs {
    const __nv_bfloat16* qk;      // hi plane of [rows][ldq]: Q at column 0, K at column c; lo plane qk_plane elements later
    long long qk_plane;
    int ldq;
    const __nv_bfloat16* vt;      // hi plane of V^T [c][ldv] (column = global row b*n + j); lo plane vt_plane later
    long long vt_plane;
    long long ldv;
    int vt_seg;                   // columns per scene in V^T (>= n; a multiple of 8 keeps every scene on the 16-byte grid)
    const __nv_bfloat16* pt;      // hi plane of Pt [B*l][c]; lo plane pt_plane later
    long long pt_plane;
    const uint8_t* mask;          // (B,l) 1 = real token, or null
    int n, l, c;
    float scale;
    float* o;                     // optional fp32 (B*n, c)
    __nv_bfloat16* o_hi;          // optional bf16 hi plane (B*n, c); lo plane o_plane later
    long long o_plane;
};

// One score tile: scale / mask / softmax of the 128 rows in TMEM columns [0, ncols), probabilities written back in place
// (chunk j of 32 keys -> 16 hi columns, 16 lo columns).  nvalid = real keys; flag (smem) != 0 marks masked keys (-1e9).
// Each of the 4 warps of a lane quarter (hh) takes every 4th chunk; the row maximum and the row sum meet in shared memory.
// softmax(scale*s) is evaluated as 2^(s*c1 - max*c1) with c1 = scale*log2(e): one FFMA and one MUFU per element.
// `m_old` (scaled running maximum of the row over the earlier key tiles, -inf for the first) makes it the step of an online softmax:
// the probabilities are relative to max(m_old, tile maximum), which is returned for the caller's bookkeeping.
template <int NWQ>
__device__ __forceinline__ float at_softmax_tile(uint32_t tmem_row, int ncols, int nvalid, const float* __restrict__ flag, float scale,
                                                 int hh, float* __restrict__ rowmax, float* __restrict__ rowsum, int row,
                                                 float m_old = -INFINITY) {
    const int nch = (ncols + 31) >> 5;
    const float c1 = scale * 1.4426950408889634f;
    float mx = -INFINITY;                                      // maximum of the UNSCALED scores of the real, unmasked keys
    bool any_masked = false;
    for (int ch = hh; ch < nch; ch += NWQ) {
        uint32_t v[32];
        at_ld32(v, tmem_row + 32 * ch);
        const int kv = nvalid - 32 * ch;                       // real keys in this chunk (>= 32: all)
        if (flag == nullptr && kv >= 32) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
        } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                if (e < kv) {
                    if (flag != nullptr && flag[32 * ch + e] != 0.f) any_masked = true;
                    else mx = fmaxf(mx, __uint_as_float(v[e]));
                }
            }
        }
    }
    // scaled maximum; a masked key contributes the literal -1e9 of masked_fill (:247)
    float ms = mx * scale;
    if (any_masked) ms = fmaxf(ms, -1e9f);
    rowmax[hh * 128 + row] = ms;
    __syncthreads();
    ms = fmaxf(rowmax[row], rowmax[128 + row]);                // finite: key 0 is a real key
    if (NWQ == 4) ms = fmaxf(ms, fmaxf(rowmax[256 + row], rowmax[384 + row]));
    ms = fmaxf(ms, m_old);
    const float mc = ms * 1.4426950408889634f;
    const float pmask = at_ex2(-1e9f * 1.4426950408889634f - mc);      // probability weight of a masked key (1 if every key is masked)
    float sum = 0.f;
    for (int ch = hh; ch < nch; ch += NWQ) {
        uint32_t v[32];
        at_ld32(v, tmem_row + 32 * ch);
        uint32_t ph[16], pl[16];
        const int kv = nvalid - 32 * ch;
        if (flag == nullptr && kv >= 32) {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                const float p0 = at_ex2(fmaf(__uint_as_float(v[e]), c1, -mc)), p1 = at_ex2(fmaf(__uint_as_float(v[e + 1]), c1, -mc));
                sum += p0 + p1;
                at_split2_fast(p0, p1, ph[e >> 1], pl[e >> 1]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                float p[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int key = 32 * ch + e + u;
                    p[u] = 0.f;
                    if (e + u < kv) p[u] = (flag != nullptr && flag[key] != 0.f) ? pmask : at_ex2(fmaf(__uint_as_float(v[e + u]), c1, -mc));
                    sum += p[u];
                }
                at_split2_fast(p[0], p[1], ph[e >> 1], pl[e >> 1]);
            }
        }
        at_st16(tmem_row + 32 * ch, ph);
        at_st16(tmem_row + 32 * ch + 16, pl);
    }
    rowsum[hh * 128 + row] = sum;
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    return ms;
}

// NWQ = 4, KT = 256: 16 warps, 256-key tiles, 512 TMEM columns (scores 256 | accumulators), one CTA per SM — any l <= 256.
// NWQ = 2, KT = 128: 8 warps, 128-key tiles, 256 TMEM columns (stage 1: scores 128 | accumulators 2 x 32; stage 2: scores lpad <= 224 |
// accumulator at 224), TWO CTAs per SM.  A CTA is one serial chain (stage tiles -> MMA -> softmax -> MMA -> epilogue, a block-wide
// barrier between the links; ncu: 25 % of the warp slots active, long-scoreboard and barrier stalls on top): two independent chains
// per SM hide each other's latencies.
template <int NWQ, int KT>
__global__ void __launch_bounds__(128 * NWQ, NWQ == 2 ? 2 : 1) proxy_attention_tc_kernel(const AtArgs a) {
    using namespace at;
    constexpr int THREADS = 128 * NWQ;
    constexpr int TMEM_COLS = NWQ == 2 ? 256 : 512;
    constexpr int ACC1 = KT;                                    // stage-1 accumulators: columns ACC1 + 32 mt
    constexpr int ACC2 = NWQ == 2 ? 224 : 256;                  // stage-2 accumulator
    constexpr int AC = 32 / NWQ;                                // accumulator columns per thread in the epilogues
    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)at_smem_raw + 1023) & ~(uintptr_t)1023);
    float* rowmax = reinterpret_cast<float*>(smem + OFF_MISC);
    float* rowsum = rowmax + 512;
    float* keyflag = rowsum + 512;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = a.n, l = a.l, c = a.c;
    const int npad = (n + 15) & ~15, lpad = (l + 15) & ~15;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(at_smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- operand tiles are K-major, SWIZZLE_128B: 16-byte chunk ch of row r sits at chunk ch ^ (r & 7); rows = [hi 32 | lo 32]
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    float* runm = keyflag + 256;                                // running maximum (scaled) / sum of every proxy row over the key tiles
    float* runl = runm + 256;
    // [Q|K] rows r0.. (column col0 of the head) -> registers (RPT uint4 per thread) -> swizzled tile at `off`: the loads of the next
    // tile are issued before the current tile's pass and land while it runs
    auto load_rows = [&](uint4 (&reg)[4], int col0, int r0, int nrows_tile, int nrows_total) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = tid + k * THREADS, r = i >> 3, ch = i & 7, half = ch >> 2, col = col0 + 8 * (ch & 3);
            reg[k] = z4;
            if (i < nrows_tile * 8 && r0 + r < nrows_total)
                reg[k] = __ldg(reinterpret_cast<const uint4*>(a.qk + (half ? a.qk_plane : 0) + ((size_t)b * n + r0 + r) * a.ldq + col));
        }
    };
    auto store_rows = [&](const uint4 (&reg)[4], int off, int nrows_tile) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = tid + k * THREADS, r = i >> 3, ch = i & 7;
            if (i < nrows_tile * 8) *reinterpret_cast<uint4*>(smem + off + r * ROW_BYTES + ((ch ^ (r & 7)) << 4)) = reg[k];
        }
    };
    // V^T planes of keys k0 .. k0 + KT - 1 (zero past n).  Scenes whose first key is on the 16-byte grid (always, behind
    // pt_proxy_block_fused) go through registers like the K / Q rows: the loads of the next tile are in flight during the passes.
    const bool vt_vec = (((size_t)b * a.vt_seg) & 7) == 0 && (a.vt_seg & 7) == 0 && (a.ldv & 7) == 0 && (a.vt_plane & 7) == 0 && (KT & 7) == 0;
    constexpr int C8 = KT / 8;                                  // 16-byte chunks (8 keys) per row of a tile
    static_assert(2 * HD * C8 == 4 * THREADS, "V^T tile = 4 chunks per thread");
    auto load_vt = [&](uint4 (&reg)[4], int k0) {
        const size_t key0 = (size_t)b * a.vt_seg + k0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = tid + k * THREADS, plane = i / (HD * C8), e = (i / C8) & 31, j8 = i % C8;
            reg[k] = z4;
            if (n - (k0 + 8 * j8) > 0)
                reg[k] = __ldg(reinterpret_cast<const uint4*>(a.vt + (plane ? a.vt_plane : 0) + (size_t)(h * HD + e) * a.ldv + key0 + 8 * j8));
        }
    };
    auto store_vt = [&](const uint4 (&reg)[4], int k0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = tid + k * THREADS, plane = i / (HD * C8), e = (i / C8) & 31, j8 = i % C8;
            uint4 v = reg[k];
            const int left = n - (k0 + 8 * j8);                 // real keys in this chunk (the rest of a scene's last chunk is padding)
            if (left > 0 && left < 8) {
                uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q2 = 0; q2 < 4; ++q2) w4[q2] = 2 * q2 >= left ? 0u : (2 * q2 + 1 >= left ? (w4[q2] & 0xffffu) : w4[q2]);
                v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            *reinterpret_cast<uint4*>(smem + OFF_VT + plane * VT_PLANE + (j8 >> 3) * VT_TILE + e * ROW_BYTES + (((j8 & 7) ^ (e & 7)) << 4)) = v;
        }
    };
    auto stage_vt_elements = [&](int k0) {                      // scenes whose first key is not 16-byte aligned (odd n behind the stand-alone entry)
        const size_t key0 = (size_t)b * a.vt_seg + k0;
        for (int i = tid; i < 2 * HD * KT; i += THREADS) {
            const int plane = i / (HD * KT), e = (i / KT) & 31, j = i % KT;
            unsigned short v = 0;
            if (k0 + j < n) v = __ldg(reinterpret_cast<const unsigned short*>(a.vt + (plane ? a.vt_plane : 0) + (size_t)(h * HD + e) * a.ldv + key0 + j));
            *reinterpret_cast<unsigned short*>(smem + OFF_VT + plane * VT_PLANE + (j >> 6) * VT_TILE + e * ROW_BYTES + ((((j >> 3) & 7) ^ (e & 7)) << 4) + (j & 7) * 2) = v;
        }
    };
    // first K and V^T tiles and the proxies: every load is issued before the first store, so the three latencies overlap
    uint4 kreg[4], qreg[4], vreg[4];
    load_rows(kreg, c + h * HD, 0, KT, n);
    if (vt_vec) load_vt(vreg, 0);
    {
        constexpr int PPT = MAXR * 8 / THREADS;                 // Pt (all l <= 256 proxies stay resident): 16-byte chunks per thread
        uint4 preg[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int i = tid + k * THREADS, r = i >> 3, ch = i & 7, half = ch >> 2, col = h * HD + 8 * (ch & 3);
            preg[k] = z4;
            if (r < l) preg[k] = __ldg(reinterpret_cast<const uint4*>(a.pt + (half ? a.pt_plane : 0) + ((size_t)b * l + r) * c + col));
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int i = tid + k * THREADS, r = i >> 3, ch = i & 7;
            *reinterpret_cast<uint4*>(smem + OFF_P + r * ROW_BYTES + ((ch ^ (r & 7)) << 4)) = preg[k];
        }
    }
    for (int i = tid; i < 256; i += THREADS) {
        keyflag[i] = (a.mask != nullptr && i < l && a.mask[(size_t)b * l + i] == 0) ? 1.f : 0.f;
        runm[i] = -INFINITY;
        runl[i] = 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const int q4 = warp & 3, hh = warp >> 2;                    // TMEM lane quarter, which of its NWQ warps
    const int row = 32 * q4 + lane;                             // row of the 128-row tile this thread owns
    const uint32_t tmem_row = tmem + ((uint32_t)(32 * q4) << 16);
    const uint32_t sK = at_smem_u32(smem + OFF_K), sQ = at_smem_u32(smem + OFF_Q), sP = at_smem_u32(smem + OFF_P);
    const uint32_t sVT = at_smem_u32(smem + OFF_VT), sPV = at_smem_u32(smem + OFF_PV);
    uint32_t phase = 0;

    // one (scores -> softmax -> values) pass over a 128-row tile
    //   sA: operand tile of the rows, sB: operand tile of the keys (nkeys_pad rows), sV: value planes (K-major over the keys)
    //   acc_col: accumulator columns; online >= 0: index of the tile's first row in runm / runl (stage 1, key tile `kt`)
    auto pass = [&](uint32_t sA, uint32_t sB, int nkeys, int nkeys_pad, uint32_t sV, const float* flag, uint32_t acc_col, int online, int kt) {
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = at_idesc(nkeys_pad);
            const uint64_t da = at_desc_sw128(sA), db = at_desc_sw128(sB);
            // k-steps of 16 inside the 128-byte rows: +0/+2 (16-byte units) = hi, +4/+6 = lo
            at_mma_ss(tmem, da + 0, db + 0, idesc, 0u);
            at_mma_ss(tmem, da + 2, db + 2, idesc, 1u);
            at_mma_ss(tmem, da + 4, db + 0, idesc, 1u);
            at_mma_ss(tmem, da + 6, db + 2, idesc, 1u);
            at_mma_ss(tmem, da + 0, db + 4, idesc, 1u);
            at_mma_ss(tmem, da + 2, db + 6, idesc, 1u);
            at_commit(bar);
        }
        at_wait(bar, phase); phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float m_old = online >= 0 ? runm[online + row] : -INFINITY;
        const float m_new = at_softmax_tile<NWQ>(tmem_row, nkeys_pad, nkeys, flag, a.scale, hh, rowmax, rowsum, row, m_old);
        float alpha = 0.f;
        if (online >= 0 && kt > 0) {                            // earlier key tiles were accumulated relative to m_old: rescale
            alpha = at_ex2((m_old - m_new) * 1.4426950408889634f);
            uint32_t v[AC];
            at_ldn(v, tmem_row + acc_col + AC * hh);
#pragma unroll
            for (int e = 0; e < AC; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
            at_stn(tmem_row + acc_col + AC * hh, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (online >= 0 && hh == 0) {
            runl[online + row] = runl[online + row] * alpha + (NWQ == 4 ? ((rowsum[row] + rowsum[128 + row]) + (rowsum[256 + row] + rowsum[384 + row]))
                                                                         : (rowsum[row] + rowsum[128 + row]));
            runm[online + row] = m_new;
        }
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = at_idesc(HD);
            for (int ks = 0; ks < nkeys_pad / 16; ++ks) {
                const uint32_t a_hi = tmem + 32 * (ks >> 1) + 8 * (ks & 1), a_lo = a_hi + 16;
                const uint32_t voff = (uint32_t)((ks >> 2) * VT_TILE + (ks & 3) * 32);
                const uint64_t dv_hi = at_desc_sw128(sV + voff), dv_lo = at_desc_sw128(sV + VT_PLANE + voff);
                at_mma_ts(tmem + acc_col, a_hi, dv_hi, idesc, (ks != 0 || kt > 0) ? 1u : 0u);
                at_mma_ts(tmem + acc_col, a_lo, dv_hi, idesc, 1u);
                at_mma_ts(tmem + acc_col, a_hi, dv_lo, idesc, 1u);
            }
            at_commit(bar);
        }
        at_wait(bar, phase); phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    // ---- stage 1: rows = proxies, keys = clusters in tiles of KT (online softmax), values = V  ->  Pv^T (normalised) as K-major
    // operand planes.  Accumulator of proxy row tile mt: columns ACC1 + 32 mt.
    const int nmt1 = (lpad + 127) >> 7;
    for (int kt = 0; kt * KT < n; ++kt) {
        const int k0 = kt * KT, nk = min(KT, n - k0), nkpad = (nk + 15) & ~15;
        store_rows(kreg, OFF_K, KT);
        if (vt_vec) store_vt(vreg, k0);
        else stage_vt_elements(k0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (k0 + KT < n) {                                                                    // next K / V^T tiles in flight during the passes
            load_rows(kreg, c + h * HD, k0 + KT, KT, n);
            if (vt_vec) load_vt(vreg, k0 + KT);
        } else load_rows(qreg, h * HD, 128 * (int)blockIdx.z, 128, n);                        // ... or the first Q tile of stage 2
        for (int mt = 0; mt < nmt1; ++mt) {
            pass(sP + mt * 128 * ROW_BYTES, sK, nk, nkpad, sVT, nullptr, ACC1 + 32 * mt, 128 * mt, kt);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();                                    // rowmax / rowsum reusable; (last mt) K / V^T tiles reusable
        }
    }
    // (Pv^T takes the place of the V^T tile: every MMA that read it has completed — the last pass waited for its commit.  Rows of
    // proxies l .. lpad - 1 are written as zeros: stage 2 contracts over lpad proxies and the tile still holds V^T there.)
    for (int mt = 0; mt < nmt1; ++mt) {
        uint32_t v[AC];
        at_ldn(v, tmem_row + ACC1 + 32 * mt + AC * hh);
        const int i = mt * 128 + row;                           // proxy index
        if (i < lpad) {
            const float inv = i < l ? 1.0f / runl[i] : 0.f;
            uint8_t* base = smem + OFF_PV + (i >> 6) * VT_TILE + (i & 7) * 2;
            const int ch = (i & 63) >> 3;
#pragma unroll
            for (int e = 0; e < AC; ++e) {
                const int ee = AC * hh + e;
                const float y = i < l ? __uint_as_float(v[e]) * inv : 0.f;
                const __nv_bfloat16 yh = __float2bfloat16_rn(y), yl = __float2bfloat16_rn(y - __bfloat162float(yh));
                uint8_t* p = base + ee * ROW_BYTES + ((ch ^ (ee & 7)) << 4);
                *reinterpret_cast<__nv_bfloat16*>(p) = yh;
                *reinterpret_cast<__nv_bfloat16*>(p + VT_PLANE) = yl;
            }
        }
    }

    // ---- stage 2: rows = clusters (Q) in tiles of 128, keys = proxies (masked), values = Pv.  The row tiles are independent: with
    // few (scene, head) pairs the launch spreads them over gridDim.z CTAs (each repeats stage 1).
    for (int mt = blockIdx.z; mt * 128 < npad; mt += gridDim.z) {
        store_rows(qreg, OFF_Q, 128);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                        // Q tile (and, first time, Pv^T) visible; accumulator drained
        if ((mt + (int)gridDim.z) * 128 < npad) load_rows(qreg, h * HD, 128 * (mt + (int)gridDim.z), 128, n);
        pass(sQ, sP, l, lpad, sPV, a.mask != nullptr ? keyflag : nullptr, ACC2, -1, 0);
        uint32_t v[AC];
        at_ldn(v, tmem_row + ACC2 + AC * hh);
        const int r = mt * 128 + row;                           // cluster index
        if (r < n) {
            const float inv = 1.0f / (NWQ == 4 ? ((rowsum[row] + rowsum[128 + row]) + (rowsum[256 + row] + rowsum[384 + row])) : (rowsum[row] + rowsum[128 + row]));
            float y[AC];
#pragma unroll
            for (int e = 0; e < AC; ++e) y[e] = __uint_as_float(v[e]) * inv;
            const size_t off = ((size_t)b * n + r) * c + h * HD + AC * hh;
#pragma unroll
            for (int e8 = 0; e8 < AC; e8 += 8) {
                if (a.o != nullptr) {
                    *reinterpret_cast<float4*>(a.o + off + e8) = make_float4(y[e8], y[e8 + 1], y[e8 + 2], y[e8 + 3]);
                    *reinterpret_cast<float4*>(a.o + off + e8 + 4) = make_float4(y[e8 + 4], y[e8 + 5], y[e8 + 6], y[e8 + 7]);
                }
                if (a.o_hi != nullptr) {
                    uint32_t wh[4], wl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) at_split2(y[e8 + 2 * e], y[e8 + 2 * e + 1], wh[e], wl[e]);
                    *reinterpret_cast<uint4*>(a.o_hi + off + e8) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                    *reinterpret_cast<uint4*>(a.o_hi + a.o_plane + off + e8) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

bool proxy_attention_tc_supported(int n, int l, int c, int heads) {
    return c % heads == 0 && c / heads == at::HD && c % 8 == 0 && n >= 1 && n <= at::MAXN && l >= 1 && l <= at::MAXR;
}

// qk_split: [2][rows][ldq] bf16 (Q | K); vt_split: [2][c][ldv] bf16 (V^T, column = b*n + j); pt_split: [2][B*l][c] bf16.
int launch_proxy_attention_tc(const void* qk_split, long long qk_plane, int ldq, const void* vt_split, long long vt_plane, long long ldv, int vt_seg,
                              const void* pt_split, long long pt_plane, const uint8_t* mask, int B, int n, int l, int c, int heads,
                              float* o, void* o_split, long long o_plane, cudaStream_t s) {
    PT_REQUIRE(proxy_attention_tc_supported(n, l, c, heads), "attention(tcgen05): n=%d l=%d c=%d heads=%d unsupported", n, l, c, heads);
    PT_REQUIRE(qk_split && vt_split && pt_split && (o || o_split), "attention(tcgen05): null operand");
    // (V^T: any pitch / plane offset — scenes whose columns are off the 16-byte grid are staged element by element)
    PT_REQUIRE(((uintptr_t)qk_split & 15) == 0 && ((uintptr_t)vt_split & 15) == 0 && ((uintptr_t)pt_split & 15) == 0 && ldq % 8 == 0 &&
                   qk_plane % 8 == 0 && pt_plane % 8 == 0 && o_plane % 8 == 0,
               "attention(tcgen05): operand planes must be 16-byte aligned");
    static bool attr[PT_MAX_DEVICES] = {};
    if (first_use_on_current_device(attr)) {
        PT_CUDA_OK(cudaFuncSetAttribute(proxy_attention_tc_kernel<4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, at::SMEM_BYTES));
        PT_CUDA_OK(cudaFuncSetAttribute(proxy_attention_tc_kernel<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, at::SMEM_BYTES));
    }
    AtArgs a;
    a.qk = (const __nv_bfloat16*)qk_split; a.qk_plane = qk_plane; a.ldq = ldq;
    PT_REQUIRE(vt_seg >= n && ldv >= (long long)B * vt_seg, "attention(tcgen05): vt_seg=%d ldv=%lld", vt_seg, ldv);
    a.vt = (const __nv_bfloat16*)vt_split; a.vt_plane = vt_plane; a.ldv = ldv; a.vt_seg = vt_seg;
    a.pt = (const __nv_bfloat16*)pt_split; a.pt_plane = pt_plane;
    a.mask = mask; a.n = n; a.l = l; a.c = c;
    a.scale = (float)(1.0 / sqrt((double)at::HD));              // python float head_dim ** -0.5 (:186), rounded to fp32 once
    a.o = o; a.o_hi = (__nv_bfloat16*)o_split; a.o_plane = o_plane;
    // Two co-resident 8-warp CTAs per SM when the proxies fit the 256-column layout (l <= 224) and there is more than one 16-warp CTA
    // per SM to run; otherwise (few (scene, head) pairs: latency of ONE chain counts, and 256-key tiles halve its passes) the 16-warp form.
    // Few pairs also spread the independent cluster row tiles of stage 2 over the otherwise idle SMs (each CTA repeats stage 1).
    static const int force = getenv("PT_ATTN_FORM") ? atoi(getenv("PT_ATTN_FORM")) : 0;         // debug: 1 = 16-warp form, 2 = 8-warp form
    const int lpad = (l + 15) & ~15;
    const bool small = lpad <= 224 && (force == 2 || (force != 1 && heads * B > 148));
    const int row_tiles = (n + 127) / 128;
    int zsplit = (small ? 296 : 148) / (heads * B);             // stay within a single wave
    zsplit = zsplit < 1 ? 1 : (zsplit > row_tiles ? row_tiles : zsplit);
    {
        ProfScope prof_(PROF_ATTENTION, s);
        if (small) proxy_attention_tc_kernel<2, 128><<<dim3(heads, B, zsplit), 256, at::SMEM_BYTES, s>>>(a);
        else proxy_attention_tc_kernel<4, 256><<<dim3(heads, B, zsplit), 512, at::SMEM_BYTES, s>>>(a);
    }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

}  // namespace pt

// Stand-alone entry point (tests, tools): the attention core on pre-split operands.
extern "C" int pt_proxy_attention_tc(const void* qk_split, long long qk_plane, int ldq, const void* vt_split, long long vt_plane,
                                     long long ldv, const void* pt_split, long long pt_plane, const uint8_t* mask, int B, int n, int l,
                                     int c, int heads, float* o, void* o_split, long long o_plane, pt_stream_t stream) {
    // (V^T columns b*n + j: scenes back to back; the block driver pads every scene to a multiple of 8 columns instead)
    return pt::launch_proxy_attention_tc(qk_split, qk_plane, ldq, vt_split, vt_plane, ldv, n, pt_split, pt_plane, mask, B, n, l, c, heads, o,
                                         o_split, o_plane, (cudaStream_t)stream);
}
