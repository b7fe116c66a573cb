// S9 image proxies, bf16 fast path: get_img_proxy (:335-342) + AttentionPool2d.forward (:154-177) in single-query form
// (algebra in imgpool.cu) for the shipped geometry C=512 channels, 15x15=225 positions, 8 heads of 32.
//
//   pass A  img_mean_bf16_kernel   per-channel spatial mean (HBM-bound stream, one warp per 8 channels = 3600 B)
//   G1-G3   tcgen05 3xBF16 GEMMs   q = W_qc xbar + q0 ; w_eff_h = q_h W_kc_h ; cterm_h = q_h . g_k     (gemm_tc.cu)
//   pass B  img_pool_mma_kernel    scores -> softmax -> attention-weighted feature sums, one persistent CTA per SM:
//             * a producer warp streams the view as eight 64-channel slabs (28.8 KB, cp.async.bulk + mbarrier) through a
//               6-deep shared-memory ring; 6 of the 8 slabs stay resident between the score and the weighted-sum phase,
//               2 are fetched again (L2 hits), so HBM sees every byte once;
//             * scores S[8 heads][225] = W_eff X and sums Y[8][512] = P X^T run on the tensor cores (mma.sync m16n8k16
//               bf16, fp32 accumulate): the fp32 operand (w_eff / probabilities) is split into bf16 hi + lo halves that
//               occupy rows 0-7 / 8-15 of the 16-row A tile, X is already bf16, so the products are exact to ~2^-17;
//             * the k index of every MMA is permuted so that the B fragments can be read straight from the raw
//               [channel][225] bf16 rows (450-byte pitch, not 16-byte aligned: no ldmatrix / UMMA layout possible)
//               without shared-memory bank conflicts.
//   G4-G5   z_h = [y_h | a_h] [W_vc_h | h_v_h]^T ; o = W_c z + b_c ; LayerNorm                      (gemm_tc.cu, dense.cu)
#include "common.cuh"
#include "gemm_tc.cuh"

#include <math.h>
#include <stdlib.h>

namespace pt {

int launch_layernorm(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                     float* out, cudaStream_t s);

namespace ip {
constexpr int C = 512, HW = 225, HEADS = 8, HD = 32, EMB = 256;
constexpr int T = HW + 1;                  // attention tokens (mean token first)
constexpr int TP = 228;                    // cterm row pitch
constexpr int YA = 768;                    // per (view, head) row of the value GEMM: 512 weighted sums + 256 probabilities
constexpr int SLAB_CH = 64, NSLAB = C / SLAB_CH;
constexpr int SLAB_BYTES = SLAB_CH * HW * 2;          // 28800
constexpr int RING = 6, REFETCH = NSLAB - RING;       // 2 slabs are streamed a second time per view
constexpr int LOADS_PER_VIEW = NSLAB + REFETCH;
constexpr int PF_DIST = 4;                 // L2 prefetch distance of the producer, in ring loads
constexpr int NT_SCORE = 29;               // 8-token score tiles (232 >= 225)
constexpr int CONSUMER_WARPS = 16, THREADS = 32 * (CONSUMER_WARPS + 1);
// shared memory carve-up (bytes)
constexpr int OFF_RING = 0;
constexpr int OFF_PAD = OFF_RING + RING * SLAB_BYTES;                 // 128 B of zeros behind the ring (fragment over-reads)
constexpr int OFF_WFRAG = OFF_PAD + 128;                              // [32 k-blocks][32 lanes][4 x u32]
constexpr int OFF_PFRAG = OFF_WFRAG + 32 * 32 * 16;                   // [16 k-blocks][32 lanes][4 x u32]
constexpr int OFF_S = OFF_PFRAG + 16 * 32 * 16;                       // fp32 scores [8][232]
constexpr int OFF_XBAR = OFF_S + HEADS * 232 * 4;                     // fp32 [2][512]  (double-buffered by view parity)
constexpr int OFF_WSTAGE = OFF_XBAR + 2 * C * 4;                      // fp32 [8][512]  next view's w_eff (bulk-copied)
constexpr int OFF_MISC = OFF_WSTAGE + HEADS * C * 4;                  // s0 partials [16 warps][8], p0 [8], softmax exchange [32]
constexpr int OFF_BAR = OFF_MISC + (128 + 8 + 32) * 4;                      // full[RING], empty[RING], wfull, wempty
constexpr int SMEM_BYTES = OFF_BAR + (2 * RING + 2) * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(OFF_WFRAG % 16 == 0 && OFF_PFRAG % 16 == 0 && OFF_BAR % 8 == 0 && OFF_WSTAGE % 16 == 0 && OFF_XBAR % 16 == 0, "alignment");
}  // namespace ip

__device__ __forceinline__ uint32_t ip_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ip_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ip_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ip_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ip_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ip_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ip_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ip_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(ip_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ip_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ip_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(ip_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void ip_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ip_consumer_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// fp32 -> (bf16 hi, bf16 lo) with hi + lo == x to ~2^-17
__device__ __forceinline__ void split_hi_lo(float x, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = (uint32_t)__bfloat16_as_ushort(h);
    lo = (uint32_t)__bfloat16_as_ushort(l);
}

// ------------------------------------------------------------------------------------------------ pass A
// One warp per group of 8 channels (1800 bf16 = 225 uint4, 16-byte aligned).  Iteration i of a lane reads uint4
// lane + 32 i, whose 8 elements belong to channel i or i+1 of the group only, so the 8 running sums are static registers.
__global__ void __launch_bounds__(256) img_mean_bf16_kernel(const uint4* __restrict__ img, long long groups,
                                                            float* __restrict__ xbar, __nv_bfloat16* __restrict__ xb_hi,
                                                            __nv_bfloat16* __restrict__ xb_lo) {
    const int lane = threadIdx.x & 31;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long grp = wg; grp < groups; grp += nw) {
        const uint4* src = img + grp * 225;
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = lane + 32 * i;
            v[i] = j < 225 ? __ldg(src + j) : make_uint4(0u, 0u, 0u, 0u);
        }
        float acc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = lane + 32 * i;
            const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            float f[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) { f[2 * e] = __uint_as_float(w[e] << 16); f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
            const float s_all = ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
            const int nlo = min(max(225 * (i + 1) - 8 * j, 0), 8);       // elements of this uint4 that belong to channel i
            if (nlo == 8) acc[i] += s_all;
            else if (nlo == 0) acc[i + 1] += s_all;
            else {
                float s_lo = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) s_lo += e < nlo ? f[e] : 0.f;
                float s_hi = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) s_hi += e < nlo ? 0.f : f[e];
                acc[i] += s_lo;
                acc[i + 1] += s_hi;
            }
        }
        // transposed butterfly: after three exchange steps lane l holds the partial of channel (l & 7) summed over the
        // lanes congruent to l mod 4... (8 values x 32 lanes -> 8 totals with 3 + 2 shuffle rounds instead of 40 shuffles)
        float r4[4], r2[2], r1;
        {
            const bool up = lane & 16;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float keep = up ? acc[k + 4] : acc[k], give = up ? acc[k] : acc[k + 4];
                r4[k] = keep + __shfl_xor_sync(FULL, give, 16);
            }
        }
        {
            const bool up = lane & 8;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float keep = up ? r4[k + 2] : r4[k], give = up ? r4[k] : r4[k + 2];
                r2[k] = keep + __shfl_xor_sync(FULL, give, 8);
            }
        }
        {
            const bool up = lane & 4;
            const float keep = up ? r2[1] : r2[0], give = up ? r2[0] : r2[1];
            r1 = keep + __shfl_xor_sync(FULL, give, 4);
        }
        r1 += __shfl_xor_sync(FULL, r1, 2);
        r1 += __shfl_xor_sync(FULL, r1, 1);
        // lane bits: 16 -> +4, 8 -> +2, 4 -> +1 of the channel index
        if ((lane & 3) == 0) {
            const int ch = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            const float m = r1 / 225.0f;
            const long long o = grp * 8 + ch;
            xbar[o] = m;
            const __nv_bfloat16 h = __float2bfloat16_rn(m);
            xb_hi[o] = h;
            xb_lo[o] = __float2bfloat16_rn(m - __bfloat162float(h));
        }
    }
}

// ------------------------------------------------------------------------------------------------ pass B
// Debug timeline (PT_POOL_DEBUG bit 8): SM-clock cycles spent by CTA 0 / warp 0 in each phase, summed over its views.
__device__ unsigned long long g_pool_trace[8];
#define POOL_TRACE(slot)                                                                   \
    do {                                                                                   \
        if ((a.debug_skip & 8) && blockIdx.x == 0 && tid == 0) {                           \
            const long long now_ = clock64();                                              \
            atomicAdd(&g_pool_trace[slot], (unsigned long long)(now_ - t_prev));           \
            t_prev = now_;                                                                 \
        }                                                                                  \
    } while (0)

struct PoolArgs {
    const uint8_t* img;          // (BV, 512, 225) bf16
    const float* w_eff;          // (BV, 8, 512) fp32
    const float* cterm;          // (BV, 8, TP) fp32: q_h . g_k[t,h]
    const float* xbar;           // (BV, 512) fp32
    __nv_bfloat16* ya_hi;        // (BV, 8, 768) bf16 hi plane: [0,512) weighted sums, [512,768) probabilities (zero padded)
    long long ya_plane;          // elements between the hi and lo planes
    int BV;
    float scale;
    int pf_dist;                 // L2 prefetch distance of the producer in ring loads (0 = off)
    int debug_skip;              // PT_POOL_DEBUG bit mask (results are garbage): 1 skip score MMAs, 4 skip sum MMAs, 2 no slab data (16-byte loads)
};

__global__ void __launch_bounds__(ip::THREADS, 1) img_pool_mma_kernel(const PoolArgs a) {
    using namespace ip;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty = full + RING;
    uint64_t* wfull = empty + RING;
    uint64_t* wempty = wfull + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int b = 0; b < RING; ++b) { ip_mbar_init(full + b, 1); ip_mbar_init(empty + b, CONSUMER_WARPS); }
        ip_mbar_init(wfull, 1);
        ip_mbar_init(wempty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) reinterpret_cast<uint32_t*>(smem + OFF_PAD)[tid] = 0u;
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ===== producer: slabs 0..7 of the view, then slabs 0..REFETCH-1 again (FIFO ring, see header) =====
        if (lane == 0) {
            unsigned cnt = 0, vi = 0;
            for (int bv = blockIdx.x; bv < a.BV; bv += gridDim.x, ++vi) {
                // per-view operands (w_eff 16 KB -> staging, xbar 2 KB -> buffer vi & 1); they are requested as soon as the
                // previous view's slab loads are all in flight, i.e. while the consumers are still in its weighted-sum phase
                ip_mbar_wait(wempty, (vi & 1u) ^ 1u);
                ip_mbar_expect_tx(wfull, HEADS * C * 4 + C * 4);
                ip_bulk_load(smem + OFF_WSTAGE, a.w_eff + (size_t)bv * HEADS * C, HEADS * C * 4, wfull);
                ip_bulk_load(smem + OFF_XBAR + (vi & 1u) * C * 4, a.xbar + (size_t)bv * C, C * 4, wfull);
                const uint8_t* view = a.img + (size_t)bv * C * HW * 2;
                for (int k = 0; k < LOADS_PER_VIEW; ++k, ++cnt) {
                    const int slab = k < NSLAB ? k : k - NSLAB;
                    const unsigned b = cnt % RING, ph = (cnt / RING) & 1u;
                    // The ring holds barely one view, so a slab can only be requested when the consumers let go of a
                    // buffer: too late to hide HBM latency.  Pull the slab that will be requested PF_DIST loads from now
                    // into L2 already (re-fetched slabs are L2-resident anyway).
                    {
                        int k2 = k + a.pf_dist, bv2 = bv;
                        if (k2 >= LOADS_PER_VIEW) { k2 -= LOADS_PER_VIEW; bv2 += gridDim.x; }
                        if (a.pf_dist > 0 && k2 < NSLAB && bv2 < a.BV) ip_prefetch_l2(a.img + (size_t)bv2 * C * HW * 2 + (size_t)k2 * SLAB_BYTES, SLAB_BYTES);
                    }
                    ip_mbar_wait(empty + b, ph ^ 1u);
                    const uint32_t nbytes = (a.debug_skip & 2) ? 16u : (uint32_t)SLAB_BYTES;
                    ip_mbar_expect_tx(full + b, nbytes);
                    ip_bulk_load(smem + OFF_RING + b * SLAB_BYTES, view + (size_t)slab * SLAB_BYTES, nbytes, full + b);
                }
            }
        }
        return;
    }

    // ===== consumers: 16 warps =====
    const int g = lane >> 2, q = lane & 3;
    uint4* wfrag = reinterpret_cast<uint4*>(smem + OFF_WFRAG);
    uint32_t* pfrag32 = reinterpret_cast<uint32_t*>(smem + OFF_PFRAG);
    float* S = reinterpret_cast<float*>(smem + OFF_S);
    const float* wstage = reinterpret_cast<const float*>(smem + OFF_WSTAGE);
    float* s0part = reinterpret_cast<float*>(smem + OFF_MISC);          // [16 warps][8 heads]
    float* p0 = s0part + 128;                                           // [8]
    float* red = p0 + 8;                                                // [2][16] softmax max / sum exchange
    float4* ypart = reinterpret_cast<float4*>(smem + OFF_WFRAG);        // [2][8 n-tiles][32 lanes], aliases Wfrag (dead in phase 3)
    // scores: warp <-> (pair of 16-token groups gp, gp + 8 ; channel parity)
    const int gp = warp & 7, par = warp >> 3;
    const int ngrp = gp + 8 < 15 ? 2 : 1;                               // 15 groups = 240 tokens >= 225
    // sums: warp <-> (8-channel tile nt of the slab ; token half kh: tokens [128 kh, 128 kh + 128))
    const int nt = warp & 7, kh = warp >> 3;
    unsigned cnt = 0, vi = 0;      // loads consumed so far (ring position / parity), views done

    long long t_prev = clock64();
    for (int bv = blockIdx.x; bv < a.BV; bv += gridDim.x, ++vi) {
        // ---- (0) per-view operands (staged by the producer): w_eff -> bf16 hi/lo A fragments, s0 = w_eff . xbar
        ip_consumer_sync();                                   // every warp is done with the previous view's Wfrag / ypart / scores
        const float* sxbar = reinterpret_cast<const float*>(smem + OFF_XBAR + (vi & 1u) * C * 4);
        POOL_TRACE(0);                                        // barrier (0)
        ip_mbar_wait(wfull, vi & 1u);
        POOL_TRACE(1);                                        // wait for the staged operands
        {
            // 8 consecutive channels ch..ch+7 of head g -> two fragment entries: k-block 4*sl+jj takes the even channels,
            // k-block 4*sl+2+jj the odd ones (k-slot s <-> channel ch + 2s + parity), see the score loop
            const int sl = warp >> 1, jj = warp & 1;
            const int ch = 64 * sl + 16 * q + 8 * jj;
            float f[8], x[8];
            *reinterpret_cast<float4*>(f) = *reinterpret_cast<const float4*>(wstage + g * C + ch);
            *reinterpret_cast<float4*>(f + 4) = *reinterpret_cast<const float4*>(wstage + g * C + ch + 4);
            *reinterpret_cast<float4*>(x) = *reinterpret_cast<const float4*>(sxbar + ch);
            *reinterpret_cast<float4*>(x + 4) = *reinterpret_cast<const float4*>(sxbar + ch + 4);
            uint32_t hi[8], lo[8];
            float dotp = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { dotp = fmaf(f[e], x[e], dotp); split_hi_lo(f[e], hi[e], lo[e]); }
            // a0 = (row g: hi, k-slots 2q,2q+1), a1 = (row g+8: lo), a2 = (row g: hi, slots 2q+8,2q+9), a3 = lo
            wfrag[(4 * sl + jj) * 32 + lane] = make_uint4(hi[0] | (hi[2] << 16), lo[0] | (lo[2] << 16), hi[4] | (hi[6] << 16), lo[4] | (lo[6] << 16));
            wfrag[(4 * sl + 2 + jj) * 32 + lane] = make_uint4(hi[1] | (hi[3] << 16), lo[1] | (lo[3] << 16), hi[5] | (hi[7] << 16), lo[5] | (lo[7] << 16));
            dotp += __shfl_xor_sync(FULL, dotp, 1);
            dotp += __shfl_xor_sync(FULL, dotp, 2);
            if (q == 0) s0part[warp * 8 + g] = dotp;
        }
        // cterm of the score columns the even-parity warps own: tokens 16 i + 4 q + {0,1,2,3} (attention token = spatial + 1)
        float ctv[2][4];
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int tok = 16 * (gp + 8 * t) + 4 * q + e;
                ctv[t][e] = (par == 0 && t < ngrp && tok < HW) ? __ldg(a.cterm + ((size_t)bv * HEADS + g) * TP + 1 + tok) : 0.f;
            }
        ip_consumer_sync();
        if (tid == 0) ip_mbar_arrive(wempty);                 // staging may be refilled with the next view's w_eff
        POOL_TRACE(2);                                        // conversion + barrier (1)

        // ---- (1) scores: S[h][tok] = sum_ch w_eff[h][ch] X[ch][tok].
        // A B register packs two k-consecutive bf16, i.e. the same token of two channels (two rows of the slab), so every
        // aligned 32-bit word read from a row carries TWO tokens: one k-block (16 channels of one parity, chosen so that the
        // words are aligned and the 32 lanes hit 32 different banks) feeds two 8-token tiles at once.  Even channels see
        // tokens (16i+2g, +1) -> tiles E_i, O_i; odd channels (rows start on an odd bf16) see (16i+2g-1, 16i+2g) -> O'_i, E_i;
        // O' is O shifted by one token (its first column, token -1, is the previous row's tail and is dropped).
        // Warps 0-7 take the even channels, warps 8-15 the odd ones; the partial scores meet in shared memory.
        float accA[2][4], accB[2][4];            // first / second token of the words
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) { accA[t][e] = 0.f; accB[t][e] = 0.f; }
        for (int sl = 0; sl < NSLAB; ++sl) {
            const unsigned k = cnt + sl, b = k % RING, ph = (k / RING) & 1u;
            ip_mbar_wait(full + b, ph);
            const uint32_t* X0 = reinterpret_cast<const uint32_t*>(smem + OFF_RING + b * SLAB_BYTES) + 1800 * q + g + 8 * gp + (par ? 112 : 0);
            if (!(a.debug_skip & 1))
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const uint4 af = wfrag[(sl * 4 + 2 * par + jj) * 32 + lane];
                const uint32_t A[4] = {af.x, af.y, af.z, af.w};
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (t < ngrp) {
                        const uint32_t* x = X0 + (4 * jj) * HW + 64 * t;
                        const uint32_t wa = x[0], wb = x[HW], wc = x[2 * HW], wd = x[3 * HW];
                        mma_bf16_16816(accA[t], A, __byte_perm(wa, wb, 0x5410), __byte_perm(wc, wd, 0x5410));
                        mma_bf16_16816(accB[t], A, __byte_perm(wa, wb, 0x7632), __byte_perm(wc, wd, 0x7632));
                    }
                }
            }
            if (sl < REFETCH) {                                // this slab is not kept: hand the buffer back
                __syncwarp();
                if (lane == 0) ip_mbar_arrive(empty + b);
            }
        }
        POOL_TRACE(3);                                        // score MMAs (incl. waiting for slabs)
        // unscaled scores (hi + lo rows of the accumulators): even-parity warps store E, O (+ cterm), then odd-parity warps
        // add their E and O' contributions
        if (par == 0) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (t < ngrp) {
                    const int tok = 16 * (gp + 8 * t) + 4 * q;
                    float* Sr = S + g * 232 + 1 + tok;
                    if (tok < HW) Sr[0] = (accA[t][0] + accA[t][2]) + ctv[t][0];
                    if (tok + 1 < HW) Sr[1] = (accB[t][0] + accB[t][2]) + ctv[t][1];
                    if (tok + 2 < HW) Sr[2] = (accA[t][1] + accA[t][3]) + ctv[t][2];
                    if (tok + 3 < HW) Sr[3] = (accB[t][1] + accB[t][3]) + ctv[t][3];
                }
            }
        }
        ip_consumer_sync();
        if (par == 1) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (t < ngrp) {
                    const int tok = 16 * (gp + 8 * t) + 4 * q;
                    float* Sr = S + g * 232 + 1 + tok;
                    // each (head, token) below is touched by exactly one thread: O' columns are tokens tok-1, tok+1
                    float o_prev = accA[t][0] + accA[t][2], o_next = accA[t][1] + accA[t][3];
                    if (tok < HW) Sr[0] += accB[t][0] + accB[t][2];
                    if (tok + 2 < HW) Sr[2] += accB[t][1] + accB[t][3];
                    if (tok >= 1 && tok - 1 < HW) Sr[-1] += o_prev;
                    if (tok + 1 < HW) Sr[1] += o_next;
                }
            }
        }
        if (warp < HEADS && lane == 0) {                        // token 0 = mean token
            float s0 = 0.f;
#pragma unroll
            for (int w = 0; w < CONSUMER_WARPS; ++w) s0 += s0part[w * 8 + warp];
            S[warp * 232] = s0 + __ldg(a.cterm + ((size_t)bv * HEADS + warp) * TP);
        }
        ip_consumer_sync();

        // ---- (2) softmax over the 226 tokens, two warps per head ; probabilities -> bf16 hi/lo A fragments + global
        {
            const int h = warp & 7, half = warp >> 3;
            float sv[4];
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = 128 * half + lane + 32 * i;
                sv[i] = t < T ? a.scale * S[h * 232 + t] : -INFINITY;
                mx = fmaxf(mx, sv[i]);
            }
            mx = warp_max(mx);
            if (lane == 0) red[warp] = mx;
            ip_consumer_sync();
            mx = fmaxf(red[h], red[h + 8]);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) { sv[i] = (128 * half + lane + 32 * i) < T ? expf(sv[i] - mx) : 0.f; sum += sv[i]; }
            sum = warp_sum(sum);
            if (lane == 0) red[16 + warp] = sum;
            ip_consumer_sync();
            const float inv = 1.0f / (red[16 + h] + red[16 + h + 8]);
            __nv_bfloat16* ya = a.ya_hi + ((size_t)bv * HEADS + h) * YA + C;
            unsigned short* pf16 = reinterpret_cast<unsigned short*>(pfrag32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = 128 * half + lane + 32 * i;      // attention token 0..255 (>= 226: zero padding)
                const float p = sv[i] * inv;
                uint32_t hi, lo;
                split_hi_lo(p, hi, lo);
                ya[t] = __ushort_as_bfloat16((unsigned short)hi);
                ya[a.ya_plane + t] = __ushort_as_bfloat16((unsigned short)lo);
                if (t == 0) p0[h] = p;
                // spatial token tau = t - 1 -> fragment slot
                const int tau = t - 1;
                if (tau >= 0) {
                    const int kb = 2 * (tau >> 5) + ((tau >> 2) & 1), fl = 4 * h + ((tau >> 3) & 3), reg = (tau & 2);
                    const int o16 = ((kb * 32 + fl) * 4 + reg) * 2 + (tau & 1);
                    pf16[o16] = (unsigned short)hi;
                    pf16[o16 + 2] = (unsigned short)lo;       // reg + 1
                }
            }
            if (half == 1 && lane == 0) {                       // tau = 255 (t = 256) is not covered by the loop above
                const int tau = 255;
                const int kb = 2 * (tau >> 5) + ((tau >> 2) & 1), fl = 4 * h + ((tau >> 3) & 3), reg = (tau & 2);
                const int o16 = ((kb * 32 + fl) * 4 + reg) * 2 + (tau & 1);
                pf16[o16] = 0; pf16[o16 + 2] = 0;
            }
        }
        ip_consumer_sync();

        POOL_TRACE(4);                                        // score exchange + softmax + barriers
        // ---- (3) weighted sums: Y[h][ch] = sum_tok P[h][tok] X[ch][tok]  (+ p0[h] xbar[ch]).
        // warp <-> (8 channels of the slab, half of the tokens); the two halves meet through ypart + a 64-thread barrier
        uint32_t PA[8][4];
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {
            const uint4 pf = reinterpret_cast<const uint4*>(pfrag32)[(8 * kh + kb) * 32 + lane];
            PA[kb][0] = pf.x; PA[kb][1] = pf.y; PA[kb][2] = pf.z; PA[kb][3] = pf.w;
        }
        const float p0g = p0[g];
        const uint32_t shift = (g & 1) * 16;                   // rows of odd channels start on an odd bf16 (225 is odd)
        __nv_bfloat16* yrow = a.ya_hi + ((size_t)bv * HEADS + g) * YA;
        for (int s2 = 0; s2 < NSLAB; ++s2) {
            // resident slabs REFETCH..7 first (FIFO release order), then the re-fetched slabs 0..REFETCH-1
            const int sl = s2 < NSLAB - REFETCH ? s2 + REFETCH : s2 - (NSLAB - REFETCH);
            const unsigned k = s2 < NSLAB - REFETCH ? cnt + sl : cnt + NSLAB + sl;
            const unsigned b = k % RING;
            if (s2 >= NSLAB - REFETCH) ip_mbar_wait(full + b, (k / RING) & 1u);
            const int cl = 8 * nt + g;                          // channel within the slab
            const uint32_t* Xw = reinterpret_cast<const uint32_t*>(smem + OFF_RING + b * SLAB_BYTES) + ((cl * HW + 8 * q) >> 1) + 64 * kh;
            float y[4] = {0.f, 0.f, 0.f, 0.f}, y2[4] = {0.f, 0.f, 0.f, 0.f};     // two independent MMA chains
            if (!(a.debug_skip & 4))
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                uint32_t w[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) w[i] = Xw[16 * p + i];
                uint32_t r[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = __funnelshift_r(w[i], w[i + 1], shift);
                if (p == 3 && kh == 1) {                       // tokens 224..255: only 224 is real, the rest must not leak NaNs
                    r[0] = q == 0 ? (r[0] & 0xffffu) : 0u;
                    r[1] = 0u; r[2] = 0u; r[3] = 0u;
                }
                mma_bf16_16816(y, PA[2 * p], r[0], r[1]);
                mma_bf16_16816(y2, PA[2 * p + 1], r[2], r[3]);
            }
            __syncwarp();
            if (lane == 0) ip_mbar_arrive(empty + b);
#pragma unroll
            for (int e = 0; e < 4; ++e) y[e] += y2[e];
            float4* yp = ypart + ((s2 & 1) * 8 + nt) * 32 + lane;
            if (kh == 1) *yp = make_float4(y[0], y[1], y[2], y[3]);
            asm volatile("bar.sync %0, 64;" ::"r"(2 + nt) : "memory");
            if (kh == 0) {
                const float4 o4 = *yp;
                // accumulator rows g (hi part) / g+8 (lo part), columns = channels 8*nt + 2q, +1 of slab sl
                const int ch = sl * SLAB_CH + 8 * nt + 2 * q;
                const float y0 = ((y[0] + o4.x) + (y[2] + o4.z)) + p0g * sxbar[ch], y1 = ((y[1] + o4.y) + (y[3] + o4.w)) + p0g * sxbar[ch + 1];
                uint32_t h0, l0, h1, l1;
                split_hi_lo(y0, h0, l0);
                split_hi_lo(y1, h1, l1);
                *reinterpret_cast<uint32_t*>(yrow + ch) = h0 | (h1 << 16);
                *reinterpret_cast<uint32_t*>(yrow + a.ya_plane + ch) = l0 | (l1 << 16);
            }
        }
        cnt += LOADS_PER_VIEW;
        POOL_TRACE(5);                                        // weighted sums
    }
}

}  // namespace pt
extern "C" int pt_debug_pool_trace(unsigned long long* out8, int reset) {
    if (out8 && cudaMemcpyFromSymbol(out8, pt::g_pool_trace, sizeof(pt::g_pool_trace)) != cudaSuccess) return PT_ERR_CUDA;
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cudaMemcpyToSymbol(pt::g_pool_trace, z, sizeof(z)) != cudaSuccess) return PT_ERR_CUDA;
    }
    return PT_OK;
}
namespace pt {

// ------------------------------------------------------------------------------------------------ host
struct ImgTcWs {
    float *xbar, *w_eff, *cterm, *o;
    __nv_bfloat16 *xbar_split, *q_split, *ya_split, *z_split;
    size_t total;
};

static ImgTcWs carve_tc(void* ws, int BV) {
    using namespace ip;
    ImgTcWs r;
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = ws ? (void*)((char*)ws + off) : nullptr; off += align_up(bytes, 256); return p; };
    r.xbar = (float*)take((size_t)BV * C * 4);
    r.xbar_split = (__nv_bfloat16*)take((size_t)2 * BV * C * 2);
    r.q_split = (__nv_bfloat16*)take((size_t)2 * BV * EMB * 2);
    r.w_eff = (float*)take((size_t)BV * HEADS * C * 4);
    r.cterm = (float*)take((size_t)BV * HEADS * TP * 4);
    r.ya_split = (__nv_bfloat16*)take((size_t)2 * BV * HEADS * YA * 2);
    r.z_split = (__nv_bfloat16*)take((size_t)2 * BV * EMB * 2);
    r.o = (float*)take((size_t)BV * EMB * 4);
    r.total = off;
    return r;
}

size_t img_attnpool_tc_ws_bytes(int BV) { return carve_tc(nullptr, BV).total; }

bool img_attnpool_tc_supported(int img_dtype, const pt_img_pool_params* p, int C, int HW, int c, int heads) {
    return img_dtype == PT_DTYPE_BF16 && C == ip::C && HW == ip::HW && c == ip::EMB && heads == ip::HEADS && p->w_qc_split &&
           p->wk_pad_split && p->gk_pad_split && p->wv_cat_split && p->cproj_split;
}

int launch_img_attnpool_tc(const void* img_feat, const pt_img_pool_params* p, int BV, float* img_proxy, void* ws, size_t ws_bytes,
                           cudaStream_t s) {
    using namespace ip;
    PT_REQUIRE(((uintptr_t)img_feat & 15) == 0, "pt_img_attnpool: img_feat must be 16-byte aligned");
    ImgTcWs w = carve_tc(ws, BV);
    if (ws_bytes < w.total) { set_error("pt_img_attnpool: workspace %zu < %zu", ws_bytes, w.total); return PT_ERR_WORKSPACE; }
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    {   // pass A
        const long long groups = (long long)BV * (C / 8);
        const long long blocks = (groups + 7) / 8;
        const int grid = (int)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
        ProfScope prof_(PROF_IMG_MEAN, s);
        img_mean_bf16_kernel<<<grid, 256, 0, s>>>((const uint4*)img_feat, groups, w.xbar, w.xbar_split, w.xbar_split + (size_t)BV * C);
    }
    PT_LAUNCH_CHECK();
    int rc;
    {   // G1: q = xbar W_qc^T + q0  -> split planes only
        GemmTc gp;
        gp.M = BV; gp.N = EMB; gp.K = C;
        gp.a_split = w.xbar_split; gp.a_rows = BV; gp.a_cols = C; gp.lda = C;
        gp.w_split = p->w_qc_split; gp.w_rows = EMB; gp.ldw = C;
        gp.bias = p->q0;
        gp.c_split = w.q_split; gp.cs_plane = (long long)BV * EMB; gp.ldcs = EMB;
        if ((rc = launch_gemm_tc_ex(gp, s))) return rc;
    }
    {   // G2: w_eff[:, h, :] = q[:, 32h:32h+32] W_kc_h   (K = 32 real + 32 columns that hit zero weights)
        GemmTc gp;
        gp.M = BV; gp.N = C; gp.K = 64; gp.batch = HEADS;
        gp.a_split = w.q_split; gp.a_rows = BV; gp.a_cols = EMB; gp.lda = EMB; gp.a_koff_z = HD;
        gp.w_split = p->wk_pad_split; gp.w_rows = HEADS * C; gp.ldw = 64; gp.w_row_z = C;
        gp.C = w.w_eff; gp.ldc = HEADS * C; gp.c_off_z = C;
        if ((rc = launch_gemm_tc_ex(gp, s))) return rc;
    }
    {   // G3: cterm[:, h, t] = q[:, 32h:32h+32] . g_k[t, 32h:32h+32]
        GemmTc gp;
        gp.M = BV; gp.N = TP; gp.K = 64; gp.batch = HEADS;
        gp.a_split = w.q_split; gp.a_rows = BV; gp.a_cols = EMB; gp.lda = EMB; gp.a_koff_z = HD;
        gp.w_split = p->gk_pad_split; gp.w_rows = HEADS * TP; gp.ldw = 64; gp.w_row_z = TP;
        gp.C = w.cterm; gp.ldc = HEADS * TP; gp.c_off_z = TP;
        if ((rc = launch_gemm_tc_ex(gp, s))) return rc;
    }
    {   // pass B
        static bool attr_set = false;
        if (!attr_set) {
            PT_CUDA_OK(cudaFuncSetAttribute(img_pool_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set = true;
        }
        PoolArgs a;
        a.img = (const uint8_t*)img_feat; a.w_eff = w.w_eff; a.cterm = w.cterm; a.xbar = w.xbar;
        a.ya_hi = w.ya_split; a.ya_plane = (long long)BV * HEADS * YA; a.BV = BV;
        a.scale = (float)(1.0 / sqrt((double)HD));
        const char* dbg = getenv("PT_POOL_DEBUG");
        a.debug_skip = dbg ? atoi(dbg) : 0;
        const char* pf = getenv("PT_POOL_PF");
        a.pf_dist = pf ? atoi(pf) : PF_DIST;
        const int grid = BV < sms ? BV : sms;
        { ProfScope prof_(PROF_IMG_POOL, s); img_pool_mma_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(a); }
        PT_LAUNCH_CHECK();
    }
    {   // G4: z[:, 32h:32h+32] = [y_h | a_h] [W_vc_h | h_v_h]^T   -> split planes only
        GemmTc gp;
        gp.M = BV; gp.N = HD; gp.K = YA; gp.batch = HEADS;
        gp.a_split = w.ya_split; gp.a_rows = BV; gp.a_cols = HEADS * YA; gp.lda = HEADS * YA; gp.a_koff_z = YA;
        gp.w_split = p->wv_cat_split; gp.w_rows = EMB; gp.ldw = YA; gp.w_row_z = HD;
        gp.c_split = w.z_split; gp.cs_plane = (long long)BV * EMB; gp.ldcs = EMB; gp.cs_off_z = HD;
        if ((rc = launch_gemm_tc_ex(gp, s))) return rc;
    }
    {   // G5: o = z W_c^T + b_c
        GemmTc gp;
        gp.M = BV; gp.N = EMB; gp.K = EMB;
        gp.a_split = w.z_split; gp.a_rows = BV; gp.a_cols = EMB; gp.lda = EMB;
        gp.w_split = p->cproj_split; gp.w_rows = EMB; gp.ldw = EMB;
        gp.bias = p->cproj_b;
        gp.C = w.o; gp.ldc = EMB;
        if ((rc = launch_gemm_tc_ex(gp, s))) return rc;
    }
    return launch_layernorm(w.o, p->ln_w, p->ln_b, nullptr, 1, BV, EMB, img_proxy, s);
}

}  // namespace pt
