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
//             * X fragments come from ldmatrix on the RAW [channel][225] rows.  A row pitch of 450 B is not 16-byte
//               aligned, but 225 = 1 (mod 8): channel c starts at element c (mod 8) of a 16-byte chunk, so the channels
//               of one residue class s = c mod 8 share their alignment.  Every MMA therefore works on one class: its
//               8 (sums: n index) or 16 (scores: k index, 8 from each slab of a pair) channels are s, s+8, s+16, ... and
//               its token axis runs over the aligned chunks u = token + s.  Chunk columns that belong to the neighbouring
//               channel meet zero probabilities (sums) or land in score columns nobody reads (scores).  The channel
//               order this induces on w_eff and on the weighted sums is absorbed into the GEMM weights on the host
//               (pt_img_pool_params: wk_pad_split rows, wv_cat_split columns).
//   G4-G5   z_h = [y_h | a_h] [W_vc_h | h_v_h]^T ; o = W_c z + b_c ; LayerNorm                      (gemm_tc.cu, dense.cu)
#include "common.cuh"
#include <cuda_fp16.h>
#include "gemm_tc.cuh"

#include <math.h>
#include <stdlib.h>

namespace pt {

int launch_layernorm(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                     float* out, cudaStream_t s);
constexpr float UMMA_FP16_WSCALE = 16.0f;       // scale of the half w_eff planes (launch_img_attnpool_tc G2 <-> imgpool_umma.cu)
int launch_img_pool_umma(const void* img_feat, bool fp16, float wscale, const __nv_bfloat16* wpl, const float* cterm, const float* xbar, __nv_bfloat16* ya_hi,
                         long long ya_plane, int BV, float* dbg, cudaStream_t s);      // imgpool_umma.cu

namespace ip {
constexpr int C = 512, HW = 225, HEADS = 8, HD = 32, EMB = 256;
constexpr int T = HW + 1;                  // attention tokens (mean token first)
constexpr int TP = 228;                    // cterm row pitch
constexpr int YA = 768;                    // per (view, head) row of the value GEMM: 512 weighted sums + 256 probabilities
constexpr int SLAB_CH = 64, NSLAB = C / SLAB_CH;
constexpr int SLAB_BYTES = SLAB_CH * HW * 2;          // 28800
constexpr int RING = 6, REFETCH = NSLAB - RING;       // 2 slabs (the first slab pair) are streamed a second time per view
constexpr int LOADS_PER_VIEW = NSLAB + REFETCH;
constexpr int PF_DIST = 4;                 // L2 prefetch distance of the producer, in ring loads
constexpr int NCHUNK = 29;                 // aligned 8-token chunks per channel row in u = token + class coordinates (232 >= 225 + 7)
constexpr int CONSUMER_WARPS = 16, THREADS = 32 * (CONSUMER_WARPS + 4);   // + a warpgroup whose first warp is the producer
constexpr int WPITCH = 528;                // w_eff plane row pitch in bf16 (512 + 16: lanes of different heads hit different banks)
constexpr int WPLANE = HEADS * WPITCH;     // elements per plane; a view's operand block is [hi plane][lo plane]
constexpr int WBYTES = 2 * WPLANE * 2;     // 16896
constexpr int PPITCH = 264;                // probability row pitch in bf16 (8-token zero margins on both sides)
constexpr int PBYTES = 2 * 2 * HEADS * PPITCH * 2;    // [hi|lo][token parity copy][head][PPITCH] = 16896
constexpr int SPITCH = 232;                // partial-score row pitch (floats)
constexpr int SBUF = HEADS * SPITCH;       // floats per partial-score buffer; 4 buffers (class s and s+4 share one)
// shared memory carve-up (bytes)
constexpr int OFF_RING = 0;
constexpr int OFF_W0 = OFF_RING + RING * SLAB_BYTES;                  // w_eff planes of even views
constexpr int OFF_P = OFF_W0 + WBYTES;                                // probabilities (bf16 hi/lo, two token alignments)
constexpr int OFF_W1 = OFF_P + PBYTES;                                // w_eff planes of odd views
constexpr int OFF_XBAR = OFF_W1 + WBYTES;                             // fp32 [2][512]  (double-buffered by view parity)
constexpr int OFF_MISC = OFF_XBAR + 2 * C * 4;                        // s0 partials [16 warps][8], p0 [8], softmax exchange [32]
constexpr int OFF_BAR = OFF_MISC + (128 + 8 + 32) * 4;                // full[RING], empty[RING], wfull[2], wempty[2]
constexpr int SMEM_BYTES = OFF_BAR + (2 * RING + 4) * 8;
// The four partial-score buffers (29.7 KB) live only between the score MMAs and the softmax; they overlay the (dead) w_eff
// planes of the current view plus the adjacent part of the (dead) probability arrays.
constexpr int SPART_BYTES = 4 * SBUF * 4;
constexpr int DBG_PER_VIEW = HEADS * 256;              // floats per view of the PT_POOL_DEBUG=64 dump (PoolArgs::dbg)
constexpr int OFF_SPART_EVEN = OFF_W0, OFF_SPART_ODD = OFF_W1 + WBYTES - SPART_BYTES;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(SPART_BYTES <= WBYTES + PBYTES && OFF_SPART_ODD >= OFF_P && OFF_SPART_ODD % 16 == 0, "partial-score overlay");
static_assert(OFF_W0 % 16 == 0 && OFF_P % 16 == 0 && OFF_W1 % 16 == 0 && OFF_XBAR % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
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
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x1(uint32_t& r, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
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
// FP16: the 16-bit elements are IEEE half instead of bfloat16 (features of an autocast backbone).
template <bool FP16>
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
            for (int e = 0; e < 4; ++e) {
                if (FP16) {
                    const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                    f[2 * e] = h2.x; f[2 * e + 1] = h2.y;
                } else {
                    f[2 * e] = __uint_as_float(w[e] << 16); f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
                }
            }
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

// Debug event log (build with PT_NVCC_DEFINES=-DPT_POOL_EVENTS, run with PT_POOL_DEBUG bit 32): (id, SM clock) pairs of
// CTA 0 for views 40..43, kept in shared memory by the three logging threads (producer lane, warp 0, warp 8) and dumped at
// the end of the kernel; decoded by tools/pool_events.py.
__device__ long long g_pool_ev[2 * 4096];
__device__ unsigned int g_pool_evn;
#ifdef PT_POOL_EVENTS
#define POOL_EV_MAX 16
#define POOL_EV_LOGGERS 17
#define POOL_EV_DECL(logger) unsigned int ev_n = 0; unsigned int* ev_buf = reinterpret_cast<unsigned int*>(smem + ip::SMEM_BYTES) + (logger) * 2 * POOL_EV_MAX
#define POOL_EV(id)                                                                                        \
    do {                                                                                                   \
        if ((a.debug_skip & 32) && blockIdx.x == 0 && vi >= 40 && vi < 43 && ev_n < POOL_EV_MAX) {         \
            ev_buf[2 * ev_n] = (unsigned int)(id); ev_buf[2 * ev_n + 1] = (unsigned int)clock64(); ++ev_n; \
        }                                                                                                  \
    } while (0)
#define POOL_EV_DUMP()                                                                                     \
    do {                                                                                                   \
        if ((a.debug_skip & 32) && blockIdx.x == 0) {                                                      \
            const unsigned int e0_ = atomicAdd(&g_pool_evn, ev_n);                                         \
            for (unsigned int i_ = 0; i_ < ev_n && e0_ + i_ < 4096; ++i_) {                                \
                g_pool_ev[2 * (e0_ + i_)] = ev_buf[2 * i_]; g_pool_ev[2 * (e0_ + i_) + 1] = ev_buf[2 * i_ + 1]; \
            }                                                                                              \
        }                                                                                                  \
    } while (0)
#define POOL_EV_SMEM (POOL_EV_LOGGERS * 2 * POOL_EV_MAX * 4)
#else
#define POOL_EV_DECL(logger)
#define POOL_EV(id)
#define POOL_EV_DUMP()
#define POOL_EV_SMEM 0
#endif

struct PoolArgs {
    const uint8_t* img;          // (BV, 512, 225) bf16
    const __nv_bfloat16* wpl;    // (BV, 2, 8, WPITCH) bf16: w_eff hi / lo planes, columns in score order (see header)
    const float* cterm;          // (BV, 8, TP) fp32: q_h . g_k[t,h]
    const float* xbar;           // (BV, 512) fp32
    __nv_bfloat16* ya_hi;        // (BV, 8, 768) bf16 hi plane: [0,512) weighted sums (sum order), [512,768) probabilities (zero padded)
    long long ya_plane;          // elements between the hi and lo planes
    int BV;
    float scale;
    int pf_dist;                 // L2 prefetch distance of the producer in ring loads (0 = off)
    float* dbg;                  // debug dump area behind the workspace (null unless the caller over-allocated it): per view
                                 // [8 heads][256] scaled scores (PT_POOL_DEBUG bit 64; tools/pool_check.py)
    int debug_skip;              // PT_POOL_DEBUG bit mask (1, 2, 4, 16 give garbage results): 1 skip score MMAs, 2 slabs are 16-byte loads
                                 // (no HBM traffic), 4 skip sum MMAs, 8 per-phase cycle trace of CTA 0, 16 skip exchange + softmax
};

// Channel orders (host side: pt_img_pool_score_order / pt_img_pool_sum_order in api.cu mirror these):
//   score column ((p*8 + s)*4 + q)*4 + e  <->  channel 128 p + 64 (e >> 1) + s + 16 q + 8 (e & 1)
//   sum   column ((sl*8 + s)*4 + q)*2 + e <->  channel 64 sl + s + 16 q + 8 e
__global__ void __launch_bounds__(ip::THREADS, 1) img_pool_mma_kernel(const PoolArgs a) {
    using namespace ip;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty = full + RING;
    uint64_t* wfull = empty + RING;      // [2]
    uint64_t* wempty = wfull + 2;        // [2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int b = 0; b < RING; ++b) { ip_mbar_init(full + b, 1); ip_mbar_init(empty + b, CONSUMER_WARPS); }
        for (int b = 0; b < 2; ++b) { ip_mbar_init(wfull + b, 1); ip_mbar_init(wempty + b, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // 5 warps per SM sub-partition cap the kernel at 96 registers per thread; the producer needs almost none, so its
    // warpgroup (setmaxnreg is warpgroup-wide: three idle warps keep the producer company and exit at once) hands its share
    // to the consumers.  The pool is what the launch allocated (640 x 96): 4 warps x 72 released registers = 16 per consumer thread
    if (warp >= CONSUMER_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;" ::: "memory");
        if (warp > CONSUMER_WARPS) return;
        // ===== producer: slabs 0..7 of the view, then slabs 0..REFETCH-1 again (FIFO ring, see header) =====
        if (lane == 0) {
            unsigned cnt = 0, vi = 0;
            POOL_EV_DECL(16);
            for (int bv = blockIdx.x; bv < a.BV; bv += gridDim.x, ++vi) {
                // per-view operands: w_eff planes (16.5 KB) + xbar (2 KB) into the buffers of this view's parity; they are
                // requested as soon as the previous view's slab loads are all in flight
                const unsigned wb = vi & 1u;
                ip_mbar_wait(wempty + wb, ((vi >> 1) & 1u) ^ 1u);
                ip_mbar_expect_tx(wfull + wb, WBYTES + C * 4);
                ip_bulk_load(smem + (wb ? OFF_W1 : OFF_W0), a.wpl + (size_t)bv * 2 * WPLANE, WBYTES, wfull + wb);
                ip_bulk_load(smem + OFF_XBAR + wb * C * 4, a.xbar + (size_t)bv * C, C * 4, wfull + wb);
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
            POOL_EV_DUMP();
        }
        return;
    }

    // ===== consumers: 16 warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
    const int g = lane >> 2, q = lane & 3;
    const int mi = lane >> 3, r8 = lane & 7;                            // ldmatrix: this lane addresses row r8 of matrix mi
    float* s0part = reinterpret_cast<float*>(smem + OFF_MISC);          // [16 warps][8 heads]
    float* p0 = s0part + 128;                                           // [8]
    float* red = p0 + 8;                                                // [2][16] softmax max / sum exchange
    const uint32_t ring_u32 = ip_smem_u32(smem + OFF_RING);
    // every MMA works on the channels of one residue class s = channel mod 8 (see header)
    const int s = warp & 7;
    // scores: warp <-> (class s ; token-chunk half nh: chunks [15 nh, 15 nh + 15) of the 29)
    const int nh = warp >> 3, i0 = 15 * nh;
    const uint32_t sc_off = 448u * s + 3600u * r8 + 16u * (i0 + (mi >> 1));   // + 32 per chunk pair; slab of the pair = mi & 1
    // sums: warp <-> (class s ; slab parity hb in processing order)
    const int hb = warp >> 3;
    const uint32_t sm_off = 448u * s + 3600u * r8 + 16u * mi;                  // + 64 per pair of k-blocks
    // softmax: warp <-> (head ; half of the 256 padded attention tokens)
    const int sh = warp & 7, shalf = warp >> 3;
    unsigned vi = 0;               // views done
    unsigned slot0 = 0, wrap0 = 0; // ring slot and wrap count of this view's load 0 (kept incrementally: no division in the loops)
    POOL_EV_DECL(warp);

    long long t_prev = clock64();
    for (int bv = blockIdx.x; bv < a.BV; bv += gridDim.x, ++vi) {
        const unsigned wb = vi & 1u;
        const uint8_t* wbuf = smem + (wb ? OFF_W1 : OFF_W0);
        const float* sxbar = reinterpret_cast<const float*>(smem + OFF_XBAR + wb * C * 4);
        float* spart = reinterpret_cast<float*>(smem + (wb ? OFF_SPART_ODD : OFF_SPART_EVEN));
        // ring slot / mbarrier parity of this view's load j (j < 10 is a compile-time constant wherever this is used)
        auto slot_of = [&](int j) -> unsigned { const unsigned x = slot0 + j; return x >= 2 * RING ? x - 2 * RING : (x >= RING ? x - RING : x); };
        auto par_of = [&](int j) -> unsigned { const unsigned x = slot0 + j; return (wrap0 + (x >= 2 * RING ? 2u : (x >= RING ? 1u : 0u))) & 1u; };
        // position terms of the attention tokens this thread owns in the softmax (global; in flight during the score phase)
        float ct[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = 128 * shalf + lane + 32 * i;
            ct[i] = t < T ? __ldg(a.cterm + ((size_t)bv * HEADS + sh) * TP + t) : 0.f;
        }
        if (lane == 0) POOL_EV(1000 * warp + 10 * ((int)vi - 40) + 0);                    // reached the top of the view
        ip_consumer_sync();                                   // every warp is done with the previous view (probabilities, s0part, red)
        POOL_TRACE(0);
        if (lane == 0) POOL_EV(1000 * warp + 10 * ((int)vi - 40) + 1);                    // past the view barrier
        ip_mbar_wait(wfull + wb, (vi >> 1) & 1u);
        POOL_TRACE(1);                                        // wait for the staged operands
        if (lane == 0) POOL_EV(1000 * warp + 10 * ((int)vi - 40) + 2);                    // operands landed

        // ---- (1) scores: S_s[h][u] = sum over the channels c of class s of w_eff[h][c] X[c][u - s].
        // k-block = 8 class-s channels of slab 2p + 8 of slab 2p+1; B fragments by ldmatrix.trans (rows = channels,
        // 16-byte chunks = 8 tokens), two token chunks per x4.
        float acc[15][4];
#pragma unroll
        for (int i = 0; i < 15; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
        float dotp = 0.f;                                      // this lane's share of s0[g] = w_eff[g] . xbar
#pragma unroll
        for (int p = 0; p < NSLAB / 2; ++p) {
            const unsigned b0 = slot_of(2 * p), b1 = slot_of(2 * p + 1);
            // A fragments straight from the bf16 planes: 4 consecutive columns = k slots 2q, 2q+1, 2q+8, 2q+9
            const int col = ((p * 8 + s) * 4 + q) * 4;
            const uint2 ah = *reinterpret_cast<const uint2*>(wbuf + (g * WPITCH + col) * 2);
            const uint2 al = *reinterpret_cast<const uint2*>(wbuf + (WPLANE + g * WPITCH + col) * 2);
            const uint32_t A[4] = {ah.x, al.x, ah.y, al.y};
            if (nh == 1) {                                     // mean-token score (the upper-half warps have one chunk less to do)
                const int ch = 128 * p + s + 16 * q;
                const float w0 = __uint_as_float(ah.x << 16) + __uint_as_float(al.x << 16), w1 = __uint_as_float(ah.x & 0xffff0000u) + __uint_as_float(al.x & 0xffff0000u);
                const float w2 = __uint_as_float(ah.y << 16) + __uint_as_float(al.y << 16), w3 = __uint_as_float(ah.y & 0xffff0000u) + __uint_as_float(al.y & 0xffff0000u);
                dotp = fmaf(w0, sxbar[ch], dotp); dotp = fmaf(w1, sxbar[ch + 8], dotp);
                dotp = fmaf(w2, sxbar[ch + 64], dotp); dotp = fmaf(w3, sxbar[ch + 72], dotp);
            }
            ip_mbar_wait(full + b0, par_of(2 * p));
            ip_mbar_wait(full + b1, par_of(2 * p + 1));
            const uint32_t base = ring_u32 + ((mi & 1) ? b1 : b0) * SLAB_BYTES + sc_off;
            if (!(a.debug_skip & 1)) {
                uint32_t bf[2][4];                             // fragment loads run one step ahead of the MMAs
                ldsm_x4_t(bf[0], base);
#pragma unroll
                for (int m = 0; m < 7; ++m) {
                    if (m < 6) ldsm_x4_t(bf[(m + 1) & 1], base + 32 * (m + 1));
                    else if (nh == 0) ldsm_x2_t(reinterpret_cast<uint32_t(&)[2]>(bf[1][0]), base + 32 * 7);   // 15th chunk of the lower half
                    mma_bf16_16816(acc[2 * m], A, bf[m & 1][0], bf[m & 1][1]);
                    mma_bf16_16816(acc[2 * m + 1], A, bf[m & 1][2], bf[m & 1][3]);
                }
                if (nh == 0) mma_bf16_16816(acc[14], A, bf[1][0], bf[1][1]);
            }
            if (p == 0) {                                      // the first pair is not kept: hand the buffers back
                __syncwarp();
                if (lane == 0) { ip_mbar_arrive(empty + b0); ip_mbar_arrive(empty + b1); }
            }
        }
        if (nh == 1) {
            dotp += __shfl_xor_sync(FULL, dotp, 1);
            dotp += __shfl_xor_sync(FULL, dotp, 2);
            if (q == 0) s0part[s * 8 + g] = dotp;             // 8 class partials per head
        }
        POOL_TRACE(2);                                        // score MMAs (incl. waiting for slabs)
        if (lane == 0) POOL_EV(1000 * warp + 10 * ((int)vi - 40) + 3);                    // score MMAs done
        ip_consumer_sync();                                   // w_eff planes are dead: the partial-score overlay may be written
        if (lane == 0 && vi == 41) POOL_EV(100000 + 1000 * warp + 20);
        if (!(a.debug_skip & 16)) {
        // hi + lo rows of the accumulators; classes 0-3 store S_s[h][u], then classes 4-7 add theirs four columns lower so
        // that buffer j holds S_j[h][u] + S_{j+4}[h][u+4]: token t sits at column t + j in buffer j
        if (s < 4) {
            float* dst = spart + s * SBUF + g * SPITCH + 2 * q;
#pragma unroll
            for (int i = 0; i < 15; ++i)
                if (i < 14 || nh == 0)
                    *reinterpret_cast<float2*>(dst + 8 * (i0 + i)) = make_float2(acc[i][0] + acc[i][2], acc[i][1] + acc[i][3]);
        }
        ip_consumer_sync();
        if (lane == 0 && vi == 41) POOL_EV(100000 + 1000 * warp + 21);
        if (s >= 4) {
            float* dst = spart + (s - 4) * SBUF + g * SPITCH + 2 * q - 4;
            float2 v[15];
#pragma unroll
            for (int i = 0; i < 15; ++i)
                if ((i < 14 || nh == 0) && (8 * (i0 + i) + 2 * q >= 4)) v[i] = *reinterpret_cast<const float2*>(dst + 8 * (i0 + i));
#pragma unroll
            for (int i = 0; i < 15; ++i)
                if ((i < 14 || nh == 0) && (8 * (i0 + i) + 2 * q >= 4))
                    *reinterpret_cast<float2*>(dst + 8 * (i0 + i)) = make_float2(v[i].x + (acc[i][0] + acc[i][2]), v[i].y + (acc[i][1] + acc[i][3]));
        }
        ip_consumer_sync();
        if (lane == 0 && vi == 41) POOL_EV(100000 + 1000 * warp + 22);

        // ---- (2) softmax over the 226 tokens, two warps per head ; probabilities -> bf16 hi/lo arrays + global
        {
            float sv[4];
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = 128 * shalf + lane + 32 * i;     // attention token; spatial token tau = t - 1
                float v = -INFINITY;
                if (t == 0) {
                    float s0 = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) s0 += s0part[w * 8 + sh];
                    v = a.scale * (s0 + ct[i]);
                } else if (t < T) {
                    const float* sp = spart + sh * SPITCH + (t - 1);
                    v = a.scale * ((((sp[0] + sp[SBUF + 1]) + sp[2 * SBUF + 2]) + sp[3 * SBUF + 3]) + ct[i]);
                }
                sv[i] = v;
                mx = fmaxf(mx, v);
            }
            if ((a.debug_skip & 64) && a.dbg) {                // debug: dump the scaled scores
                float* dbg = a.dbg + (size_t)bv * DBG_PER_VIEW + sh * 256;
                for (int i = 0; i < 4; ++i) dbg[128 * shalf + lane + 32 * i] = sv[i];
            }
            mx = warp_max(mx);
            if (lane == 0) red[warp] = mx;
            ip_consumer_sync();                                // all partial scores have been read
        if (lane == 0 && vi == 41) POOL_EV(100000 + 1000 * warp + 23);
            if (tid == 0) ip_mbar_arrive(wempty + wb);         // the w_eff buffer (and the overlay) may be refilled
            {   // zero the margins of the probability rows (the overlay clobbered them): words [0,5) and [116,132) of 32 rows
                const int row = tid >> 4, j = tid & 15;        // 512 threads = 32 rows x 16
                uint32_t* prow = reinterpret_cast<uint32_t*>(smem + OFF_P) + row * (PPITCH / 2);
                prow[116 + j] = 0u;
                if (j < 5) prow[j] = 0u;
            }
            mx = fmaxf(red[sh], red[sh + 8]);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) { sv[i] = (128 * shalf + lane + 32 * i) < T ? expf(sv[i] - mx) : 0.f; sum += sv[i]; }
            sum = warp_sum(sum);
            if (lane == 0) red[16 + warp] = sum;
            ip_consumer_sync();
        if (lane == 0 && vi == 41) POOL_EV(100000 + 1000 * warp + 24);
            const float inv = 1.0f / (red[16 + sh] + red[16 + sh + 8]);
            __nv_bfloat16* ya = a.ya_hi + ((size_t)bv * HEADS + sh) * YA + C;
            // copy e of plane pl: element tau + 8 + e of row ((pl*2 + e)*8 + head) holds token tau, so that both token
            // parities can be fetched as aligned 32-bit pairs
            unsigned short* pq = reinterpret_cast<unsigned short*>(smem + OFF_P) + sh * PPITCH;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = 128 * shalf + lane + 32 * i;      // attention token 0..255 (>= 226: zero padding)
                const float pr = sv[i] * inv;
                uint32_t hi, lo;
                split_hi_lo(pr, hi, lo);
                ya[t] = __ushort_as_bfloat16((unsigned short)hi);
                ya[a.ya_plane + t] = __ushort_as_bfloat16((unsigned short)lo);
                if (t == 0) p0[sh] = pr;
                if (t >= 1 && t < T) {
                    const int x = t - 1 + 8;
                    pq[x] = (unsigned short)hi;                               // hi, even copy
                    pq[HEADS * PPITCH + x + 1] = (unsigned short)hi;           // hi, odd copy
                    pq[2 * HEADS * PPITCH + x] = (unsigned short)lo;           // lo, even copy
                    pq[3 * HEADS * PPITCH + x + 1] = (unsigned short)lo;       // lo, odd copy
                }
            }
        }
        } else if (tid == 0) ip_mbar_arrive(wempty + wb);
        ip_consumer_sync();
        POOL_TRACE(3);                                        // score exchange + softmax + barriers
        if (lane == 0) POOL_EV(1000 * warp + 10 * ((int)vi - 40) + 4);                    // softmax done

        // ---- (3) weighted sums: Y[h][c] = sum_tok P[h][tok] X[c][tok]  (+ p0[h] xbar[c]) for the 8 class-s channels of a slab
        // per MMA column tile; k runs over the aligned chunks u = tok + s, so the A fragments are the probabilities shifted
        // by s (zero outside [0,225)), held in registers for the whole view.
        uint32_t PA[15][4];
        {
            const int e = s & 1;
            const uint32_t* ph = reinterpret_cast<const uint32_t*>(smem + OFF_P) + (e * HEADS + g) * (PPITCH / 2) + q + ((8 + e - s) >> 1);
            const uint32_t* pl = ph + 2 * HEADS * (PPITCH / 2);
#pragma unroll
            for (int kb = 0; kb < 15; ++kb) {
                PA[kb][0] = ph[8 * kb]; PA[kb][1] = pl[8 * kb]; PA[kb][2] = ph[8 * kb + 4]; PA[kb][3] = pl[8 * kb + 4];
            }
        }
        const float p0g = p0[g];
        if (lane == 0 && vi == 41) POOL_EV(100000 + 1000 * warp + 25);
        __nv_bfloat16* yrow = a.ya_hi + ((size_t)bv * HEADS + g) * YA;
        // Processing order: resident slabs 2..7 (FIFO release order), then the re-fetched slabs 0, 1 (loads 8, 9).  A warp
        // reads every other slab of that order (parity hb); the slots of the slabs it does not read are handed back up
        // front (the arrival only counts, the readers still hold the slot), the re-fetched one after its load has landed
        // so that the arrival cannot fall into the slot's previous phase.
#pragma unroll
        for (int s2 = 0; s2 < NSLAB - REFETCH; ++s2)
            if ((s2 & 1) != hb && lane == 0) ip_mbar_arrive(empty + slot_of(s2 + REFETCH));
#pragma unroll
        for (int s2 = 0; s2 < NSLAB; ++s2) {
            const int sl = s2 < NSLAB - REFETCH ? s2 + REFETCH : s2 - (NSLAB - REFETCH);
            const int j = s2 < NSLAB - REFETCH ? sl : NSLAB + sl;            // load index within the view
            const unsigned b = slot_of(j);
            if ((s2 & 1) != hb) {
                if (s2 >= NSLAB - REFETCH) {
                    ip_mbar_wait(full + b, par_of(j));
                    __syncwarp();
                    if (lane == 0) ip_mbar_arrive(empty + b);
                }
                continue;
            }
            if (s2 >= NSLAB - REFETCH) ip_mbar_wait(full + b, par_of(j));
            float y0[4] = {0.f, 0.f, 0.f, 0.f}, y1[4] = {0.f, 0.f, 0.f, 0.f}, y2[4] = {0.f, 0.f, 0.f, 0.f};   // independent MMA chains
            if (!(a.debug_skip & 4)) {
                const uint32_t base = ring_u32 + b * SLAB_BYTES + sm_off;
                uint32_t bf[2][4];                             // fragment loads run one step ahead of the MMAs
                ldsm_x4(bf[0], base);
#pragma unroll
                for (int m = 0; m < 7; ++m) {
                    if (m < 6) ldsm_x4(bf[(m + 1) & 1], base + 64 * (m + 1));
                    else ldsm_x1(bf[1][0], base + 64 * 7);     // chunk 28; the k-block's upper half (u >= 232) is empty
                    const uint32_t* f = bf[m & 1];
                    if (m % 3 == 0) { mma_bf16_16816(y0, PA[2 * m], f[0], f[1]); mma_bf16_16816(y1, PA[2 * m + 1], f[2], f[3]); }
                    else if (m % 3 == 1) { mma_bf16_16816(y2, PA[2 * m], f[0], f[1]); mma_bf16_16816(y0, PA[2 * m + 1], f[2], f[3]); }
                    else { mma_bf16_16816(y1, PA[2 * m], f[0], f[1]); mma_bf16_16816(y2, PA[2 * m + 1], f[2], f[3]); }
                }
                mma_bf16_16816(y0, PA[14], bf[1][0], 0u);
            }
            __syncwarp();
            if (lane == 0) ip_mbar_arrive(empty + b);
            // accumulator rows g (hi part) / g+8 (lo part), columns 2q, 2q+1 = channels s + 16 q, s + 16 q + 8 of slab sl
            const int ch = sl * SLAB_CH + s + 16 * q;
            const float v0 = (((y0[0] + y1[0]) + y2[0]) + ((y0[2] + y1[2]) + y2[2])) + p0g * sxbar[ch];
            const float v1 = (((y0[1] + y1[1]) + y2[1]) + ((y0[3] + y1[3]) + y2[3])) + p0g * sxbar[ch + 8];
            uint32_t h0, l0, h1, l1;
            split_hi_lo(v0, h0, l0);
            split_hi_lo(v1, h1, l1);
            const int col = ((sl * 8 + s) * 4 + q) * 2;
            *reinterpret_cast<uint32_t*>(yrow + col) = h0 | (h1 << 16);
            *reinterpret_cast<uint32_t*>(yrow + a.ya_plane + col) = l0 | (l1 << 16);
        }
        slot0 += LOADS_PER_VIEW - RING;                        // 10 loads per view on a ring of 6
        wrap0 += 1;
        if (slot0 >= RING) { slot0 -= RING; wrap0 += 1; }
        POOL_TRACE(4);                                        // weighted sums
    }
    if (lane == 0) POOL_EV_DUMP();
}



}  // namespace pt
extern "C" int pt_debug_pool_events(long long* out, int max_events) {
    unsigned int n = 0;
    if (cudaMemcpyFromSymbol(&n, pt::g_pool_evn, sizeof(n)) != cudaSuccess) return PT_ERR_CUDA;
    if (n > 4096u) n = 4096u;
    if ((int)n > max_events) n = (unsigned int)max_events;
    if (n && cudaMemcpyFromSymbol(out, pt::g_pool_ev, (size_t)n * 16) != cudaSuccess) return PT_ERR_CUDA;
    const unsigned int z = 0;
    if (cudaMemcpyToSymbol(pt::g_pool_evn, &z, sizeof(z)) != cudaSuccess) return PT_ERR_CUDA;
    return (int)n;
}
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
    float *xbar, *cterm, *o;
    __nv_bfloat16 *xbar_split, *q_split, *wpl, *ya_split, *z_split;
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
    r.wpl = (__nv_bfloat16*)take((size_t)BV * 2 * WPLANE * 2);
    r.cterm = (float*)take((size_t)BV * HEADS * TP * 4);
    r.ya_split = (__nv_bfloat16*)take((size_t)2 * BV * HEADS * YA * 2);
    r.z_split = (__nv_bfloat16*)take((size_t)2 * BV * EMB * 2);
    r.o = (float*)take((size_t)BV * EMB * 4);
    r.total = off;
    return r;
}

size_t img_attnpool_tc_ws_bytes(int BV) { return carve_tc(nullptr, BV).total; }

bool img_attnpool_tc_supported(int img_dtype, const pt_img_pool_params* p, int C, int HW, int c, int heads) {
    // fp16 features: the tcgen05 kernel only (its feature tiles are the A operand of both contractions; kind::f16 takes an f16 A
    // next to bf16 B operands), the mma.sync kernel is bf16 x bf16
    const bool dtype_ok = img_dtype == PT_DTYPE_BF16 || (img_dtype == PT_DTYPE_F16 && p->variant == PT_POOL_VARIANT_UMMA);
    return dtype_ok && C == ip::C && HW == ip::HW && c == ip::EMB && heads == ip::HEADS && p->w_qc_split &&
           p->wk_pad_split && p->gk_pad_split && p->wv_cat_split && p->cproj_split;
}

int launch_img_attnpool_tc(const void* img_feat, int img_dtype, const pt_img_pool_params* p, int BV, float* img_proxy, void* ws, size_t ws_bytes,
                           int stages, cudaStream_t s) {
    const bool fp16 = img_dtype == PT_DTYPE_F16;
    using namespace ip;
    PT_REQUIRE(((uintptr_t)img_feat & 15) == 0, "pt_img_attnpool: img_feat must be 16-byte aligned");
    ImgTcWs w = carve_tc(ws, BV);
    if (ws_bytes < w.total) { set_error("pt_img_attnpool: workspace %zu < %zu", ws_bytes, w.total); return PT_ERR_WORKSPACE; }
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int rc;
    GemmProfTagScope img_gemms(PROF_GEMM_IMG);
    const bool umma = p->variant == PT_POOL_VARIANT_UMMA;
    if (stages & PT_IMG_STAGE_FRONT) {
    {   // pass A
        const long long groups = (long long)BV * (C / 8);
        const long long blocks = (groups + 7) / 8;
        // CTAs per SM: 8 saturate HBM when the kernel runs alone; fewer leave room for the geometry kernels that the host
        // module runs concurrently on another stream (they are latency / ALU bound and barely touch HBM)
        static const int per_sm = [] { const char* e = getenv("PT_MEAN_CTAS"); const int v = e ? atoi(e) : 8; return v >= 1 && v <= 8 ? v : 8; }();
        const int grid = (int)(blocks < (long long)sms * per_sm ? blocks : (long long)sms * per_sm);
        ProfScope prof_(PROF_IMG_MEAN, s);
        if (fp16) img_mean_bf16_kernel<true><<<grid, 256, 0, s>>>((const uint4*)img_feat, groups, w.xbar, w.xbar_split, w.xbar_split + (size_t)BV * C);
        else img_mean_bf16_kernel<false><<<grid, 256, 0, s>>>((const uint4*)img_feat, groups, w.xbar, w.xbar_split, w.xbar_split + (size_t)BV * C);
    }
    PT_LAUNCH_CHECK();
    {   // G1: q = xbar W_qc^T + q0  -> split planes only
        GemmTc gp;
        gp.M = BV; gp.N = EMB; gp.K = C;
        gp.a_split = w.xbar_split; gp.a_rows = BV; gp.a_cols = C; gp.lda = C;
        gp.w_split = p->w_qc_split; gp.w_rows = EMB; gp.ldw = C;
        gp.bias = p->q0;
        gp.c_split = w.q_split; gp.cs_plane = (long long)BV * EMB; gp.ldcs = EMB;
        if ((rc = launch_gemm_tc_ex(gp, s))) return rc;
    }
    {   // G2: w_eff[:, h, :] = q[:, 32h:32h+32] W_kc_h   (K = 32 real + 32 columns that hit zero weights), emitted as the
        // bf16 hi / lo planes the pool kernel bulk-copies per view: [view][hi|lo][head][WPITCH], columns in score order
        GemmTc gp;
        gp.M = BV; gp.N = C; gp.K = 64; gp.batch = HEADS;
        gp.a_split = w.q_split; gp.a_rows = BV; gp.a_cols = EMB; gp.lda = EMB; gp.a_koff_z = HD;
        gp.w_split = p->wk_pad_split; gp.w_rows = HEADS * C; gp.ldw = 64; gp.w_row_z = C;
        // (the tcgen05 pool kernel reads the planes through a tensor map: unpadded rows of 512)
        const int wpitch = umma ? C : WPITCH;
        gp.c_split = w.wpl; gp.cs_plane = HEADS * wpitch; gp.ldcs = 2 * HEADS * wpitch; gp.cs_off_z = wpitch;
        // fp16 features: kind::f16 wants both MMA operands in the same 16-bit format, so the planes are IEEE half (x 16: |w_eff| < 4094,
        // lo halves clear of the half subnormals down to |w_eff| ~ 1e-2; the pool kernel folds the 1/16 into its score scale)
        gp.cs_fp16 = fp16; gp.cs_scale = fp16 ? UMMA_FP16_WSCALE : 1.0f;
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
    }
    if (!(stages & PT_IMG_STAGE_BACK)) return PT_OK;
    if (umma) {   // pass B on tcgen05 tensor cores, TMA-fed (imgpool_umma.cu)
        const char* dbg = getenv("PT_POOL_DEBUG");
        float* dump = dbg && (atoi(dbg) & 64) && ws_bytes >= w.total + (size_t)BV * DBG_PER_VIEW * 4 ? reinterpret_cast<float*>((char*)ws + w.total) : nullptr;
        if ((rc = launch_img_pool_umma(img_feat, fp16, fp16 ? UMMA_FP16_WSCALE : 1.0f, w.wpl, w.cterm, w.xbar, w.ya_split, (long long)BV * HEADS * YA, BV, dump, s))) return rc;
    } else {   // pass B, mma.sync form
        static bool attr_set[PT_MAX_DEVICES] = {};
        if (first_use_on_current_device(attr_set))
            PT_CUDA_OK(cudaFuncSetAttribute(img_pool_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + POOL_EV_SMEM));
        PoolArgs a;
        a.img = (const uint8_t*)img_feat; a.wpl = w.wpl; a.cterm = w.cterm; a.xbar = w.xbar;
        a.ya_hi = w.ya_split; a.ya_plane = (long long)BV * HEADS * YA; a.BV = BV;
        a.scale = (float)(1.0 / sqrt((double)HD));
        const char* dbg = getenv("PT_POOL_DEBUG");
        a.debug_skip = dbg ? atoi(dbg) : 0;
        a.dbg = ws_bytes >= w.total + (size_t)BV * DBG_PER_VIEW * 4 ? reinterpret_cast<float*>((char*)ws + w.total) : nullptr;
        const char* pf = getenv("PT_POOL_PF");
        a.pf_dist = pf ? atoi(pf) : PF_DIST;
        int grid = BV < sms ? BV : sms;
        if (const char* ge = getenv("PT_POOL_GRID")) { const int gv = atoi(ge); if (gv >= 1 && gv < grid) grid = gv; }   // probes: fewer persistent CTAs
        { ProfScope prof_(PROF_IMG_POOL, s); img_pool_mma_kernel<<<grid, THREADS, SMEM_BYTES + POOL_EV_SMEM, s>>>(a); }
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
