// Geometric stages of the preshape path: grid prior (S1), first-K ball query (S2/S4), offset network (S3),
// cluster dropout (S5) and the PointNet point encoder (S6).  HBM/latency-bound fp32 CUDA-core kernels.
// Reference semantics: embodiedscan/models/necks/preshape_norm_reverse_drop.py (":line" below).
#include "common.cuh"

#include <math.h>

namespace pt {

// ------------------------------------------------------------------------------------------------ S1
// :33-51.  Stage 1: per-block partial min/max of a slab of points; stage 2: finish the reduction and emit centres.
constexpr int MM_THREADS = 256;
constexpr int MM_POINTS_PER_BLOCK = 4096;

__global__ void __launch_bounds__(MM_THREADS) minmax_partial_kernel(const float* __restrict__ points, int N,
                                                                    float* __restrict__ partial) {
    const int b = blockIdx.y, nblk = gridDim.x;
    const float* P = points + (size_t)b * N * 3;
    const int p0 = blockIdx.x * MM_POINTS_PER_BLOCK;
    const int p1 = min(N, p0 + MM_POINTS_PER_BLOCK);
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int p_scalar = p0;
    if ((reinterpret_cast<uintptr_t>(P) & 15) == 0) {           // 4 points = 3 x 16 bytes per thread and step
        const int ngrp = (p1 - p0) >> 2;
        const float4* P4 = reinterpret_cast<const float4*>(P + (size_t)p0 * 3);
        auto fold = [&](const float4& a, const float4& b4, const float4& c) {
            const float x[4] = {a.x, a.w, b4.z, c.y}, y[4] = {a.y, b4.x, b4.w, c.z}, z[4] = {a.z, b4.y, c.x, c.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mn[0] = fminf(mn[0], x[i]); mx[0] = fmaxf(mx[0], x[i]);
                mn[1] = fminf(mn[1], y[i]); mx[1] = fmaxf(mx[1], y[i]);
                mn[2] = fminf(mn[2], z[i]); mx[2] = fmaxf(mx[2], z[i]);
            }
        };
        if (ngrp == MM_POINTS_PER_BLOCK / 4) {                    // full block: all 12 loads of the thread in flight at once
            constexpr int GPT = MM_POINTS_PER_BLOCK / 4 / MM_THREADS;
            float4 v[GPT][3];
#pragma unroll
            for (int k = 0; k < GPT; ++k) {
                const int gq = threadIdx.x + k * MM_THREADS;
                v[k][0] = __ldg(P4 + 3 * gq); v[k][1] = __ldg(P4 + 3 * gq + 1); v[k][2] = __ldg(P4 + 3 * gq + 2);
            }
#pragma unroll
            for (int k = 0; k < GPT; ++k) fold(v[k][0], v[k][1], v[k][2]);
        } else {
            for (int gq = threadIdx.x; gq < ngrp; gq += MM_THREADS) fold(__ldg(P4 + 3 * gq), __ldg(P4 + 3 * gq + 1), __ldg(P4 + 3 * gq + 2));
        }
        p_scalar = p0 + 4 * ngrp;
    }
    for (int p = p_scalar + threadIdx.x; p < p1; p += MM_THREADS) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float v = __ldg(P + (size_t)p * 3 + d);
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    }
    __shared__ float s[MM_THREADS / 32][6];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        mn[d] = warp_min(mn[d]);
        mx[d] = warp_max(mx[d]);
    }
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { s[w][d] = mn[d]; s[w][3 + d] = mx[d]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[0][threadIdx.x];
        for (int i = 1; i < MM_THREADS / 32; ++i) v = threadIdx.x < 3 ? fminf(v, s[i][threadIdx.x]) : fmaxf(v, s[i][threadIdx.x]);
        partial[((size_t)b * nblk + blockIdx.x) * 6 + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) centres_kernel(const float* __restrict__ partial, int nblk, int gs, int M,
                                                      const float* __restrict__ lin, float margin, float* __restrict__ mn_out,
                                                      float* __restrict__ mx_out, float* __restrict__ centres) {
    const int b = blockIdx.y;
    __shared__ float s[6];
    if (threadIdx.x < 6) {
        const float* pp = partial + (size_t)b * nblk * 6 + threadIdx.x;
        float v = pp[0];
        for (int i = 1; i < nblk; ++i) v = threadIdx.x < 3 ? fminf(v, pp[(size_t)i * 6]) : fmaxf(v, pp[(size_t)i * 6]);
        s[threadIdx.x] = v;
        if (blockIdx.x == 0) {
            if (threadIdx.x < 3) mn_out[b * 3 + threadIdx.x] = v;
            else mx_out[b * 3 + threadIdx.x - 3] = v;
        }
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int g[3] = {j / (gs * gs), (j / gs) % gs, j % gs};   // 'ij' meshgrid: x slowest (:44-45)
    const float two_m = __fmul_rn(2.0f, margin);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        // ((mn + margin) + grid * ((mx - mn) - 2*margin)), no contraction (:48)
        float ext = __fsub_rn(__fsub_rn(s[3 + d], s[d]), two_m);
        centres[((size_t)b * M + j) * 3 + d] = __fadd_rn(__fadd_rn(s[d], margin), __fmul_rn(__ldg(lin + g[d]), ext));
    }
}

// ------------------------------------------------------------------------------------------------ S2 / S4
// pytorch3d ball_query semantics (:56,:65): first K point indices in ascending index with d2 < r2, -1 padding.
// One warp per centre, no shared memory and no block barriers: every centre scans the same point stream from index 0,
// so the stream is L1/L2-resident and each warp stops exactly when ITS centre has K hits.  A lane tests 4 consecutive
// points per step (3 x LDG.128 = 48 contiguous bytes, next step prefetched while the current one is tested); hits are
// appended in index order from 4 ballots: slot = count + (hits of lower lanes) + (earlier hits of this lane).
constexpr int BQ_WARPS = 8;

__device__ __forceinline__ void bq_unpack(const float4& a, const float4& b, const float4& c, float (&x)[4], float (&y)[4], float (&z)[4]) {
    x[0] = a.x; y[0] = a.y; z[0] = a.z;
    x[1] = a.w; y[1] = b.x; z[1] = b.y;
    x[2] = b.z; y[2] = b.w; z[2] = c.x;
    x[3] = c.y; y[3] = c.z; z[3] = c.w;
}

__global__ void __launch_bounds__(BQ_WARPS * 32) ball_query_kernel(const float* __restrict__ centres,
                                                                   const float* __restrict__ points, int M, int N, int K,
                                                                   float r2, int32_t* __restrict__ idx,
                                                                   int32_t* __restrict__ pad_counts, long long total) {
    const int lane = threadIdx.x & 31;
    const long long cm = (long long)blockIdx.x * BQ_WARPS + (threadIdx.x >> 5);      // flat (scene, centre)
    if (cm >= total) return;
    const int b = (int)(cm / M);
    const float* P = points + (size_t)b * N * 3;
    const float cx = __ldg(centres + cm * 3), cy = __ldg(centres + cm * 3 + 1), cz = __ldg(centres + cm * 3 + 2);
    int32_t* out = idx + cm * K;
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    const int nfull = N / 128;                        // full 128-point steps (vector loads need 4 whole points per lane)
    const bool vec_ok = (((uintptr_t)P) & 15) == 0;   // scene base 16-byte aligned (N*12 bytes per scene)
    int step = 0;
    if (vec_ok && nfull > 0) {
        const float4* P4 = reinterpret_cast<const float4*>(P) + 3 * lane;
        // one 128-point step on registers (a, bq, c); returns true when the centre has its K hits
        auto test = [&](const float4& a, const float4& bq, const float4& c, int st) -> bool {
            float x[4], y[4], z[4];
            bq_unpack(a, bq, c, x, y, z);
            unsigned m[4];
            bool h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                h[i] = dist2_rn(cx, cy, cz, x[i], y[i], z[i]) < r2;
                m[i] = __ballot_sync(FULL, h[i]);
            }
            if ((m[0] | m[1] | m[2] | m[3]) != 0u) {
                int slot = cnt + __popc(m[0] & lt) + __popc(m[1] & lt) + __popc(m[2] & lt) + __popc(m[3] & lt);
                const int j0 = st * 128 + 4 * lane;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (h[i]) {
                        if (slot < K) out[slot] = j0 + i;
                        ++slot;
                    }
                }
                cnt += __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
                if (cnt >= K) return true;
            }
            return false;
        };
        // two steps per trip on two register sets, each reloaded (for the step after next) right after it has been tested: the
        // single-set form spent 13 of its ~93 instructions per step copying the prefetched registers (the kernel is issue-bound:
        // `not_selected` is its top stall)
        float4 a0 = __ldg(P4), b0 = __ldg(P4 + 1), c0 = __ldg(P4 + 2);
        float4 a1 = a0, b1 = b0, c1 = c0;
        if (nfull > 1) { a1 = __ldg(P4 + 96); b1 = __ldg(P4 + 97); c1 = __ldg(P4 + 98); }
        for (; step < nfull; step += 2) {
            const bool done0 = test(a0, b0, c0, step);
            if (done0 || step + 1 >= nfull) break;
            if (step + 2 < nfull) {
                const float4* q4 = P4 + (size_t)(step + 2) * 96;
                a0 = __ldg(q4); b0 = __ldg(q4 + 1); c0 = __ldg(q4 + 2);
            }
            const bool done1 = test(a1, b1, c1, step + 1);
            if (done1) break;
            if (step + 3 < nfull) {
                const float4* q4 = P4 + (size_t)(step + 3) * 96;
                a1 = __ldg(q4); b1 = __ldg(q4 + 1); c1 = __ldg(q4 + 2);
            }
        }
    }
    if (cnt < K) {                                     // scalar tail (and the unaligned / tiny-N case): 32 points per step
        for (int j0 = (vec_ok ? nfull * 128 : 0); j0 < N && cnt < K; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            if (j < N) hit = dist2_rn(cx, cy, cz, __ldg(P + (size_t)j * 3), __ldg(P + (size_t)j * 3 + 1), __ldg(P + (size_t)j * 3 + 2)) < r2;
            const unsigned mask = __ballot_sync(FULL, hit);
            if (hit) {
                const int slot = cnt + __popc(mask & lt);
                if (slot < K) out[slot] = j;
            }
            cnt += __popc(mask);
        }
    }
    const int filled = min(cnt, K);
    for (int k = filled + lane; k < K; k += 32) out[k] = -1;
    if (pad_counts != nullptr && lane == 0) pad_counts[cm] = K - filled;
}

// ------------------------------------------------------------------------------------------------ S3 / S6
// Shared body of OffsetNetwork (:87-107) and SimplifiedPointNet (:126-142): gather K neighbours, 6 features,
// conv1x1 6->256 + BN(eval) + ReLU, then mean (offset net) or max (encoder) over K.  One warp per cluster; lane l owns
// channels l, l+32, ..., l+224, so the 6x256 weights live in registers and features are broadcast by shuffle.
constexpr int MLP_H = 256;
constexpr int MLP_CPL = MLP_H / 32;   // channels per lane
constexpr int MLP_WARPS = 8;

template <bool ENCODER>
__global__ void __launch_bounds__(MLP_WARPS * 32) cluster_mlp_kernel(
    const float* __restrict__ points, const int32_t* __restrict__ idx, const float* __restrict__ centres,
    const float* __restrict__ mn, const float* __restrict__ mx, const float* __restrict__ conv_w,
    const float* __restrict__ conv_b, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
    const float* __restrict__ map_w, int B, int M, int N, int K, float margin, float* __restrict__ out,
    float* __restrict__ raw_offsets) {
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * MLP_WARPS + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * MLP_WARPS;
    float w[MLP_CPL][6], bb[MLP_CPL], sc[MLP_CPL], sh[MLP_CPL];
#pragma unroll
    for (int i = 0; i < MLP_CPL; ++i) {
        const int ch = lane + 32 * i;
#pragma unroll
        for (int f = 0; f < 6; ++f) w[i][f] = __ldg(conv_w + ch * 6 + f);
        bb[i] = __ldg(conv_b + ch);
        sc[i] = __ldg(bn_scale + ch);
        sh[i] = __ldg(bn_shift + ch);
    }
    // The gather (idx -> point: two dependent global loads) of the NEXT cluster's first 32 neighbours is issued before the
    // arithmetic of the current one, so its latency hides behind ~2000 FMAs instead of stalling the warp.
    auto gather = [&](int cm, int kk, float& px, float& py, float& pz) {
        px = 0.f; py = 0.f; pz = 0.f;
        if (cm < B * M && kk < K) {
            const float* P = points + (size_t)(cm / M) * N * 3;
            const int id = __ldg(idx + (size_t)cm * K + kk);
            if (id >= 0) { px = __ldg(P + (size_t)id * 3); py = __ldg(P + (size_t)id * 3 + 1); pz = __ldg(P + (size_t)id * 3 + 2); }
        }
    };
    float nx, ny, nz;
    gather(warp_global, lane, nx, ny, nz);
    for (int cm = warp_global; cm < B * M; cm += n_warps) {
        const int b = cm / M;
        const float cx = __ldg(centres + (size_t)cm * 3), cy = __ldg(centres + (size_t)cm * 3 + 1),
                    cz = __ldg(centres + (size_t)cm * 3 + 2);
        float red[MLP_CPL];
#pragma unroll
        for (int i = 0; i < MLP_CPL; ++i) red[i] = ENCODER ? -INFINITY : 0.0f;
        for (int k0 = 0; k0 < K; k0 += 32) {
            float px, py, pz;
            if (k0 == 0) {
                px = nx; py = ny; pz = nz;
                gather(cm + n_warps, lane, nx, ny, nz);
            } else {
                gather(cm, k0 + lane, px, py, pz);
            }
            // padding is detected on the gathered COORDINATES, not on idx (:94, :132); -0.0 compares equal to 0
            const bool pad = (px == 0.0f) && (py == 0.0f) && (pz == 0.0f);
            const float rx = pad ? 0.0f : __fsub_rn(px, cx), ry = pad ? 0.0f : __fsub_rn(py, cy),
                        rz = pad ? 0.0f : __fsub_rn(pz, cz);
            const int kn = min(32, K - k0);
            for (int t = 0; t < kn; ++t) {
                const float f0 = __shfl_sync(FULL, rx, t), f1 = __shfl_sync(FULL, ry, t), f2 = __shfl_sync(FULL, rz, t);
                const float f3 = __shfl_sync(FULL, px, t), f4 = __shfl_sync(FULL, py, t), f5 = __shfl_sync(FULL, pz, t);
#pragma unroll
                for (int i = 0; i < MLP_CPL; ++i) {
                    float a = bb[i];
                    a = fmaf(w[i][0], f0, a); a = fmaf(w[i][1], f1, a); a = fmaf(w[i][2], f2, a);
                    a = fmaf(w[i][3], f3, a); a = fmaf(w[i][4], f4, a); a = fmaf(w[i][5], f5, a);
                    const float y = fmaxf(fmaf(a, sc[i], sh[i]), 0.0f);
                    red[i] = ENCODER ? fmaxf(red[i], y) : red[i] + y;
                }
            }
        }
        if (ENCODER) {
#pragma unroll
            for (int i = 0; i < MLP_CPL; ++i) out[(size_t)cm * MLP_H + lane + 32 * i] = red[i];
        } else {
            float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < MLP_CPL; ++i) {
                const float z = __fdiv_rn(red[i], (float)K);   // torch.mean = sum / K (:102)
#pragma unroll
                for (int d = 0; d < 3; ++d) o[d] = fmaf(__ldg(map_w + d * MLP_H + lane + 32 * i), z, o[d]);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) o[d] = warp_sum(o[d]);
            if (lane < 3) {
                const float raw = lane == 0 ? o[0] : (lane == 1 ? o[1] : o[2]);
                const float c0 = lane == 0 ? cx : (lane == 1 ? cy : cz);
                if (raw_offsets != nullptr) raw_offsets[(size_t)cm * 3 + lane] = raw;
                const float c1 = __fadd_rn(c0, __fmul_rn(tanhf(raw), margin));            // :59-61
                out[(size_t)cm * 3 + lane] = fmaxf(fminf(c1, __ldg(mx + b * 3 + lane)), __ldg(mn + b * 3 + lane));  // :62
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ S5
// dynamic_cluster_dropout (:352-420), one CTA per scene.  Everything is index work on <= a few thousand clusters:
// pad counts -> stable counting sort (keys 0..K) -> FPS over the first keep1 centres (registers hold the running
// minimum distances, one block-wide arg-max per round with a packed (dist,~index) key so the first maximum wins) ->
// ordered complement -> gathers.
constexpr int FPS_PER = 8;

// The FPS rounds of one scene, run by the first TF threads of the CTA (a multiple of 32; they meet at named barrier 1).
// A round is one dependent chain (distances to the last pick -> running minima -> arg-max -> next pick) and the kernel is
// bound by its latency (ncu: ~5 cycles per issued instruction with two warps per scheduler), so the body is straight-line
// code with NPER independent distance chains per thread and as few instructions as possible: slots past keep1 keep a
// running minimum of 0 and therefore lose every comparison (a valid point with distance 0 still wins on the index half),
// no per-round validity test.  Arg-max key = (distance bits, ~index): non-negative floats order like their bit patterns and
// ~index makes the FIRST maximum win (pytorch3d / the in-tree pin :577-614).
template <int NPER>
__device__ __forceinline__ void fps_rounds(const float* ux, const float* uy, const float* uz, int keep1, int n_drop, int TF,
                                           unsigned long long* red, int* sel) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwf = TF >> 5;
    float px[NPER], py[NPER], pz[NPER], dmin[NPER];
#pragma unroll
    for (int r = 0; r < NPER; ++r) {
        const int p = tid + r * TF;
        const bool ok = p < keep1;
        px[r] = ok ? ux[p] : 0.f; py[r] = ok ? uy[p] : 0.f; pz[r] = ok ? uz[p] : 0.f;
        dmin[r] = ok ? INFINITY : 0.f;
    }
    const unsigned nt = 0xffffffffu - (unsigned)tid;
    int last = 0;
    if (tid == 0) sel[0] = 0;
    for (int k = 1; k < n_drop; ++k) {
        const float lx = ux[last], ly = uy[last], lz = uz[last];
        unsigned kh[NPER];
#pragma unroll
        for (int r = 0; r < NPER; ++r) {
            const float d = dist2_rn(lx, ly, lz, px[r], py[r], pz[r]);
            dmin[r] = fminf(d, dmin[r]);
            kh[r] = __float_as_uint(dmin[r]);
        }
        unsigned bh = kh[0], bl = nt;                         // lower slots hold lower indices: strict > keeps the first maximum
#pragma unroll
        for (int r = 1; r < NPER; ++r) {
            const bool gt = kh[r] > bh;
            bh = gt ? kh[r] : bh;
            bl = gt ? nt - (unsigned)(r * TF) : bl;
        }
        unsigned long long* slot = red + (k & 1) * 32;
        {
            // warp arg-max: one redux on the distance bits; the index half needs a second one only when several lanes hold the
            // maximal distance (duplicated centres), otherwise the single winner lane publishes its own key
            const unsigned mh = __reduce_max_sync(FULL, bh);
            const unsigned tie = __ballot_sync(FULL, bh == mh);
            if (__popc(tie) == 1) {
                if (bh == mh) slot[wid] = ((unsigned long long)bh << 32) | bl;
            } else {
                const unsigned ml = __reduce_max_sync(FULL, bh == mh ? bl : 0u);
                if (lane == 0) slot[wid] = ((unsigned long long)mh << 32) | ml;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(TF) : "memory");
        unsigned long long v;
        if (nwf <= 8) {                                        // every thread combines the 4 / 8 warp results itself
            const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(slot);
            const ulonglong2 a = s2[0], b2 = s2[1];
            unsigned long long m0 = a.x > a.y ? a.x : a.y, m1 = b2.x > b2.y ? b2.x : b2.y;
            if (nwf == 8) {
                const ulonglong2 c = s2[2], d = s2[3];
                const unsigned long long m2 = c.x > c.y ? c.x : c.y, m3 = d.x > d.y ? d.x : d.y;
                m0 = m0 > m2 ? m0 : m2; m1 = m1 > m3 ? m1 : m3;
            }
            v = m0 > m1 ? m0 : m1;
        } else {
            const unsigned long long w = lane < nwf ? slot[lane] : 0ull;
            const unsigned wh = (unsigned)(w >> 32), wl = (unsigned)w;
            const unsigned mh = __reduce_max_sync(FULL, wh);
            const unsigned ml = __reduce_max_sync(FULL, wh == mh ? wl : 0u);
            v = ((unsigned long long)mh << 32) | ml;
        }
        last = (int)(0xffffffffu - (unsigned)(v & 0xffffffffull));
        if (tid == 0) sel[k] = last;
    }
}

// MAXT = 256: launches of 256 threads (keep1 <= 2048) get the register budget for up to 10 slots per thread, so that up to 1 280
// candidates (the shipped config keeps 1 210) run their rounds on 4 warps instead of 8 (cheaper cross-warp step).
template <int MAXT>
__global__ void __launch_bounds__(MAXT) cluster_dropout_kernel(const float* __restrict__ centres, const int32_t* __restrict__ idx, int M, int K,
                                       int keep1, int n_keep, int32_t* __restrict__ kept_src,
                                       float* __restrict__ kept_centres, int32_t* __restrict__ kept_idx,
                                       int32_t* __restrict__ drop_idx, int32_t* __restrict__ fps_sel) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_drop = keep1 - n_keep;
    unsigned long long* red = reinterpret_cast<unsigned long long*>(smem_raw);          // [2][32]
    float* ux = reinterpret_cast<float*>(red + 64);
    float* uy = ux + keep1;
    float* uz = uy + keep1;
    int* order = reinterpret_cast<int*>(uz + keep1);                                     // [M]
    int* start = order + M;                                                              // [K+2]
    int* sel = start + (K + 2);                                                          // [n_drop]
    int* kl = sel + n_drop;                                                              // [n_keep]
    unsigned short* pcs = reinterpret_cast<unsigned short*>(kl + n_keep);                // [M]
    unsigned char* flag = reinterpret_cast<unsigned char*>(pcs + M);                     // [keep1]

    const int32_t* I = idx + (size_t)b * M * K;
    const float* C = centres + (size_t)b * M * 3;
    for (int i = tid; i < K + 2; i += T) start[i] = 0;
    __syncthreads();
    for (int m = tid; m < M; m += T) {                 // :372 padding_counts
        int pc = 0;
        for (int k = 0; k < K; ++k) pc += (__ldg(I + (size_t)m * K + k) == -1);
        pcs[m] = (unsigned short)pc;
        atomicAdd(&start[pc + 1], 1);
    }
    __syncthreads();
    if (tid == 0) for (int k = 1; k <= K + 1; ++k) start[k] += start[k - 1];   // start[key] = first slot of key
    __syncthreads();
    if (wid == 0) {                                    // :378 argsort, pinned stable: ties keep ascending cluster id
        const unsigned lt = (1u << lane) - 1u;
        for (int m0 = 0; m0 < M; m0 += 32) {
            const int m = m0 + lane;
            const unsigned key = m < M ? pcs[m] : 0xffffu;
            const unsigned grp = __match_any_sync(FULL, key);
            const int rank = __popc(grp & lt);
            if (m < M) order[start[key] + rank] = m;
            __syncwarp();
            if (m < M && rank == 0) start[key] += __popc(grp);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = tid; i < keep1; i += T) {             // :379-385 keep the keep1 fullest clusters
        const int m = order[i];
        ux[i] = __ldg(C + m * 3); uy[i] = __ldg(C + m * 3 + 1); uz[i] = __ldg(C + m * 3 + 2);
        flag[i] = 0;
    }
    __syncthreads();
    // :393 farthest point sampling of n_drop centres (pytorch3d semantics; in-tree pin :577-614).  A round is a dependent
    // chain (distances -> arg-max -> next centre), so its latency is what counts: only the first TF threads take part
    // (TF = 128 / 256 / T by keep1: fewer warps make the cross-warp step cheaper, nper slots per thread actually used) and
    // they meet at a named barrier of their own.
    constexpr int PER128 = MAXT <= 256 ? 10 : FPS_PER;     // slots per thread up to which 4 warps take the rounds
    const int TF = keep1 <= 128 * PER128 ? 128 : (keep1 <= 256 * FPS_PER ? 256 : T);
    if (tid < TF) {
        switch ((keep1 + TF - 1) / TF) {               // slots per thread, compile-time inside the rounds
            case 9: if (MAXT <= 256) { fps_rounds<9>(ux, uy, uz, keep1, n_drop, TF, red, sel); break; }
            case 10: if (MAXT <= 256) { fps_rounds<10>(ux, uy, uz, keep1, n_drop, TF, red, sel); break; }
            case 1: fps_rounds<1>(ux, uy, uz, keep1, n_drop, TF, red, sel); break;
            case 2: fps_rounds<2>(ux, uy, uz, keep1, n_drop, TF, red, sel); break;
            case 3: fps_rounds<3>(ux, uy, uz, keep1, n_drop, TF, red, sel); break;
            case 4: fps_rounds<4>(ux, uy, uz, keep1, n_drop, TF, red, sel); break;
            case 5: fps_rounds<5>(ux, uy, uz, keep1, n_drop, TF, red, sel); break;
            case 6: fps_rounds<6>(ux, uy, uz, keep1, n_drop, TF, red, sel); break;
            case 7: fps_rounds<7>(ux, uy, uz, keep1, n_drop, TF, red, sel); break;
            default: fps_rounds<FPS_PER>(ux, uy, uz, keep1, n_drop, TF, red, sel); break;
        }
    }
    __syncthreads();
    for (int k = tid; k < n_drop; k += T) flag[sel[k]] = 1;
    __syncthreads();
    if (wid == 0) {                                    // :395-408 ascending complement, truncated to n_keep
        const unsigned lt = (1u << lane) - 1u;
        int cnt = 0;
        for (int i0 = 0; i0 < keep1 && cnt < n_keep; i0 += 32) {
            const int i = i0 + lane;
            const bool kf = i < keep1 && !flag[i];
            const unsigned mask = __ballot_sync(FULL, kf);
            const int slot = cnt + __popc(mask & lt);
            if (kf && slot < n_keep) kl[slot] = i;
            cnt += __popc(mask);
        }
    }
    __syncthreads();
    for (int s = tid; s < n_keep; s += T) {            // :414-416
        const int i = kl[s];
        kept_src[(size_t)b * n_keep + s] = order[i];
        float* kc = kept_centres + ((size_t)b * n_keep + s) * 3;
        kc[0] = ux[i]; kc[1] = uy[i]; kc[2] = uz[i];
    }
    for (int e = tid; e < n_keep * K; e += T) {
        const int s = e / K, k = e - s * K;
        kept_idx[(size_t)b * n_keep * K + e] = __ldg(I + (size_t)order[kl[s]] * K + k);
    }
    for (int e = tid; e < n_drop * K; e += T) {        // :417-418
        const int s = e / K, k = e - s * K;
        drop_idx[(size_t)b * n_drop * K + e] = __ldg(I + (size_t)order[sel[s]] * K + k);
    }
    if (fps_sel != nullptr) for (int k = tid; k < n_drop; k += T) fps_sel[(size_t)b * n_drop + k] = sel[k];
}

}  // namespace pt

using namespace pt;

extern "C" size_t pt_minmax_ws_bytes(int B, int N) {
    return (size_t)B * ceil_div(N, MM_POINTS_PER_BLOCK) * 6 * sizeof(float);
}

extern "C" int pt_minmax_centres(const float* points, int B, int N, int gs, const float* lin, float margin, float* mn,
                                 float* mx, float* centres, void* ws, size_t ws_bytes, pt_stream_t stream) {
    PT_REQUIRE(B > 0 && N > 0 && gs > 0, "pt_minmax_centres: B=%d N=%d gs=%d", B, N, gs);
    PT_REQUIRE(points && lin && mn && mx && centres && ws, "pt_minmax_centres: null pointer");
    if (ws_bytes < pt_minmax_ws_bytes(B, N)) { set_error("pt_minmax_centres: workspace %zu < %zu", ws_bytes, pt_minmax_ws_bytes(B, N)); return PT_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const int nblk = ceil_div(N, MM_POINTS_PER_BLOCK), M = gs * gs * gs;
    { ProfScope prof_(PROF_MINMAX, s); minmax_partial_kernel<<<dim3(nblk, B), MM_THREADS, 0, s>>>(points, N, (float*)ws); }
    PT_LAUNCH_CHECK();
    { ProfScope prof_(PROF_CENTRES, s); centres_kernel<<<dim3(ceil_div(M, 256), B), 256, 0, s>>>((const float*)ws, nblk, gs, M, lin, margin, mn, mx, centres); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

extern "C" int pt_ball_query_firstk(const float* centres, const float* points, int B, int M, int N, int K, float radius,
                                    int32_t* idx, int32_t* pad_counts, pt_stream_t stream) {
    PT_REQUIRE(B > 0 && M > 0 && N > 0 && K > 0, "pt_ball_query_firstk: B=%d M=%d N=%d K=%d", B, M, N, K);
    PT_REQUIRE(centres && points && idx, "pt_ball_query_firstk: null pointer");
    const long long total = (long long)B * M;
    { ProfScope prof_(PROF_BALL_QUERY, (cudaStream_t)stream); ball_query_kernel<<<(unsigned)((total + BQ_WARPS - 1) / BQ_WARPS), BQ_WARPS * 32, 0, (cudaStream_t)stream>>>(
        centres, points, M, N, K, radius * radius, idx, pad_counts, total); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

static int mlp_grid(int clusters) {
    // ~4 clusters per warp amortises the per-warp weight load; cap at a few waves of the 148 SMs
    int blocks = ceil_div(ceil_div(clusters, 4), MLP_WARPS);
    return blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks);
}

extern "C" int pt_offset_net_fused(const float* points, const int32_t* idx, const float* centres0, const float* mn,
                                   const float* mx, const float* conv_w, const float* conv_b, const float* bn_scale,
                                   const float* bn_shift, const float* map_w, int B, int M, int N, int K, int H,
                                   float margin, float* centres_out, float* raw_offsets, pt_stream_t stream) {
    PT_REQUIRE(H == MLP_H, "pt_offset_net_fused: hidden width %d unsupported (reference hard-wires 256, :31)", H);
    PT_REQUIRE(B > 0 && M > 0 && N > 0 && K > 0, "pt_offset_net_fused: bad shape");
    PT_REQUIRE(points && idx && centres0 && mn && mx && conv_w && conv_b && bn_scale && bn_shift && map_w && centres_out,
               "pt_offset_net_fused: null pointer");
    { ProfScope prof_(PROF_OFFSET_NET, (cudaStream_t)stream); cluster_mlp_kernel<false><<<mlp_grid(B * M), MLP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        points, idx, centres0, mn, mx, conv_w, conv_b, bn_scale, bn_shift, map_w, B, M, N, K, margin, centres_out, raw_offsets); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

extern "C" int pt_point_encoder_fused(const float* points, const int32_t* kept_idx, const float* kept_centres,
                                      const float* conv_w, const float* conv_b, const float* bn_scale,
                                      const float* bn_shift, int B, int n, int N, int K, int H, float* point_proxy,
                                      pt_stream_t stream) {
    PT_REQUIRE(H == MLP_H, "pt_point_encoder_fused: width %d unsupported (reference hard-wires 256, :110,:302)", H);
    PT_REQUIRE(B > 0 && n > 0 && N > 0 && K > 0, "pt_point_encoder_fused: bad shape");
    PT_REQUIRE(points && kept_idx && kept_centres && conv_w && conv_b && bn_scale && bn_shift && point_proxy,
               "pt_point_encoder_fused: null pointer");
    { ProfScope prof_(PROF_ENCODER, (cudaStream_t)stream); cluster_mlp_kernel<true><<<mlp_grid(B * n), MLP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        points, kept_idx, kept_centres, nullptr, nullptr, conv_w, conv_b, bn_scale, bn_shift, nullptr, B, n, N, K, 0.f,
        point_proxy, nullptr); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

extern "C" int pt_cluster_dropout(const float* centres, const int32_t* idx, int B, int M, int K, int keep1, int n_keep,
                                  int32_t* kept_src, float* kept_centres, int32_t* kept_idx, int32_t* drop_idx,
                                  int32_t* fps_sel, pt_stream_t stream) {
    PT_REQUIRE(B > 0 && M > 0 && K > 0 && K < 65535, "pt_cluster_dropout: bad shape");
    PT_REQUIRE(keep1 <= M && n_keep >= 1 && keep1 - n_keep >= 1,
               "pt_cluster_dropout: need 1 <= n_keep < keep1 <= M (got keep1=%d n_keep=%d M=%d); the reference's FPS needs n_drop >= 1 (:390-393)",
               keep1, n_keep, M);
    PT_REQUIRE(keep1 <= 1024 * FPS_PER, "pt_cluster_dropout: keep1=%d > %d unsupported", keep1, 1024 * FPS_PER);
    PT_REQUIRE(centres && idx && kept_src && kept_centres && kept_idx && drop_idx, "pt_cluster_dropout: null pointer");
    const int n_drop = keep1 - n_keep;
    const int T = keep1 <= 256 * FPS_PER ? 256 : 1024;
    size_t smem = 64 * sizeof(unsigned long long) + (size_t)3 * keep1 * sizeof(float) +
                  ((size_t)M + (K + 2) + n_drop + n_keep) * sizeof(int) + (size_t)M * sizeof(unsigned short) + keep1;
    smem = align_up(smem, 16);
    PT_REQUIRE(smem <= 227 * 1024, "pt_cluster_dropout: M=%d needs %zu B shared memory", M, smem);
    if (smem > 48 * 1024) {
        PT_CUDA_OK(cudaFuncSetAttribute(cluster_dropout_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PT_CUDA_OK(cudaFuncSetAttribute(cluster_dropout_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    auto kern = T == 256 ? cluster_dropout_kernel<256> : cluster_dropout_kernel<1024>;
    { ProfScope prof_(PROF_DROPOUT, (cudaStream_t)stream); kern<<<B, T, smem, (cudaStream_t)stream>>>(centres, idx, M, K, keep1, n_keep, kept_src, kept_centres,
                                                                 kept_idx, drop_idx, fps_sel); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}
