// S9 image proxies: get_img_proxy (:335-342) + AttentionPool2d.forward (:154-177) in single-query form.
//
// The reference runs a 1x1 conv (C=512 -> c=256) and a full 226-token multi-head attention per view and keeps only
// token 0 (:177).  Token 0's output depends on the image features only through (a) the per-channel spatial mean and
// (b) per-head softmax-weighted sums of the RAW feature columns, because the conv, k- and v-projections are linear:
//     q      = Wq(Wc xbar + bc + pos_0) + bq                         = W_qc xbar + q0
//     s_h,t  = scale * q_h . (Wk_h (Wc x_t + bc + pos_t) + bk_h)      = scale * (w_eff_h . x_t + q_h . g_k[t,h]) (+const)
//     out_h  = sum_t a_h,t (Wv_h (Wc x_t + bc + pos_t) + bv_h)        = W_vc_h y_h + sum_t a_h,t h_v[t,h]
// with x_0 = xbar, y_h = sum_t a_h,t x_t.  So the (BV, C, HW) tensor is streamed twice (mean; scores+weighted sum) and
// everything else is small GEMMs on (BV, *) matrices.  HBM-bound stage: 2*C*HW*e bytes per view.
#include "common.cuh"

#include <math.h>

namespace pt {

int launch_gemm_f32_strided(const float* A, const float* W, const float* bias, const float* residual, int act, int M, int N,
                            int K, float* C, int lda, int ldw, int ldc, int batch, long long bsA, long long bsW,
                            long long bsC, long long bsBias, cudaStream_t s);
int launch_layernorm(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                     float* out, cudaStream_t s);

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------ pass A: spatial mean
// fp32: one warp per channel row.  bf16: one warp per channel PAIR (2*HW elements = HW aligned 32-bit words; word w holds
// elements 2w,2w+1 of the pair stream, element e belongs to the first channel iff e < HW).
template <bool BF16>
__global__ void __launch_bounds__(256) img_mean_kernel(const void* __restrict__ img, int C, int HW, long long rows,
                                                       float* __restrict__ xbar) {
    const int lane = threadIdx.x & 31;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    if (!BF16) {
        const float* X = (const float*)img;
        for (long long r = wg; r < rows; r += nw) {
            float s = 0.f;
            for (int j = lane; j < HW; j += 32) s += __ldg(X + r * HW + j);
            s = warp_sum(s);
            if (lane == 0) xbar[r] = s / (float)HW;
        }
    } else {
        const uint32_t* X = (const uint32_t*)img;
        for (long long pr = wg; pr < rows / 2; pr += nw) {
            float s0 = 0.f, s1 = 0.f;
            for (int w = lane; w < HW; w += 32) {
                const uint32_t v = __ldg(X + pr * HW + w);
                const float lo = bf16_lo(v), hi = bf16_hi(v);
                if (2 * w < HW) s0 += lo; else s1 += lo;
                if (2 * w + 1 < HW) s0 += hi; else s1 += hi;
            }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (lane == 0) { xbar[2 * pr] = s0 / (float)HW; xbar[2 * pr + 1] = s1 / (float)HW; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ pass B: pool
// One CTA per view.  Scores for the HW spatial tokens (thread <-> position, channels streamed through shared memory in
// slabs of PB_CH), score of the mean token, per-head softmax over T = HW+1 tokens (warp <-> head), then the weighted
// feature sums y[h][ch] (thread <-> (channel, position-quarter)).  HEADS is fixed at 8 (one warp per head).
constexpr int PB_THREADS = 256;
constexpr int PB_CH = 64;
constexpr int PB_HEADS = 8;

template <bool BF16>
__global__ void __launch_bounds__(PB_THREADS) img_pool_kernel(const void* __restrict__ img, const float* __restrict__ xbar,
                                                              const float* __restrict__ w_eff, const float* __restrict__ cterm,
                                                              int C, int HW, int Tp, float scale, float* __restrict__ y,
                                                              float* __restrict__ attn) {
    extern __shared__ __align__(16) float sm[];
    const int HWp = HW | 1;                       // odd pitch: conflict-free both along positions and along channels
    const int T = HW + 1;
    float* xs = sm;                               // [PB_CH][HWp]
    float* ws = xs + (size_t)PB_CH * HWp;         // [PB_CH][8]   w_eff slab, head-contiguous
    float* sa = ws + PB_CH * PB_HEADS;            // [Tp][8]      scores, then probabilities (token-major, head-contiguous)
    float* red = sa + (size_t)Tp * PB_HEADS;      // [4][8][PB_CH] partial y / [8][8] partial dots
    const long long bv = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* we = w_eff + bv * PB_HEADS * C;  // (8, C)
    const float* xb = xbar + bv * C;
    const size_t tile = (size_t)bv * C * HW;

    auto load_slab = [&](int c0) {
        if (BF16) {
            const uint32_t* X = (const uint32_t*)img + (tile + (size_t)c0 * HW) / 2;   // (C*HW even) -> aligned
            for (int w = tid; w < PB_CH * HW / 2; w += PB_THREADS) {
                const uint32_t v = __ldg(X + w);
                const int e0 = 2 * w, ch0 = e0 / HW, j0 = e0 - ch0 * HW;
                xs[ch0 * HWp + j0] = bf16_lo(v);
                const int e1 = e0 + 1, ch1 = e1 / HW, j1 = e1 - ch1 * HW;
                xs[ch1 * HWp + j1] = bf16_hi(v);
            }
        } else {
            const float* X = (const float*)img + tile + (size_t)c0 * HW;
            for (int i = tid; i < PB_CH * HW; i += PB_THREADS) {
                const int ch = i / HW, j = i - ch * HW;
                xs[ch * HWp + j] = __ldg(X + i);
            }
        }
        for (int i = tid; i < PB_CH * PB_HEADS; i += PB_THREADS) {
            const int ch = i >> 3, h = i & 7;
            ws[i] = __ldg(we + (size_t)h * C + c0 + ch);
        }
    };

    // ---- scores of the spatial tokens
    float s[PB_HEADS];
#pragma unroll
    for (int h = 0; h < PB_HEADS; ++h) s[h] = 0.f;
    for (int c0 = 0; c0 < C; c0 += PB_CH) {
        __syncthreads();
        load_slab(c0);
        __syncthreads();
        for (int j = tid; j < HW; j += PB_THREADS) {     // HW <= PB_THREADS: at most one position per thread
            for (int ch = 0; ch < PB_CH; ++ch) {
                const float x = xs[ch * HWp + j];
                const float4 w0 = *reinterpret_cast<const float4*>(ws + ch * 8), w1 = *reinterpret_cast<const float4*>(ws + ch * 8 + 4);
                s[0] = fmaf(w0.x, x, s[0]); s[1] = fmaf(w0.y, x, s[1]); s[2] = fmaf(w0.z, x, s[2]); s[3] = fmaf(w0.w, x, s[3]);
                s[4] = fmaf(w1.x, x, s[4]); s[5] = fmaf(w1.y, x, s[5]); s[6] = fmaf(w1.z, x, s[6]); s[7] = fmaf(w1.w, x, s[7]);
            }
        }
    }
    if (tid < HW) {
#pragma unroll
        for (int h = 0; h < PB_HEADS; ++h) sa[(tid + 1) * 8 + h] = scale * (s[h] + __ldg(cterm + (bv * PB_HEADS + h) * Tp + tid + 1));
    }
    // ---- score of the mean token (token 0): w_eff_h . xbar
    {
        float d[PB_HEADS];
#pragma unroll
        for (int h = 0; h < PB_HEADS; ++h) d[h] = 0.f;
        for (int ch = tid; ch < C; ch += PB_THREADS) {
            const float x = __ldg(xb + ch);
#pragma unroll
            for (int h = 0; h < PB_HEADS; ++h) d[h] = fmaf(__ldg(we + (size_t)h * C + ch), x, d[h]);
        }
#pragma unroll
        for (int h = 0; h < PB_HEADS; ++h) d[h] = warp_sum(d[h]);
        if (lane == 0)
#pragma unroll
            for (int h = 0; h < PB_HEADS; ++h) red[wid * 8 + h] = d[h];
        __syncthreads();
        if (tid < PB_HEADS) {
            float v = 0.f;
            for (int w = 0; w < PB_THREADS / 32; ++w) v += red[w * 8 + tid];
            sa[tid] = scale * (v + __ldg(cterm + (bv * PB_HEADS + tid) * Tp));
        }
    }
    __syncthreads();
    // ---- softmax over the T tokens, warp h <-> head h
    {
        const int h = wid;
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, sa[t * 8 + h]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int t = lane; t < T; t += 32) { const float p = expf(sa[t * 8 + h] - mx); sa[t * 8 + h] = p; sum += p; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        float* arow = attn + (bv * PB_HEADS + h) * Tp;
        for (int t = lane; t < Tp; t += 32) {
            const float p = t < T ? sa[t * 8 + h] * inv : 0.f;
            if (t < T) sa[t * 8 + h] = p;
            arow[t] = p;
        }
    }
    // ---- weighted feature sums
    const int ch = tid & (PB_CH - 1), js = tid / PB_CH;          // 4 position quarters
    const int jq = (HW + 3) / 4, jb = js * jq, je = min(HW, jb + jq);
    for (int c0 = 0; c0 < C; c0 += PB_CH) {
        __syncthreads();
        load_slab(c0);
        __syncthreads();
        float a[PB_HEADS];
#pragma unroll
        for (int h = 0; h < PB_HEADS; ++h) a[h] = 0.f;
        for (int j = jb; j < je; ++j) {
            const float x = xs[ch * HWp + j];
            const float4 p0 = *reinterpret_cast<const float4*>(sa + (j + 1) * 8), p1 = *reinterpret_cast<const float4*>(sa + (j + 1) * 8 + 4);
            a[0] = fmaf(p0.x, x, a[0]); a[1] = fmaf(p0.y, x, a[1]); a[2] = fmaf(p0.z, x, a[2]); a[3] = fmaf(p0.w, x, a[3]);
            a[4] = fmaf(p1.x, x, a[4]); a[5] = fmaf(p1.y, x, a[5]); a[6] = fmaf(p1.z, x, a[6]); a[7] = fmaf(p1.w, x, a[7]);
        }
#pragma unroll
        for (int h = 0; h < PB_HEADS; ++h) red[(js * 8 + h) * PB_CH + ch] = a[h];
        __syncthreads();
        for (int i = tid; i < PB_HEADS * PB_CH; i += PB_THREADS) {
            const int h = i / PB_CH, cc = i - h * PB_CH;
            const float v = ((red[(0 * 8 + h) * PB_CH + cc] + red[(1 * 8 + h) * PB_CH + cc]) +
                             (red[(2 * 8 + h) * PB_CH + cc] + red[(3 * 8 + h) * PB_CH + cc])) +
                            sa[h] * __ldg(xb + c0 + cc);                     // token 0 contributes a_h,0 * xbar
            y[(bv * PB_HEADS + h) * C + c0 + cc] = v;
        }
    }
}

size_t img_attnpool_tc_ws_bytes(int BV);
bool img_attnpool_tc_supported(int img_dtype, const pt_img_pool_params* p, int C, int HW, int c, int heads);
int launch_img_attnpool_tc(const void* img_feat, int img_dtype, const pt_img_pool_params* p, int BV, float* img_proxy, void* ws, size_t ws_bytes, int stages,
                           cudaStream_t s);

static size_t pool_smem_bytes(int HW, int Tp) {
    return ((size_t)PB_CH * (HW | 1) + PB_CH * PB_HEADS + (size_t)Tp * PB_HEADS + 4 * PB_HEADS * PB_CH) * sizeof(float);
}

struct ImgWs {
    float *xbar, *q, *w_eff, *cterm, *y, *attn, *z, *o;
    size_t total;
};

static ImgWs carve(void* ws, int BV, int C, int HW, int c, int heads) {
    const int Tp = (HW + 1 + 3) / 4 * 4;
    ImgWs r;
    size_t off = 0;
    auto take = [&](size_t n) { float* p = ws ? (float*)((char*)ws + off) : nullptr; off += align_up(n * sizeof(float), 256); return p; };
    r.xbar = take((size_t)BV * C);
    r.q = take((size_t)BV * c);
    r.w_eff = take((size_t)BV * heads * C);
    r.cterm = take((size_t)BV * heads * Tp);
    r.y = take((size_t)BV * heads * C);
    r.attn = take((size_t)BV * heads * Tp);
    r.z = take((size_t)BV * c);
    r.o = take((size_t)BV * c);
    r.total = off;
    return r;
}

}  // namespace pt

using namespace pt;

extern "C" size_t pt_img_attnpool_ws_bytes(int BV, int C, int HW, int c, int heads) {
    const size_t a = carve(nullptr, BV, C, HW, c, heads).total, b = img_attnpool_tc_ws_bytes(BV);
    return a > b ? a : b;
}

// Host-prepared folded weights (see pt_img_pool_params): w_qc (c,C); q0 (c); w_kc is consumed TRANSPOSED PER HEAD as
// w_kcT (heads, C, hd) so that w_eff_h = q_h @ w_kcT[h]^T is an NT GEMM; g_k (T,c) padded to Tp rows of zeros;
// w_vc (c,C); h_v consumed TRANSPOSED PER HEAD as h_vT (heads, hd, Tp) zero-padded.
extern "C" int pt_img_attnpool(const void* img_feat, int img_dtype, const pt_img_pool_params* p, int BV, int C, int HW,
                               int c, int heads, float* img_proxy, void* ws, size_t ws_bytes, pt_stream_t stream) {
    return pt_img_attnpool_stage(img_feat, img_dtype, p, BV, C, HW, c, heads, img_proxy, ws, ws_bytes,
                                 PT_IMG_STAGE_FRONT | PT_IMG_STAGE_BACK, stream);
}

extern "C" int pt_img_attnpool_stage(const void* img_feat, int img_dtype, const pt_img_pool_params* p, int BV, int C, int HW,
                                     int c, int heads, float* img_proxy, void* ws, size_t ws_bytes, int stages, pt_stream_t stream) {
    PT_REQUIRE(img_feat && p && img_proxy && ws, "pt_img_attnpool: null pointer");
    PT_REQUIRE(img_dtype == PT_DTYPE_F32 || img_dtype == PT_DTYPE_BF16 || img_dtype == PT_DTYPE_F16, "pt_img_attnpool: dtype %d", img_dtype);
    PT_REQUIRE(stages >= 1 && stages <= 3, "pt_img_attnpool_stage: stages=%d", stages);
    if (BV > 0 && img_attnpool_tc_supported(img_dtype, p, C, HW, c, heads))      // 16-bit tensor-core fast path (imgpool_tc.cu)
        return launch_img_attnpool_tc(img_feat, img_dtype, p, BV, img_proxy, ws, ws_bytes, stages, (cudaStream_t)stream);
    PT_REQUIRE(img_dtype != PT_DTYPE_F16, "pt_img_attnpool: fp16 features need the tcgen05 pooling path (shipped geometry C=512, 15x15, c=256, 8 heads, "
                                          "split weights, variant PT_POOL_VARIANT_UMMA); convert to fp32 otherwise");
    if (!(stages & PT_IMG_STAGE_BACK)) return PT_OK;                              // generic path: everything runs in the back stage
    PT_REQUIRE(heads == PB_HEADS, "pt_img_attnpool: heads=%d unsupported (8)", heads);
    PT_REQUIRE(BV > 0 && C % PB_CH == 0 && c % heads == 0 && HW >= 1 && HW <= PB_THREADS,
               "pt_img_attnpool: BV=%d C=%d HW=%d c=%d unsupported", BV, C, HW, c);
    PT_REQUIRE(img_dtype == PT_DTYPE_F32 || ((size_t)C * HW) % 2 == 0, "pt_img_attnpool: bf16 needs an even C*HW");
    const int hd = c / heads, T = HW + 1, Tp = (T + 3) / 4 * 4;
    PT_REQUIRE(hd % 4 == 0, "pt_img_attnpool: head_dim %d", hd);
    ImgWs w = carve(ws, BV, C, HW, c, heads);
    if (ws_bytes < w.total) { set_error("pt_img_attnpool: workspace %zu < %zu", ws_bytes, w.total); return PT_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const bool bf = img_dtype == PT_DTYPE_BF16;
    const long long rows = (long long)BV * C;
    {   // pass A
        const long long units = bf ? rows / 2 : rows;
        int grid = (int)((units + 7) / 8 < 148 * 8 ? (units + 7) / 8 : 148 * 8);
        if (bf) { ProfScope prof_(PROF_IMG_MEAN, s); img_mean_kernel<true><<<grid, 256, 0, s>>>(img_feat, C, HW, rows, w.xbar); }
        else { ProfScope prof_(PROF_IMG_MEAN, s); img_mean_kernel<false><<<grid, 256, 0, s>>>(img_feat, C, HW, rows, w.xbar); }
        PT_LAUNCH_CHECK();
    }
    int rc;
    // q = xbar @ w_qc^T + q0
    if ((rc = launch_gemm_f32_strided(w.xbar, p->w_qc, p->q0, nullptr, 0, BV, c, C, w.q, C, C, c, 1, 0, 0, 0, 0, s))) return rc;
    // w_eff[:, h, :] = q[:, h] @ w_kcT[h]^T      (batch over heads)
    if ((rc = launch_gemm_f32_strided(w.q, p->w_kc, nullptr, nullptr, 0, BV, C, hd, w.w_eff, c, hd, heads * C, heads, hd,
                                      (long long)C * hd, C, 0, s))) return rc;
    // cterm[:, h, t] = q[:, h] . g_k[t, h]
    if ((rc = launch_gemm_f32_strided(w.q, p->g_k, nullptr, nullptr, 0, BV, Tp, hd, w.cterm, c, c, heads * Tp, heads, hd, hd,
                                      Tp, 0, s))) return rc;
    {   // pass B
        const size_t smem = pool_smem_bytes(HW, Tp);
        const float scale = (float)(1.0 / sqrt((double)hd));
        if (bf) {
            PT_CUDA_OK(cudaFuncSetAttribute(img_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            { ProfScope prof_(PROF_IMG_POOL, s); img_pool_kernel<true><<<BV, PB_THREADS, smem, s>>>(img_feat, w.xbar, w.w_eff, w.cterm, C, HW, Tp, scale, w.y, w.attn); }
        } else {
            PT_CUDA_OK(cudaFuncSetAttribute(img_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            { ProfScope prof_(PROF_IMG_POOL, s); img_pool_kernel<false><<<BV, PB_THREADS, smem, s>>>(img_feat, w.xbar, w.w_eff, w.cterm, C, HW, Tp, scale, w.y, w.attn); }
        }
        PT_LAUNCH_CHECK();
    }
    // z[:, h] = y[:, h, :] @ w_vc[h rows]^T ; then += attn[:, h, :] @ h_vT[h]^T
    if ((rc = launch_gemm_f32_strided(w.y, p->w_vc, nullptr, nullptr, 0, BV, hd, C, w.z, heads * C, C, c, heads, C,
                                      (long long)hd * C, hd, 0, s))) return rc;
    if ((rc = launch_gemm_f32_strided(w.attn, p->h_v, nullptr, w.z, 0, BV, hd, Tp, w.z, heads * Tp, Tp, c, heads, Tp,
                                      (long long)hd * Tp, hd, 0, s))) return rc;
    if ((rc = launch_gemm_f32_strided(w.z, p->cproj_w, p->cproj_b, nullptr, 0, BV, c, c, w.o, c, c, c, 1, 0, 0, 0, 0, s))) return rc;
    return launch_layernorm(w.o, p->ln_w, p->ln_b, nullptr, 1, BV, c, img_proxy, s);
}
