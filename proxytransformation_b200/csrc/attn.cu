// Two-stage proxy attention core of ProxyAttention.forward (:225-252), one CTA per (scene, head).
//   stage 1 (proxy as query, :232-238):  Pv = softmax_n((Pt*scale) K^T) V          (l x hd), unmasked
//   stage 2 (proxy as key,   :241-250):  O  = softmax_l(mask((Q*scale) Pt^T)) Pv   (n x hd)
// K, V, Pt, Pv of the head stay in shared memory; cost is O(n*l*hd), never n^2.  fp32 CUDA cores, exact expf.
#include "common.cuh"

#include <math.h>

namespace pt {

constexpr int AT_THREADS = 256;
constexpr int AT_WARPS = AT_THREADS / 32;

// qkv: (B*n, 3c) rows [Q | K | V]; pt: (B*l, c); mask: (B,l) uint8 or null; o: (B*n, c)
template <int HD>
__global__ void __launch_bounds__(AT_THREADS) proxy_attention_kernel(const float* __restrict__ qkv,
                                                                     const float* __restrict__ pt_tok,
                                                                     const uint8_t* __restrict__ mask, int n, int l, int c,
                                                                     float scale, float* __restrict__ o) {
    extern __shared__ __align__(16) float sm[];
    constexpr int LD = HD + 1;
    const int b = blockIdx.y, h = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float* sK = sm;                     // [n][LD]
    float* sV = sK + (size_t)n * LD;    // [n][LD]
    float* sP = sV + (size_t)n * LD;    // [l][LD]  proxy tokens of this head
    float* sPv = sP + (size_t)l * LD;   // [l][LD]
    float* sRow = sPv + (size_t)l * LD; // [AT_WARPS][max(n,l)] score / probability scratch per warp
    float* sMask = sRow + (size_t)AT_WARPS * max(n, l);   // [l] additive? no: 1 = keep, 0 = masked
    const int nmax = max(n, l);

    for (int i = tid; i < n * HD; i += AT_THREADS) {
        const int j = i / HD, e = i - j * HD;
        const float* row = qkv + ((size_t)b * n + j) * 3 * c + h * HD + e;
        sK[j * LD + e] = row[c];
        sV[j * LD + e] = row[2 * c];
    }
    for (int i = tid; i < l * HD; i += AT_THREADS) {
        const int j = i / HD, e = i - j * HD;
        sP[j * LD + e] = pt_tok[((size_t)b * l + j) * c + h * HD + e];
    }
    for (int i = tid; i < l; i += AT_THREADS) sMask[i] = (mask == nullptr || mask[(size_t)b * l + i]) ? 1.f : 0.f;
    __syncthreads();

    float* row = sRow + (size_t)wid * nmax;
    // ---- stage 1: one warp per proxy token i
    for (int i = wid; i < l; i += AT_WARPS) {
        float q[HD];
#pragma unroll
        for (int e = 0; e < HD; ++e) q[e] = sP[i * LD + e] * scale;      // (proxy_tokens * scale) (:232)
        float mx = -INFINITY;
        for (int j = lane; j < n; j += 32) {
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < HD; ++e) s = fmaf(q[e], sK[j * LD + e], s);
            row[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < n; j += 32) {
            const float p = expf(row[j] - mx);
            row[j] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        __syncwarp();
        const float inv = 1.0f / sum;
        // Pv[i][e] = sum_j p_j V[j][e]; lanes own e (and split j when HD < 32)
        constexpr int JS = 32 / (HD < 32 ? HD : 32);      // j-splits per warp
        const int e = lane % (HD < 32 ? HD : 32), js = lane / (HD < 32 ? HD : 32);
        for (int e0 = 0; e0 < HD; e0 += 32) {
            float acc = 0.f;
            for (int j = js; j < n; j += JS) acc = fmaf(row[j], sV[j * LD + e0 + e], acc);
#pragma unroll
            for (int off = 16; off >= (HD < 32 ? HD : 32); off >>= 1) acc += __shfl_xor_sync(FULL, acc, off);
            if (js == 0) sPv[i * LD + e0 + e] = acc * inv;
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- stage 2: one warp per point-proxy row j
    for (int j = wid; j < n; j += AT_WARPS) {
        const float* qrow = qkv + ((size_t)b * n + j) * 3 * c + h * HD;
        float q[HD];
#pragma unroll
        for (int e = 0; e < HD; ++e) q[e] = __ldg(qrow + e) * scale;      // (q * scale) (:241)
        float mx = -INFINITY;
        for (int i = lane; i < l; i += 32) {
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < HD; ++e) s = fmaf(q[e], sP[i * LD + e], s);
            if (sMask[i] == 0.f) s = -1e9f;                                // masked_fill(~mask, -1e9) (:247)
            row[i] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int i = lane; i < l; i += 32) {
            const float p = expf(row[i] - mx);
            row[i] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        __syncwarp();
        const float inv = 1.0f / sum;
        constexpr int JS = 32 / (HD < 32 ? HD : 32);
        const int e = lane % (HD < 32 ? HD : 32), js = lane / (HD < 32 ? HD : 32);
        for (int e0 = 0; e0 < HD; e0 += 32) {
            float acc = 0.f;
            for (int i = js; i < l; i += JS) acc = fmaf(row[i], sPv[i * LD + e0 + e], acc);
#pragma unroll
            for (int off = 16; off >= (HD < 32 ? HD : 32); off >>= 1) acc += __shfl_xor_sync(FULL, acc, off);
            if (js == 0) o[((size_t)b * n + j) * c + h * HD + e0 + e] = acc * inv;
        }
        __syncwarp();
    }
}

size_t attention_smem_bytes(int n, int l, int hd) {
    const int ld = hd + 1;
    return ((size_t)2 * n * ld + (size_t)2 * l * ld + (size_t)AT_WARPS * (n > l ? n : l) + l) * sizeof(float);
}

int launch_proxy_attention(const float* qkv, const float* pt_tok, const uint8_t* mask, int B, int n, int l, int c,
                           int heads, float* o, cudaStream_t s) {
    PT_REQUIRE(c % heads == 0, "attention: c=%d not divisible by heads=%d", c, heads);
    const int hd = c / heads;
    const size_t smem = attention_smem_bytes(n, l, hd);
    PT_REQUIRE(smem <= 227 * 1024, "attention: n=%d l=%d needs %zu B of shared memory", n, l, smem);
    const float scale = (float)(1.0 / sqrt((double)hd));   // python float head_dim ** -0.5 (:186), rounded to fp32 once
    dim3 grid(heads, B);
#define PT_AT_CASE(HD)                                                                                                   \
    case HD:                                                                                                             \
        if (smem > 48 * 1024) PT_CUDA_OK(cudaFuncSetAttribute(proxy_attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        { ProfScope prof_(PROF_ATTENTION, s); proxy_attention_kernel<HD><<<grid, AT_THREADS, smem, s>>>(qkv, pt_tok, mask, n, l, c, scale, o); }                 \
        break;
    switch (hd) {
        PT_AT_CASE(8) PT_AT_CASE(16) PT_AT_CASE(32) PT_AT_CASE(64)
        default: PT_REQUIRE(false, "attention: head_dim=%d unsupported (8/16/32/64)", hd);
    }
#undef PT_AT_CASE
    PT_LAUNCH_CHECK();
    return PT_OK;
}

}  // namespace pt
