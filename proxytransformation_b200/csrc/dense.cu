// Dense building blocks of the ProxyBlock stage (S7/S8): LayerNorm(+position bias), fp32 CUDA-core NT GEMM with fused
// bias/GELU/residual epilogue, bf16 hi/lo operand splitting for the tensor-core path, position-bias table, heads.
// Reference: embodiedscan/models/necks/preshape_norm_reverse_drop.py :206-276, :326-330, :445-455.
#include "common.cuh"

#include <math.h>

namespace pt {

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row, row held in registers (c <= 1024), two-pass mean/variance like torch (biased variance, eps 1e-5).
template <int CPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, const float* __restrict__ add,
                                                        int add_rows, int rows, float* __restrict__ out,
                                                        __nv_bfloat16* __restrict__ out_hi, long long out_plane) {
    constexpr int C = CPL * 32;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float v[CPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { v[i] = x[(size_t)row * C + lane + 32 * i]; s += v[i]; }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
    const float* ad = add ? add + (size_t)(row % add_rows) * C : nullptr;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const int ch = lane + 32 * i;
        float y = fmaf((v[i] - mean) * rstd, __ldg(w + ch), __ldg(b + ch));
        if (ad) y += __ldg(ad + ch);
        if (out != nullptr) out[(size_t)row * C + ch] = y;
        if (out_hi != nullptr) {                  // bf16 hi/lo planes: operand of the following tensor-core GEMM
            const __nv_bfloat16 h = __float2bfloat16_rn(y);
            out_hi[(size_t)row * C + ch] = h;
            out_hi[out_plane + (size_t)row * C + ch] = __float2bfloat16_rn(y - __bfloat162float(h));
        }
    }
}

// ------------------------------------------------------------------------------------------------ fp32 GEMM (NT)
// C[M,N] = act(A[M,K] * W[N,K]^T + bias) + residual.  128x128x16 tiles, 256 threads, 8x8 register micro-tiles,
// register-prefetched double buffering.  This is the CUDA-core path (exact fp32); the tcgen05 3xBF16 path lives in
// gemm_tc.cu and is selected when split weights are supplied.
constexpr int G_BM = 128, G_BN = 128, G_BK = 16, G_THREADS = 256;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Strided/batched form: blockIdx.z selects a batch (per-head slices of the image-pool projections); lda/ldw/ldc are
// row strides in elements, bs* the per-batch element offsets.  residual shares C's layout (and may alias C).
__global__ void __launch_bounds__(G_THREADS, 2) gemm_nt_f32_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                                const float* __restrict__ bias,
                                                                const float* residual, int act, int M, int N,
                                                                int K, float* C, int lda, int ldw, int ldc,
                                                                long long bsA, long long bsW, long long bsC, long long bsBias) {
    A += blockIdx.z * bsA; W += blockIdx.z * bsW; C += blockIdx.z * bsC;
    if (residual) residual += blockIdx.z * bsC;
    if (bias) bias += blockIdx.z * bsBias;
    __shared__ __align__(16) float As[2][G_BK][G_BM + 4];
    __shared__ __align__(16) float Ws[2][G_BK][G_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * G_BM, n0 = blockIdx.x * G_BN;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads, each 8 (m) x 8 (n)
    // global->smem: each thread moves 2 float4 of A and 2 of W per k-tile: row = (tid>>2) + 64*i, kq = (tid&3)*4
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    float4 ra[2], rw[2];
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lrow + 64 * i, k = k0 + lk;
            ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            rw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < M) {
                const float* p = A + (size_t)(m0 + r) * lda + k;
                if (k + 3 < K) ra[i] = *reinterpret_cast<const float4*>(p);
                else { if (k < K) ra[i].x = p[0]; if (k + 1 < K) ra[i].y = p[1]; if (k + 2 < K) ra[i].z = p[2]; }
            }
            if (n0 + r < N) {
                const float* p = W + (size_t)(n0 + r) * ldw + k;
                if (k + 3 < K) rw[i] = __ldg(reinterpret_cast<const float4*>(p));
                else { if (k < K) rw[i].x = p[0]; if (k + 1 < K) rw[i].y = p[1]; if (k + 2 < K) rw[i].z = p[2]; }
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lrow + 64 * i;
            As[buf][lk][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
            Ws[buf][lk][r] = rw[i].x; Ws[buf][lk + 1][r] = rw[i].y; Ws[buf][lk + 2][r] = rw[i].z; Ws[buf][lk + 3][r] = rw[i].w;
        }
    };
    const int nk = (K + G_BK - 1) / G_BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * G_BK);
#pragma unroll
        for (int kk = 0; kk < G_BK; ++kk) {
            float a[8], w[8];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
            *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float y = acc[i][jh * 4 + j];
                if (n + j < N) {
                    if (bias) y += __ldg(bias + n + j);
                    if (act == 1) y = gelu_erf(y);
                    if (residual) y += residual[(size_t)m * ldc + n + j];
                }
                v[j] = y;
            }
            if (n + 3 < N && (ldc & 3) == 0) *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
            else
                for (int j = 0; j < 4; ++j) if (n + j < N) C[(size_t)m * ldc + n + j] = v[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------ bf16 hi/lo split
__global__ void split_bf16_kernel(const float* __restrict__ x, int64_t count, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// ------------------------------------------------------------------------------------------------ position bias (:212-215)
// F.interpolate(pb (1,n,4,4), size=(s,s), mode='bilinear', align_corners=False) + (pc (n,s,1) + pr (n,1,s)).
__global__ void position_bias_kernel(const float* __restrict__ pb, const float* __restrict__ pc,
                                     const float* __restrict__ pr, int n, int s, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * s * s) return;
    const int m = i / (s * s), r = (i / s) % s, q = i % s;
    const float scale = 4.0f / (float)s;
    // ATen area_pixel_compute_source_index (align_corners=False): src = max(scale*(dst+0.5)-0.5, 0)
    float sr = fmaxf(scale * ((float)r + 0.5f) - 0.5f, 0.f), sq = fmaxf(scale * ((float)q + 0.5f) - 0.5f, 0.f);
    const int r0 = (int)sr, q0 = (int)sq;
    const int r1 = r0 + (r0 < 3 ? 1 : 0), q1 = q0 + (q0 < 3 ? 1 : 0);
    const float lr1 = sr - (float)r0, lq1 = sq - (float)q0, lr0 = 1.f - lr1, lq0 = 1.f - lq1;
    const float* P = pb + (size_t)m * 16;
    const float v = lr0 * (lq0 * P[r0 * 4 + q0] + lq1 * P[r0 * 4 + q1]) + lr1 * (lq0 * P[r1 * 4 + q0] + lq1 * P[r1 * 4 + q1]);
    out[i] = v + (pc[(size_t)m * s + r] + pr[(size_t)m * s + q]);
}

// ------------------------------------------------------------------------------------------------ heads (:445-446,:454-455)
// One warp per row: o (<= 16) dot products of length c, then the eval-mode BatchNorm1d affine.
__global__ void __launch_bounds__(256) heads_kernel(const float* __restrict__ g, const float* __restrict__ lw,
                                                    const float* __restrict__ lb, const float* __restrict__ sc,
                                                    const float* __restrict__ sh, int rows, int c, int o,
                                                    float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    for (int j = 0; j < o; ++j) {
        float a = 0.f;
        for (int ch = lane; ch < c; ch += 32) a = fmaf(g[(size_t)row * c + ch], __ldg(lw + (size_t)j * c + ch), a);
        a = warp_sum(a);
        if (lane == 0) out[(size_t)row * o + j] = fmaf(a + __ldg(lb + j), __ldg(sc + j), __ldg(sh + j));
    }
}

// The shipped shapes (c = 256, o = 3 translate / 9 transform): the row is read once (two coalesced 16-byte loads per lane), the O
// dot products are independent accumulators and their O warp reductions interleave — the generic kernel above re-reads the row
// and runs one dependent load -> FMA -> 5-shuffle chain per output (26 us for 16 384 rows x 9 where 17 MB of input take 3 us).
template <int O>
__global__ void __launch_bounds__(256) heads256_kernel(const float* __restrict__ g, const float* __restrict__ lw,
                                                       const float* __restrict__ lb, const float* __restrict__ sc,
                                                       const float* __restrict__ sh, int rows, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* gr = reinterpret_cast<const float4*>(g + (size_t)row * 256);
    const float4 x0 = gr[lane], x1 = gr[32 + lane];
    float acc[O];
#pragma unroll
    for (int j = 0; j < O; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(lw + (size_t)j * 256);
        const float4 w0 = __ldg(wr + lane), w1 = __ldg(wr + 32 + lane);
        // same order of products inside a lane as the generic kernel's strided loop is NOT kept: the sum of 256 products is
        // re-associated (|difference| ~1e-7 relative), well inside the path's 1e-4 coordinate bar
        float a = x0.x * w0.x;
        a = fmaf(x0.y, w0.y, a); a = fmaf(x0.z, w0.z, a); a = fmaf(x0.w, w0.w, a);
        a = fmaf(x1.x, w1.x, a); a = fmaf(x1.y, w1.y, a); a = fmaf(x1.z, w1.z, a); a = fmaf(x1.w, w1.w, a);
        acc[j] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < O; ++j) acc[j] += __shfl_xor_sync(FULL, acc[j], o);
    }
    float y = 0.f;
#pragma unroll
    for (int j = 0; j < O; ++j) y = lane == j ? acc[j] : y;
    if (lane < O) out[(size_t)row * O + lane] = fmaf(y + __ldg(lb + lane), __ldg(sc + lane), __ldg(sh + lane));
}

int launch_layernorm_split(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                           float* out, __nv_bfloat16* out_hi, long long out_plane, cudaStream_t s);
int launch_layernorm(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                     float* out, cudaStream_t s) {
    return launch_layernorm_split(x, w, b, add, add_rows, rows, c, out, nullptr, 0, s);
}

int launch_layernorm_split(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                           float* out, __nv_bfloat16* out_hi, long long out_plane, cudaStream_t s) {
    PT_REQUIRE(c % 32 == 0 && c >= 32 && c <= 1024, "layernorm: c=%d unsupported", c);
    const int wpb = 8, grid = ceil_div(rows, wpb);
    switch (c / 32) {
#define PT_LN_CASE(CPL) case CPL: { ProfScope prof_(PROF_LAYERNORM, s); layernorm_kernel<CPL><<<grid, wpb * 32, 0, s>>>(x, w, b, add, add_rows, rows, out, out_hi, out_plane); } break;
        PT_LN_CASE(1) PT_LN_CASE(2) PT_LN_CASE(4) PT_LN_CASE(8) PT_LN_CASE(16) PT_LN_CASE(32)
#undef PT_LN_CASE
        default: PT_REQUIRE(false, "layernorm: c=%d unsupported (c/32 must be a power of two)", c);
    }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

int launch_gemm_f32_strided(const float* A, const float* W, const float* bias, const float* residual, int act, int M, int N,
                            int K, float* C, int lda, int ldw, int ldc, int batch, long long bsA, long long bsW,
                            long long bsC, long long bsBias, cudaStream_t s) {
    PT_REQUIRE(M > 0 && N > 0 && K > 0 && (K % 4) == 0 && (lda % 4) == 0 && (ldw % 4) == 0 && (bsA % 4) == 0 && (bsW % 4) == 0,
               "gemm: M=%d N=%d K=%d lda=%d ldw=%d (K, strides must be multiples of 4)", M, N, K, lda, ldw);
    { ProfScope prof_(PROF_GEMM_F32, s); gemm_nt_f32_kernel<<<dim3(ceil_div(N, G_BN), ceil_div(M, G_BM), batch), G_THREADS, 0, s>>>(
        A, W, bias, residual, act, M, N, K, C, lda, ldw, ldc, bsA, bsW, bsC, bsBias); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

int launch_gemm_f32(const float* A, const float* W, const float* bias, const float* residual, int act, int M, int N, int K,
                    float* C, cudaStream_t s) {
    return launch_gemm_f32_strided(A, W, bias, residual, act, M, N, K, C, K, K, N, 1, 0, 0, 0, 0, s);
}

}  // namespace pt

using namespace pt;

extern "C" int pt_layernorm(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                            float* out, pt_stream_t stream) {
    PT_REQUIRE(x && w && b && out && rows > 0, "pt_layernorm: bad argument");
    PT_REQUIRE(add == nullptr || add_rows > 0, "pt_layernorm: add_rows");
    return launch_layernorm(x, w, b, add, add_rows, rows, c, out, (cudaStream_t)stream);
}

extern "C" int pt_split_bf16(const float* x, int64_t count, void* out, pt_stream_t stream) {
    PT_REQUIRE(x && out && count > 0, "pt_split_bf16: bad argument");
    __nv_bfloat16* hi = (__nv_bfloat16*)out;
    int grid = (int)((count + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    { ProfScope prof_(PROF_SPLIT, (cudaStream_t)stream); split_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, count, hi, hi + count); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

extern "C" int pt_position_bias(const float* pb, const float* pc, const float* pr, int n, int s, float* out,
                                pt_stream_t stream) {
    PT_REQUIRE(pb && pc && pr && out && n > 0 && s > 0, "pt_position_bias: bad argument");
    { ProfScope prof_(PROF_MISC, (cudaStream_t)stream); position_bias_kernel<<<ceil_div(n * s * s, 256), 256, 0, (cudaStream_t)stream>>>(pb, pc, pr, n, s, out); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

extern "C" int pt_heads(const float* guide, const float* lin_w, const float* lin_b, const float* bn_scale,
                        const float* bn_shift, int rows, int c, int o, float* out, pt_stream_t stream) {
    PT_REQUIRE(guide && lin_w && lin_b && bn_scale && bn_shift && out && rows > 0 && c > 0 && o > 0 && o <= 16,
               "pt_heads: bad argument");
    {
        ProfScope prof_(PROF_HEADS, (cudaStream_t)stream);
        const bool vec = c == 256 && (((uintptr_t)guide | (uintptr_t)lin_w) & 15) == 0;
        if (vec && o == 3) heads256_kernel<3><<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(guide, lin_w, lin_b, bn_scale, bn_shift, rows, out);
        else if (vec && o == 9) heads256_kernel<9><<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(guide, lin_w, lin_b, bn_scale, bn_shift, rows, out);
        else heads_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(guide, lin_w, lin_b, bn_scale, bn_shift, rows, c, o, out);
    }
    PT_LAUNCH_CHECK();
    return PT_OK;
}
