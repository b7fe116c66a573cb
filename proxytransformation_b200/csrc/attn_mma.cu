// Tensor-core form of the two-stage proxy attention core of ProxyAttention.forward (:225-252), one CTA per (scene, head):
//   stage 1 (proxy as query, :232-238):  Pv = softmax_n((Pt*scale) K^T) V          (l x hd), unmasked
//   stage 2 (proxy as key,   :241-250):  O  = softmax_l(mask((Q*scale) Pt^T)) Pv   (n x hd)
// Both stages are the same "rows x keys -> softmax -> values" pattern, run flash-style per 16-row tile with an online
// softmax over 64-key chunks; the contractions use mma.sync m16n8k16 bf16 with fp32 accumulation and 3xBF16 operand
// splitting (hi*hi + lo*hi + hi*lo) so scores and outputs keep ~2^-17 relative accuracy.
// K, V^T, Pt, Pv^T of the head live in shared memory as bf16 hi/lo planes whose pitches make every fragment load
// bank-conflict free; probabilities go from the score accumulators straight into A fragments (no smem round trip).
// Cost is O(n*l*hd), never n^2.
#include "common.cuh"

#include <math.h>

namespace pt {

constexpr int AM_THREADS = 512, AM_WARPS = 16, AM_HD = 32, AM_KP = 40;   // KP: bf16 pitch of the [rows][32] operand planes

__device__ __forceinline__ void am_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x0, x1) -> packed bf16x2 hi and lo words (element 0 in the low half)
__device__ __forceinline__ void am_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

struct AmFrag { uint32_t hi[2][4], lo[2][4]; };      // 16 x 32 A operand (two k-steps), hi and lo planes

// A fragments of rows [r0, r0+16) of a [rows][AM_KP] bf16 plane pair in shared memory
__device__ __forceinline__ void am_load_a(AmFrag& A, const __nv_bfloat16* hi, const __nv_bfloat16* lo, int r0, int g, int q) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const int c0 = 16 * ks + 2 * q;
        const int o0 = (r0 + g) * AM_KP + c0, o1 = (r0 + g + 8) * AM_KP + c0;
        A.hi[ks][0] = *reinterpret_cast<const uint32_t*>(hi + o0);
        A.hi[ks][1] = *reinterpret_cast<const uint32_t*>(hi + o1);
        A.hi[ks][2] = *reinterpret_cast<const uint32_t*>(hi + o0 + 8);
        A.hi[ks][3] = *reinterpret_cast<const uint32_t*>(hi + o1 + 8);
        A.lo[ks][0] = *reinterpret_cast<const uint32_t*>(lo + o0);
        A.lo[ks][1] = *reinterpret_cast<const uint32_t*>(lo + o1);
        A.lo[ks][2] = *reinterpret_cast<const uint32_t*>(lo + o0 + 8);
        A.lo[ks][3] = *reinterpret_cast<const uint32_t*>(lo + o1 + 8);
    }
}

// One 16-row tile against all keys: O (16 x 32, unnormalised) and the per-thread partial row sums.
//   keys   Bhi/Blo [nkpad][AM_KP]   (rows >= nk are zero and are excluded from the softmax)
//   values Vhi/Vlo [32][vp]         (V^T: element (e, key); columns >= nk are zero)
//   kflag  per-key float or null: != 0 -> score is replaced by -1e9 (masked_fill, :247)
__device__ __forceinline__ void am_flash_tile(const AmFrag& A, const __nv_bfloat16* __restrict__ Bhi, const __nv_bfloat16* __restrict__ Blo,
                                              const __nv_bfloat16* __restrict__ Vhi, const __nv_bfloat16* __restrict__ Vlo, int vp,
                                              int nk, int nkpad, const float* __restrict__ kflag, float post_scale, int g, int q,
                                              float (&O)[4][4], float (&lsum)[2]) {
    float mrun[2] = {-INFINITY, -INFINITY};
    lsum[0] = 0.f; lsum[1] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) O[t][e] = 0.f;
    for (int j0 = 0; j0 < nkpad; j0 += 64) {
        const int ntc = min(8, (nkpad - j0) >> 3);          // 8-key tiles in this chunk (even: nkpad % 16 == 0)
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
            if (nt < ntc) {
                const int row = (j0 + 8 * nt + g) * AM_KP + 2 * q;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(Bhi + row + 16 * ks);
                    const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(Bhi + row + 16 * ks + 8);
                    const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(Blo + row + 16 * ks);
                    const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(Blo + row + 16 * ks + 8);
                    am_mma(s[nt], A.hi[ks], bh0, bh1);
                    am_mma(s[nt], A.lo[ks], bh0, bh1);
                    am_mma(s[nt], A.hi[ks], bl0, bl1);
                }
            }
        }
        // scale, mask, chunk maximum of rows g (elements 0,1) and g+8 (elements 2,3)
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            if (nt < ntc) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int key = j0 + 8 * nt + 2 * q + e;
                    float v0 = s[nt][e] * post_scale, v1 = s[nt][2 + e] * post_scale;
                    if (key >= nk) { v0 = -INFINITY; v1 = -INFINITY; }
                    else if (kflag != nullptr && kflag[key] != 0.f) { v0 = -1e9f; v1 = -1e9f; }
                    s[nt][e] = v0; s[nt][2 + e] = v1;
                    mx[0] = fmaxf(mx[0], v0); mx[1] = fmaxf(mx[1], v1);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(FULL, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(FULL, mx[r], 2));
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float mnew = fmaxf(mrun[r], mx[r]);       // finite: every chunk's first key is a real key
            corr[r] = expf(mrun[r] - mnew);
            mrun[r] = mnew;
            lsum[r] *= corr[r];
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) { O[t][0] *= corr[0]; O[t][1] *= corr[0]; O[t][2] *= corr[1]; O[t][3] *= corr[1]; }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            if (nt < ntc) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float p0 = expf(s[nt][e] - mrun[0]), p1 = expf(s[nt][2 + e] - mrun[1]);
                    s[nt][e] = p0; s[nt][2 + e] = p1;
                    lsum[0] += p0; lsum[1] += p1;
                }
            }
        }
        // O += P V : probabilities of key tiles (2kk, 2kk+1) are exactly the A fragment of k-step kk
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            if (2 * kk < ntc) {
                uint32_t ph[4], pl[4];
                am_split2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
                am_split2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
                am_split2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
                am_split2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int off = (8 * t + g) * vp + j0 + 16 * kk + 2 * q;
                    const uint32_t vh0 = *reinterpret_cast<const uint32_t*>(Vhi + off), vh1 = *reinterpret_cast<const uint32_t*>(Vhi + off + 8);
                    const uint32_t vl0 = *reinterpret_cast<const uint32_t*>(Vlo + off), vl1 = *reinterpret_cast<const uint32_t*>(Vlo + off + 8);
                    am_mma(O[t], ph, vh0, vh1);
                    am_mma(O[t], pl, vh0, vh1);
                    am_mma(O[t], ph, vl0, vl1);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        lsum[r] += __shfl_xor_sync(FULL, lsum[r], 1);
        lsum[r] += __shfl_xor_sync(FULL, lsum[r], 2);
    }
}

struct AmSmem {
    int npad, lpad, vpn, vpl;
    size_t khi, klo, vthi, vtlo, pthi, ptlo, pvhi, pvlo, flag, total;
};
__host__ __device__ inline AmSmem am_layout(int n, int l) {
    AmSmem s;
    s.npad = (n + 15) & ~15; s.lpad = (l + 15) & ~15;
    s.vpn = s.npad + 8; s.vpl = s.lpad + 8;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    s.khi = take((size_t)s.npad * AM_KP * 2); s.klo = take((size_t)s.npad * AM_KP * 2);
    s.vthi = take((size_t)AM_HD * s.vpn * 2); s.vtlo = take((size_t)AM_HD * s.vpn * 2);
    s.pthi = take((size_t)s.lpad * AM_KP * 2); s.ptlo = take((size_t)s.lpad * AM_KP * 2);
    s.pvhi = take((size_t)AM_HD * s.vpl * 2); s.pvlo = take((size_t)AM_HD * s.vpl * 2);
    s.flag = take((size_t)s.lpad * 4);
    s.total = off;
    return s;
}

// qkv: (B*n, 3c) rows [Q | K | V]; pt: (B*l, c); mask: (B,l) uint8 (1 = real token) or null.
// Output: o (B*n, c) fp32 and/or bf16 hi/lo planes o_hi / o_hi + o_plane (operand of the proj GEMM).
__global__ void __launch_bounds__(AM_THREADS) proxy_attention_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ pt_tok,
                                                                         const uint8_t* __restrict__ mask, int n, int l, int c, float scale,
                                                                         float* __restrict__ o, __nv_bfloat16* __restrict__ o_hi,
                                                                         long long o_plane) {
    extern __shared__ __align__(16) uint8_t sm[];
    const AmSmem L = am_layout(n, l);
    __nv_bfloat16* Khi = reinterpret_cast<__nv_bfloat16*>(sm + L.khi);
    __nv_bfloat16* Klo = reinterpret_cast<__nv_bfloat16*>(sm + L.klo);
    __nv_bfloat16* Vthi = reinterpret_cast<__nv_bfloat16*>(sm + L.vthi);
    __nv_bfloat16* Vtlo = reinterpret_cast<__nv_bfloat16*>(sm + L.vtlo);
    __nv_bfloat16* Pthi = reinterpret_cast<__nv_bfloat16*>(sm + L.pthi);
    __nv_bfloat16* Ptlo = reinterpret_cast<__nv_bfloat16*>(sm + L.ptlo);
    __nv_bfloat16* Pvhi = reinterpret_cast<__nv_bfloat16*>(sm + L.pvhi);
    __nv_bfloat16* Pvlo = reinterpret_cast<__nv_bfloat16*>(sm + L.pvlo);
    float* kflag = reinterpret_cast<float*>(sm + L.flag);
    const int b = blockIdx.y, h = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);

    // ---- stage the head's K (row-major), V (transposed), Pt (row-major) as bf16 hi/lo planes; zero the padding
    for (int i = tid; i < L.npad * 8; i += AM_THREADS) {
        const int j = i >> 3, e4 = (i & 7) * 4;
        float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
        if (j < n) {
            const float* row = qkv + ((size_t)b * n + j) * 3 * c + h * AM_HD + e4;
            kv = *reinterpret_cast<const float4*>(row + c);
            vv = *reinterpret_cast<const float4*>(row + 2 * c);
        }
        uint32_t h0, l0, h1, l1;
        am_split2(kv.x, kv.y, h0, l0);
        am_split2(kv.z, kv.w, h1, l1);
        *reinterpret_cast<uint2*>(Khi + j * AM_KP + e4) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(Klo + j * AM_KP + e4) = make_uint2(l0, l1);
        const float vf[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat16 hh = __float2bfloat16_rn(vf[e]);
            Vthi[(e4 + e) * L.vpn + j] = hh;
            Vtlo[(e4 + e) * L.vpn + j] = __float2bfloat16_rn(vf[e] - __bfloat162float(hh));
        }
    }
    for (int i = tid; i < L.lpad * 8; i += AM_THREADS) {
        const int j = i >> 3, e4 = (i & 7) * 4;
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < l) pv = *reinterpret_cast<const float4*>(pt_tok + ((size_t)b * l + j) * c + h * AM_HD + e4);
        uint32_t h0, l0, h1, l1;
        am_split2(pv.x, pv.y, h0, l0);
        am_split2(pv.z, pv.w, h1, l1);
        *reinterpret_cast<uint2*>(Pthi + j * AM_KP + e4) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(Ptlo + j * AM_KP + e4) = make_uint2(l0, l1);
    }
    for (int i = tid; i < AM_HD * 8; i += AM_THREADS) {          // the 8 pad columns behind V^T / Pv^T rows
        const int e = i >> 3, k = i & 7;
        Vthi[e * L.vpn + L.npad + k] = zero; Vtlo[e * L.vpn + L.npad + k] = zero;
        Pvhi[e * L.vpl + L.lpad + k] = zero; Pvlo[e * L.vpl + L.lpad + k] = zero;
    }
    for (int i = tid; i < L.lpad; i += AM_THREADS) kflag[i] = (mask != nullptr && i < l && mask[(size_t)b * l + i] == 0) ? 1.f : 0.f;
    __syncthreads();

    // ---- stage 1: rows = proxy tokens, keys = K, values = V  ->  Pv^T (normalised) as bf16 hi/lo planes
    for (int mt = warp; mt < L.lpad / 16; mt += AM_WARPS) {
        AmFrag A;
        am_load_a(A, Pthi, Ptlo, 16 * mt, g, q);
        float O[4][4], ls[2];
        am_flash_tile(A, Khi, Klo, Vthi, Vtlo, L.vpn, n, L.npad, nullptr, scale, g, q, O, ls);
        const float inv0 = 1.0f / ls[0], inv1 = 1.0f / ls[1];
        const int i0 = 16 * mt + g, i1 = i0 + 8;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ee = 8 * t + 2 * q + e;
                // rows >= l are padding proxies: store zeros so that stage 2's P*V never multiplies garbage
                const float v0 = i0 < l ? O[t][e] * inv0 : 0.f, v1 = i1 < l ? O[t][2 + e] * inv1 : 0.f;
                const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                Pvhi[ee * L.vpl + i0] = h0; Pvlo[ee * L.vpl + i0] = __float2bfloat16_rn(v0 - __bfloat162float(h0));
                Pvhi[ee * L.vpl + i1] = h1; Pvlo[ee * L.vpl + i1] = __float2bfloat16_rn(v1 - __bfloat162float(h1));
            }
        }
    }
    __syncthreads();

    // ---- stage 2: rows = point proxies (Q * scale), keys = Pt (masked), values = Pv
    for (int mt = warp; mt < L.npad / 16; mt += AM_WARPS) {
        AmFrag A;
        const int r0 = 16 * mt + g, r1 = r0 + 8;
        const float* q0p = qkv + ((size_t)b * n + min(r0, n - 1)) * 3 * c + h * AM_HD;
        const float* q1p = qkv + ((size_t)b * n + min(r1, n - 1)) * 3 * c + h * AM_HD;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int c0 = 16 * ks + 2 * q;
            const float2 a00 = __ldg(reinterpret_cast<const float2*>(q0p + c0)), a01 = __ldg(reinterpret_cast<const float2*>(q0p + c0 + 8));
            const float2 a10 = __ldg(reinterpret_cast<const float2*>(q1p + c0)), a11 = __ldg(reinterpret_cast<const float2*>(q1p + c0 + 8));
            am_split2(a00.x * scale, a00.y * scale, A.hi[ks][0], A.lo[ks][0]);       // (q * scale) (:241)
            am_split2(a10.x * scale, a10.y * scale, A.hi[ks][1], A.lo[ks][1]);
            am_split2(a01.x * scale, a01.y * scale, A.hi[ks][2], A.lo[ks][2]);
            am_split2(a11.x * scale, a11.y * scale, A.hi[ks][3], A.lo[ks][3]);
        }
        float O[4][4], ls[2];
        am_flash_tile(A, Pthi, Ptlo, Pvhi, Pvlo, L.vpl, l, L.lpad, kflag, 1.0f, g, q, O, ls);
        const float inv0 = 1.0f / ls[0], inv1 = 1.0f / ls[1];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int col = h * AM_HD + 8 * t + 2 * q;
            const float y00 = O[t][0] * inv0, y01 = O[t][1] * inv0, y10 = O[t][2] * inv1, y11 = O[t][3] * inv1;
            if (r0 < n) {
                const size_t off = ((size_t)b * n + r0) * c + col;
                if (o != nullptr) *reinterpret_cast<float2*>(o + off) = make_float2(y00, y01);
                if (o_hi != nullptr) {
                    uint32_t hh, ll;
                    am_split2(y00, y01, hh, ll);
                    *reinterpret_cast<uint32_t*>(o_hi + off) = hh;
                    *reinterpret_cast<uint32_t*>(o_hi + o_plane + off) = ll;
                }
            }
            if (r1 < n) {
                const size_t off = ((size_t)b * n + r1) * c + col;
                if (o != nullptr) *reinterpret_cast<float2*>(o + off) = make_float2(y10, y11);
                if (o_hi != nullptr) {
                    uint32_t hh, ll;
                    am_split2(y10, y11, hh, ll);
                    *reinterpret_cast<uint32_t*>(o_hi + off) = hh;
                    *reinterpret_cast<uint32_t*>(o_hi + o_plane + off) = ll;
                }
            }
        }
    }
}

bool proxy_attention_mma_supported(int n, int l, int c, int heads) {
    return c % heads == 0 && c / heads == AM_HD && c % 4 == 0 && n >= 1 && l >= 1 && am_layout(n, l).total <= 227 * 1024;
}

int launch_proxy_attention_mma(const float* qkv, const float* pt_tok, const uint8_t* mask, int B, int n, int l, int c, int heads,
                               float* o, void* o_split, long long o_plane, cudaStream_t s) {
    PT_REQUIRE(proxy_attention_mma_supported(n, l, c, heads), "attention(mma): n=%d l=%d c=%d heads=%d unsupported", n, l, c, heads);
    const size_t smem = am_layout(n, l).total;
    if (smem > 48 * 1024)      // per device and size-dependent: set on every launch (a host-side call of a few hundred ns)
        PT_CUDA_OK(cudaFuncSetAttribute(proxy_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float scale = (float)(1.0 / sqrt((double)AM_HD));     // python float head_dim ** -0.5 (:186), rounded to fp32 once
    { ProfScope prof_(PROF_ATTENTION, s); proxy_attention_mma_kernel<<<dim3(heads, B), AM_THREADS, smem, s>>>(qkv, pt_tok, mask, n, l, c, scale, o, (__nv_bfloat16*)o_split, o_plane); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

}  // namespace pt
