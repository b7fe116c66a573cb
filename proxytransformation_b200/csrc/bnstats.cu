// Training-mode BatchNorm statistics for the forward pass (forward only).
//
// In train() mode the four BatchNorm layers of the path normalise with BATCH statistics (nn.BatchNorm2d at :72 / :112 over
// (B, M, K) per channel, nn.BatchNorm1d at :326-330 over (B, n) per channel) and update their running statistics.  The fused
// kernels of the eval path take a per-channel (scale, shift) pair, so train mode is: one statistics pass that re-evaluates
// the layer's pre-BN output and accumulates sum / sum of squares per channel in fp64 (ATen's CPU kernel accumulates in
// double as well), pt_bn_batch_affine -> (scale, shift) + running-statistics update, then the unchanged fused kernel.
#include "common.cuh"

#include <math.h>

namespace pt {

constexpr int BS_H = 256, BS_CPL = BS_H / 32, BS_WARPS = 8;

// Pre-BN output of the 1x1 conv of OffsetNetwork / SimplifiedPointNet (:87-101, :126-139), same gather, feature
// construction and FMA order as cluster_mlp_kernel (geom.cu); one warp per cluster, lane l owns channels l + 32 i.
__global__ void __launch_bounds__(BS_WARPS * 32) cluster_conv_stats_kernel(
    const float* __restrict__ points, const int32_t* __restrict__ idx, const float* __restrict__ centres,
    const float* __restrict__ conv_w, const float* __restrict__ conv_b, int B, int M, int N, int K, double* __restrict__ sums) {
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * BS_WARPS + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * BS_WARPS;
    float w[BS_CPL][6], bb[BS_CPL];
#pragma unroll
    for (int i = 0; i < BS_CPL; ++i) {
        const int ch = lane + 32 * i;
#pragma unroll
        for (int f = 0; f < 6; ++f) w[i][f] = __ldg(conv_w + ch * 6 + f);
        bb[i] = __ldg(conv_b + ch);
    }
    double s1[BS_CPL], s2[BS_CPL];
#pragma unroll
    for (int i = 0; i < BS_CPL; ++i) { s1[i] = 0.0; s2[i] = 0.0; }
    for (int cm = warp_global; cm < B * M; cm += n_warps) {
        const float* P = points + (size_t)(cm / M) * N * 3;
        const float cx = __ldg(centres + (size_t)cm * 3), cy = __ldg(centres + (size_t)cm * 3 + 1),
                    cz = __ldg(centres + (size_t)cm * 3 + 2);
        float c1[BS_CPL], c2[BS_CPL];                  // per-cluster partials in fp32 (K terms), folded into fp64 per cluster
#pragma unroll
        for (int i = 0; i < BS_CPL; ++i) { c1[i] = 0.f; c2[i] = 0.f; }
        for (int k0 = 0; k0 < K; k0 += 32) {
            float px = 0.f, py = 0.f, pz = 0.f;
            if (k0 + lane < K) {
                const int id = __ldg(idx + (size_t)cm * K + k0 + lane);
                if (id >= 0) { px = __ldg(P + (size_t)id * 3); py = __ldg(P + (size_t)id * 3 + 1); pz = __ldg(P + (size_t)id * 3 + 2); }
            }
            const bool pad = (px == 0.0f) && (py == 0.0f) && (pz == 0.0f);          // :94, :132
            const float rx = pad ? 0.0f : __fsub_rn(px, cx), ry = pad ? 0.0f : __fsub_rn(py, cy),
                        rz = pad ? 0.0f : __fsub_rn(pz, cz);
            const int kn = min(32, K - k0);
            for (int t = 0; t < kn; ++t) {
                const float f0 = __shfl_sync(FULL, rx, t), f1 = __shfl_sync(FULL, ry, t), f2 = __shfl_sync(FULL, rz, t);
                const float f3 = __shfl_sync(FULL, px, t), f4 = __shfl_sync(FULL, py, t), f5 = __shfl_sync(FULL, pz, t);
#pragma unroll
                for (int i = 0; i < BS_CPL; ++i) {
                    float a = bb[i];
                    a = fmaf(w[i][0], f0, a); a = fmaf(w[i][1], f1, a); a = fmaf(w[i][2], f2, a);
                    a = fmaf(w[i][3], f3, a); a = fmaf(w[i][4], f4, a); a = fmaf(w[i][5], f5, a);
                    c1[i] += a;
                    c2[i] = fmaf(a, a, c2[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < BS_CPL; ++i) { s1[i] += (double)c1[i]; s2[i] += (double)c2[i]; }
    }
#pragma unroll
    for (int i = 0; i < BS_CPL; ++i) {
        atomicAdd(sums + lane + 32 * i, s1[i]);
        atomicAdd(sums + BS_H + lane + 32 * i, s2[i]);
    }
}

// Pre-BN output of the heads' Linear (:445, :454): one warp per row, lane j keeps the sums of output column j (o <= 16).
__global__ void __launch_bounds__(256) linear_stats_kernel(const float* __restrict__ g, const float* __restrict__ lw,
                                                           const float* __restrict__ lb, int rows, int c, int o,
                                                           double* __restrict__ sums) {
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    double s1 = 0.0, s2 = 0.0;
    for (int row = warp_global; row < rows; row += n_warps) {
        for (int j = 0; j < o; ++j) {
            float a = 0.f;
            for (int ch = lane; ch < c; ch += 32) a = fmaf(g[(size_t)row * c + ch], __ldg(lw + (size_t)j * c + ch), a);
            a = warp_sum(a);
            const float z = a + __ldg(lb + j);
            if (lane == j) { s1 += (double)z; s2 += (double)z * (double)z; }
        }
    }
    if (lane < o) {
        atomicAdd(sums + lane, s1);
        atomicAdd(sums + o + lane, s2);
    }
}

// sums -> batch mean / biased variance -> y = x * scale + shift, and the running statistics after this step
// (nn.BatchNorm: momentum on the batch mean and on the UNBIASED batch variance).
__global__ void bn_batch_affine_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ running_mean,
                                       float* __restrict__ running_var, float momentum, float eps, int C,
                                       float* __restrict__ scale, float* __restrict__ shift) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= C) return;
    const double mean = sums[ch] / count;
    double var = sums[C + ch] / count - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = invstd * gamma[ch];
    scale[ch] = sc;
    shift[ch] = beta[ch] - (float)mean * sc;
    if (running_mean != nullptr) running_mean[ch] = (1.0f - momentum) * running_mean[ch] + momentum * (float)mean;
    if (running_var != nullptr) {
        const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
        running_var[ch] = (1.0f - momentum) * running_var[ch] + momentum * (float)unbiased;
    }
}

}  // namespace pt

using namespace pt;

extern "C" int pt_cluster_conv_bn_stats(const float* points, const int32_t* idx, const float* centres, const float* conv_w,
                                        const float* conv_b, int B, int M, int N, int K, int H, double* sums,
                                        pt_stream_t stream) {
    PT_REQUIRE(H == BS_H, "pt_cluster_conv_bn_stats: hidden width %d unsupported (reference hard-wires 256, :31,:110)", H);
    PT_REQUIRE(B > 0 && M > 0 && N > 0 && K > 0, "pt_cluster_conv_bn_stats: bad shape");
    PT_REQUIRE(points && idx && centres && conv_w && conv_b && sums, "pt_cluster_conv_bn_stats: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    PT_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)2 * H * sizeof(double), s));
    int blocks = ceil_div(ceil_div(B * M, 4), BS_WARPS);
    blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
    { ProfScope prof_(PROF_MISC, s); cluster_conv_stats_kernel<<<blocks, BS_WARPS * 32, 0, s>>>(points, idx, centres, conv_w, conv_b, B, M, N, K, sums); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

extern "C" int pt_linear_bn_stats(const float* guide, const float* lin_w, const float* lin_b, int rows, int c, int o,
                                  double* sums, pt_stream_t stream) {
    PT_REQUIRE(guide && lin_w && lin_b && sums && rows > 0 && c > 0 && o > 0 && o <= 16, "pt_linear_bn_stats: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    PT_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)2 * o * sizeof(double), s));
    int blocks = ceil_div(rows, 8 * 4);
    blocks = blocks < 1 ? 1 : (blocks > 148 * 4 ? 148 * 4 : blocks);
    { ProfScope prof_(PROF_MISC, s); linear_stats_kernel<<<blocks, 256, 0, s>>>(guide, lin_w, lin_b, rows, c, o, sums); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

extern "C" int pt_bn_batch_affine(const double* sums, long long count, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float momentum, float eps, int C, float* scale,
                                  float* shift, pt_stream_t stream) {
    PT_REQUIRE(sums && gamma && beta && scale && shift && count > 0 && C > 0, "pt_bn_batch_affine: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    { ProfScope prof_(PROF_MISC, s); bn_batch_affine_kernel<<<ceil_div(C, 128), 128, 0, s>>>(sums, (double)count, gamma, beta, running_mean, running_var, momentum, eps, C, scale, shift); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}
