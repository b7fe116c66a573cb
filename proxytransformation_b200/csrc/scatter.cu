// S10-S12: per-cluster affine (:459-462), duplicate-resolving scatter (pt_replace :472-498) and point removal
// (remove_points_by_index :501-525) fused into one mark pass + one ordered compaction over the N points.
// The transformed coordinates are never materialised per (m,k): the winning slot of a point is resolved with
// atomicMax on the flat index m*K+k (the pinned "last writer wins" rule) and the affine is applied while compacting,
// because a valid slot's source coordinate is the point itself.  HBM-bound: ~12N read + 12N' write + 8N scratch / scene.
#include "common.cuh"

namespace pt {

constexpr int SC_DROP = 0x7fffffff;
constexpr int SC_BLOCK = 1024;

__global__ void mark_kernel(const int32_t* __restrict__ kept_idx, const int32_t* __restrict__ drop_idx, int B, int N, int nK,
                            int n_drop_entries, int* __restrict__ winner) {
    const long long total_keep = (long long)B * nK, total = total_keep + (long long)B * n_drop_entries;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (i < total_keep) {
            const int b = (int)(i / nK), e = (int)(i - (long long)b * nK);
            const int id = __ldg(kept_idx + i);
            if (id >= 0) atomicMax(winner + (size_t)b * N + id, e);
        } else {
            const long long k = i - total_keep;
            const int b = (int)(k / n_drop_entries);
            const int id = __ldg(drop_idx + k);
            if (id >= 0) atomicMax(winner + (size_t)b * N + id, SC_DROP);
        }
    }
}

__global__ void __launch_bounds__(SC_BLOCK) count_kernel(const int* __restrict__ winner, int N, int* __restrict__ blockcnt) {
    const int b = blockIdx.y, i = blockIdx.x * SC_BLOCK + threadIdx.x;
    const bool keep = i < N && winner[(size_t)b * N + i] != SC_DROP;
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) blockcnt[(size_t)b * gridDim.x + blockIdx.x] = c;
}

__global__ void __launch_bounds__(SC_BLOCK) compact_kernel(const float* __restrict__ points, const int* __restrict__ winner,
                                                           const int* __restrict__ blockcnt, const float* __restrict__ centres,
                                                           const float* __restrict__ transform, const float* __restrict__ translate,
                                                           int N, int n, int K, float* __restrict__ out, int32_t* __restrict__ counts) {
    __shared__ int warp_tot[SC_BLOCK / 32];
    __shared__ int base_s;
    const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (wid == 0) {                                   // exclusive prefix over the preceding blocks of this scene
        int s = 0;
        for (int j = lane; j < blk; j += 32) s += blockcnt[(size_t)b * nblk + j];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if (lane == 0) base_s = s;
    }
    const int i = blk * SC_BLOCK + tid;
    int w = SC_DROP;
    if (i < N) w = winner[(size_t)b * N + i];
    const bool keep = w != SC_DROP;
    const unsigned mask = __ballot_sync(FULL, keep);
    if (lane == 0) warp_tot[wid] = __popc(mask);
    __syncthreads();
    int off = base_s;
    for (int j = 0; j < wid; ++j) off += warp_tot[j];
    if (keep) {
        const float* p = points + ((size_t)b * N + i) * 3;
        float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
        if (w >= 0) {
            const int m = w / K;
            const float* T = transform + ((size_t)b * n + m) * 9;
            const float* c = centres + ((size_t)b * n + m) * 3;
            const float* t = translate + ((size_t)b * n + m) * 3;
            const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
            const float rx = __fsub_rn(x, cx), ry = __fsub_rn(y, cy), rz = __fsub_rn(z, cz);
            // ((T @ rel) + centre) + translate (:462)
            x = __fadd_rn(__fadd_rn(fmaf(__ldg(T + 2), rz, fmaf(__ldg(T + 1), ry, __fmul_rn(__ldg(T + 0), rx))), cx), __ldg(t));
            y = __fadd_rn(__fadd_rn(fmaf(__ldg(T + 5), rz, fmaf(__ldg(T + 4), ry, __fmul_rn(__ldg(T + 3), rx))), cy), __ldg(t + 1));
            z = __fadd_rn(__fadd_rn(fmaf(__ldg(T + 8), rz, fmaf(__ldg(T + 7), ry, __fmul_rn(__ldg(T + 6), rx))), cz), __ldg(t + 2));
        }
        float* o = out + ((size_t)b * N + off + __popc(mask & ((1u << lane) - 1u))) * 3;
        o[0] = x; o[1] = y; o[2] = z;
    }
    if (blk == nblk - 1 && tid == SC_BLOCK - 1) counts[b] = off + __popc(mask);
}

// N1 hand-off to the sparse backbone (detectors/sparse_featfusion_grounder_preshape.py:388-391):
//   ME.utils.batch_sparse_collate([(p[:, :3] / voxel_size, p) for p in points]) -> coordinates (T,4) int32 [batch, x, y, z],
//   features (T,3) fp32, T = sum of the per-scene counts, scenes in order.
// The float -> int32 step is the tensor assignment `bcoords[s:s+n, 1:] = coord` of MinkowskiEngine's sparse_collate, i.e.
// truncation toward zero (floor is offered for callers that quantise with sparse_quantize).  The division follows torch:
// `tensor / python_float` is a true IEEE division on the CPU and a multiplication by the fp32 reciprocal in torch's CUDA
// kernel (div_true_kernel_cuda, CPU-scalar divisor); `recip` selects which one is reproduced bit for bit.
__global__ void __launch_bounds__(256) collate_kernel(const float* __restrict__ packed, const int32_t* __restrict__ counts, int B,
                                                      int N, float voxel_size, float inv_voxel, int recip, int use_floor,
                                                      int32_t* __restrict__ coords, float* __restrict__ feats,
                                                      int32_t* __restrict__ total) {
    __shared__ long long base_s;
    const int b = blockIdx.y;
    if (threadIdx.x < 32) {
        long long s = 0, all = 0;
        for (int j = threadIdx.x; j < B; j += 32) { const int c = __ldg(counts + j); all += c; if (j < b) s += c; }
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(FULL, s, o); all += __shfl_xor_sync(FULL, all, o); }
        if (threadIdx.x == 0) { base_s = s; if (b == 0 && blockIdx.x == 0) *total = (int32_t)all; }
    }
    __syncthreads();
    const int cnt = __ldg(counts + b);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= cnt) return;
    const float* p = packed + ((size_t)b * N + i) * 3;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    const float qx = recip ? __fmul_rn(x, inv_voxel) : __fdiv_rn(x, voxel_size);
    const float qy = recip ? __fmul_rn(y, inv_voxel) : __fdiv_rn(y, voxel_size);
    const float qz = recip ? __fmul_rn(z, inv_voxel) : __fdiv_rn(z, voxel_size);
    const long long r = base_s + i;
    int4 c;
    c.x = b;
    c.y = use_floor ? __float2int_rd(qx) : __float2int_rz(qx);
    c.z = use_floor ? __float2int_rd(qy) : __float2int_rz(qy);
    c.w = use_floor ? __float2int_rd(qz) : __float2int_rz(qz);
    reinterpret_cast<int4*>(coords)[r] = c;
    float* f = feats + r * 3;
    f[0] = x; f[1] = y; f[2] = z;
}

// N3 input side: AggregateMultiViewPoints (datasets/transforms/multiview.py:224-241) + the gather of PointSample
// (datasets/transforms/points.py:411-417) fused: only the SAMPLED points are moved to the global frame.
//   points_cat (T,3): the per-view ego-frame points back to back, view v owns rows [view_off[v], view_off[v+1])
//   ego2global (V,16): row-major 4x4 inverse of the view's `extrinsic` (the reference solves extrinsic . x = [p;1] per view)
//   choices (n): indices into the concatenation (the data loader's np.random.choice: RNG stays on the host)
//   out (n,3) = (ego2global[view(choices[i])] . [p;1])[:3], in the order of `choices`
__global__ void __launch_bounds__(256) aggregate_sample_kernel(const float* __restrict__ pts, const long long* __restrict__ view_off, int V,
                                                               const float* __restrict__ ego2global, const long long* __restrict__ choices,
                                                               long long n, float* __restrict__ out) {
    extern __shared__ long long s_off[];              // V + 1 offsets
    for (int i = threadIdx.x; i <= V; i += blockDim.x) s_off[i] = view_off[i];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long c = __ldg(choices + i);
    int lo = 0, hi = V;                               // largest v with s_off[v] <= c
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= c) lo = mid; else hi = mid; }
    const float* M = ego2global + (size_t)lo * 16;
    const float x = __ldg(pts + c * 3), y = __ldg(pts + c * 3 + 1), z = __ldg(pts + c * 3 + 2);
#pragma unroll
    for (int r = 0; r < 3; ++r)
        out[i * 3 + r] = fmaf(__ldg(M + 4 * r + 2), z, fmaf(__ldg(M + 4 * r + 1), y, fmaf(__ldg(M + 4 * r), x, __ldg(M + 4 * r + 3))));
}

}  // namespace pt

using namespace pt;

extern "C" int pt_aggregate_sample(const float* points_cat, const long long* view_off, int V, const float* ego2global,
                                   const long long* choices, long long n, float* out, pt_stream_t stream) {
    PT_REQUIRE(points_cat && view_off && ego2global && choices && out, "pt_aggregate_sample: null pointer");
    PT_REQUIRE(V >= 1 && V <= 4096 && n >= 1, "pt_aggregate_sample: V=%d n=%lld", V, n);
    cudaStream_t s = (cudaStream_t)stream;
    { ProfScope prof_(PROF_MISC, s);
      aggregate_sample_kernel<<<(unsigned)((n + 255) / 256), 256, (size_t)(V + 1) * sizeof(long long), s>>>(points_cat, view_off, V, ego2global, choices, n, out); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

extern "C" int pt_sparse_collate(const float* packed, const int32_t* counts, int B, int N, float voxel_size, int flags,
                                 int32_t* coords, float* feats, int32_t* total, pt_stream_t stream) {
    PT_REQUIRE(packed && counts && coords && feats && total, "pt_sparse_collate: null pointer");
    PT_REQUIRE(B > 0 && N > 0 && voxel_size > 0.f && (flags & ~3) == 0, "pt_sparse_collate: bad argument");
    PT_REQUIRE(((uintptr_t)coords & 15) == 0, "pt_sparse_collate: coords must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    { ProfScope prof_(PROF_MISC, s);
      collate_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, s>>>(packed, counts, B, N, voxel_size, 1.0f / voxel_size, flags & PT_COLLATE_RECIPROCAL,
                                                               flags & PT_COLLATE_FLOOR, coords, feats, total); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}

extern "C" size_t pt_scatter_ws_bytes(int B, int N) {
    return align_up((size_t)B * N * sizeof(int), 256) + align_up((size_t)B * ceil_div(N, SC_BLOCK) * sizeof(int), 256);
}

extern "C" int pt_affine_scatter_compact(const float* points, const int32_t* kept_idx, const int32_t* drop_idx,
                                         const float* kept_centres, const float* transform, const float* translate, int B,
                                         int N, int n, int K, int n_drop_entries, float* out, int32_t* counts, void* ws,
                                         size_t ws_bytes, pt_stream_t stream) {
    PT_REQUIRE(points && kept_idx && drop_idx && kept_centres && transform && translate && out && counts && ws,
               "pt_affine_scatter_compact: null pointer");
    PT_REQUIRE(B > 0 && N > 0 && n > 0 && K > 0 && n_drop_entries >= 0 && (long long)n * K < SC_DROP,
               "pt_affine_scatter_compact: bad shape");
    if (ws_bytes < pt_scatter_ws_bytes(B, N)) { set_error("pt_affine_scatter_compact: workspace %zu < %zu", ws_bytes, pt_scatter_ws_bytes(B, N)); return PT_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    int* winner = (int*)ws;
    int* blockcnt = (int*)((char*)ws + align_up((size_t)B * N * sizeof(int), 256));
    PT_CUDA_OK(cudaMemsetAsync(winner, 0xff, (size_t)B * N * sizeof(int), s));   // -1 = untouched
    const long long total = (long long)B * ((long long)n * K + n_drop_entries);
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    { ProfScope prof_(PROF_MARK, s); mark_kernel<<<grid, 256, 0, s>>>(kept_idx, drop_idx, B, N, n * K, n_drop_entries, winner); }
    PT_LAUNCH_CHECK();
    const int nblk = ceil_div(N, SC_BLOCK);
    { ProfScope prof_(PROF_COUNT, s); count_kernel<<<dim3(nblk, B), SC_BLOCK, 0, s>>>(winner, N, blockcnt); }
    PT_LAUNCH_CHECK();
    { ProfScope prof_(PROF_COMPACT, s); compact_kernel<<<dim3(nblk, B), SC_BLOCK, 0, s>>>(points, winner, blockcnt, kept_centres, transform, translate, N, n, K, out, counts); }
    PT_LAUNCH_CHECK();
    return PT_OK;
}
