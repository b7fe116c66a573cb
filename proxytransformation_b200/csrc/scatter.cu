// S10-S12: per-cluster affine (:459-462), duplicate-resolving scatter (pt_replace :472-498) and point removal
// (remove_points_by_index :501-525) fused into one mark pass + one ordered compaction over the N points.
// The transformed coordinates are never materialised per (m,k): the winning slot of a point is resolved with
// atomicMax on the flat index m*K+k (the pinned "last writer wins" rule) and the affine is applied while compacting,
// because a valid slot's source coordinate is the point itself.  HBM-bound: ~12N read + 12N' write + 8N scratch / scene.
#include "common.cuh"

namespace pt {

constexpr int SC_DROP = 0x7fffffff;
constexpr int SC_BLOCK = 1024;

// `blockdrop[scene][block of SC_BLOCK points]` counts the DISTINCT dropped points of a block (the first SC_DROP mark of a point
// is the one whose atomicMax returns something else), so the compaction needs no counting pass over the winner array.
__global__ void mark_kernel(const int32_t* __restrict__ kept_idx, const int32_t* __restrict__ drop_idx, int B, int N, int nK,
                            int n_drop_entries, int nblk, int* __restrict__ winner, int* __restrict__ blockdrop) {
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
            if (id >= 0 && atomicMax(winner + (size_t)b * N + id, SC_DROP) != SC_DROP) atomicAdd(blockdrop + (size_t)b * nblk + id / SC_BLOCK, 1);
        }
    }
}

// One CTA (SC_THREADS threads, 4 consecutive points each) per block of SC_BLOCK points of a scene: coordinates and winners come
// in with 16-byte loads, the survivors (transformed where a cluster claimed them) are packed in shared memory and leave as one
// contiguous, fully coalesced run at the scene's running output offset (prefix of the preceding blocks' survivor counts).
constexpr int SC_THREADS = SC_BLOCK / 4;

__device__ __forceinline__ void sc_affine(float& x, float& y, float& z, int w, int K, const float* __restrict__ centres,
                                          const float* __restrict__ transform, const float* __restrict__ translate, size_t bn) {
    const int m = w / K;
    const float* T = transform + (bn + m) * 9;
    const float* c = centres + (bn + m) * 3;
    const float* t = translate + (bn + m) * 3;
    const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    const float rx = __fsub_rn(x, cx), ry = __fsub_rn(y, cy), rz = __fsub_rn(z, cz);
    // ((T @ rel) + centre) + translate (:462)
    x = __fadd_rn(__fadd_rn(fmaf(__ldg(T + 2), rz, fmaf(__ldg(T + 1), ry, __fmul_rn(__ldg(T + 0), rx))), cx), __ldg(t));
    y = __fadd_rn(__fadd_rn(fmaf(__ldg(T + 5), rz, fmaf(__ldg(T + 4), ry, __fmul_rn(__ldg(T + 3), rx))), cy), __ldg(t + 1));
    z = __fadd_rn(__fadd_rn(fmaf(__ldg(T + 8), rz, fmaf(__ldg(T + 7), ry, __fmul_rn(__ldg(T + 6), rx))), cz), __ldg(t + 2));
}

__global__ void __launch_bounds__(SC_THREADS) compact_kernel(const float* __restrict__ points, const int* __restrict__ winner,
                                                             const int* __restrict__ blockdrop, const float* __restrict__ centres,
                                                             const float* __restrict__ transform, const float* __restrict__ translate,
                                                             int N, int n, int K, float* __restrict__ out, int32_t* __restrict__ counts) {
    __shared__ float sout[3 * SC_BLOCK];
    __shared__ int warp_tot[SC_THREADS / 32];
    __shared__ int base_s;
    const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int i0 = blk * SC_BLOCK + 4 * tid;          // this thread's first point
    if (wid == 0) {                                   // survivors of the preceding (full) blocks of this scene
        int s = 0;
        for (int j = lane; j < blk; j += 32) s += SC_BLOCK - blockdrop[(size_t)b * nblk + j];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if (lane == 0) base_s = s;
    }
    const float* src = points + ((size_t)b * N + i0) * 3;
    const int* wsrc = winner + (size_t)b * N + i0;
    float x[4], y[4], z[4];
    int w[4] = {SC_DROP, SC_DROP, SC_DROP, SC_DROP};
    if (i0 + 3 < N && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(wsrc) & 15) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b4 = __ldg(reinterpret_cast<const float4*>(src) + 1),
                     c = __ldg(reinterpret_cast<const float4*>(src) + 2);
        x[0] = a.x; y[0] = a.y; z[0] = a.z; x[1] = a.w; y[1] = b4.x; z[1] = b4.y;
        x[2] = b4.z; y[2] = b4.w; z[2] = c.x; x[3] = c.y; y[3] = c.z; z[3] = c.w;
        const int4 w4 = *reinterpret_cast<const int4*>(wsrc);
        w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            x[e] = y[e] = z[e] = 0.f;
            if (i0 + e < N) { x[e] = __ldg(src + 3 * e); y[e] = __ldg(src + 3 * e + 1); z[e] = __ldg(src + 3 * e + 2); w[e] = wsrc[e]; }
        }
    }
    int mine = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) mine += w[e] != SC_DROP;
    int incl = mine;                                  // inclusive scan of the per-thread survivor counts within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int off = incl - mine, cnt = 0;
#pragma unroll
    for (int j = 0; j < SC_THREADS / 32; ++j) { const int t = warp_tot[j]; off += j < wid ? t : 0; cnt += t; }
    const size_t bn = (size_t)b * n;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (w[e] != SC_DROP) {
            if (w[e] >= 0) sc_affine(x[e], y[e], z[e], w[e], K, centres, transform, translate, bn);
            sout[3 * off] = x[e]; sout[3 * off + 1] = y[e]; sout[3 * off + 2] = z[e];
            ++off;
        }
    }
    __syncthreads();
    float* dst = out + ((size_t)b * N + base_s) * 3;
    for (int j = tid; j < 3 * cnt; j += SC_THREADS) dst[j] = sout[j];
    if (blk == nblk - 1 && tid == 0) counts[b] = base_s + cnt;
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

extern "C" int pt_affine_scatter_compact_stage(const float* points, const int32_t* kept_idx, const int32_t* drop_idx,
                                               const float* kept_centres, const float* transform, const float* translate, int B,
                                               int N, int n, int K, int n_drop_entries, float* out, int32_t* counts, void* ws,
                                               size_t ws_bytes, int stages, pt_stream_t stream) {
    PT_REQUIRE((stages & ~(PT_SCATTER_STAGE_MARK | PT_SCATTER_STAGE_COMPACT)) == 0 && stages != 0, "pt_affine_scatter_compact_stage: stages=%d", stages);
    PT_REQUIRE(kept_idx && drop_idx && ws, "pt_affine_scatter_compact: null pointer");
    PT_REQUIRE(!(stages & PT_SCATTER_STAGE_COMPACT) || (points && kept_centres && transform && translate && out && counts),
               "pt_affine_scatter_compact: null pointer");
    PT_REQUIRE(B > 0 && N > 0 && n > 0 && K > 0 && n_drop_entries >= 0 && (long long)n * K < SC_DROP,
               "pt_affine_scatter_compact: bad shape");
    if (ws_bytes < pt_scatter_ws_bytes(B, N)) { set_error("pt_affine_scatter_compact: workspace %zu < %zu", ws_bytes, pt_scatter_ws_bytes(B, N)); return PT_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    int* winner = (int*)ws;
    int* blockcnt = (int*)((char*)ws + align_up((size_t)B * N * sizeof(int), 256));
    const int nblk = ceil_div(N, SC_BLOCK);
    if (stages & PT_SCATTER_STAGE_MARK) {       // index work only: who writes each point, how many points of a block are dropped
        PT_CUDA_OK(cudaMemsetAsync(winner, 0xff, (size_t)B * N * sizeof(int), s));   // -1 = untouched
        PT_CUDA_OK(cudaMemsetAsync(blockcnt, 0, (size_t)B * nblk * sizeof(int), s));  // distinct dropped points per block
        const long long total = (long long)B * ((long long)n * K + n_drop_entries);
        int grid = (int)((total + 255) / 256);
        if (grid > 148 * 16) grid = 148 * 16;
        { ProfScope prof_(PROF_MARK, s); mark_kernel<<<grid, 256, 0, s>>>(kept_idx, drop_idx, B, N, n * K, n_drop_entries, nblk, winner, blockcnt); }
        PT_LAUNCH_CHECK();
    }
    if (stages & PT_SCATTER_STAGE_COMPACT) {
        { ProfScope prof_(PROF_COMPACT, s); compact_kernel<<<dim3(nblk, B), SC_THREADS, 0, s>>>(points, winner, blockcnt, kept_centres, transform, translate, N, n, K, out, counts); }
        PT_LAUNCH_CHECK();
    }
    return PT_OK;
}

extern "C" int pt_affine_scatter_compact(const float* points, const int32_t* kept_idx, const int32_t* drop_idx,
                                         const float* kept_centres, const float* transform, const float* translate, int B,
                                         int N, int n, int K, int n_drop_entries, float* out, int32_t* counts, void* ws,
                                         size_t ws_bytes, pt_stream_t stream) {
    return pt_affine_scatter_compact_stage(points, kept_idx, drop_idx, kept_centres, transform, translate, B, N, n, K, n_drop_entries, out,
                                           counts, ws, ws_bytes, PT_SCATTER_STAGE_MARK | PT_SCATTER_STAGE_COMPACT, stream);
}
