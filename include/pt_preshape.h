/*
 * pt_preshape.h — C ABI of the B200 (sm_100a) implementation of the proxy-attention point-cloud
 * preshaping hot path of pqh22/ProxyTransformation.
 *
 * The reference has no FFI for this path: it is a Python nn.Module
 * (embodiedscan/models/necks/preshape_norm_reverse_drop.py, cited below as ":line") whose native
 * kernels live in pytorch3d / ATen.  Every entry point here replaces one stage of that module's
 * forward (:424-469) and is what the Python host module (proxytransformation_b200/necks/preshape.py,
 * a drop-in for ProxyTransformationNormReverse) binds through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - tensors are dense, row-major, fp32 unless stated; indices are int32 (the reference's int64
 *     indices never leave the module); B scenes, N points/scene, M = grid_size^3 centres,
 *     K = num_sub neighbours, n = kept clusters, c = embed_dim, l = proxy tokens;
 *   - launches go to `stream` (a cudaStream_t passed as void*); nothing synchronises, nothing allocates;
 *     scratch comes from the caller (`ws`, sized by the matching *_ws_bytes function);
 *   - return value: 0 = ok, negative = error (PT_ERR_*); pt_last_error_string() describes the last one
 *     on the calling thread.
 */
#ifndef PT_PRESHAPE_H_
#define PT_PRESHAPE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_OK 0
#define PT_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define PT_ERR_CUDA (-2)      /* a CUDA runtime call or launch failed */
#define PT_ERR_WORKSPACE (-3) /* workspace too small */

#define PT_DTYPE_F32 0
#define PT_DTYPE_BF16 1
#define PT_DTYPE_F16 2   /* image features only: tensor-core pooling path (tcgen05 kernel), what the reference's --amp backbone emits */

typedef void* pt_stream_t; /* cudaStream_t */

int pt_abi_version(void);
const char* pt_last_error_string(void);
/* Number of kernels launched through this library by the calling process so far (bench.py's gpu_launches). */
int64_t pt_launch_count(void);

/* Per-kernel device timing for the benchmark: while enabled every kernel launched through this library is bracketed
 * by CUDA events on its stream.  pt_profile_enable(on) also clears the records; pt_profile_read synchronises on the
 * recorded events and returns the summed duration and the number of launches of one kernel kind. */
int pt_profile_enable(int on);
int pt_profile_num_tags(void);
const char* pt_profile_tag_name(int tag);
int pt_profile_read(int tag, double* total_ms, int64_t* launches);
/* Every recorded launch in launch order: tag, start and end in milliseconds after the first record's start (event timestamps, so
 * records of different streams share one time axis) — which kernels of the two streams of a forward actually ran side by side. */
int pt_profile_timeline(int* tags, double* start_ms, double* end_ms, int max_records, int* n_records);

/* ---- S1 grid prior — DeformablePointCluster.init_uniform_cluster_center (:33-51) --------------------
 * mn/mx (B,3) per-axis min/max over the scene; centres (B,M,3) = (mn + margin) + lin3[j] * ((mx - mn) - 2*margin),
 * j = ix*gs^2 + iy*gs + iz ('ij' meshgrid).  `lin` is torch.linspace(0,1,gs) computed by the host (fp32).
 * ws: pt_minmax_ws_bytes(B, N). */
size_t pt_minmax_ws_bytes(int B, int N);
int pt_minmax_centres(const float* points, int B, int N, int gs, const float* lin, float margin, float* mn, float* mx,
                      float* centres, void* ws, size_t ws_bytes, pt_stream_t stream);

/* ---- S2/S4 ball query — pytorch3d.ops.ball_query call sites (:56, :65) ------------------------------
 * For every centre the first K point indices, in ascending index, with ((dx*dx)+(dy*dy))+(dz*dz) < radius^2
 * (fp32, no FMA contraction); unfilled slots -1.  pad_counts (B,M) = number of -1 slots, may be NULL. */
int pt_ball_query_firstk(const float* centres, const float* points, int B, int M, int N, int K, float radius,
                         int32_t* idx, int32_t* pad_counts, pt_stream_t stream);

/* ---- S3 offset network — OffsetNetwork.forward (:87-107) + tanh*margin, add, clamp (:58-62) ---------
 * Gathers the K neighbours of `idx` from `points` (pad slot -> (0,0,0), masked_gather :627-672), builds the
 * 6 features [rel (zeroed where the gathered point == (0,0,0)), abs], conv1x1 6->H + BatchNorm(eval, given as
 * per-channel scale/shift) + ReLU, mean over K, H->3 map (no bias), tanh*margin, + centre, clamp to [mn,mx].
 * conv_w (H,6), conv_b/bn_scale/bn_shift (H), map_w (3,H); H must be 256.  raw_offsets (B,M,3) may be NULL. */
int pt_offset_net_fused(const float* points, const int32_t* idx, const float* centres0, const float* mn,
                        const float* mx, const float* conv_w, const float* conv_b, const float* bn_scale,
                        const float* bn_shift, const float* map_w, int B, int M, int N, int K, int H, float margin,
                        float* centres_out, float* raw_offsets, pt_stream_t stream);

/* ---- S5 cluster dropout — dynamic_cluster_dropout (:352-420) ----------------------------------------
 * Per scene: pad count per cluster -> STABLE ascending order (pinned tie rule) -> first keep1 -> farthest point
 * sampling of n_drop = keep1 - n_keep centres (start 0, first arg-max) = clusters to drop -> ascending
 * complement truncated to n_keep.  Outputs: kept_src (B,n_keep) original cluster id of each kept slot,
 * kept_centres (B,n_keep,3), kept_idx (B,n_keep,K), drop_idx (B,n_drop,K) (rows of the FPS-selected clusters,
 * in FPS order), fps_sel (B,n_drop) FPS picks as positions in the keep1 list (may be NULL). */
int pt_cluster_dropout(const float* centres, const int32_t* idx, int B, int M, int K, int keep1, int n_keep,
                       int32_t* kept_src, float* kept_centres, int32_t* kept_idx, int32_t* drop_idx,
                       int32_t* fps_sel, pt_stream_t stream);

/* ---- S6 point proxies — SimplifiedPointNet.forward (:126-142) ----------------------------------------
 * Same features/conv/BN/ReLU as S3 with the kept centres, max over K -> point_proxy (B,n,H), H = 256. */
int pt_point_encoder_fused(const float* points, const int32_t* kept_idx, const float* kept_centres,
                           const float* conv_w, const float* conv_b, const float* bn_scale, const float* bn_shift,
                           int B, int n, int N, int K, int H, float* point_proxy, pt_stream_t stream);

/* ---- S7 ProxyBlock (:273-276) + ProxyAttention (:206-257) + timm Mlp + trailing LayerNorm (:443/:452) ----
 * out = LN_out( x1 + fc2(GELU_erf(fc1(LN2(x1)))) ),  x1 = x + proj(attn(LN1(x) + pos_bias, proxy, mask)).
 * pos_bias (n,c) is the parameter-only table bilinear(pb 4x4 -> s x s) + pc + pr (:212-215), see pt_position_bias.
 * mask (B,l) uint8, 1 = real token, NULL = no mask (image branch).  Dense layers run on tcgen05 tensor cores
 * with fp32-accurate 3xBF16 operand splitting when the *_split weight pointers are set, else on fp32 CUDA cores. */
typedef struct pt_proxy_block_params {
    const float *ln1_w, *ln1_b;       /* (c) */
    const float* pos_bias;            /* (n,c) */
    const float* qkv_w;               /* (3c,c) */
    const float* qkv_b;               /* (3c) or NULL (qkv_bias=False, every shipped config) */
    const float *pp_w, *pp_b;         /* proxy_proj (c,c),(c) */
    const float *proj_w, *proj_b;     /* (c,c),(c) */
    const float *ln2_w, *ln2_b;       /* (c) */
    const float *fc1_w, *fc1_b;       /* (hid,c),(hid) */
    const float *fc2_w, *fc2_b;       /* (c,hid),(c) */
    const float *lno_w, *lno_b;       /* trailing text_norm[i] / img_norm[i] (c) */
    /* optional bf16 hi/lo splits of the four big weights for the tensor-core path: each points at
     * [2][rows][cols] bf16 (hi plane then lo plane), or NULL */
    const void *qkv_w_split, *proj_w_split, *fc1_w_split, *fc2_w_split, *pp_w_split;
} pt_proxy_block_params;

size_t pt_proxy_block_ws_bytes(int B, int n, int l, int c, int hidden);
int pt_proxy_block_fused(const float* x, const float* proxy, const uint8_t* mask, const pt_proxy_block_params* p,
                         int B, int n, int l, int c, int heads, int hidden, float* out, void* ws, size_t ws_bytes,
                         pt_stream_t stream);

/* Position-bias table of one ProxyAttention: out (n, s*s) = bilinear_{4x4 -> s x s, align_corners=False}(pb[m])
 * + pc[m,r] + pr[m,q]   (:212-215).  pb (n,4,4), pc (n,s), pr (n,s). */
int pt_position_bias(const float* pb, const float* pc, const float* pr, int n, int s, float* out, pt_stream_t stream);

/* The two-stage proxy attention core (:225-252) alone, tcgen05 / TMEM form, on pre-split bf16 hi/lo operand planes
 * (heads of 32 channels, any n <= 1024 — the clusters are streamed in key tiles of 256 / row tiles of 128 — and l <= 256):
 * qk_split [rows][ldq] holds Q at column 0 and K at column c, vt_split [c][ldv] holds V^T (column = scene*n + cluster; scenes whose
 * first column is off the 16-byte grid are staged element by element — pt_proxy_block_fused pads every scene to a multiple of 8
 * columns instead), pt_split [B*l][c] the projected proxies; every lo plane lies *_plane elements behind its hi plane.  Writes o
 * (fp32, optional) and / or o_split hi/lo planes, (B*n, c).  pt_proxy_block_fused uses it when the shape fits and the mma.sync
 * kernel otherwise (other head sizes, l > 256).  More than 148 (scene, head) pairs with l <= 224 run as two 8-warp CTAs per SM
 * (key tiles of 128), anything else as one 16-warp CTA per SM (key tiles of 256); same results to rounding. */
int pt_proxy_attention_tc(const void* qk_split, long long qk_plane, int ldq, const void* vt_split, long long vt_plane, long long ldv,
                          const void* pt_split, long long pt_plane, const uint8_t* mask, int B, int n, int l, int c, int heads,
                          float* o, void* o_split, long long o_plane, pt_stream_t stream);

/* ---- S8 heads — Linear + BatchNorm1d(eval) (:445-446, :454-455) ---------------------------------------
 * out (rows,o) = (guide (rows,c) @ lin_w(o,c)^T + lin_b) * bn_scale + bn_shift, o <= 16. */
int pt_heads(const float* guide, const float* lin_w, const float* lin_b, const float* bn_scale, const float* bn_shift,
             int rows, int c, int o, float* out, pt_stream_t stream);

/* ---- training-mode BatchNorm statistics (forward only, no backward yet) ---------------
 * In train() mode the BatchNorm2d of OffsetNetwork / SimplifiedPointNet (:72, :112) and the BatchNorm1d of the two heads
 * (:326-330) normalise with batch statistics.  The fused stage kernels above take a per-channel (scale, shift), so a
 * train-mode forward is: a statistics pass over the layer's pre-BN output (fp64 sums, sums[0..C) = sum, sums[C..2C) = sum
 * of squares; the call zeroes them first), pt_bn_batch_affine, then the unchanged stage kernel.
 *   pt_cluster_conv_bn_stats: pre-BN conv output of pt_offset_net_fused / pt_point_encoder_fused, count = B*M*K (padded
 *                             slots included, as in the reference)
 *   pt_linear_bn_stats:       pre-BN Linear output of pt_heads, count = rows
 *   pt_bn_batch_affine:       scale = gamma / sqrt(var_biased + eps), shift = beta - mean * scale; when running_mean /
 *                             running_var are given they become (1-momentum) * running + momentum * batch (unbiased
 *                             variance), what nn.BatchNorm does in train() mode. */
int pt_cluster_conv_bn_stats(const float* points, const int32_t* idx, const float* centres, const float* conv_w,
                             const float* conv_b, int B, int M, int N, int K, int H, double* sums, pt_stream_t stream);
int pt_linear_bn_stats(const float* guide, const float* lin_w, const float* lin_b, int rows, int c, int o, double* sums,
                       pt_stream_t stream);
int pt_bn_batch_affine(const double* sums, long long count, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float momentum, float eps, int C, float* scale, float* shift, pt_stream_t stream);

/* ---- S9 image proxies — get_img_proxy (:335-342) + AttentionPool2d.forward (:154-177) -----------------
 * Only token 0 of the 226-token attention is kept (:177), so the stage is evaluated in single-query form:
 * one pass for the per-channel spatial mean, folded q/k projections, one pass for scores -> softmax ->
 * attention-weighted feature sum, folded v/c projections, LayerNorm.  img_feat (BV, C, HW) fp32, bf16 or fp16 (PT_DTYPE_F16:
 * tensor-core path of PT_POOL_VARIANT_UMMA only, i.e. the shipped geometry with the split weights below).
 * The folded weights are produced once per weight load by the host (see pt_img_pool_params). */
typedef struct pt_img_pool_params {
    const float* w_qc;   /* (c,C)   = Wq @ Wc                                             */
    const float* q0;     /* (c)     = Wq @ (bc + pos[0]) + bq                              */
    const float* w_kc;   /* (c,C)   = Wk @ Wc                                             */
    const float* g_k;    /* (T,c)   = (pos + bc) @ Wk^T        (T = HW+1 tokens)           */
    const float* w_vc;   /* (c,C)   = Wv @ Wc                                             */
    const float* h_v;    /* (T,c)   = (pos + bc) @ Wv^T + bv                               */
    const float *cproj_w, *cproj_b; /* (c,c),(c) */
    const float *ln_w, *ln_b;       /* norm_img (c) */
    /* Optional bf16 hi/lo planes ([2][rows][cols], see pt_split_bf16) for the tensor-core fast path taken when img_feat
     * is bf16 / fp16 and (C, HW, c, heads) = (512, 225, 256, 8); all five or none:
     *   w_qc_split   (c, C)            W_qc
     *   wk_pad_split (heads*C, 64)     row h*C + j, col e < hd: (Wk Wc)[h*hd+e][score_order[j]]; cols >= hd zero
     *   gk_pad_split (heads*228, 64)   row h*228 + t, col e < hd: g_k[t][h*hd+e]; rows t >= T and cols >= hd zero
     *   wv_cat_split (c, 768)          cols j < C: (Wv Wc)[row][sum_order[j]] ; cols C + t: h_v[t][row] (t < T), rest zero
     *   cproj_split  (c, c)            c_proj weight
     * The pool kernel reads the raw 450-byte-pitch channel rows with ldmatrix, which forces it to walk the channels by
     * residue class s = channel mod 8 (225 = 1 mod 8: a class shares its 16-byte alignment).  The channel orders that
     * follow are absorbed into the two weight matrices above:
     *   score_order[((p*8 + s)*4 + q)*4 + e] = 128 p + 64 (e >> 1) + s + 16 q + 8 (e & 1)      p<4, s<8, q<4, e<4
     *   sum_order[((sl*8 + s)*4 + q)*2 + e]  = 64 sl + s + 16 q + 8 e                          sl<8, s<8, q<4, e<2 */
    const void *w_qc_split, *wk_pad_split, *gk_pad_split, *wv_cat_split, *cproj_split;
    /* Which pooling kernel the two channel orders above were folded for:
     *   PT_POOL_VARIANT_MMA  (0) img_pool_mma_kernel (mma.sync + ldmatrix on the raw rows), orders as stated above
     *   PT_POOL_VARIANT_UMMA (1) img_pool_umma_kernel (tcgen05 / TMEM, TMA-fed; csrc/imgpool_umma.cu; what the Python module packs by
     *                            default): score_order[64 s + r] = sum_order[64 s + r] = s + 8 r   (residue class s, row r of the class) */
    int variant;
} pt_img_pool_params;
#define PT_POOL_VARIANT_MMA 0
#define PT_POOL_VARIANT_UMMA 1

size_t pt_img_attnpool_ws_bytes(int BV, int C, int HW, int c, int heads);
int pt_img_attnpool(const void* img_feat, int img_dtype, const pt_img_pool_params* p, int BV, int C, int HW, int c,
                    int heads, float* img_proxy, void* ws, size_t ws_bytes, pt_stream_t stream);

/* The same stage in two calls, so that the host can run the first one concurrently with the geometric stages on another
 * stream: FRONT = pass over the features for the spatial means + the query-side projections (HBM-bound, few SM resources),
 * BACK = pooling pass + value-side projections + LayerNorm.  Both calls get the same arguments and the same workspace, whose
 * contents carry the state between them; stages = FRONT | BACK is pt_img_attnpool. */
#define PT_IMG_STAGE_FRONT 1
#define PT_IMG_STAGE_BACK 2
int pt_img_attnpool_stage(const void* img_feat, int img_dtype, const pt_img_pool_params* p, int BV, int C, int HW, int c,
                          int heads, float* img_proxy, void* ws, size_t ws_bytes, int stages, pt_stream_t stream);

/* Debug only: per-phase SM-clock cycles of CTA 0 of the bf16 image-pool kernel, accumulated while PT_POOL_DEBUG has bit 8
 * set: [0] view barrier, [1] operand wait, [2] score MMAs, [3] score exchange + softmax, [4] weighted sums. */
int pt_debug_pool_trace(unsigned long long* out8, int reset);
/* Per-role SM-clock trace of CTA 0 of img_pool_umma_kernel (PT_UMMA_DEBUG bit 16; slots in csrc/imgpool_umma.cu). */
int pt_debug_umma_trace(unsigned long long* out16, int reset);
/* Debug only (PT_POOL_DEBUG bit 64): when the workspace handed to pt_img_attnpool[_stage] is at least
 * pt_img_attnpool_ws_bytes() + BV * 8 * 256 * 4 bytes, the bf16 pool kernel also writes the scaled scores [BV][8 heads][256]
 * (fp32, tokens 0..225) behind the regular workspace; tools/pool_check.py compares them with a float64 evaluation. */
/* Debug only (PT_POOL_DEBUG bit 32): (id, SM clock) event pairs logged by CTA 0 for views 40..43; returns the count and
 * clears the log.  Decoded by tools/pool_events.py. */
int pt_debug_pool_events(long long* out, int max_events);

/* ---- S10-S12 affine (:459-462) + pt_replace (:472-498) + remove_points_by_index (:501-525) ------------
 * new = (T[m] @ (p - centre[m]) + centre[m]) + t[m] for every valid (m,k); duplicate destinations resolved by the
 * pinned rule "largest flat m*K+k wins"; points listed in drop_idx (>= 0) are removed; survivors are written in
 * ascending original order, packed per scene at out + b*N*3; counts (B) = survivors per scene.
 * ws: pt_scatter_ws_bytes(B, N). */
size_t pt_scatter_ws_bytes(int B, int N);
int pt_affine_scatter_compact(const float* points, const int32_t* kept_idx, const int32_t* drop_idx,
                              const float* kept_centres, const float* transform, const float* translate, int B, int N,
                              int n, int K, int n_drop_entries, float* out, int32_t* counts, void* ws, size_t ws_bytes,
                              pt_stream_t stream);
/* The same in two calls with the same ws: PT_SCATTER_STAGE_MARK is index work only (who writes each point, how many points of a block
 * are dropped: needs kept_idx / drop_idx, the other pointers may be null) and can be issued as soon as the cluster dropout is done;
 * PT_SCATTER_STAGE_COMPACT applies the affine maps and writes the survivors. */
#define PT_SCATTER_STAGE_MARK 1
#define PT_SCATTER_STAGE_COMPACT 2
int pt_affine_scatter_compact_stage(const float* points, const int32_t* kept_idx, const int32_t* drop_idx,
                                    const float* kept_centres, const float* transform, const float* translate, int B, int N,
                                    int n, int K, int n_drop_entries, float* out, int32_t* counts, void* ws, size_t ws_bytes,
                                    int stages, pt_stream_t stream);

/* ---- N3 input side: AggregateMultiViewPoints (datasets/transforms/multiview.py:224-241) + the gather of PointSample
 * (datasets/transforms/points.py:411-417), fused so that only the sampled points are transformed.
 * points_cat (T,3): per-view ego-frame points back to back; view_off (V+1) int64 row offsets; ego2global (V,16) row-major
 * inverses of the views' `extrinsic` matrices (the reference solves extrinsic x = [p;1]); choices (n) int64 indices into
 * the concatenation, drawn by the data loader (np.random.choice stays on the host); out (n,3) in the order of choices. */
int pt_aggregate_sample(const float* points_cat, const long long* view_off, int V, const float* ego2global,
                        const long long* choices, long long n, float* out, pt_stream_t stream);

/* ---- N1 hand-off to the sparse backbone (detectors/sparse_featfusion_grounder_preshape.py:388-391) ------------
 * ME.utils.batch_sparse_collate([(p[:, :3] / voxel_size, p) for p in points]) on the packed result of
 * pt_affine_scatter_compact: coords (T,4) int32 rows [scene, x, y, z], feats (T,3) fp32 rows = the coordinates themselves,
 * scenes in order, T = sum(counts) written to *total.  coords / feats must hold B*N rows.  float -> int32 truncates
 * toward zero (tensor assignment in MinkowskiEngine's sparse_collate) unless PT_COLLATE_FLOOR; the quotient is the IEEE
 * division torch performs on the CPU, or with PT_COLLATE_RECIPROCAL the multiplication by fp32(1/voxel_size) of torch's
 * CUDA kernel.  MinkowskiEngine is not vendored by the reference (version unpinned): its semantics are restated, parity for
 * this entry point is pinned against the torch expression only. */
#define PT_COLLATE_RECIPROCAL 1
#define PT_COLLATE_FLOOR 2
int pt_sparse_collate(const float* packed, const int32_t* counts, int B, int N, float voxel_size, int flags,
                      int32_t* coords, float* feats, int32_t* total, pt_stream_t stream);

/* ---- building blocks exposed for tests ------------------------------------------------------------------ */
/* C (M,N) = act(A (M,K) @ W (N,K)^T + bias) + residual ; act: 0 none, 1 GELU(erf).  bias/residual may be NULL.
 * w_split: optional [2][N][K] bf16 hi/lo planes of W -> tcgen05 path (A is split on the fly); NULL -> fp32 CUDA cores. */
size_t pt_gemm_ws_bytes(int M, int N, int K);
int pt_gemm_nt(const float* A, const float* W, const void* w_split, const float* bias, const float* residual, int act,
               int M, int N, int K, float* C, void* ws, size_t ws_bytes, pt_stream_t stream);
/* General (batched, pre-split operands, optional split output) form of the tensor-core GEMM:
 *   C_z[M,N] = act(A[:, z*a_koff_z : +K] W_z[N,K]^T + bias_z) + residual_z,  z in [0,batch)
 * A: bf16 planes [2][a_rows][lda] (hi then lo), a_cols readable columns (reads past a_cols are zero);
 * W: bf16 planes [2][w_rows][ldw], batch z uses rows [z*w_row_z, z*w_row_z+N);
 * C (fp32, optional): element (row,n) of batch z at C + z*c_off_z + row*ldc + n; residual is laid out like C;
 * c_split (optional): bf16 hi plane of the result with pitch ldcs, lo plane cs_plane elements later.
 * K % 64 == 0, N % 4 == 0; bn = tile width (32/64/128/256) or 0 = automatic. */
typedef struct pt_gemm_tc_desc {
    int M, N, K, batch;
    const void* a_split; int a_rows, a_cols, lda, a_koff_z;
    const void* w_split; int w_rows, ldw, w_row_z;
    const float* bias; long long bias_off_z;
    const float* residual; int act;
    float* C; int ldc; long long c_off_z;
    void* c_split; long long cs_plane; int ldcs; long long cs_off_z;
    int bn;
} pt_gemm_tc_desc;
int pt_gemm_tc(const pt_gemm_tc_desc* desc, pt_stream_t stream);
/* bf16 hi/lo split of an fp32 matrix: out [2][rows][cols] bf16, hi = bf16(x), lo = bf16(x - hi). */
int pt_split_bf16(const float* x, int64_t count, void* out, pt_stream_t stream);
/* out (rows,c) = LayerNorm(x) * w + b (+ add[row % add_rows]) ; eps 1e-5 ; c % 32 == 0, c <= 1024. */
int pt_layernorm(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                 float* out, pt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PT_PRESHAPE_H_ */
