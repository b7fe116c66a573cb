"""Stage-level Python entry points over the C ABI (include/pt_preshape.h).

Each function takes CUDA torch tensors (torch is used for device memory and the current stream only), checks
dtype/contiguity, and launches the corresponding sm_100a kernels.  No function here has a CPU path.
Reference stages: embodiedscan/models/necks/preshape_norm_reverse_drop.py (":line" in the C header).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import ImgPoolParams, ProxyBlockParams, check

# image feature dtypes consumed without conversion (fp16 is what the reference's --amp backbone emits, tools/train.py:93-105;
# it needs the tcgen05 pooling kernel and the shipped geometry, see img_attnpool)
IMG_FEAT_DTYPES = (torch.float32, torch.bfloat16, torch.float16)
RADIUS = 3.0   # DeformablePointCluster(radius=3)  (:23)
MARGIN = 4.0   # DeformablePointCluster(margin=4)  (:23)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: Optional[torch.Tensor], dtype, name: str, optional: bool = False):
    if t is None:
        if optional:
            return None
        raise ValueError(f"{name} is required")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the preshape path has no CPU implementation)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t.data_ptr()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def linspace01(gs: int, device) -> torch.Tensor:
    """torch.linspace(0, 1, gs) exactly as the reference builds it (:41), fp32."""
    return torch.linspace(0, 1, gs, dtype=torch.float32).to(device)


def minmax_centres(points: torch.Tensor, gs: int, lin: Optional[torch.Tensor] = None, margin: float = MARGIN):
    """S1 (:33-51).  points (B,N,3) -> (mn (B,3), mx (B,3), centres (B,gs^3,3))."""
    L = _lib.load()
    B, N, _ = points.shape
    dev = points.device
    lin = linspace01(gs, dev) if lin is None else lin
    mn = torch.empty(B, 3, dtype=torch.float32, device=dev)
    mx = torch.empty_like(mn)
    centres = torch.empty(B, gs ** 3, 3, dtype=torch.float32, device=dev)
    ws = _ws(L.pt_minmax_ws_bytes(B, N), dev)
    check(L.pt_minmax_centres(_chk(points, torch.float32, "points"), B, N, gs, _chk(lin, torch.float32, "lin"), margin,
                              mn.data_ptr(), mx.data_ptr(), centres.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
          "pt_minmax_centres")
    return mn, mx, centres


def ball_query(centres: torch.Tensor, points: torch.Tensor, K: int, radius: float = RADIUS):
    """S2/S4 (:56,:65).  -> idx (B,M,K) int32 (-1 padded), pad_counts (B,M) int32."""
    L = _lib.load()
    B, M, _ = centres.shape
    N = points.shape[1]
    idx = torch.empty(B, M, K, dtype=torch.int32, device=points.device)
    pc = torch.empty(B, M, dtype=torch.int32, device=points.device)
    check(L.pt_ball_query_firstk(_chk(centres, torch.float32, "centres"), _chk(points, torch.float32, "points"), B, M, N, K,
                                 radius, idx.data_ptr(), pc.data_ptr(), _stream()), "pt_ball_query_firstk")
    return idx, pc


def offset_net(points, idx, centres0, mn, mx, w: Dict[str, torch.Tensor], margin: float = MARGIN, want_raw: bool = False):
    """S3 (:87-107, :58-62).  w: conv_w (256,6), conv_b, bn_scale, bn_shift (256), map_w (3,256)."""
    L = _lib.load()
    B, M, K = idx.shape
    N = points.shape[1]
    out = torch.empty(B, M, 3, dtype=torch.float32, device=points.device)
    raw = torch.empty_like(out) if want_raw else None
    f = torch.float32
    check(L.pt_offset_net_fused(_chk(points, f, "points"), _chk(idx, torch.int32, "idx"), _chk(centres0, f, "centres0"),
                                _chk(mn, f, "mn"), _chk(mx, f, "mx"), _chk(w["conv_w"], f, "conv_w"), _chk(w["conv_b"], f, "conv_b"),
                                _chk(w["bn_scale"], f, "bn_scale"), _chk(w["bn_shift"], f, "bn_shift"), _chk(w["map_w"], f, "map_w"),
                                B, M, N, K, w["conv_w"].shape[0], margin, out.data_ptr(), raw.data_ptr() if want_raw else None,
                                _stream()), "pt_offset_net_fused")
    return (out, raw) if want_raw else out


def cluster_dropout(centres, idx, keep1: int, n_keep: int):
    """S5 (:352-420).  -> kept_src (B,n), kept_centres (B,n,3), kept_idx (B,n,K), drop_idx (B,n_drop*K), fps_sel (B,n_drop)."""
    L = _lib.load()
    B, M, K = idx.shape
    dev = idx.device
    n_drop = keep1 - n_keep
    kept_src = torch.empty(B, n_keep, dtype=torch.int32, device=dev)
    kept_centres = torch.empty(B, n_keep, 3, dtype=torch.float32, device=dev)
    kept_idx = torch.empty(B, n_keep, K, dtype=torch.int32, device=dev)
    drop_idx = torch.empty(B, max(n_drop, 0) * K, dtype=torch.int32, device=dev)
    fps_sel = torch.empty(B, max(n_drop, 0), dtype=torch.int32, device=dev)
    check(L.pt_cluster_dropout(_chk(centres, torch.float32, "centres"), _chk(idx, torch.int32, "idx"), B, M, K, keep1, n_keep,
                               kept_src.data_ptr(), kept_centres.data_ptr(), kept_idx.data_ptr(), drop_idx.data_ptr(),
                               fps_sel.data_ptr(), _stream()), "pt_cluster_dropout")
    return kept_src, kept_centres, kept_idx, drop_idx, fps_sel


def point_encoder(points, kept_idx, kept_centres, w: Dict[str, torch.Tensor]):
    """S6 (:126-142) -> point_proxy (B,n,256)."""
    L = _lib.load()
    B, n, K = kept_idx.shape
    N = points.shape[1]
    H = w["conv_w"].shape[0]
    out = torch.empty(B, n, H, dtype=torch.float32, device=points.device)
    f = torch.float32
    check(L.pt_point_encoder_fused(_chk(points, f, "points"), _chk(kept_idx, torch.int32, "kept_idx"),
                                   _chk(kept_centres, f, "kept_centres"), _chk(w["conv_w"], f, "conv_w"),
                                   _chk(w["conv_b"], f, "conv_b"), _chk(w["bn_scale"], f, "bn_scale"),
                                   _chk(w["bn_shift"], f, "bn_shift"), B, n, N, K, H, out.data_ptr(), _stream()),
          "pt_point_encoder_fused")
    return out


def bn_batch_affine_cluster_conv(points, idx, centres, w: Dict[str, torch.Tensor], bn: "torch.nn.modules.batchnorm._BatchNorm"):
    """Train-mode BatchNorm2d of OffsetNetwork / SimplifiedPointNet (:72, :112): batch statistics of the pre-BN conv output
    over (B, M, K) -> (scale, shift) for the fused stage kernel; updates bn.running_mean / running_var in place (N4)."""
    L = _lib.load()
    B, M, K = idx.shape
    H = w["conv_w"].shape[0]
    f = torch.float32
    sums = torch.empty(2 * H, dtype=torch.float64, device=points.device)
    check(L.pt_cluster_conv_bn_stats(_chk(points, f, "points"), _chk(idx, torch.int32, "idx"), _chk(centres, f, "centres"),
                                     _chk(w["conv_w"], f, "conv_w"), _chk(w["conv_b"], f, "conv_b"), B, M, points.shape[1], K, H,
                                     sums.data_ptr(), _stream()), "pt_cluster_conv_bn_stats")
    return _bn_batch_affine(sums, B * M * K, bn, H)


def bn_batch_affine_linear(guide, lin_w, lin_b, bn: "torch.nn.modules.batchnorm._BatchNorm"):
    """Train-mode BatchNorm1d of the heads (:326-330, :445-446, :454-455): batch statistics of the Linear output over the
    rows -> (scale, shift) for pt_heads; updates the running statistics in place (N4)."""
    L = _lib.load()
    c = guide.shape[-1]
    rows = guide.numel() // c
    o = lin_w.shape[0]
    f = torch.float32
    sums = torch.empty(2 * o, dtype=torch.float64, device=guide.device)
    check(L.pt_linear_bn_stats(_chk(guide, f, "guide"), _chk(lin_w, f, "lin_w"), _chk(lin_b, f, "lin_b"), rows, c, o,
                               sums.data_ptr(), _stream()), "pt_linear_bn_stats")
    return _bn_batch_affine(sums, rows, bn, o)


def _bn_batch_affine(sums, count: int, bn, C: int):
    L = _lib.load()
    f = torch.float32
    dev = sums.device
    scale = torch.empty(C, dtype=f, device=dev)
    shift = torch.empty(C, dtype=f, device=dev)
    if count <= 1:
        raise ValueError(f"Expected more than 1 value per channel when training, got {count}")     # as nn.BatchNorm raises
    track = bn.track_running_stats and bn.running_mean is not None
    # momentum=None is torch's cumulative moving average: factor 1 / num_batches_tracked AFTER this step's increment
    momentum = 1.0 / float(int(bn.num_batches_tracked) + 1) if bn.momentum is None and track else float(bn.momentum or 0.0)
    check(L.pt_bn_batch_affine(sums.data_ptr(), count, _chk(bn.weight.detach(), f, "bn.weight"), _chk(bn.bias.detach(), f, "bn.bias"),
                               _chk(bn.running_mean, f, "running_mean") if track else None,
                               _chk(bn.running_var, f, "running_var") if track else None, momentum, float(bn.eps), C,
                               scale.data_ptr(), shift.data_ptr(), _stream()), "pt_bn_batch_affine")
    if track:
        bn.num_batches_tracked += 1
    return scale, shift


_BLOCK_F32 = ("ln1_w", "ln1_b", "pos_bias", "qkv_w", "pp_w", "pp_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b",
              "fc2_w", "fc2_b", "lno_w", "lno_b")
_BLOCK_SPLIT = ("qkv_w_split", "proj_w_split", "fc1_w_split", "fc2_w_split", "pp_w_split")


def make_block_params(w: Dict[str, torch.Tensor]) -> ProxyBlockParams:
    p = ProxyBlockParams()
    for k in _BLOCK_F32:
        setattr(p, k, _chk(w[k], torch.float32, k))
    qb = w.get("qkv_b")
    p.qkv_b = _chk(qb, torch.float32, "qkv_b") if qb is not None else None
    for k in _BLOCK_SPLIT:
        t = w.get(k)
        setattr(p, k, _chk(t, torch.bfloat16, k) if t is not None else None)
    return p


def proxy_block(x, proxy, mask, w: Dict[str, torch.Tensor], heads: int, ws: Optional[torch.Tensor] = None,
                params: Optional[ProxyBlockParams] = None, out: Optional[torch.Tensor] = None):
    """S7 (:206-276 + trailing norm :443/:452).  x (B,n,c), proxy (B,l,c), mask (B,l) bool/uint8 or None."""
    L = _lib.load()
    B, n, c = x.shape
    l = proxy.shape[1]
    hidden = w["fc1_w"].shape[0]
    p = params if params is not None else make_block_params(w)
    if mask is not None and mask.dtype == torch.bool:
        mask = mask.to(torch.uint8)
    need = L.pt_proxy_block_ws_bytes(B, n, l, c, hidden)
    if ws is None or ws.numel() < need:
        ws = _ws(need, x.device)
    out = torch.empty_like(x) if out is None else out
    check(L.pt_proxy_block_fused(_chk(x, torch.float32, "x"), _chk(proxy, torch.float32, "proxy"),
                                 _chk(mask, torch.uint8, "mask", optional=True), ctypes.byref(p), B, n, l, c, heads, hidden,
                                 out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "pt_proxy_block_fused")
    return out


def position_bias(pb, pc, pr):
    """:212-215.  pb (1,n,4,4), pc (1,n,s,1), pr (1,n,1,s) -> (n, s*s)."""
    L = _lib.load()
    n, s = pb.shape[1], pc.shape[2]
    out = torch.empty(n, s * s, dtype=torch.float32, device=pb.device)
    f = torch.float32
    check(L.pt_position_bias(_chk(pb.contiguous(), f, "pb"), _chk(pc.contiguous(), f, "pc"), _chk(pr.contiguous(), f, "pr"),
                             n, s, out.data_ptr(), _stream()), "pt_position_bias")
    return out


def heads(guide, lin_w, lin_b, bn_scale, bn_shift):
    """S8 (:445-446, :454-455).  guide (..., c) -> (..., o)."""
    L = _lib.load()
    c = guide.shape[-1]
    rows = guide.numel() // c
    o = lin_w.shape[0]
    out = torch.empty(*guide.shape[:-1], o, dtype=torch.float32, device=guide.device)
    f = torch.float32
    check(L.pt_heads(_chk(guide, f, "guide"), _chk(lin_w, f, "lin_w"), _chk(lin_b, f, "lin_b"), _chk(bn_scale, f, "bn_scale"),
                     _chk(bn_shift, f, "bn_shift"), rows, c, o, out.data_ptr(), _stream()), "pt_heads")
    return out


_IMG_SPLIT_KEYS = ("w_qc_split", "wk_pad_split", "gk_pad_split", "wv_cat_split", "cproj_split")
_IMG_KEYS = ("w_qc", "q0", "w_kc", "g_k", "w_vc", "h_v", "cproj_w", "cproj_b", "ln_w", "ln_b")


def make_img_params(w: Dict[str, torch.Tensor]) -> ImgPoolParams:
    p = ImgPoolParams()
    for k in _IMG_KEYS:
        setattr(p, k, _chk(w[k], torch.float32, k))
    for k in _IMG_SPLIT_KEYS:      # tensor-core fast path operands (all or none)
        t = w.get(k)
        setattr(p, k, _chk(t, torch.bfloat16, k) if t is not None else None)
    p.variant = int(w.get("variant", _lib.PT_POOL_VARIANT_MMA))
    return p


def img_pool_variant() -> int:
    """Pooling kernel of the 16-bit image path: the tcgen05 / TMEM kernel fed by TMA in class-aligned coordinates
    (csrc/imgpool_umma.cu, 0.60 ms per 64 scenes at the benchmark shape) unless PT_POOL_KERNEL=mma selects the mma.sync kernel
    fed by bulk copies (csrc/imgpool_tc.cu, 0.73 ms).  Both are parity-tested.  Read when the weights are packed (the folded
    channel orders depend on it)."""
    import os
    return _lib.PT_POOL_VARIANT_MMA if os.environ.get("PT_POOL_KERNEL", "umma") == "mma" else _lib.PT_POOL_VARIANT_UMMA


def img_pool_channel_orders(device=None, variant: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Channel orders of the 16-bit image-pool kernels (C=512), see include/pt_preshape.h.
    variant 1 (tcgen05 kernel, csrc/imgpool_umma.cu): position 64 s + r <-> channel s + 8 r for both.
    variant 0 (mma.sync kernel, csrc/imgpool_tc.cu):
    score_order[((p*8 + s)*4 + q)*4 + e] = 128 p + 64 (e >> 1) + s + 16 q + 8 (e & 1)   (w_eff columns)
    sum_order[((sl*8 + s)*4 + q)*2 + e]  = 64 sl + s + 16 q + 8 e                       (weighted-sum columns)."""
    if variant == _lib.PT_POOL_VARIANT_UMMA:
        j = torch.arange(512)
        order = (j >> 6) + 8 * (j & 63)
        assert sorted(order.tolist()) == list(range(512))
        return order.to(device), order.clone().to(device)
    p, s, q, e = torch.meshgrid(torch.arange(4), torch.arange(8), torch.arange(4), torch.arange(4), indexing="ij")
    score = (128 * p + 64 * (e >> 1) + s + 16 * q + 8 * (e & 1)).reshape(-1)
    sl, s, q, e = torch.meshgrid(torch.arange(8), torch.arange(8), torch.arange(4), torch.arange(2), indexing="ij")
    sums = (64 * sl + s + 16 * q + 8 * e).reshape(-1)
    assert sorted(score.tolist()) == list(range(512)) and sorted(sums.tolist()) == list(range(512))
    return score.to(device), sums.to(device)


IMG_STAGE_FRONT, IMG_STAGE_BACK = 1, 2


def img_attnpool(img_feat, w: Dict[str, torch.Tensor], heads: int, ws: Optional[torch.Tensor] = None,
                 params: Optional[ImgPoolParams] = None, stages: int = IMG_STAGE_FRONT | IMG_STAGE_BACK,
                 out: Optional[torch.Tensor] = None):
    """S9 (:335-342, :154-177).  img_feat (B,V,C,H,W) fp32/bf16 -> (B,V,c).
    ``stages``: FRONT (spatial means + query-side projections) and BACK (pooling + value side + LayerNorm) may be issued
    as two calls with the same ``ws`` / ``out`` (pt_img_attnpool_stage); returns (out, ws)."""
    L = _lib.load()
    B, V, C, H, W = img_feat.shape
    c = w["cproj_w"].shape[0]
    if img_feat.dtype == torch.float32:
        dt = _lib.PT_DTYPE_F32
    elif img_feat.dtype == torch.bfloat16:
        dt = _lib.PT_DTYPE_BF16
    elif img_feat.dtype == torch.float16:
        dt = _lib.PT_DTYPE_F16
    else:
        raise ValueError(f"img_feat must be fp32, bf16 or fp16, got {img_feat.dtype}")
    p = params if params is not None else make_img_params(w)
    if dt == _lib.PT_DTYPE_F16 and not (p.variant == _lib.PT_POOL_VARIANT_UMMA and p.w_qc_split and (C, H * W, c, heads) == (512, 225, 256, 8)):
        img_feat, dt = img_feat.float(), _lib.PT_DTYPE_F32      # no 16-bit tensor-core path for this geometry / kernel choice
    need = L.pt_img_attnpool_ws_bytes(B * V, C, H * W, c, heads)
    if ws is None or ws.numel() < need:
        ws = _ws(need, img_feat.device)
    if out is None:
        out = torch.empty(B, V, c, dtype=torch.float32, device=img_feat.device)
    check(L.pt_img_attnpool_stage(_chk(img_feat, img_feat.dtype, "img_feat"), dt, ctypes.byref(p), B * V, C, H * W, c, heads,
                                  out.data_ptr(), ws.data_ptr(), ws.numel(), stages, _stream()), "pt_img_attnpool_stage")
    if stages == (IMG_STAGE_FRONT | IMG_STAGE_BACK):
        return out
    return out, ws


SCATTER_STAGE_MARK, SCATTER_STAGE_COMPACT = 1, 2


def affine_scatter_mark(points, kept_idx, drop_idx):
    """Index half of S10-S12 (needs the indices only): -> the workspace to hand to affine_scatter_compact(..., ws=, marked=True)."""
    L = _lib.load()
    B, N, _ = points.shape
    n, K = kept_idx.shape[1], kept_idx.shape[2]
    ws = _ws(L.pt_scatter_ws_bytes(B, N), points.device)
    check(L.pt_affine_scatter_compact_stage(None, _chk(kept_idx, torch.int32, "kept_idx"), _chk(drop_idx, torch.int32, "drop_idx"), None, None,
                                            None, B, N, n, K, drop_idx.shape[1], None, None, ws.data_ptr(), ws.numel(), SCATTER_STAGE_MARK,
                                            _stream()), "pt_affine_scatter_compact_stage")
    return ws


def affine_scatter_compact(points, kept_idx, drop_idx, kept_centres, transform, translate, ws: Optional[torch.Tensor] = None,
                           marked: bool = False):
    """S10-S12 (:459-525).  -> out (B,N,3) packed per scene, counts (B,) int32 (device).  ``marked``: ``ws`` comes from
    affine_scatter_mark on the same indices (only the affine + compaction half is left to do)."""
    L = _lib.load()
    B, N, _ = points.shape
    n, K = kept_idx.shape[1], kept_idx.shape[2]
    nde = drop_idx.shape[1]
    dev = points.device
    out = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    need = L.pt_scatter_ws_bytes(B, N)
    if ws is None or ws.numel() < need:
        if marked:
            raise ValueError("affine_scatter_compact(marked=True) needs the workspace returned by affine_scatter_mark")
        ws = _ws(need, dev)
    f = torch.float32
    stages = SCATTER_STAGE_COMPACT if marked else SCATTER_STAGE_MARK | SCATTER_STAGE_COMPACT
    check(L.pt_affine_scatter_compact_stage(_chk(points, f, "points"), _chk(kept_idx, torch.int32, "kept_idx"),
                                            _chk(drop_idx, torch.int32, "drop_idx"), _chk(kept_centres, f, "kept_centres"),
                                            _chk(transform, f, "transform"), _chk(translate, f, "translate"), B, N, n, K, nde,
                                            out.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(), stages, _stream()),
          "pt_affine_scatter_compact_stage")
    return out, counts


def proxy_attention_tc(q, k, v, pt, mask, heads: int):
    """The tcgen05 / TMEM attention core on fp32 inputs (test / tool entry): q, k, v (B,n,c), pt (B,l,c), mask (B,l) uint8 or
    None -> o (B,n,c) fp32.  Splits the operands into the bf16 hi/lo planes pt_proxy_attention_tc expects."""
    L = _lib.load()
    B, n, c = q.shape
    l = pt.shape[1]
    rows = B * n
    qk = split_bf16(torch.cat([q, k], -1).reshape(rows, 2 * c).contiguous())                  # (2, rows, 2c)
    vt = split_bf16(v.reshape(rows, c).t().contiguous())                                        # (2, c, rows)
    pts = split_bf16(pt.reshape(B * l, c).contiguous())                                         # (2, B*l, c)
    o = torch.empty(B, n, c, dtype=torch.float32, device=q.device)
    check(L.pt_proxy_attention_tc(qk.data_ptr(), rows * 2 * c, 2 * c, vt.data_ptr(), c * rows, rows, pts.data_ptr(), B * l * c,
                                  _chk(mask, torch.uint8, "mask", optional=True), B, n, l, c, heads, o.data_ptr(), None, 0, _stream()),
          "pt_proxy_attention_tc")
    return o


def aggregate_sample(view_points, extrinsics: torch.Tensor, choices: torch.Tensor) -> torch.Tensor:
    """N3 input side: ``AggregateMultiViewPoints`` (datasets/transforms/multiview.py:224-241) followed by the gather of
    ``PointSample`` (points.py:411-417) on the device.  view_points: list of V (n_v, >=3) ego-frame tensors (or one
    concatenated (T,3) tensor plus ``extrinsics`` per view and offsets inferred from the list); extrinsics (V,4,4) the
    reference's ``depth2img['extrinsic']`` (global -> ego); choices (n,) int64 indices into the concatenation, drawn on the
    host exactly as the reference draws them.  Returns (n,3) fp32 global-frame points in the order of ``choices``."""
    L = _lib.load()
    dev = choices.device
    V = len(view_points)
    cat = torch.cat([p[:, :3].to(dev, torch.float32) for p in view_points], 0).contiguous()
    sizes = torch.tensor([0] + [int(p.shape[0]) for p in view_points], dtype=torch.int64)
    off = torch.cumsum(sizes, 0).to(dev)
    inv = torch.linalg.inv(extrinsics.to(torch.float64)).to(dev, torch.float32).reshape(V, 16).contiguous()
    ch = choices.to(dev, torch.int64).contiguous()
    if ch.numel() and (int(ch.min()) < 0 or int(ch.max()) >= cat.shape[0]):
        raise IndexError("choices out of range")
    out = torch.empty(ch.numel(), 3, dtype=torch.float32, device=dev)
    check(L.pt_aggregate_sample(cat.data_ptr(), off.data_ptr(), V, inv.data_ptr(), ch.data_ptr(), ch.numel(), out.data_ptr(), _stream()),
          "pt_aggregate_sample")
    return out


COLLATE_RECIPROCAL, COLLATE_FLOOR = 1, 2


def sparse_collate(out, counts, voxel_size: float, reciprocal: bool = True, floor: bool = False):
    """N1 hand-off (detectors/sparse_featfusion_grounder_preshape.py:388-391): packed (B,N,3) + counts (B,) ->
    coords (B*N,4) int32 [scene,x,y,z], feats (B*N,3) fp32, total (1,) int32 (device); rows >= total are unspecified.
    ``reciprocal``: quotient as torch's CUDA kernel computes ``p / voxel_size`` (p * fp32(1/voxel_size)); False = IEEE
    division (torch CPU).  ``floor``: round down instead of the truncation of MinkowskiEngine's tensor assignment."""
    L = _lib.load()
    B, N, _ = out.shape
    dev = out.device
    coords = torch.empty(B * N, 4, dtype=torch.int32, device=dev)
    feats = torch.empty(B * N, 3, dtype=torch.float32, device=dev)
    total = torch.empty(1, dtype=torch.int32, device=dev)
    flags = (COLLATE_RECIPROCAL if reciprocal else 0) | (COLLATE_FLOOR if floor else 0)
    check(L.pt_sparse_collate(_chk(out, torch.float32, "out"), _chk(counts, torch.int32, "counts"), B, N, float(voxel_size), flags,
                              coords.data_ptr(), feats.data_ptr(), total.data_ptr(), _stream()), "pt_sparse_collate")
    return coords, feats, total


def gemm_nt(A, W, bias=None, residual=None, act: int = 0, w_split=None):
    """C = act(A @ W^T + bias) + residual.  w_split: (2,N,K) bf16 hi/lo planes -> tcgen05 3xBF16 path."""
    L = _lib.load()
    M, K = A.shape
    N = W.shape[0] if W is not None else w_split.shape[1]
    C = torch.empty(M, N, dtype=torch.float32, device=A.device)
    ws = _ws(L.pt_gemm_ws_bytes(M, N, K), A.device)
    f = torch.float32
    check(L.pt_gemm_nt(_chk(A, f, "A"), _chk(W, f, "W", optional=True), _chk(w_split, torch.bfloat16, "w_split", optional=True),
                       _chk(bias, f, "bias", optional=True), _chk(residual, f, "residual", optional=True), act, M, N, K,
                       C.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "pt_gemm_nt")
    return C


def gemm_tc(a_split, w_split, M, N, K, *, batch=1, a_koff_z=0, w_row_z=0, bias=None, bias_off_z=0, residual=None, act=0,
            C=None, ldc=0, c_off_z=0, c_split=None, ldcs=0, cs_off_z=0, bn=0):
    """General tensor-core GEMM (pt_gemm_tc): a_split (2,a_rows,lda) bf16, w_split (2,w_rows,ldw) bf16."""
    L = _lib.load()
    d = _lib.GemmTcDesc()
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.a_split, d.a_rows, d.a_cols, d.lda, d.a_koff_z = _chk(a_split, torch.bfloat16, "a_split"), a_split.shape[1], a_split.shape[2], a_split.shape[2], a_koff_z
    d.w_split, d.w_rows, d.ldw, d.w_row_z = _chk(w_split, torch.bfloat16, "w_split"), w_split.shape[1], w_split.shape[2], w_row_z
    d.bias, d.bias_off_z = _chk(bias, torch.float32, "bias", optional=True), bias_off_z
    d.residual, d.act = _chk(residual, torch.float32, "residual", optional=True), act
    d.C, d.ldc, d.c_off_z = _chk(C, torch.float32, "C", optional=True), ldc, c_off_z
    if c_split is not None:
        d.c_split, d.cs_plane, d.ldcs, d.cs_off_z = _chk(c_split, torch.bfloat16, "c_split"), c_split[0].numel(), ldcs, cs_off_z
    d.bn = bn
    check(L.pt_gemm_tc(ctypes.byref(d), _stream()), "pt_gemm_tc")


def split_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 (...) -> bf16 (2, ...) hi/lo planes, hi = bf16(x), lo = bf16(x - hi)."""
    L = _lib.load()
    out = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    check(L.pt_split_bf16(_chk(x, torch.float32, "x"), x.numel(), out.data_ptr(), _stream()), "pt_split_bf16")
    return out


def layernorm(x, w, b, add=None):
    L = _lib.load()
    c = x.shape[-1]
    rows = x.numel() // c
    out = torch.empty_like(x)
    f = torch.float32
    check(L.pt_layernorm(_chk(x, f, "x"), _chk(w, f, "w"), _chk(b, f, "b"), _chk(add, f, "add", optional=True),
                         add.shape[0] if add is not None else 1, rows, c, out.data_ptr(), _stream()), "pt_layernorm")
    return out
