"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Stage-by-stage CPU restatement of ``ProxyTransformationNormReverse.forward``
(reference file ``embodiedscan/models/necks/preshape_norm_reverse_drop.py``,
cited below as ``:line``).  Functional: weights come in as a ``state_dict`` with
the reference's keys; every stage returns plain tensors and the driver records
all intermediates in a ``trace`` dict so each CUDA kernel can be checked on
identical inputs.

Pinned semantics where the reference is order-dependent:
  * ``torch.argsort`` at :378 -> ``stable=True`` (ties broken by cluster index);
  * ``pt_replace`` :495 (index_put_ with duplicate destinations) -> the write
    with the largest flat ``(m, k)`` index wins;
  * eval mode, fp32, no autograd — the mode the CUDA path implements.

Training mode (oracle side only so far): ``forward(..., train={})`` evaluates the BatchNorm layers with
batch statistics (:72,:112,:326-330 in ``train()``), returns their updated running statistics in ``train["bn"]`` and keeps
the whole computation differentiable for torch autograd (index work is done on detached tensors, the scatter is an
index_put on unique destinations), with every stochastic layer at rate 0 (Dropout / DropPath need the reference's RNG
stream and are not restated).  Pinned against the reference in ``train()`` mode by tests/golden/c1_train.npz.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import build as _build

RADIUS = 3.0       # :23 DeformablePointCluster(radius=3)
MARGIN = 4.0       # :23 margin=4
EMPTY_DROP = 0.3   # :352 empty_drop=0.3
BN_EPS = 1e-5
LN_EPS = 1e-5

_lib = None


def _geom():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_build.build())
        i64, fp, ip = ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p
        lib.oracle_ball_query.argtypes = [fp, fp, i64, i64, i64, i64, ctypes.c_float, ip, fp, ip]
        lib.oracle_ball_query.restype = None
        lib.oracle_fps.argtypes = [fp, i64, i64, i64, ip, fp]
        lib.oracle_fps.restype = None
        _lib = lib
    return _lib


# ----------------------------------------------------------------------------- pytorch3d boundary
def ball_query(p1: torch.Tensor, p2: torch.Tensor, K: int, radius: float = RADIUS, want_scanned: bool = False):
    """pytorch3d.ops.ball_query(p1, p2, K, radius) semantics (:56, :65): first K
    indices of p2 in ascending order with d2 < r^2; idx pad -1; knn pad 0.0.
    Index work on detached tensors: no gradient flows through the search (idx is an integer output and the gathered
    points come from p2, which carries no gradient on this path)."""
    p1 = p1.detach().contiguous().float()
    p2 = p2.detach().contiguous().float()
    B, M, _ = p1.shape
    N = p2.shape[1]
    idx = torch.empty(B, M, K, dtype=torch.int64)
    knn = torch.empty(B, M, K, 3, dtype=torch.float32)
    scanned = torch.empty(B, M, dtype=torch.int64) if want_scanned else None
    _geom().oracle_ball_query(p1.data_ptr(), p2.data_ptr(), B, M, N, K, float(radius), idx.data_ptr(), knn.data_ptr(),
                              scanned.data_ptr() if want_scanned else None)
    return (idx, knn, scanned) if want_scanned else (idx, knn)


def ball_query_torch(p1, p2, K, radius=RADIUS):
    """Pure-torch cross-check of ``ball_query`` for small sizes (same op order:
    ((dx*dx)+(dy*dy))+(dz*dz), torch elementwise kernels do not contract to FMA)."""
    d = p1[:, :, None, :] - p2[:, None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    hit = d2 < radius * radius
    rank = hit.cumsum(-1) - 1
    B, M, N = hit.shape
    idx = torch.full((B, M, K), -1, dtype=torch.int64)
    sel = hit & (rank < K)
    b, m, j = sel.nonzero(as_tuple=True)
    idx[b, m, rank[b, m, j]] = j
    return idx, masked_gather(p2, idx)


def masked_gather(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """:627-672 — gather with -1 -> 0.0 padding, (B,P,3) x (B,M,K) -> (B,M,K,3)."""
    B, M, K = idx.shape
    safe = idx.clamp(min=0).reshape(B, M * K)
    out = torch.gather(points, 1, safe[..., None].expand(-1, -1, 3)).reshape(B, M, K, 3)
    return out.masked_fill((idx < 0)[..., None], 0.0)


def farthest_point_indices(pts: torch.Tensor, K: int) -> torch.Tensor:
    """pytorch3d.ops.sample_farthest_points(points, K=K)[1] (:393; in-tree pin :527-625)."""
    pts = pts.detach().contiguous().float()
    B, P, _ = pts.shape
    out = torch.empty(B, K, dtype=torch.int64)
    scratch = torch.empty(P, dtype=torch.float32)
    _geom().oracle_fps(pts.data_ptr(), B, P, K, out.data_ptr(), scratch.data_ptr())
    return out


# ----------------------------------------------------------------------------- S1 grid prior
def grid_prior(P: torch.Tensor, gs: int):
    """:33-51.  C0 = (mn + 4) + grid * ((mx - mn) - 8), grid in 'ij' order."""
    mn = P.min(dim=1, keepdim=True)[0]
    mx = P.max(dim=1, keepdim=True)[0]
    lin = torch.linspace(0, 1, gs)
    gx, gy, gz = torch.meshgrid(lin, lin, lin, indexing="ij")
    grid = torch.stack([gx, gy, gz], dim=-1).reshape(-1, 3).unsqueeze(0)
    c0 = mn + MARGIN + grid * (mx - mn - 2 * MARGIN)
    return c0, mn, mx


# ----------------------------------------------------------------------------- S3 / S6 shared feature + conv-bn-relu
def _cluster_features(centre: torch.Tensor, cluster: torch.Tensor) -> torch.Tensor:
    """:93-99 / :131-137 -> (b,m,k,6) = [rel (zeroed where the gathered point is exactly (0,0,0)), abs]."""
    rel = cluster - centre.unsqueeze(2)
    pad = (cluster == 0).all(dim=-1)
    rel = rel.masked_fill(pad[..., None], 0.0)
    return torch.cat([rel, cluster], dim=-1)


def _batch_norm(sd: Dict[str, torch.Tensor], prefix: str, y: torch.Tensor, train: Optional[dict]) -> torch.Tensor:
    """BatchNorm over dim 1.  eval (``train is None``): running statistics.  train(): batch statistics (biased variance)
    and the running statistics after this step (momentum 0.1, unbiased variance, num_batches_tracked + 1) recorded in
    ``train["bn"]`` under the state_dict keys — what nn.BatchNorm{1,2}d does in train() mode."""
    if train is None:
        return F.batch_norm(y, sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"], sd[f"{prefix}.weight"],
                            sd[f"{prefix}.bias"], training=False, eps=BN_EPS)
    rm, rv = sd[f"{prefix}.running_mean"].detach().clone(), sd[f"{prefix}.running_var"].detach().clone()
    out = F.batch_norm(y, rm, rv, sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], training=True, momentum=0.1, eps=BN_EPS)
    upd = train.setdefault("bn", {})
    upd[f"{prefix}.running_mean"], upd[f"{prefix}.running_var"] = rm, rv
    upd[f"{prefix}.num_batches_tracked"] = sd[f"{prefix}.num_batches_tracked"] + 1
    return out


def _conv_bn_relu(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, train: Optional[dict] = None) -> torch.Tensor:
    """nn.Sequential(Conv2d(6,256,1), BatchNorm2d(256), ReLU) on (b,m,k,6) -> (b,256,m,k)."""
    y = F.conv2d(x.permute(0, 3, 1, 2), sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"])
    return F.relu(_batch_norm(sd, f"{prefix}.1", y, train))


def offset_network(sd, centre, cluster, prefix="get_deformable_cluster.get_offsets", train=None):
    """:87-107 -> raw offsets (b,m,3) (before tanh*4)."""
    y = _conv_bn_relu(sd, f"{prefix}.mlp", _cluster_features(centre, cluster), train)
    z = y.mean(dim=-1)                                       # (b,256,m), padded slots included
    return F.conv1d(z, sd[f"{prefix}.channel_mapper.weight"]).transpose(-2, -1)


def deform_cluster(sd, P, gs, K, trace=None, train=None):
    """:53-67 DeformablePointCluster.forward."""
    c0, mn, mx = grid_prior(P, gs)
    idx1, knn1 = ball_query(c0, P, K)
    raw = offset_network(sd, c0, knn1, train=train)
    off = raw.tanh() * MARGIN
    c1 = c0 + off
    cc = torch.max(torch.min(c1, mx), mn)
    idx2, cl2 = ball_query(cc, P, K)
    if trace is not None:
        trace.update(mn=mn, mx=mx, c0=c0, idx1=idx1, raw_offsets=raw, centres=cc, idx2=idx2)
    return cc, cl2, idx2


# ----------------------------------------------------------------------------- S5 cluster dropout
def cluster_dropout(cluster, centre, idx, ddr: float, trace=None):
    """:352-420 with the stable-argsort pin.  Returns (cluster', centre', idx', drop_idx)."""
    B, M, K, _ = cluster.shape
    pc = (idx == -1).sum(dim=2)
    keep1_n = M - int(M * EMPTY_DROP)
    keep1 = torch.argsort(pc, dim=1, stable=True)[:, :keep1_n]
    bi = torch.arange(B)[:, None]
    u_c, u_cl, u_idx = centre[bi, keep1], cluster[bi, keep1], idx[bi, keep1]
    n = int(M * (1 - ddr))
    n_drop = keep1_n - n
    fps = farthest_point_indices(u_c, n_drop)                       # clusters to DROP
    keep2 = []
    for b in range(B):
        dropped = torch.zeros(keep1_n, dtype=torch.bool)
        dropped[fps[b]] = True
        keep2.append((~dropped).nonzero(as_tuple=True)[0][:n])     # ascending complement, truncated (:405-406)
    keep2 = torch.stack(keep2, 0)
    new_c, new_cl, new_idx = u_c[bi, keep2], u_cl[bi, keep2], u_idx[bi, keep2]
    drop_idx = u_idx[bi, fps].reshape(B, -1)
    if trace is not None:
        trace.update(pad_counts=pc, keep1=keep1, fps=fps, keep2=keep2, kept_src=keep1.gather(1, keep2),
                     kept_centres=new_c, kept_idx=new_idx, drop_idx=drop_idx)
    return new_cl, new_c, new_idx, drop_idx


# ----------------------------------------------------------------------------- S6 point proxies
def point_encoder(sd, centre, cluster, prefix="simple_encoder", train=None):
    """:126-142 -> (b,n,256), max over K of ReLU(BN(conv))."""
    y = _conv_bn_relu(sd, f"{prefix}.mlp", _cluster_features(centre, cluster), train)
    return y.permute(0, 2, 3, 1).max(dim=2)[0]


# ----------------------------------------------------------------------------- S7 proxy block
def position_bias(sd, prefix: str, s: int) -> torch.Tensor:
    """:212-215 -> (n, s*s): bilinear(pb 4x4 -> s x s, align_corners=False) + (pc + pr)."""
    pb = F.interpolate(sd[f"{prefix}.pb_bias"], size=(s, s), mode="bilinear")
    n = pb.shape[1]
    return (pb.reshape(1, n, -1) + (sd[f"{prefix}.pc_bias"] + sd[f"{prefix}.pr_bias"]).reshape(1, n, -1))[0]


def proxy_attention(sd, prefix, x, proxy, mask, num_heads):
    """:206-257.  x is LN1(x) on entry; mask True = real token (or None)."""
    b, n, c = x.shape
    l = proxy.shape[1]
    hd = c // num_heads
    scale = hd ** -0.5
    x = x + position_bias(sd, prefix, int(c ** 0.5))[None]
    qkv = F.linear(x, sd[f"{prefix}.qkv.weight"], sd.get(f"{prefix}.qkv.bias")).reshape(b, n, 3, c).permute(2, 0, 1, 3)   # bias iff qkv_bias (:199)
    pt = F.linear(proxy, sd[f"{prefix}.proxy_proj.weight"], sd[f"{prefix}.proxy_proj.bias"])
    q, k, v = (t.reshape(b, n, num_heads, hd).permute(0, 2, 1, 3) for t in (qkv[0], qkv[1], qkv[2]))
    pt = pt.reshape(b, l, num_heads, hd).permute(0, 2, 1, 3)
    a1 = torch.softmax((pt * scale) @ k.transpose(-2, -1), dim=-1)         # (b,h,l,n) — unmasked (:232-236)
    pv = a1 @ v                                                            # (b,h,l,hd)
    s2 = (q * scale) @ pt.transpose(-2, -1)                                # (b,h,n,l)
    if mask is not None:
        s2 = s2.masked_fill((~mask)[:, None, None, :], -1e9)               # :242-247
    o = torch.softmax(s2, dim=-1) @ pv
    o = o.transpose(1, 2).reshape(b, n, c)
    return F.linear(o, sd[f"{prefix}.proj.weight"], sd[f"{prefix}.proj.bias"])


def proxy_block(sd, prefix, x, proxy, mask, num_heads):
    """:273-276 (DropPath/Dropout are identity in eval); timm Mlp = fc1 -> GELU(erf) -> fc2."""
    c = x.shape[-1]
    u = F.layer_norm(x, (c,), sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"], LN_EPS)
    x = x + proxy_attention(sd, f"{prefix}.attn", u, proxy, mask, num_heads)
    h = F.layer_norm(x, (c,), sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"], LN_EPS)
    h = F.linear(F.gelu(F.linear(h, sd[f"{prefix}.mlp.fc1.weight"], sd[f"{prefix}.mlp.fc1.bias"])),
                 sd[f"{prefix}.mlp.fc2.weight"], sd[f"{prefix}.mlp.fc2.bias"])
    return x + h


def branch(sd, stack, norm, n_blocks, pp, proxy, mask, num_heads, faithful_cost=False):
    """:441-443 / :450-452 — every block is fed ``pp``; only the last block's output survives."""
    c = pp.shape[-1]
    out = None
    for i in (range(n_blocks) if faithful_cost else [n_blocks - 1]):
        out = proxy_block(sd, f"{stack}.{i}", pp, proxy, mask, num_heads)
        out = F.layer_norm(out, (c,), sd[f"{norm}.{i}.weight"], sd[f"{norm}.{i}.bias"], LN_EPS)
    return out


def head(sd, lin, bn, g, train=None):
    """:445-446 / :454-455 — Linear then BatchNorm1d over the channel dim."""
    y = F.linear(g, sd[f"{lin}.weight"], sd[f"{lin}.bias"])
    return _batch_norm(sd, bn, y.transpose(-2, -1), train).transpose(-2, -1)


# ----------------------------------------------------------------------------- S9 image proxies
def image_proxies(sd, img_feat: torch.Tensor, num_heads: int, faithful_cost: bool = False) -> torch.Tensor:
    """:335-342 + AttentionPool2d :154-177.  With ``faithful_cost`` all 226 query
    tokens go through the attention as in the reference (F.multi_head_attention_forward)
    and row 0 is kept; otherwise only the token-0 query is evaluated (same value)."""
    B, V, C, H, W = img_feat.shape
    x = F.conv2d(img_feat.reshape(B * V, C, H, W).float(), sd["channel_mapper.weight"], sd["channel_mapper.bias"])
    c = x.shape[1]
    x = x.reshape(B * V, c, H * W).permute(2, 0, 1)                         # (225, BV, c)
    x = torch.cat([x.mean(dim=0, keepdim=True), x], dim=0) + sd["attn_pool2d.positional_embedding"][:, None, :]
    hd = c // num_heads
    xq = x if faithful_cost else x[:1]
    q = F.linear(xq, sd["attn_pool2d.q_proj.weight"], sd["attn_pool2d.q_proj.bias"]) * hd ** -0.5
    k = F.linear(x, sd["attn_pool2d.k_proj.weight"], sd["attn_pool2d.k_proj.bias"])
    v = F.linear(x, sd["attn_pool2d.v_proj.weight"], sd["attn_pool2d.v_proj.bias"])
    T, Tq, N = x.shape[0], xq.shape[0], B * V
    q = q.reshape(Tq, N * num_heads, hd).transpose(0, 1)
    k = k.reshape(T, N * num_heads, hd).transpose(0, 1)
    v = v.reshape(T, N * num_heads, hd).transpose(0, 1)
    a = torch.softmax(q @ k.transpose(1, 2), dim=-1)
    o = (a @ v).transpose(0, 1).reshape(Tq, N, c)
    o = F.linear(o, sd["attn_pool2d.c_proj.weight"], sd["attn_pool2d.c_proj.bias"])[0]
    o = F.layer_norm(o, (c,), sd["norm_img.weight"], sd["norm_img.bias"], LN_EPS)
    return o.reshape(B, V, c)


# ----------------------------------------------------------------------------- S10-S12
def affine(transform, translate, centre, cluster):
    """:459-462 — ((T @ (p - c)) + c) + t for every slot (padded ones too)."""
    b, m = centre.shape[:2]
    T = transform.reshape(b, m, 3, 3)
    tc = centre.unsqueeze(-2)
    return (T @ (cluster - tc).transpose(-2, -1)).transpose(-2, -1) + tc + translate.unsqueeze(-2)


def _scatter_numpy(P: torch.Tensor, idx: torch.Tensor, new: torch.Tensor) -> torch.Tensor:
    out = P.clone().numpy()
    idx_np, new_np = idx.numpy(), new.numpy()
    for b in range(P.shape[0]):
        flat = idx_np[b].reshape(-1)
        valid = flat != -1
        out[b][flat[valid]] = new_np[b].reshape(-1, 3)[valid]
    return torch.from_numpy(out)


class _ScatterTrain(torch.autograd.Function):
    """Training-mode form of the scatter: forward = the pinned last-writer-wins result; backward = what autograd does for
    the reference's ``p2[batch, idx] = cluster`` (:495, index_put_ without accumulate): EVERY valid source slot receives
    the gradient of the destination it was written to — also the slots a later duplicate overwrote — and the overwritten
    destinations of ``p2`` receive none."""

    @staticmethod
    def forward(ctx, P, idx, new):
        ctx.save_for_backward(idx)
        return _scatter_numpy(P.detach(), idx, new.detach())

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        B, M, K = idx.shape
        flat = idx.reshape(B, M * K)
        valid = flat != -1
        src = torch.gather(g, 1, flat.clamp(min=0)[..., None].expand(-1, -1, 3))
        g_new = src.masked_fill(~valid[..., None], 0.0).reshape(B, M, K, 3)
        g_P = g.clone()
        for b in range(B):
            g_P[b, flat[b][valid[b]]] = 0.0
        return g_P, None, g_new


def scatter_last_writer_wins(P: torch.Tensor, idx: torch.Tensor, new: torch.Tensor) -> torch.Tensor:
    """:472-498 with the pinned duplicate rule: numpy fancy assignment applies
    repeated indices in order, so the largest flat (m,k) wins."""
    if new.requires_grad or P.requires_grad:
        return _ScatterTrain.apply(P, idx, new)
    return _scatter_numpy(P, idx, new)


def remove_points(P: torch.Tensor, drop_idx: torch.Tensor) -> List[torch.Tensor]:
    """:501-525 — ascending survivors; -1 in drop_idx never matches a point."""
    out = []
    for b in range(P.shape[0]):
        keep = torch.ones(P.shape[1], dtype=torch.bool)
        d = drop_idx[b]
        keep[d[d >= 0]] = False
        out.append(P[b][keep])
    return out


# ----------------------------------------------------------------------------- driver
@torch.no_grad()
def aggregate_sample(view_points: List[torch.Tensor], extrinsics: torch.Tensor, choices: torch.Tensor) -> torch.Tensor:
    """Input side, restated from the reference's data pipeline: ``AggregateMultiViewPoints.transform``
    (datasets/transforms/multiview.py:224-241: per view ``torch.linalg.solve(global2ego, [p;1]^T)^T``, first three
    columns, views concatenated in order) followed by ``points[choices]`` of ``PointSample._points_random_sampling``
    (datasets/transforms/points.py:411-417; ``choices`` = the np.random.choice the data loader draws)."""
    glob = []
    for v, p in enumerate(view_points):
        point = torch.cat([p[:, :3].float(), p.new_ones(p.shape[0], 1).float()], dim=1)
        global2ego = extrinsics[v].to(point.dtype)
        glob.append(torch.linalg.solve(global2ego, point.transpose(0, 1)).transpose(0, 1)[:, :3])
    return torch.cat(glob)[choices.long()]


def batch_sparse_collate(points: List[torch.Tensor], voxel_size: float, reciprocal: bool = False, floor: bool = False):
    """Caller hand-off (detectors/sparse_featfusion_grounder_preshape.py:388-391):
    ``ME.utils.batch_sparse_collate([(p[:, :3] / voxel_size, p) for p in points])`` -> (coordinates (T,4) int32, features (T,3)).
    MinkowskiEngine (un-vendored, unpinned) is restated from its published ``sparse_collate``: an int32 (T, 1+D) tensor whose
    column 0 is the batch index and whose other columns receive the float coordinates by tensor assignment (= truncation toward
    zero); features are concatenated.  ``reciprocal`` restates torch's CUDA division by a Python scalar (multiply by the fp32
    reciprocal); the default is the CPU kernel's IEEE division, which is what ``p / voxel_size`` evaluates to here."""
    n = sum(len(p) for p in points)
    coords = torch.zeros(n, 4, dtype=torch.int32)
    s = 0
    for b, p in enumerate(points):
        xyz = p[:, :3].float()
        quot = xyz * (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(voxel_size, dtype=torch.float32)) if reciprocal else xyz / voxel_size
        if floor:
            quot = torch.floor(quot)
        coords[s:s + len(p), 1:] = quot          # the assignment MinkowskiEngine performs: float -> int32 truncates
        coords[s:s + len(p), 0] = b
        s += len(p)
    return coords, torch.cat([p.float() for p in points], 0)


def forward(sd: Dict[str, torch.Tensor], points: List[torch.Tensor], text_dict, img_feat: Optional[torch.Tensor], *,
            grid_size: int, dynamic_drop_radio: float, text_blocks: int, img_blocks: int, num_sub: int = 30,
            num_heads: int = 8, img_proxy: Optional[torch.Tensor] = None, faithful_cost: bool = False,
            trace: Optional[dict] = None, train: Optional[dict] = None) -> List[torch.Tensor]:
    """:424-469.  ``img_proxy`` (B,V,256) may be given instead of ``img_feat`` to
    time/check the core region.  ``train`` (a dict): train() mode semantics, see the module docstring."""
    P = torch.stack([p.float() for p in points], 0)                                       # :426-427
    centre, cluster, idx = deform_cluster(sd, P, grid_size, num_sub, trace, train)        # :430
    cluster, centre, idx, drop_idx = cluster_dropout(cluster, centre, idx, dynamic_drop_radio, trace)   # :433
    pp = point_encoder(sd, centre, cluster, train=train)                                  # :437
    text, mask = tuple(text_dict.values())                                                # :332-333, :440
    tg = branch(sd, "textformer", "text_norm", text_blocks, pp, text.float(), mask, num_heads, faithful_cost)
    translate = head(sd, "text_trans", "text_trans_norm", tg, train)                      # :445-446
    if img_proxy is None:
        img_proxy = image_proxies(sd, img_feat, num_heads, faithful_cost)                 # :449
    ig = branch(sd, "imgformer", "img_norm", img_blocks, pp, img_proxy, None, num_heads, faithful_cost)
    transform = head(sd, "img_trans", "img_trans_norm", ig, train)                        # :454-455
    new = affine(transform, translate, centre, cluster)                                   # :459-462
    P2 = scatter_last_writer_wins(P, idx, new)                                            # :465
    out = remove_points(P2, drop_idx)                                                     # :467
    if trace is not None:
        trace.update(point_proxy=pp, text_guide=tg, img_guide=ig, img_proxy=img_proxy, translate=translate,
                     transform=transform, new_clusters=new, scattered=P2)
    return out
