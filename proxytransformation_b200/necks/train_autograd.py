"""train()-mode forward WITH autograd for ProxyTransformationNormReverse.

What a gradient never flows through — both ball queries (:56,:65), the dropout selection (pad-count sort + farthest point sampling,
:352-420) and the duplicate rule of the scatter (:495) — runs on the sm_100a kernels of the eval path.  Everything a gradient does
flow through (offset network :87-107, point encoder :126-142, the ProxyBlocks :206-276 with their Dropout / DropPath layers, the
heads :445-455, the attention pool :154-177,:335-342, the affine map :459-462, the scatter's backward rule) is expressed with torch
ops on the same device, on the module's own parameters, so torch autograd is the backward pass (ATen / cuBLAS kernels: there are no
hand-written backward kernels).  This is the training path only; it is not the measured hot path (eval, `forward_packed`).

Semantics pinned by tests/golden/c1_train.npz (captured from the unmodified reference in train() mode): outputs, the four BatchNorm
layers' running statistics after the step, the norm of every parameter gradient and the full gradients of 24 small parameters.
Two behaviours of the reference's backward that this reproduces: (i) `p2[b, idx] = cluster` (:495) is an index_put_ WITHOUT accumulate,
so every valid source slot receives the gradient of its destination (also slots that a later duplicate overwrote) and an overwritten
destination passes no gradient to the input points; (ii) the centres receive gradients through the relative coordinates of the
encoder (:131) and the affine map (:462).

Stochastic layers: Dropout / DropPath use torch's generator with the reference's call order (attn_drop after each of the two softmaxes,
proj_drop, DropPath on the attention branch, the two Mlp dropouts, DropPath on the Mlp branch; every block of a stack runs, only the
last one's output survives :441-452), so a run is reproducible under torch.manual_seed; layers at rate 0 draw nothing.  1x1 convolutions
are evaluated as fp32 matmuls (cuDNN would be free to use TF32)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F

from .. import ops

MARGIN = 4.0       # :23 DeformablePointCluster(margin=4)


def masked_gather(P: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """pytorch3d masked_gather: P (B,N,3), idx (B,M,K) with -1 padding -> (B,M,K,3), padded slots = 0."""
    B, M, K = idx.shape
    g = P.gather(1, idx.clamp(min=0).long().reshape(B, M * K, 1).expand(-1, -1, 3)).reshape(B, M, K, 3)
    return g.masked_fill((idx < 0)[..., None], 0.0)


def _conv1x1_bn_relu(seq, x: torch.Tensor) -> torch.Tensor:
    """nn.Sequential(Conv2d(6,256,1), BatchNorm2d(256), ReLU) on (B,M,K,6) -> (B,256,M,K); the conv as a matmul, the BatchNorm module
    itself (batch statistics + running-statistics update in train mode)."""
    conv, bn = seq[0], seq[1]
    y = F.linear(x, conv.weight.reshape(conv.out_channels, conv.in_channels), conv.bias)        # (B,M,K,256)
    return F.relu(bn(y.permute(0, 3, 1, 2)))


def _cluster_features(centre: torch.Tensor, cluster: torch.Tensor) -> torch.Tensor:
    """:93-99 / :131-137 -> (B,M,K,6) = [relative (zero where the gathered point is exactly (0,0,0)), absolute]."""
    rel = cluster - centre.unsqueeze(2)
    pad = (cluster == 0).all(dim=-1)
    return torch.cat([rel.masked_fill(pad[..., None], 0.0), cluster], dim=-1)


def _drop_path(x: torch.Tensor, p: float) -> torch.Tensor:
    """Stochastic depth per sample: keep with probability 1 - p, scaled by 1 / (1 - p)."""
    if p <= 0.0:
        return x
    keep = 1.0 - p
    m = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
    if keep > 0.0:
        m.div_(keep)
    return x * m


def _position_bias(attn) -> torch.Tensor:
    """:212-217 -> (1, n, c)."""
    s = attn.pc_bias.shape[2]
    n = attn.pb_bias.shape[1]
    pb = F.interpolate(attn.pb_bias, size=(s, s), mode="bilinear")
    return pb.reshape(1, n, -1) + (attn.pc_bias + attn.pr_bias).reshape(1, n, -1)


def _proxy_attention(attn, x, proxy, mask, heads: int, attn_drop: float, proj_drop: float) -> torch.Tensor:
    """:206-257; x = norm1(x) on entry, mask (B,l) True = real token or None."""
    b, n, c = x.shape
    l = proxy.shape[1]
    hd = c // heads
    scale = hd ** -0.5
    x = x + _position_bias(attn)
    qkv = attn.qkv(x).reshape(b, n, 3, c).permute(2, 0, 1, 3)
    pt = attn.proxy_proj(proxy).reshape(b, l, heads, hd).permute(0, 2, 1, 3)
    q, k, v = (t.reshape(b, n, heads, hd).permute(0, 2, 1, 3) for t in (qkv[0], qkv[1], qkv[2]))
    a1 = F.dropout(torch.softmax((pt * scale) @ k.transpose(-2, -1), dim=-1), attn_drop, True)      # proxy as query, unmasked
    pv = a1 @ v
    s2 = (q * scale) @ pt.transpose(-2, -1)                                                          # proxy as key
    if mask is not None:
        s2 = s2.masked_fill((~mask.bool())[:, None, None, :], -1e9)
    o = F.dropout(torch.softmax(s2, dim=-1), attn_drop, True) @ pv
    o = o.transpose(1, 2).reshape(b, n, c)
    return F.dropout(attn.proj(o), proj_drop, True)


def _proxy_block(blk, x, proxy, mask, heads: int, drop: float, attn_drop: float, drop_path: float) -> torch.Tensor:
    """:273-276; timm Mlp = fc1 -> GELU -> drop -> fc2 -> drop."""
    x = x + _drop_path(_proxy_attention(blk.attn, blk.norm1(x), proxy, mask, heads, attn_drop, drop), drop_path)
    h = F.dropout(F.gelu(blk.mlp.fc1(blk.norm2(x))), drop, True)
    h = F.dropout(blk.mlp.fc2(h), drop, True)
    return x + _drop_path(h, drop_path)


def _branch(m, blocks, norms, pp, proxy, mask, drop_path_rate: float) -> torch.Tensor:
    """:441-443 / :450-452: every block is fed the point proxies, the last block's (normalised) output is what is used.  With every
    rate at 0 the earlier blocks (no effect on the result, no gradient, no random numbers) are skipped."""
    n = len(blocks)
    dpr = [float(v) for v in torch.linspace(0, drop_path_rate, n)]                                 # :298-299
    stochastic = m.drop_rate > 0 or m.attn_drop_rate > 0 or drop_path_rate > 0
    out = None
    for i in (range(n) if stochastic else [n - 1]):
        out = norms[i](_proxy_block(blocks[i], pp, proxy, mask, m.num_heads, m.drop_rate, m.attn_drop_rate, dpr[i]))
    return out


def _head(lin, bn, g: torch.Tensor) -> torch.Tensor:
    """:445-446 / :454-455: Linear, then BatchNorm1d over the channel dim."""
    return bn(lin(g).transpose(-2, -1)).transpose(-2, -1)


def image_proxies(m, img_feat: torch.Tensor) -> torch.Tensor:
    """:335-342 + AttentionPool2d :154-177 in its single-query form (only token 0 of the attention output is used, :177)."""
    B, V, C, H, W = img_feat.shape
    cm, ap = m.channel_mapper, m.attn_pool2d
    x = F.linear(img_feat.reshape(B * V, C, H * W).float().transpose(1, 2), cm.weight.reshape(cm.out_channels, C), cm.bias)   # (BV,HW,c)
    c = x.shape[-1]
    x = torch.cat([x.mean(dim=1, keepdim=True), x], dim=1) + ap.positional_embedding[None]        # (BV, HW+1, c)
    heads, hd = m.num_heads, c // m.num_heads
    q = (ap.q_proj(x[:, :1]) * hd ** -0.5).reshape(B * V, 1, heads, hd).transpose(1, 2)
    k = ap.k_proj(x).reshape(B * V, -1, heads, hd).transpose(1, 2)
    v = ap.v_proj(x).reshape(B * V, -1, heads, hd).transpose(1, 2)
    o = (torch.softmax(q @ k.transpose(-2, -1), dim=-1) @ v).transpose(1, 2).reshape(B * V, c)
    return m.norm_img(ap.c_proj(o)).reshape(B, V, c)


class _ScatterLastWriterWins(torch.autograd.Function):
    """p2[b, idx] = new (:472-498).  Forward: the largest flat (cluster, slot) position wins a duplicated destination (the pinned
    behaviour of the reference's advanced-index assignment).  Backward: the autograd rule of index_put_ without accumulate — every
    valid source slot receives the gradient of its destination, written destinations pass nothing to the input points."""

    @staticmethod
    def forward(ctx, P, idx, new):
        B, N, _ = P.shape
        flat = idx.reshape(B, -1).long()
        S = flat.shape[1]
        valid = flat >= 0
        dev = P.device
        dest = flat + torch.arange(B, device=dev)[:, None] * N
        slot = torch.arange(S, device=dev).expand(B, S)
        win = torch.full((B * N,), -1, dtype=torch.long, device=dev)
        win.scatter_reduce_(0, dest[valid], slot[valid], reduce="amax", include_self=True)
        written = win >= 0
        rows = torch.arange(B * N, device=dev)[written]
        out = P.detach().clone().reshape(B * N, 3)
        out[rows] = new.detach().reshape(B * S, 3)[win[written] + (rows // N) * S]
        ctx.save_for_backward(flat, valid, rows)
        ctx.new_shape = new.shape
        return out.reshape(B, N, 3)

    @staticmethod
    def backward(ctx, g):
        flat, valid, rows = ctx.saved_tensors
        B, N, _ = g.shape
        src = torch.gather(g, 1, flat.clamp(min=0)[..., None].expand(-1, -1, 3))
        g_new = src.masked_fill(~valid[..., None], 0.0).reshape(ctx.new_shape)
        g_P = g.clone().reshape(B * N, 3)
        g_P[rows] = 0.0
        return g_P.reshape(B, N, 3), None, g_new


def forward_train(m, points: Sequence[torch.Tensor], text_dict, img_feat: torch.Tensor) -> List[torch.Tensor]:
    """The reference's forward (:424-469) in train() mode, differentiable.  points: B (N,3) tensors, text_dict.values() =
    (text_feats (B,L,c), mask (B,L) bool), img_feat (B,V,C,H,W); returns the B transformed, thinned clouds."""
    dev = next(m.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("ProxyTransformationNormReverse (B200) has no CPU path: move the module to a CUDA device")
    P = torch.stack([p.to(dev, torch.float32) for p in points], 0).contiguous()
    text, mask = tuple(text_dict.values())
    text, mask = text.to(dev, torch.float32), (mask.to(dev) if mask is not None else None)
    img_feat = img_feat.to(dev)
    K, n = m.num_sub, m.real_cluster_num
    with torch.cuda.device(dev):
        Pd = P.detach()
        # S1-S4 (:53-67): the grid prior and both ball queries are index work; the offsets carry gradients
        mn, mx, c0 = ops.minmax_centres(Pd, m.grid_size, None)
        idx1, _ = ops.ball_query(c0, Pd, K)
        on = m.get_deformable_cluster.get_offsets
        y = _conv1x1_bn_relu(on.mlp, _cluster_features(c0, masked_gather(P, idx1)))                # (B,256,M,K)
        raw = F.linear(y.mean(dim=-1).transpose(1, 2), on.channel_mapper.weight.reshape(3, -1))    # (B,M,3), padded slots included (:102)
        centres = torch.max(torch.min(c0 + raw.tanh() * MARGIN, mx[:, None, :]), mn[:, None, :])   # :59-62
        idx2, _ = ops.ball_query(centres.detach().contiguous(), Pd, K)
        # S5 (:352-420): which clusters survive is index work; their centres are gathered differentiably
        kept_src, _, kidx, drop_idx, _ = ops.cluster_dropout(centres.detach().contiguous(), idx2, m.keep1, n)
        kc = centres.gather(1, kept_src.long()[..., None].expand(-1, -1, 3))
        cluster = masked_gather(P, kidx)
        # S6 (:126-142)
        pp = _conv1x1_bn_relu(m.simple_encoder.mlp, _cluster_features(kc, cluster)).permute(0, 2, 3, 1).max(dim=2)[0]
        # S7 / S8 (:440-455)
        tg = _branch(m, m.textformer, m.text_norm, pp, text, mask, m.drop_path_rate)
        translate = _head(m.text_trans, m.text_trans_norm, tg)
        ig = _branch(m, m.imgformer, m.img_norm, pp, image_proxies(m, img_feat), None, m.drop_path_rate)
        transform = _head(m.img_trans, m.img_trans_norm, ig)
        # S10-S12 (:459-467)
        B = P.shape[0]
        tc = kc.unsqueeze(-2)
        new = (transform.reshape(B, n, 3, 3) @ (cluster - tc).transpose(-2, -1)).transpose(-2, -1) + tc + translate.unsqueeze(-2)
        P2 = _ScatterLastWriterWins.apply(P, kidx, new)
        keep = torch.ones(B, P.shape[1], dtype=torch.bool, device=dev)
        d = drop_idx.long()
        keep.view(-1)[(d + torch.arange(B, device=dev)[:, None] * P.shape[1])[d >= 0]] = False
        return [P2[b][keep[b]] for b in range(B)]
