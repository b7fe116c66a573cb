"""Drop-in for the reference's ``ProxyTransformationNormReverse``
(embodiedscan/models/necks/preshape_norm_reverse_drop.py:280-469).

Same registry name, constructor signature, ``forward(points, text_dict, img_feat)`` contract and ``state_dict`` keys
(released checkpoints load with ``strict=True``); the forward pass itself runs entirely in the sm_100a kernels behind
``include/pt_preshape.h`` — the sub-modules below are parameter containers that mirror the reference's key layout and
are never called on that path.  Eval / no-grad is the measured mode.  train() mode: under no_grad with every drop rate at 0 a
batch-statistics forward on the same kernels (BatchNorm batch statistics + running-statistics update); with autograd
(or Dropout / DropPath) the differentiable path of necks/train_autograd.py — index work on the kernels, the arithmetic as
torch ops on these sub-modules' parameters, torch autograd as the backward pass.

Algorithmic differences from the reference, all output-preserving:
  * only the LAST block of ``textformer`` / ``imgformer`` is evaluated — every block is fed ``point_proxy`` and only the
    last iteration's result is used (:441-443, :450-452);
  * ``get_img_proxy`` is evaluated in single-query form (only token 0 of the attention pool survives, :177);
  * ``argsort`` ties (:378) and duplicate scatter destinations (:495) follow the pinned rules of SURVEY.md §8c
    (stable order; largest flat (m,k) wins) where the reference's own result depends on thread scheduling.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from .. import ops
from ..registry import MODELS

_BN_EPS = 1e-5


# --------------------------------------------------------------------------- parameter containers (state_dict layout)
class _ConvBnRelu(nn.Sequential):
    def __init__(self, cin: int, cout: int):
        super().__init__(nn.Conv2d(cin, cout, 1), nn.BatchNorm2d(cout), nn.ReLU())


class _OffsetNetworkParams(nn.Module):                       # :69-77
    def __init__(self, in_features: int = 6, hidden_features: int = 256):
        super().__init__()
        self.mlp = _ConvBnRelu(in_features, hidden_features)
        self.channel_mapper = nn.Conv1d(hidden_features, 3, kernel_size=1, bias=False)


class _DeformClusterParams(nn.Module):                       # :22-31 (radius=3, margin=4, hidden=256 are its defaults)
    def __init__(self):
        super().__init__()
        self.get_offsets = _OffsetNetworkParams(6, 256)


class _PointNetParams(nn.Module):                            # :109-116
    def __init__(self):
        super().__init__()
        self.mlp = _ConvBnRelu(6, 256)


class _AttnPoolParams(nn.Module):                            # :144-152
    def __init__(self, spacial_dim: int, embed_dim: int):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.randn(spacial_dim ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, embed_dim)


class _ProxyAttentionParams(nn.Module):                      # :179-204
    def __init__(self, dim: int, n_tokens: int, qkv_bias: bool):
        super().__init__()
        s = int(dim ** 0.5)
        self.pb_bias = nn.Parameter(torch.zeros(1, n_tokens, 4, 4))
        self.pc_bias = nn.Parameter(torch.zeros(1, n_tokens, s, 1))
        self.pr_bias = nn.Parameter(torch.zeros(1, n_tokens, 1, s))
        for p in (self.pb_bias, self.pc_bias, self.pr_bias):
            nn.init.trunc_normal_(p, std=0.02)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proxy_proj = nn.Linear(dim, dim)
        self.proj = nn.Linear(dim, dim)


class _MlpParams(nn.Module):                                 # timm.models.layers.Mlp key layout (fc1 / fc2)
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _ProxyBlockParams(nn.Module):                          # :259-271
    def __init__(self, dim: int, n_tokens: int, mlp_ratio: float, qkv_bias: bool, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _ProxyAttentionParams(dim, n_tokens, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _MlpParams(dim, int(dim * mlp_ratio))


def _bn_affine(bn: nn.modules.batchnorm._BatchNorm):
    """Eval-mode BatchNorm as y = x*scale + shift, computed like ATen's CPU kernel (invstd*weight; bias - mean*alpha)."""
    invstd = 1.0 / torch.sqrt(bn.running_var.float() + bn.eps)
    scale = invstd * bn.weight.float()
    shift = bn.bias.float() - bn.running_mean.float() * scale
    return scale.contiguous(), shift.contiguous()


@MODELS.register_module()
class ProxyTransformationNormReverse(nn.Module):
    """See module docstring.  Constructor mirrors :282-285 (including the reference's spelling of ``*_radio``)."""

    def __init__(self, embed_dim=256, num_heads=8, n_points=100000, grid_size=4, text_blocks=1, img_blocks=1,
                 dynamic_drop_radio=0.8, mlp_radio=4, qkv_bias=False, drop_rate=0.2, attn_drop_rate=0.2,
                 drop_path_rate=0.2, act_layer=nn.GELU, norm_layer=nn.LayerNorm, num_sub=30, drop_radio=0.2,
                 input_dim=512, img_spacial_dim=15):
        super().__init__()
        if act_layer is not nn.GELU or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("the CUDA path implements the shipped configuration: GELU(erf) + LayerNorm")
        if embed_dim != 256:
            # the reference hard-wires 256 in SimplifiedPointNet()/OffsetNetwork (:31,:110,:302): any other embed_dim
            # fails at norm1 there as well
            raise ValueError("embed_dim must be 256 (reference :302 hard-wires the point encoder width)")
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.grid_size = grid_size
        self.num_cluster = grid_size ** 3
        self.num_sub = num_sub or n_points // self.num_cluster      # :291
        self.input_dim = input_dim
        self.img_spacial_dim = img_spacial_dim
        self.drop_radio = drop_radio
        self.text_blocks = text_blocks
        self.img_blocks = img_blocks
        self.dynamic_drop_radio = dynamic_drop_radio
        self.mlp_radio = mlp_radio
        # rates only matter in training mode, which this implementation does not run
        self.drop_rate, self.attn_drop_rate, self.drop_path_rate = drop_rate, attn_drop_rate, drop_path_rate
        self.real_cluster_num = int(self.num_cluster * (1 - dynamic_drop_radio))      # :195, :389
        self.keep1 = self.num_cluster - int(self.num_cluster * 0.3)                    # :374-376, empty_drop=0.3
        if text_blocks < 1 or img_blocks < 1:
            raise ValueError("text_blocks and img_blocks must be >= 1 (the reference's forward needs one iteration)")

        self.get_deformable_cluster = _DeformClusterParams()
        self.simple_encoder = _PointNetParams()
        self.channel_mapper = nn.Conv2d(input_dim, embed_dim, kernel_size=1)
        self.attn_pool2d = _AttnPoolParams(img_spacial_dim, embed_dim)
        self.norm_img = nn.LayerNorm(embed_dim)
        n = self.real_cluster_num
        self.textformer = nn.ModuleList([_ProxyBlockParams(embed_dim, n, mlp_radio, qkv_bias, norm_layer) for _ in range(text_blocks)])
        self.text_norm = nn.ModuleList([norm_layer(embed_dim) for _ in range(text_blocks)])
        self.imgformer = nn.ModuleList([_ProxyBlockParams(embed_dim, n, mlp_radio, qkv_bias, norm_layer) for _ in range(img_blocks)])
        self.img_norm = nn.ModuleList([norm_layer(embed_dim) for _ in range(img_blocks)])
        self.text_trans = nn.Linear(embed_dim, 3)
        self.img_trans = nn.Linear(embed_dim, 9)
        self.text_trans_norm = nn.BatchNorm1d(3)
        self.img_trans_norm = nn.BatchNorm1d(9)

        self._packed: Optional[dict] = None
        self._packed_key = None
        self._ws: Dict[str, torch.Tensor] = {}
        self.use_tensor_cores = True     # 3xBF16 tcgen05 GEMMs for the dense layers (falls back per shape inside the C side)
        # first half of the image stage (feature means + query-side projections) on a side stream, concurrent with the
        # geometric stages.  (The whole image stage on a side stream measured slower, 3.04 vs 2.86 ms/step: the persistent
        # pool kernel owns every SM's shared memory and the small kernels queue behind it.)
        self.overlap_mean_pass = os.environ.get("PT_OVERLAP_MEAN", "0") != "0"    # measured: 2.735 vs 2.765 ms/step at best, off by default
        # the WHOLE image stage on a second stream next to the geometric stages + text branch (measured: 2.475 -> 2.410 ms per
        # 64-scene step; shipped config at batch 4 as a CUDA graph 0.931 -> 0.888 ms, eager 1.018 -> 1.044 ms: the extra stream
        # hand-offs cost host time that only matters when the forward is launch bound).  "auto": batches of at least
        # `overlap_img_min_batch` scenes, or while a CUDA graph is being captured
        self.overlap_img_stage = os.environ.get("PT_OVERLAP_IMG", "auto")
        self.overlap_img_min_batch = 2            # (with the image branch on the side stream as well, eager batches of 2-4 scenes gain ~9 %)
        self.parallel_branch_max_rows = int(os.environ.get("PT_PARALLEL_BRANCH_ROWS", str(1 << 30)))   # B * n up to which the image branch runs beside the text branch (measured: every size gains)
        self._streams: Dict[str, torch.cuda.Stream] = {}
        self.host_chunk_scenes = 8       # scenes per pipeline chunk when forward() is fed host tensors
        self.cuda_graphs = os.environ.get("PT_CUDA_GRAPHS", "0") != "0"     # replay small device-resident batches as one CUDA graph
        self.cuda_graph_max_bytes = 256 << 20
        self._graphs: Dict[tuple, tuple] = {}

    # ------------------------------------------------------------------ reference helper API (same names, :332-350)
    def get_text_proxy(self, text_dict):
        return text_dict.values()

    def get_img_proxy(self, img_feat: torch.Tensor) -> torch.Tensor:
        with torch.cuda.device(img_feat.device):
            w = self._weights(img_feat.device)
            return ops.img_attnpool(self._img_feat_dtype(img_feat).contiguous(), w["img"], self.num_heads, params=w["img_struct"])

    def get_point_proxy(self, center, cluster_idx, points):
        with torch.cuda.device(points.device):
            w = self._weights(points.device)
            return ops.point_encoder(points, cluster_idx, center, w["encoder"])

    @staticmethod
    def _img_feat_dtype(img_feat: torch.Tensor) -> torch.Tensor:
        """fp32, bf16 and fp16 feature maps are consumed as they are; anything else is converted to fp32."""
        return img_feat if img_feat.dtype in ops.IMG_FEAT_DTYPES else img_feat.float()

    # ------------------------------------------------------------------ weight packing
    def _weight_key(self, device):
        return (str(device), self.use_tensor_cores, ops.img_pool_variant()) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def refresh_weights(self):
        """Force re-packing (folded BN, position-bias tables, folded image-pool projections) on the next forward."""
        self._packed = None
        self._graphs.clear()          # captured graphs hold pointers into the packed weights

    def _weights(self, device) -> dict:
        key = self._weight_key(device)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        self._graphs.clear()          # captured graphs hold pointers into the packed weights that are replaced below
        with torch.cuda.device(device):
            return self._pack_weights(device, key)

    def _pack_weights(self, device, key) -> dict:
        for name, mod in self.named_modules():
            if isinstance(mod, nn.LayerNorm) and abs(mod.eps - 1e-5) > 1e-12:
                raise NotImplementedError(f"{name}: LayerNorm eps {mod.eps} (the kernels implement the reference's default 1e-5)")
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        with torch.no_grad():
            def conv_bn(seq):
                sc, sh = _bn_affine(seq[1])
                return dict(conv_w=f32(seq[0].weight.reshape(seq[0].weight.shape[0], -1)), conv_b=f32(seq[0].bias),
                            bn_scale=f32(sc), bn_shift=f32(sh))

            off = conv_bn(self.get_deformable_cluster.get_offsets.mlp)
            off["map_w"] = f32(self.get_deformable_cluster.get_offsets.channel_mapper.weight.reshape(3, -1))
            enc = conv_bn(self.simple_encoder.mlp)

            def block(blk: _ProxyBlockParams, out_norm: nn.LayerNorm):
                a = blk.attn
                d = dict(ln1_w=f32(blk.norm1.weight), ln1_b=f32(blk.norm1.bias),
                         pos_bias=ops.position_bias(f32(a.pb_bias), f32(a.pc_bias), f32(a.pr_bias)),
                         qkv_w=f32(a.qkv.weight), pp_w=f32(a.proxy_proj.weight), pp_b=f32(a.proxy_proj.bias),
                         proj_w=f32(a.proj.weight), proj_b=f32(a.proj.bias), ln2_w=f32(blk.norm2.weight),
                         ln2_b=f32(blk.norm2.bias), fc1_w=f32(blk.mlp.fc1.weight), fc1_b=f32(blk.mlp.fc1.bias),
                         fc2_w=f32(blk.mlp.fc2.weight), fc2_b=f32(blk.mlp.fc2.bias), lno_w=f32(out_norm.weight),
                         lno_b=f32(out_norm.bias))
                if a.qkv.bias is not None:                               # qkv_bias=True (:199; no shipped config sets it)
                    d["qkv_b"] = f32(a.qkv.bias)
                if self.use_tensor_cores:
                    for k in ("qkv_w", "proj_w", "fc1_w", "fc2_w", "pp_w"):
                        d[k + "_split"] = ops.split_bf16(d[k])
                return d

            tb = block(self.textformer[-1], self.text_norm[-1])      # only the last block of each stack is live
            ib = block(self.imgformer[-1], self.img_norm[-1])

            def head(lin: nn.Linear, bn: nn.BatchNorm1d):
                sc, sh = _bn_affine(bn)
                return dict(lin_w=f32(lin.weight), lin_b=f32(lin.bias), bn_scale=f32(sc), bn_shift=f32(sh))

            img = self._fold_img_pool(device)
        self._packed = dict(offset=off, encoder=enc, text=tb, text_struct=ops.make_block_params(tb), imgb=ib,
                            imgb_struct=ops.make_block_params(ib), text_head=head(self.text_trans, self.text_trans_norm),
                            img_head=head(self.img_trans, self.img_trans_norm), img=img, img_struct=ops.make_img_params(img),
                            lin=ops.linspace01(self.grid_size, device))
        self._packed_key = key
        return self._packed

    def _fold_img_pool(self, device) -> dict:
        """Single-query folding of channel_mapper + AttentionPool2d (:154-177, :338-340); fp64 on the device, once per
        weight load.  See csrc/imgpool.cu for the algebra."""
        d64 = lambda t: t.detach().to(device=device, dtype=torch.float64)
        c, C, heads = self.embed_dim, self.input_dim, self.num_heads
        hd = c // heads
        ap = self.attn_pool2d
        Wc, bc = d64(self.channel_mapper.weight).reshape(c, C), d64(self.channel_mapper.bias)
        pos = d64(ap.positional_embedding)
        Wq, bq, Wk, Wv, bv = d64(ap.q_proj.weight), d64(ap.q_proj.bias), d64(ap.k_proj.weight), d64(ap.v_proj.weight), d64(ap.v_proj.bias)
        T = pos.shape[0]
        Tp = (T + 3) // 4 * 4
        posb = pos + bc
        g_k = torch.zeros(Tp, c, dtype=torch.float64, device=device)
        g_k[:T] = posb @ Wk.T                                              # k-bias dropped: constant over tokens
        h_v = torch.zeros(heads, hd, Tp, dtype=torch.float64, device=device)
        h_v[:, :, :T] = (posb @ Wv.T + bv).T.reshape(heads, hd, T)
        w_kc = (Wk @ Wc).reshape(heads, hd, C).transpose(1, 2)             # (heads, C, hd): NT operand per head
        f = lambda t: t.to(torch.float32).contiguous()
        out = dict(w_qc=f(Wq @ Wc), q0=f(Wq @ posb[0] + bq), w_kc=f(w_kc), g_k=f(g_k), w_vc=f(Wv @ Wc), h_v=f(h_v),
                   cproj_w=f(d64(ap.c_proj.weight)), cproj_b=f(d64(ap.c_proj.bias)), ln_w=f(d64(self.norm_img.weight)),
                   ln_b=f(d64(self.norm_img.bias)))
        if self.use_tensor_cores and (C, self.img_spacial_dim, c, heads) == (512, 15, 256, 8):
            # operands of the bf16 tensor-core fast path (csrc/imgpool_tc.cu), layouts in include/pt_preshape.h
            TPc = 228
            # the pool kernel walks the channels residue class by residue class (channel mod 8, csrc/imgpool_tc.cu): w_eff
            # columns and weighted-sum columns come in the orders below, absorbed here into the GEMM weights
            variant = ops.img_pool_variant()
            score_order, sum_order = ops.img_pool_channel_orders(device, variant)
            wk_pad = torch.zeros(heads * C, 64, dtype=torch.float64, device=device)
            wk_pad[:, :hd] = w_kc[:, score_order, :].reshape(heads * C, hd)
            gk_pad = torch.zeros(heads, TPc, 64, dtype=torch.float64, device=device)
            gk_pad[:, :T, :hd] = (posb @ Wk.T).reshape(T, heads, hd).transpose(0, 1)
            wv_cat = torch.zeros(c, 768, dtype=torch.float64, device=device)
            wv_cat[:, :C] = (Wv @ Wc)[:, sum_order]
            wv_cat[:, C:C + T] = (posb @ Wv.T + bv).T
            out.update(w_qc_split=ops.split_bf16(out["w_qc"]), wk_pad_split=ops.split_bf16(f(wk_pad)),
                       gk_pad_split=ops.split_bf16(f(gk_pad.reshape(heads * TPc, 64))), wv_cat_split=ops.split_bf16(f(wv_cat)),
                       cproj_split=ops.split_bf16(out["cproj_w"]), variant=variant)
        return out

    # ------------------------------------------------------------------ forward (:424-469)
    def forward(self, points: Sequence[torch.Tensor], text_dict, img_feat: torch.Tensor, *, img_proxy: Optional[torch.Tensor] = None,
                trace: Optional[dict] = None) -> List[torch.Tensor]:
        """points: list of B (N,3) fp32 tensors (equal N); text_dict: values() = (text_feats (B,L,c), mask (B,L) bool);
        img_feat: (B,V,input_dim,H,W) fp32 or bf16.  Returns the list of (N'_b,3) tensors on the inputs' device.
        Host tensors are accepted (pinned memory makes the copies asynchronous); results then come back on the host.
        ``img_proxy`` (B,V,c) may replace ``img_feat`` (core region of the benchmark); ``trace`` collects intermediates.

        train() mode: under torch.no_grad() with every drop rate at 0 the forward runs on the sm_100a kernels
        with BatchNorm batch statistics (the four layers update their running statistics as nn.BatchNorm does).  With autograd enabled,
        or with Dropout / DropPath rates above 0, it runs as necks/train_autograd.py: index work on the same kernels, every
        differentiable operation as torch ops on the module's parameters, so ``loss.backward()`` fills their ``.grad`` like the
        reference's (pinned against the reference's gradients by tests/golden/c1_train.npz)."""
        self._check_mode()
        if self.training and (torch.is_grad_enabled() or self.drop_rate or self.attn_drop_rate or self.drop_path_rate):
            if img_proxy is not None or trace is not None:
                raise NotImplementedError("ProxyTransformationNormReverse (B200): img_proxy= / trace= are eval-mode arguments")
            from . import train_autograd
            in_dev = points[0].device
            out = train_autograd.forward_train(self, points, text_dict, img_feat)
            self._packed_key = None                  # parameters / running statistics are about to change: re-fold for the next eval()
            return out if in_dev.type == "cuda" else [o.to(in_dev) for o in out]
        with torch.no_grad():
            return self._forward_impl(points, text_dict, img_feat, img_proxy, trace)

    def _check_mode(self):
        """eval() under autograd: one warning (the eval path carries no gradients)."""
        if not self.training and torch.is_grad_enabled() and not getattr(self, "_warned_no_grad", False) and \
                any(p.requires_grad for p in self.parameters()):
            import warnings
            self._warned_no_grad = True
            warnings.warn("ProxyTransformationNormReverse (B200): eval() mode runs under torch.no_grad() and its outputs carry no "
                          "gradient into the parameters or the image / text backbones; use train() mode for a differentiable "
                          "forward, or wrap the call in torch.no_grad() to silence this warning", stacklevel=3)

    def _forward_impl(self, points, text_dict, img_feat, img_proxy, trace) -> List[torch.Tensor]:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("ProxyTransformationNormReverse (B200) has no CPU path: move the module to a CUDA device")
        in_dev = points[0].device
        if in_dev.type != "cuda":
            return self._forward_from_host(points, text_dict, img_feat, img_proxy, dev, trace)
        P = self._stack_points(points, dev)                                                     # :426-427
        text, mask = tuple(self.get_text_proxy(text_dict))                                      # :440
        text = text.to(dev, torch.float32, non_blocking=True).contiguous()
        mask = mask.to(dev, non_blocking=True).to(torch.uint8).contiguous() if mask is not None else None
        if img_proxy is None:
            img_feat = img_feat.to(dev, non_blocking=True)
            img_feat = self._img_feat_dtype(img_feat).contiguous()
        else:
            img_proxy = img_proxy.to(dev, torch.float32, non_blocking=True).contiguous()
        if self._use_graph(P, img_feat, img_proxy, trace):
            out, counts = self._forward_graphed(P, text, mask, img_feat)
        else:
            out, counts = self.forward_packed(P, text, mask, img_feat, img_proxy=img_proxy, trace=trace)
        cnt = counts.cpu().tolist()                                                             # the one D2H sync
        return [out[b, :cnt[b]] for b in range(len(cnt))]

    # ------------------------------------------------------------------ CUDA graphs for small batches
    def _use_graph(self, P, img_feat, img_proxy, trace) -> bool:
        """Small batches are launch-latency bound (about 40 dependent kernels of a few microseconds each): replaying the whole
        forward as one CUDA graph removes the gaps.  The inputs have to be copied into the graph's static buffers, so the
        mode is limited to batches whose inputs are small (``cuda_graph_max_bytes``); off unless ``cuda_graphs`` is set."""
        if not self.cuda_graphs or trace is not None or img_proxy is not None or torch.cuda.is_current_stream_capturing() or self.training:
            return False
        return P.numel() * 4 + img_feat.numel() * img_feat.element_size() <= self.cuda_graph_max_bytes

    def _forward_graphed(self, P, text, mask, img_feat):
        key = (tuple(P.shape), tuple(text.shape), mask is not None, tuple(img_feat.shape), img_feat.dtype, self._weight_key(P.device))
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= 8:                       # shapes normally repeat; do not hoard graph memory pools
                self._graphs.clear()
            static = [torch.empty_like(P), torch.empty_like(text), torch.empty_like(mask) if mask is not None else None,
                      torch.empty_like(img_feat)]
            for dst, src in zip(static, (P, text, mask, img_feat)):
                if dst is not None:
                    dst.copy_(src)
            cur = torch.cuda.current_stream(P.device)
            side = self._side_stream(P.device, "graph-warmup")
            side.wait_stream(cur)
            with torch.cuda.stream(side):                    # warm-up outside the capture: weight packing, function attributes
                for _ in range(2):
                    self.forward_packed(*static)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out, counts = self.forward_packed(*static)
            entry = self._graphs[key] = (graph, static, out, counts)
        graph, static, out, counts = entry
        for dst, src in zip(static, (P, text, mask, img_feat)):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        graph.replay()
        return out.clone(), counts.clone()                   # the graph owns its outputs: hand out copies

    def _forward_from_host(self, points, text_dict, img_feat, img_proxy, dev, trace) -> List[torch.Tensor]:
        """Host inputs -> host results.  The batch is cut into chunks of ``host_chunk_scenes`` scenes that flow through three
        streams: H2D copies of chunk j+1 (pinned inputs make them asynchronous) overlap the kernels of chunk j and the D2H
        copy of chunk j-1, so a call costs about the PCIe time of its inputs instead of copy + compute + copy.  One host
        synchronisation at the end; the results are views of one pinned host block."""
        B = len(points)
        N = points[0].shape[0]
        for p in points:
            if p.dim() != 2 or p.shape[1] != 3:
                raise ValueError(f"points must be (N,3) xyz per scene, got {tuple(p.shape)}")
            if p.shape[0] != N:
                raise RuntimeError(f"all scenes must have the same number of points (got {p.shape[0]} vs {N})")
        text, mask = tuple(self.get_text_proxy(text_dict))                                      # :440
        if img_proxy is None:
            img_feat = self._img_feat_dtype(img_feat)
        cur = torch.cuda.current_stream(dev)
        h2d, d2h = self._side_stream(dev, "h2d"), self._side_stream(dev, "d2h")
        h2d.wait_stream(cur)
        d2h.wait_stream(cur)
        host_out = torch.empty(B, N, 3, dtype=torch.float32, pin_memory=True)
        host_cnt = torch.empty(B, dtype=torch.int32, pin_memory=True)
        chunk = max(1, min(B, int(self.host_chunk_scenes)))
        if self.training:
            chunk = B                                 # batch statistics are over the whole batch
        traces = [] if trace is not None else None
        for s0 in range(0, B, chunk):
            s1 = min(B, s0 + chunk)
            with torch.cuda.stream(h2d):
                P = torch.empty(s1 - s0, N, 3, dtype=torch.float32, device=dev)
                for b in range(s0, s1):
                    P[b - s0].copy_(points[b], non_blocking=True)                               # :426-427
                tx = text[s0:s1].to(dev, non_blocking=True).float().contiguous()
                mk = mask[s0:s1].to(dev, non_blocking=True).to(torch.uint8).contiguous() if mask is not None else None
                if img_proxy is None:
                    im, ip = img_feat[s0:s1].to(dev, non_blocking=True).contiguous(), None
                else:
                    im, ip = None, img_proxy[s0:s1].to(dev, non_blocking=True).float().contiguous()
                ready = torch.cuda.Event()
                ready.record(h2d)
            cur.wait_event(ready)
            for t in (P, tx, mk, im, ip):
                if t is not None:
                    t.record_stream(cur)
            tr = {} if traces is not None else None
            out, counts = self.forward_packed(P, tx, mk, im, img_proxy=ip, trace=tr)
            if traces is not None:
                traces.append(tr)
            done = torch.cuda.Event()
            done.record(cur)
            d2h.wait_event(done)
            out.record_stream(d2h)
            counts.record_stream(d2h)
            with torch.cuda.stream(d2h):
                host_out[s0:s1].copy_(out, non_blocking=True)
                host_cnt[s0:s1].copy_(counts, non_blocking=True)
        d2h.synchronize()                                                                       # the one host sync
        if trace is not None:
            for k in traces[0]:
                trace[k] = torch.cat([t[k] for t in traces], 0)
        cnt = host_cnt.tolist()
        return [host_out[b, :cnt[b]] for b in range(B)]

    @torch.no_grad()
    def forward_sparse(self, points, text_dict, img_feat, voxel_size: float, *, reciprocal: bool = True, floor: bool = False):
        """forward() followed by the caller's hand-off to the sparse backbone
        (detectors/sparse_featfusion_grounder_preshape.py:385-391, ``use_xyz_feat=True``):
        ``ME.utils.batch_sparse_collate([(p[:, :3] / voxel_size, p) for p in self(points, ...)])`` without the per-scene Python
        list in between.  Device inputs; returns (coordinates (T,4) int32 [scene,x,y,z], features (T,3) fp32) on the device.
        ``reciprocal=True`` reproduces the quotient of torch's CUDA kernel (the reference's production path)."""
        dev = next(self.parameters()).device
        P = self._stack_points(points, dev)
        text, mask = tuple(self.get_text_proxy(text_dict))
        text = text.to(dev, torch.float32).contiguous()
        mask = mask.to(dev).to(torch.uint8).contiguous() if mask is not None else None
        img_feat = img_feat.to(dev)
        out, counts = self.forward_packed(P, text, mask, self._img_feat_dtype(img_feat).contiguous())
        coords, feats, total = ops.sparse_collate(out, counts, voxel_size, reciprocal=reciprocal, floor=floor)
        t = int(total.item())                                                                    # the one D2H sync
        return coords[:t], feats[:t]

    @staticmethod
    def _stack_points(points, dev) -> torch.Tensor:
        if isinstance(points, torch.Tensor):          # already (B,N,3)
            P = points
        else:
            n0 = points[0].shape
            for p in points:
                if p.shape != n0:
                    raise RuntimeError(f"all scenes must have the same number of points (got {tuple(p.shape)} vs {tuple(n0)})")
            P = torch.stack([p.to(dev, non_blocking=True) for p in points], 0)
        if P.dim() != 3 or P.shape[-1] != 3:
            raise ValueError(f"points must be (N,3) xyz per scene, got {tuple(P.shape)}")
        return P.to(dev, torch.float32, non_blocking=True).contiguous()

    def _side_stream(self, device, name: str = "img") -> "torch.cuda.Stream":
        key = f"{device}/{name}"
        if key not in self._streams:
            self._streams[key] = torch.cuda.Stream(device=device)
        return self._streams[key]

    def forward_packed(self, P, text, mask, img_feat, *, img_proxy=None, trace=None):
        """Device-resident form: P (B,N,3), text (B,L,c), mask (B,L) uint8|None, img_feat (B,V,C,H,W) ->
        (out (B,N,3) packed per scene, counts (B,) int32), no host synchronisation.  Runs on P's device and that device's
        current stream whatever the calling thread's current device is."""
        with torch.cuda.device(P.device):
            return self._forward_packed(P, text, mask, img_feat, img_proxy, trace)

    def _forward_packed(self, P, text, mask, img_feat, img_proxy, trace):
        self._check_mode()
        w = self._weights(P.device)
        K, n = self.num_sub, self.real_cluster_num
        train = self.training                        # batch-statistics BatchNorm (forward only), see forward()
        # S9 image proxies (:449) depend on nothing but img_feat.  Their first half (pass over the features for the spatial
        # means + query-side projections) is HBM-bound and light on SM resources, the geometric stages S1-S6 are latency /
        # ALU bound and barely touch HBM: the two run concurrently on two streams.  The second half (the persistent pooling
        # kernel, which owns every SM's shared memory) stays on the main stream.
        side = img_state = None
        img_side = None
        ov = self.overlap_img_stage
        if img_proxy is None and not train and (ov == "1" or (ov == "auto" and (P.shape[0] >= self.overlap_img_min_batch or
                                                                                  torch.cuda.is_current_stream_capturing()))):
            cur = torch.cuda.current_stream(P.device)
            img_side = self._side_stream(P.device)
            # The result is allocated from the CURRENT stream's pool and handed to the side stream: the current stream waits for
            # the side stream before it reads the proxies, so every later reuse of the block (and of the caller's img_feat) is
            # ordered after the side stream's work without record_stream().  A result allocated on the side stream and
            # record_stream()-ed to the current one cannot be reused until its event has completed; with the host running
            # several steps ahead that meant a fresh cudaMalloc per step in flight (seen as 8-11 ms steps at the start of a run).
            img_proxy = torch.empty(img_feat.shape[0], img_feat.shape[1], self.embed_dim, dtype=torch.float32, device=P.device)
            img_side.wait_stream(cur)
            with torch.cuda.stream(img_side):
                ops.img_attnpool(img_feat, w["img"], self.num_heads, params=w["img_struct"], out=img_proxy)
        elif img_proxy is None and self.overlap_mean_pass:
            cur = torch.cuda.current_stream(P.device)
            side = self._side_stream(P.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                img_state = ops.img_attnpool(img_feat, w["img"], self.num_heads, params=w["img_struct"], stages=ops.IMG_STAGE_FRONT)
            img_feat.record_stream(side)
            for t in img_state:                     # allocated on the side stream, finished on the main one
                t.record_stream(cur)
        # S1-S4 deformable clustering (:53-67)
        mn, mx, c0 = ops.minmax_centres(P, self.grid_size, w["lin"])
        idx1, _ = ops.ball_query(c0, P, K)
        w_off = w["offset"]
        if train:                                    # :72 BatchNorm2d over (B, M, K)
            sc, sh = ops.bn_batch_affine_cluster_conv(P, idx1, c0, w_off, self.get_deformable_cluster.get_offsets.mlp[1])
            w_off = dict(w_off, bn_scale=sc, bn_shift=sh)
        centres = ops.offset_net(P, idx1, c0, mn, mx, w_off)
        idx2, _ = ops.ball_query(centres, P, K)
        # S5 dropout (:352-420)
        kept_src, kc, kidx, drop_idx, fps = ops.cluster_dropout(centres, idx2, self.keep1, n)
        # index half of S10-S12 (who writes each point, dropped points per block) now, off the tail of the step
        scatter_ws = ops.affine_scatter_mark(P, kidx, drop_idx)
        # S6 point proxies (:437)
        w_enc = w["encoder"]
        if train:                                    # :112 BatchNorm2d over (B, n, K)
            sc, sh = ops.bn_batch_affine_cluster_conv(P, kidx, kc, w_enc, self.simple_encoder.mlp[1])
            w_enc = dict(w_enc, bn_scale=sc, bn_shift=sh)
        pp = ops.point_encoder(P, kidx, kc, w_enc)
        # The two branches (:440-446 text, :449-455 image) only share the point proxies: the image branch follows the image
        # stage on ITS stream and runs next to the text branch; the main stream joins before the scatter.  At small B * n the
        # GEMMs under-fill the GPU (the shipped config at batch 4 has 22 row tiles for 148 SMs: 0.851 -> 0.783 ms per forward
        # as a CUDA graph); at the benchmark's 16 384 rows the tails of one branch's kernels fill with the other's (2.20 -> 2.14 ms).
        ig = transform = None
        if img_side is not None and P.shape[0] * n <= self.parallel_branch_max_rows:
            cur = torch.cuda.current_stream(P.device)
            pp_ready = torch.cuda.Event()
            pp_ready.record(cur)
            with torch.cuda.stream(img_side):
                img_side.wait_event(pp_ready)
                ig = ops.proxy_block(pp, img_proxy, None, w["imgb"], self.num_heads, params=w["imgb_struct"])
                ih = w["img_head"]
                transform = ops.heads(ig, ih["lin_w"], ih["lin_b"], ih["bn_scale"], ih["bn_shift"])
        # S7/S8 text branch -> translate (:440-446)
        tg = ops.proxy_block(pp, text, mask, w["text"], self.num_heads, params=w["text_struct"])
        th = w["text_head"]
        t_sc, t_sh = (ops.bn_batch_affine_linear(tg, th["lin_w"], th["lin_b"], self.text_trans_norm) if train
                      else (th["bn_scale"], th["bn_shift"]))                                # :328 BatchNorm1d over (B, n)
        translate = ops.heads(tg, th["lin_w"], th["lin_b"], t_sc, t_sh)
        # S9 image proxies (:449) and image branch -> transform (:450-455)
        if img_side is not None:
            torch.cuda.current_stream(P.device).wait_stream(img_side)
        elif side is not None:
            torch.cuda.current_stream(P.device).wait_stream(side)
            img_proxy = ops.img_attnpool(img_feat, w["img"], self.num_heads, params=w["img_struct"], stages=ops.IMG_STAGE_BACK,
                                         out=img_state[0], ws=img_state[1])[0]
        elif img_proxy is None:
            img_proxy = ops.img_attnpool(img_feat, w["img"], self.num_heads, params=w["img_struct"])
        if transform is None:
            ig = ops.proxy_block(pp, img_proxy, None, w["imgb"], self.num_heads, params=w["imgb_struct"])
            ih = w["img_head"]
            i_sc, i_sh = (ops.bn_batch_affine_linear(ig, ih["lin_w"], ih["lin_b"], self.img_trans_norm) if train
                          else (ih["bn_scale"], ih["bn_shift"]))                            # :330 BatchNorm1d over (B, n)
            transform = ops.heads(ig, ih["lin_w"], ih["lin_b"], i_sc, i_sh)
        if train:
            self._packed_key = None                  # running statistics changed behind torch's version counters: re-fold for eval
        # S10-S12 (:459-467)
        out, counts = ops.affine_scatter_compact(P, kidx, drop_idx, kc, transform, translate, ws=scatter_ws, marked=True)
        if trace is not None:
            trace.update(mn=mn, mx=mx, c0=c0, idx1=idx1, centres=centres, idx2=idx2, kept_src=kept_src, kept_centres=kc,
                         kept_idx=kidx, drop_idx=drop_idx, fps=fps, point_proxy=pp, text_guide=tg, translate=translate,
                         img_proxy=img_proxy, img_guide=ig, transform=transform, counts=counts, out=out)
        return out, counts
