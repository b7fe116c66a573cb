"""Deterministic synthetic scenes and weights for the preshape hot path.

The reference ships neither data nor tests for this path, so both
sides of every parity check (oracle and CUDA path) are fed from the generators
here.  Shapes follow the reference's contracts:

* points: list of B fp32 ``(N, 3)`` tensors, equal N, random order
  (``embodiedscan/datasets/transforms/points.py:290,411``),
* ``text_dict``: ``{'text_feats': (B,L,256) f32, 'text_token_mask': (B,L) bool}``
  in that insertion order (``detectors/sparse_featfusion_grounder_preshape.py:671-673``),
* ``img_feat``: ``(B, V, 512, 15, 15)`` (``necks/preshape_norm_reverse_drop.py:336``),
* state_dict keys/shapes of ``ProxyTransformationNormReverse``
  (``necks/preshape_norm_reverse_drop.py:282-330``).

Everything is generated on the CPU with a seeded ``torch.Generator`` so the same
tensors appear in this container (golden generation) and on the GPU box.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class PreshapeConfig:
    """Constructor kwargs of the module + the input shapes of one workload."""
    name: str
    n_points: int
    grid_size: int
    dynamic_drop_radio: float
    text_blocks: int = 3
    img_blocks: int = 3
    num_sub: int = 30
    embed_dim: int = 256
    num_heads: int = 8
    input_dim: int = 512
    img_spacial_dim: int = 15
    n_text: int = 64          # L
    n_views: int = 196        # V
    box: Tuple[float, float, float] = (24.0, 24.0, 24.0)
    qkv_bias: bool = False    # :285 (no shipped config sets it)

    @property
    def num_cluster(self) -> int:          # M
        return self.grid_size ** 3

    @property
    def keep1(self) -> int:                # :374-376
        return self.num_cluster - int(self.num_cluster * 0.3)

    @property
    def real_cluster_num(self) -> int:     # n, :195, :389
        return int(self.num_cluster * (1 - self.dynamic_drop_radio))

    @property
    def n_drop(self) -> int:               # :390
        return self.keep1 - self.real_cluster_num

    def module_kwargs(self) -> dict:
        return dict(embed_dim=self.embed_dim, num_heads=self.num_heads, n_points=self.n_points,
                    grid_size=self.grid_size, text_blocks=self.text_blocks, img_blocks=self.img_blocks,
                    dynamic_drop_radio=self.dynamic_drop_radio, num_sub=self.num_sub,
                    input_dim=self.input_dim, img_spacial_dim=self.img_spacial_dim, qkv_bias=self.qkv_bias)

    def replace(self, **kw) -> "PreshapeConfig":
        d = asdict(self)
        d.update(kw)
        return PreshapeConfig(**d)


# BASELINE.json configs mapped to concrete shapes (SURVEY.md §8 preamble).
C1 = PreshapeConfig("C1", n_points=4096, grid_size=4, dynamic_drop_radio=0.75, text_blocks=1, img_blocks=1,
                    n_text=16, n_views=4, box=(12.0, 12.0, 12.0))
C2_WIDE = PreshapeConfig("C2-wide", n_points=100000, grid_size=8, dynamic_drop_radio=0.5, n_text=64, n_views=196,
                         box=(24.0, 24.0, 24.0))
C2_ROOM = C2_WIDE.replace(name="C2-room", box=(8.0, 6.0, 3.0))
C3 = PreshapeConfig("C3", n_points=100000, grid_size=12, dynamic_drop_radio=0.6, n_text=32, n_views=50,
                    box=(8.0, 6.0, 3.0))
C3_WIDE = C3.replace(name="C3-wide", box=(40.0, 40.0, 40.0))
CONFIGS = {c.name: c for c in (C1, C2_WIDE, C2_ROOM, C3, C3_WIDE)}


def state_dict_spec(cfg: PreshapeConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every entry of the module's state_dict, in module
    registration order.  kind drives the synthetic distribution."""
    c, n, s = cfg.embed_dim, cfg.real_cluster_num, int(cfg.embed_dim ** 0.5)
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(prefix, ch):
        spec.extend([(f"{prefix}.weight", (ch,), "gamma"), (f"{prefix}.bias", (ch,), "beta"),
                     (f"{prefix}.running_mean", (ch,), "rmean"), (f"{prefix}.running_var", (ch,), "rvar"),
                     (f"{prefix}.num_batches_tracked", (), "count")])

    def ln(prefix):
        spec.extend([(f"{prefix}.weight", (c,), "gamma"), (f"{prefix}.bias", (c,), "beta")])

    def lin(prefix, out_f, in_f, bias=True, extra=()):
        spec.append((f"{prefix}.weight", (out_f, in_f) + tuple(extra), "w"))
        if bias:
            spec.append((f"{prefix}.bias", (out_f,), "b"))

    # :301 DeformablePointCluster.get_offsets = OffsetNetwork(6, 256)  (:31 — hidden is 256 regardless of embed_dim)
    lin("get_deformable_cluster.get_offsets.mlp.0", 256, 6, extra=(1, 1))
    bn("get_deformable_cluster.get_offsets.mlp.1", 256)
    lin("get_deformable_cluster.get_offsets.channel_mapper", 3, 256, bias=False, extra=(1,))
    # :302 SimplifiedPointNet()
    lin("simple_encoder.mlp.0", 256, 6, extra=(1, 1))
    bn("simple_encoder.mlp.1", 256)
    # :304-306
    lin("channel_mapper", c, cfg.input_dim, extra=(1, 1))
    spec.append(("attn_pool2d.positional_embedding", (cfg.img_spacial_dim ** 2 + 1, c), "pos"))
    for nm in ("k_proj", "q_proj", "v_proj", "c_proj"):
        lin(f"attn_pool2d.{nm}", c, c)
    ln("norm_img")

    def blocks(stack, count):
        for i in range(count):
            p = f"{stack}.{i}"
            ln(f"{p}.norm1")
            spec.append((f"{p}.attn.pb_bias", (1, n, 4, 4), "tn"))
            spec.append((f"{p}.attn.pc_bias", (1, n, s, 1), "tn"))
            spec.append((f"{p}.attn.pr_bias", (1, n, 1, s), "tn"))
            lin(f"{p}.attn.qkv", 3 * c, c, bias=cfg.qkv_bias)
            lin(f"{p}.attn.proxy_proj", c, c)
            lin(f"{p}.attn.proj", c, c)
            ln(f"{p}.norm2")
            lin(f"{p}.mlp.fc1", 4 * c, c)
            lin(f"{p}.mlp.fc2", c, 4 * c)

    blocks("textformer", cfg.text_blocks)
    for i in range(cfg.text_blocks):
        ln(f"text_norm.{i}")
    blocks("imgformer", cfg.img_blocks)
    for i in range(cfg.img_blocks):
        ln(f"img_norm.{i}")
    lin("text_trans", 3, c)
    lin("img_trans", 9, c)
    bn("text_trans_norm", 3)
    bn("img_trans_norm", 9)
    return spec


def make_state_dict(cfg: PreshapeConfig, seed: int = 0, bf16_round: bool = False) -> Dict[str, torch.Tensor]:
    """Synthetic weights with NON-trivial BN statistics, LN affine and position
    biases (default torch init leaves them at identity/zero, which would hide
    folding bugs).  Loadable into the reference module with ``strict=True``."""
    g = torch.Generator().manual_seed(10007 * seed + 17)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in state_dict_spec(cfg):
        if kind == "w":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = (torch.rand(shape, generator=g) * 2 - 1) * (fan_in ** -0.5) * 1.5
        elif kind == "b":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == "gamma":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif kind == "beta":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == "rmean":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.2
        elif kind == "rvar":
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "count":
            t = torch.tensor(100, dtype=torch.long)
        elif kind == "pos":
            t = torch.randn(shape, generator=g) / shape[1] ** 0.5
        elif kind == "tn":
            t = torch.randn(shape, generator=g).clamp_(-2, 2) * 0.02
        else:  # pragma: no cover
            raise KeyError(kind)
        if bf16_round and t.is_floating_point():
            t = t.to(torch.bfloat16).to(torch.float32)
        sd[key] = t
    return sd


def make_points(cfg: PreshapeConfig, scene_id: int) -> torch.Tensor:
    """``torch.rand(N,3) * box`` with generator seed 1234+scene_id."""
    g = torch.Generator().manual_seed(1234 + scene_id)
    return torch.rand(cfg.n_points, 3, generator=g) * torch.tensor(cfg.box, dtype=torch.float32)


def make_inputs(cfg: PreshapeConfig, batch: int, first_scene: int = 0, img_dtype: torch.dtype = torch.float32,
                with_img: bool = True):
    """-> (points list, text_dict, img_feat).  The text mask keeps the first
    ``L - (scene_id % 8)`` tokens (True = real token)."""
    pts = [make_points(cfg, first_scene + b) for b in range(batch)]
    g = torch.Generator().manual_seed(99991 + first_scene)
    text = torch.randn(batch, cfg.n_text, cfg.embed_dim, generator=g)
    mask = torch.zeros(batch, cfg.n_text, dtype=torch.bool)
    for b in range(batch):
        mask[b, : max(1, cfg.n_text - ((first_scene + b) % 8))] = True
    text_dict = {"text_feats": text, "text_token_mask": mask}
    img = None
    if with_img:
        h = cfg.img_spacial_dim
        img = torch.randn(batch, cfg.n_views, cfg.input_dim, h, h, generator=g)
        # ResNet layer4 outputs are post-ReLU; keep the scale but make them non-negative-ish and sparse
        img = torch.relu(img) * 1.5
        if img_dtype != torch.float32:
            img = img.to(img_dtype)
    return pts, text_dict, img
