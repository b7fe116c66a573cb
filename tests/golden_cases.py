"""Golden-case table shared by tests/golden/make_golden.py (generator, build
container only) and the parity tests.  Inputs are regenerated from seeds on both
sides; the fixtures hold what the unmodified reference produced for them."""
from __future__ import annotations

import os

import numpy as np
import torch

from proxytransformation_b200 import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (config, batch, first_scene, weight seed, mutate(points) or None, store_full_output)
def _collapse(pts):      # extent < 8 m on every axis: inverted/collapsed grid
    return [p * torch.tensor([0.2, 0.15, 0.1]) for p in pts]


def _origin_point(pts):  # a real point at exactly (0,0,0): pad-detection collision (:94, :132)
    out = []
    for p in pts:
        p = p.clone() - torch.tensor([6.0, 6.0, 6.0])
        p[7] = 0.0
        p[1234 % p.shape[0]] = torch.tensor([0.0, -0.0, 0.0])
        out.append(p)
    return out


def _sparse(pts):        # many centres with < K or zero hits: all-pad clusters, -1 inside drop_idx
    return [p * 4.0 for p in pts]


def _very_sparse(pts):   # > 30 % of the centres see no point at all: all-pad clusters survive the pad-count cut
    return [p * 10.0 for p in pts]


def _dups(pts):          # all points in one tiny blob far from most centres -> duplicated clamped centres
    return [p * 0.01 + torch.tensor([1.0, 2.0, 3.0]) for p in pts]


C1 = syn.C1
CASES = {
    "c1_b2": (C1, 2, 0, 0, None, True),
    "c1_collapsed": (C1, 1, 3, 1, _collapse, True),
    "c1_origin": (C1, 2, 5, 2, _origin_point, True),
    "c1_sparse": (C1, 2, 8, 3, _sparse, True),
    "c1_very_sparse": (C1, 2, 9, 3, _very_sparse, True),
    "c1_dups": (C1, 1, 11, 4, _dups, True),
    "c1_blocks3": (C1.replace(name="C1-b3", text_blocks=3, img_blocks=2, n_text=9, n_views=3), 1, 13, 5, None, True),
    "gs5_ragged": (C1.replace(name="gs5", n_points=5003, grid_size=5, dynamic_drop_radio=0.6, n_text=7, n_views=5,
                              num_sub=17), 3, 17, 6, None, True),
    "c1_qkv_bias": (C1.replace(name="C1-qkvb", qkv_bias=True, text_blocks=2, n_text=11, n_views=3), 2, 21, 10, None, True),
    "c2_wide_b1": (syn.C2_WIDE.replace(n_views=8), 1, 0, 7, None, False),
    "c2_room_b1": (syn.C2_ROOM.replace(n_views=8), 1, 1, 8, None, False),
    "c3_wide_b1": (syn.C3_WIDE.replace(n_views=6), 1, 2, 9, None, False),
}


def load_case(name: str):
    """-> (cfg, state_dict, points, text_dict, img_feat, golden npz dict)."""
    cfg, batch, first, wseed, mutate, _full = CASES[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))
    sd = syn.make_state_dict(cfg, wseed)
    pts, text_dict, img = syn.make_inputs(cfg, batch, first)
    if mutate is not None:
        pts = mutate(pts)
    return cfg, sd, pts, text_dict, img, g


SMALL_CASES = [k for k, v in CASES.items() if v[5]]
LARGE_CASES = [k for k, v in CASES.items() if not v[5]]


# ---- training-mode pin (oracle side): tests/golden/make_golden_train.py -> c1_train.npz
TRAIN_CASE = (C1.replace(name="C1-train", text_blocks=2, img_blocks=2, n_text=9, n_views=3), 2, 31, 12)   # cfg, batch, first scene, weight seed


def train_loss_weights(counts):
    """R_b of the fixed scalar loss L = sum_b <out_b, R_b>."""
    return [torch.randn(int(n), 3, generator=torch.Generator().manual_seed(777 + b)) for b, n in enumerate(counts)]


def FULL_GRAD_KEYS(cfg):
    """Parameters whose full gradients are stored (the small ones); every other gradient is pinned by its norm."""
    t, i = cfg.text_blocks - 1, cfg.img_blocks - 1
    return {"text_trans.weight", "text_trans.bias", "img_trans.weight", "img_trans.bias", "text_trans_norm.weight",
            "text_trans_norm.bias", "img_trans_norm.weight", "img_trans_norm.bias",
            "get_deformable_cluster.get_offsets.channel_mapper.weight", "get_deformable_cluster.get_offsets.mlp.0.weight",
            "get_deformable_cluster.get_offsets.mlp.1.weight", "get_deformable_cluster.get_offsets.mlp.1.bias",
            "simple_encoder.mlp.0.weight", "simple_encoder.mlp.0.bias", "simple_encoder.mlp.1.weight",
            f"textformer.{t}.norm1.weight", f"textformer.{t}.attn.pc_bias", f"imgformer.{i}.norm2.bias",
            f"imgformer.{i}.attn.proxy_proj.bias", f"text_norm.{t}.weight", f"img_norm.{i}.bias", "norm_img.weight",
            "channel_mapper.bias", "attn_pool2d.q_proj.bias"}
