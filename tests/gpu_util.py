"""Helpers shared by the -m gpu parity tests (CUDA path vs oracle / golden fixtures)."""
import numpy as np
import torch

from oracle import preshape_oracle as po
from proxytransformation_b200 import ProxyTransformationNormReverse

DEV = "cuda"


def build_module(cfg, sd, tensor_cores=True):
    m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
    m.load_state_dict(sd, strict=True)
    m.use_tensor_cores = tensor_cores
    return m.to(DEV)


def cu(t, dtype=None):
    t = torch.as_tensor(t)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV).contiguous()


def conv_bn_weights(sd, prefix):
    """Packed conv/BN weights for ops.offset_net / ops.point_encoder from a reference-layout state_dict."""
    inv = 1.0 / torch.sqrt(sd[f"{prefix}.1.running_var"] + 1e-5)
    scale = inv * sd[f"{prefix}.1.weight"]
    shift = sd[f"{prefix}.1.bias"] - sd[f"{prefix}.1.running_mean"] * scale
    return dict(conv_w=cu(sd[f"{prefix}.0.weight"].reshape(256, 6)), conv_b=cu(sd[f"{prefix}.0.bias"]), bn_scale=cu(scale),
                bn_shift=cu(shift))


def oracle_forward(cfg, sd, pts, text_dict, img, **kw):
    trace = {}
    out = po.forward(sd, pts, text_dict, img, grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio,
                     text_blocks=cfg.text_blocks, img_blocks=cfg.img_blocks, num_sub=cfg.num_sub, num_heads=cfg.num_heads,
                     trace=trace, **kw)
    return out, trace


def np_(t):
    return t.detach().cpu().numpy()
