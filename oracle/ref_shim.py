"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Imports the reference's hot-path file UNMODIFIED under three import shims (SURVEY.md §8c recipe) so the restatement
in ``preshape_oracle.py`` can be pinned against the reference itself and golden vectors can be generated.  The module is
loaded from ``/root/reference`` in the build container, or — on the GPU box, where that tree does not exist — from the
bytecode ``oracle/build.py::build_ref`` compiled from it into ``oracle/_ref/`` (git-ignored; travels like a built .so).
Nothing in the ``-m gpu`` tests or ``smoke()`` depends on it; ``bench.py`` uses it for the CPU arm only
(``cpu_baseline.kind = "reference"``) and falls back to the restatement (``"port"``) when it is absent.

Shims (none of these packages is installed here, and ``import embodiedscan``
proper fails as shipped — ``embodiedscan/utils/__init__.py:2`` needs a file that
is missing from the tree):
  1. ``embodiedscan.registry.MODELS``   -> object with a ``register_module()`` decorator;
  2. ``timm.models.layers.{Mlp,DropPath,trunc_normal_}`` -> fc1/act/drop1/fc2/drop2 Mlp with
     timm's attribute names, identity DropPath in eval, torch's trunc_normal_;
  3. ``pytorch3d.ops.ball_query`` -> the C restatement in geom.c (namedtuple ``(dists, idx, knn)``);
     ``pytorch3d.ops.sample_farthest_points`` -> the reference's OWN in-tree
     ``sample_farthest_points_naive`` (:527-625), assigned after import.
Pins applied when ``pinned=True``: ``torch.argsort(..., stable=True)`` inside the
reference module's namespace (:378) and single-threaded ``index_put_`` (:495).
"""
from __future__ import annotations

import collections
import importlib.machinery
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = "/root/reference"
REF_FILE = os.path.join(REF_ROOT, "embodiedscan/models/necks/preshape_norm_reverse_drop.py")


REF_PYC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "preshape_norm_reverse_drop.bytecode")


def available() -> bool:
    return os.path.exists(REF_FILE) or os.path.exists(REF_PYC)


class _Registry:
    def __init__(self):
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        return deco(module) if module is not None else deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self.module_dict[cfg.pop("type")](**cfg)


class _Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.training and self.drop_prob > 0:
            raise RuntimeError("oracle shim: DropPath is eval-only")
        return x


_BallQuery = collections.namedtuple("_BallQuery", "dists idx knn")
_module = None


def load():
    """-> the reference module object (cached)."""
    global _module
    if _module is not None:
        return _module
    if not available():
        raise FileNotFoundError(f"{REF_FILE} (or the bytecode {REF_PYC} built from it by oracle/build.py)")
    from . import preshape_oracle as po

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    saved = {k: sys.modules.get(k) for k in ("embodiedscan", "embodiedscan.registry", "timm", "timm.models",
                                             "timm.models.layers", "pytorch3d", "pytorch3d.ops")}
    reg = mod("embodiedscan.registry")
    mod("embodiedscan")
    reg.MODELS = _Registry()
    mod("timm"); mod("timm.models")
    tl = mod("timm.models.layers")
    tl.Mlp, tl.DropPath, tl.trunc_normal_ = _Mlp, _DropPath, torch.nn.init.trunc_normal_
    mod("pytorch3d")
    ops = mod("pytorch3d.ops")

    def ball_query(p1, p2, K, radius, **kw):
        idx, knn = po.ball_query(p1, p2, K, radius)
        return _BallQuery(None, idx, knn)

    ops.ball_query = ball_query
    ops.sample_farthest_points = None   # replaced below with the in-tree naive copy
    if os.path.exists(REF_FILE):
        spec = importlib.util.spec_from_file_location("_ref_preshape_norm_reverse_drop", REF_FILE)
    else:
        spec = importlib.util.spec_from_loader("_ref_preshape_norm_reverse_drop",
                                               importlib.machinery.SourcelessFileLoader("_ref_preshape_norm_reverse_drop", REF_PYC))
    m = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    m.sample_farthest_points = m.sample_farthest_points_naive
    m._registry = reg.MODELS
    _module = m
    return m


class _PinnedTorch:
    """Proxy for the ``torch`` name inside the reference module: argsort becomes stable."""
    def __init__(self):
        self._t = torch

    def __getattr__(self, k):
        return getattr(self._t, k)

    def argsort(self, x, dim=-1, descending=False, stable=False):
        return torch.argsort(x, dim=dim, descending=descending, stable=True)


def use_native_fps(on: bool = True):
    """Timing mode: ``sample_farthest_points`` -> the C restatement of pytorch3d's CPU loop (what the reference executes in
    production) instead of the reference's in-tree pure-Python copy (:527-625), which is what the parity pins use."""
    m = load()
    if on:
        from . import preshape_oracle as po

        def sample_farthest_points(points, lengths=None, K=50, random_start_point=False):
            idx = po.farthest_point_indices(points, int(K))
            return torch.gather(points, 1, idx[..., None].expand(-1, -1, points.shape[-1])), idx
        m.sample_farthest_points = sample_farthest_points
    else:
        m.sample_farthest_points = m.sample_farthest_points_naive


def build_module(cfg_kwargs: dict, state_dict=None, pinned: bool = True):
    m = load()
    m.torch = _PinnedTorch() if pinned else torch
    net = m.ProxyTransformationNormReverse(**cfg_kwargs).eval()
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    return net


@torch.no_grad()
def run_reference(net, points, text_dict, img_feat, pinned: bool = True):
    """Forward of the unmodified reference module; with ``pinned`` the duplicate
    scatter at :495 runs single-threaded (last writer in flat order wins)."""
    nt = torch.get_num_threads()
    if pinned:
        torch.set_num_threads(1)
    try:
        return net(points, text_dict, img_feat)
    finally:
        torch.set_num_threads(nt)
