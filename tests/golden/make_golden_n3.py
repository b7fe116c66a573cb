"""Generates tests/golden/n3_aggregate.npz: the UNMODIFIED reference data transforms of the input side,
``AggregateMultiViewPoints.transform`` (embodiedscan/datasets/transforms/multiview.py:224-251) followed by
``PointSample._points_random_sampling`` (embodiedscan/datasets/transforms/points.py:373-419), run on seeded multi-view
ego-frame points.  The two files and the reference's own point structures (embodiedscan/structures/points/*.py) are imported
by path under import shims for what is absent here (mmcv, the registry, pytorch3d-dependent box utilities); nothing of the
reference is modified or copied.  Pins oracle.aggregate_sample and, on the GPU, pt_aggregate_sample.

Run in the build container only:    python tests/golden/make_golden_n3.py
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/embodiedscan"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_transforms():
    """-> (AggregateMultiViewPoints, PointSample, DepthPoints) classes of the unmodified reference."""
    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    class BaseTransform:                       # mmcv.transforms.BaseTransform: __call__ -> transform
        def __call__(self, results):
            return self.transform(results)

    _stub("mmcv", imresize=None)
    _stub("mmcv.transforms", BaseTransform=BaseTransform, Compose=object)
    pkg = _stub("embodiedscan"); pkg.__path__ = []
    _stub("embodiedscan.registry", TRANSFORMS=_Registry())
    st = _stub("embodiedscan.structures"); st.__path__ = [os.path.join(REF, "structures")]       # real sub-packages, no __init__
    bb = _stub("embodiedscan.structures.bbox_3d", points_cam2img=None, points_img2cam=None); bb.__path__ = []
    _stub("embodiedscan.structures.bbox_3d.utils", rotation_3d_in_axis=None, rotation_3d_in_euler=None)                      # (needs pytorch3d; unused here)
    points_pkg = importlib.import_module("embodiedscan.structures.points")                        # the reference's own files

    def by_path(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    mv = by_path("embodiedscan.datasets.transforms.multiview", "datasets/transforms/multiview.py")
    pt = by_path("embodiedscan.datasets.transforms.points", "datasets/transforms/points.py")
    return mv.AggregateMultiViewPoints, pt.PointSample, points_pkg.DepthPoints


def make_views(seed: int, n_views: int, dims: int):
    """Seeded ego-frame views (n_v, dims) fp32 and global->ego extrinsics (n_views, 4, 4) fp32 (rigid, as the dataset stores them)."""
    g = torch.Generator().manual_seed(seed)
    views, ext = [], np.zeros((n_views, 4, 4), np.float32)
    for v in range(n_views):
        n_v = int(torch.randint(300, 900, (1,), generator=g))
        views.append(torch.cat([torch.rand(n_v, 3, generator=g) * 6 - 3, torch.rand(n_v, dims - 3, generator=g)], 1).float())
        a = torch.rand(3, generator=g) * 6.2831853
        rz = torch.tensor([[torch.cos(a[0]), -torch.sin(a[0]), 0], [torch.sin(a[0]), torch.cos(a[0]), 0], [0, 0, 1.0]])
        ry = torch.tensor([[torch.cos(a[1]), 0, torch.sin(a[1])], [0, 1.0, 0], [-torch.sin(a[1]), 0, torch.cos(a[1])]])
        rx = torch.tensor([[1.0, 0, 0], [0, torch.cos(a[2]), -torch.sin(a[2])], [0, torch.sin(a[2]), torch.cos(a[2])]])
        e = torch.eye(4)
        e[:3, :3] = rz @ ry @ rx
        e[:3, 3] = torch.rand(3, generator=g) * 10 - 5
        ext[v] = e.numpy()
    return views, ext


CASES = {"xyz": (11, 5, 3, 2048), "xyzrgb": (12, 7, 6, 1500), "replace": (13, 2, 3, 4000)}   # seed, views, point dims, samples


def main():
    Agg, Sample, DepthPoints = load_reference_transforms()
    out = {}
    for name, (seed, n_views, dims, n_samples) in CASES.items():
        views, ext = make_views(seed, n_views, dims)
        attr = dict(color=[3, 4, 5]) if dims == 6 else None
        results = {"points": [DepthPoints(v.clone(), points_dim=dims, attribute_dims=attr) for v in views], "depth2img": {"extrinsic": [e for e in ext]}}
        results = Agg(coord_type="DEPTH")(results)                                   # multiview.py:224-251
        np.random.seed(1000 + seed)
        sampled, choices = Sample(num_points=n_samples)._points_random_sampling(results["points"], n_samples, return_choices=True)   # points.py:373-419
        out[f"{name}.views"] = torch.cat(views).numpy()
        out[f"{name}.sizes"] = np.array([len(v) for v in views], np.int64)
        out[f"{name}.extrinsics"] = ext
        out[f"{name}.choices"] = np.asarray(choices, np.int64)
        out[f"{name}.aggregated"] = results["points"].tensor.numpy()
        out[f"{name}.sampled"] = sampled.tensor.numpy()
    np.savez_compressed(os.path.join(HERE, "n3_aggregate.npz"), **out)
    print("wrote n3_aggregate.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
