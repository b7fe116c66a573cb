"""Generates tests/golden/*.npz by running the UNMODIFIED reference module
(/root/reference, imported under the shims of oracle/ref_shim.py) on the
deterministic synthetic scenes/weights of proxytransformation_b200.synthetic.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py [case names ...]
Each fixture stores the seeds/config needed to regenerate the inputs, every
intermediate of the path captured from the reference itself, and either the
full outputs (small cases) or their digests (100k-point cases).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from proxytransformation_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

from tests.golden_cases import CASES  # noqa: E402


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def capture(net, ref_mod, pts, text_dict, img):
    """Run the reference forward and capture its intermediates by wrapping its own methods."""
    cap = {}
    orig = dict(gpc=net.get_point_cluster, dcd=net.dynamic_cluster_dropout, gpp=net.get_point_proxy,
                gip=net.get_img_proxy, ptr=ref_mod.pt_replace)

    def gpc(points):
        c, cl, idx = orig["gpc"](points)
        cap.update(centres=c.clone(), idx2=idx.clone())
        return c, cl, idx

    def dcd(cluster, center, idx, empty_drop=0.3):
        r = orig["dcd"](cluster, center, idx, empty_drop)
        cap.update(kept_centres=r[1].clone(), kept_idx=r[2].clone(), drop_idx=r[3].clone())
        return r

    def gpp(center, cluster):
        r = orig["gpp"](center, cluster)
        cap["point_proxy"] = r.clone()
        return r

    def gip(img_feat):
        r = orig["gip"](img_feat)
        cap["img_proxy"] = r.clone()
        return r

    def ptr(p2, idx, cluster):
        cap["new_clusters"] = cluster.clone()
        r = orig["ptr"](p2, idx, cluster)
        cap["scattered"] = r.clone()
        return r

    hooks = [net.text_trans_norm.register_forward_hook(lambda m, i, o: cap.__setitem__("translate", o.transpose(-2, -1).clone())),
             net.img_trans_norm.register_forward_hook(lambda m, i, o: cap.__setitem__("transform", o.transpose(-2, -1).clone()))]
    net.get_point_cluster, net.dynamic_cluster_dropout, net.get_point_proxy, net.get_img_proxy = gpc, dcd, gpp, gip
    ref_mod.pt_replace = ptr
    try:
        out = ref_shim.run_reference(net, pts, text_dict, img, pinned=True)
    finally:
        ref_mod.pt_replace = orig["ptr"]
        for h in hooks:
            h.remove()
        del net.get_point_cluster, net.dynamic_cluster_dropout, net.get_point_proxy, net.get_img_proxy
    return out, cap


def main():
    ref_mod = ref_shim.load()
    only = set(sys.argv[1:])          # optional: generate just the named cases (fixtures of the others stay untouched)
    for name, (cfg, batch, first, wseed, mutate, full) in CASES.items():
        if only and name not in only:
            continue
        sd = syn.make_state_dict(cfg, wseed)
        pts, text_dict, img = syn.make_inputs(cfg, batch, first)
        if mutate is not None:
            pts = mutate(pts)
        net = ref_shim.build_module(cfg.module_kwargs(), sd, pinned=True)
        out, cap = capture(net, ref_mod, pts, text_dict, img)
        rec = {
            "cfg_name": np.array(cfg.name), "batch": np.array(batch), "first_scene": np.array(first),
            "weight_seed": np.array(wseed), "mutate": np.array(mutate.__name__ if mutate else ""),
            "cfg_kwargs": np.array(repr(dict(n_points=cfg.n_points, grid_size=cfg.grid_size,
                                              dynamic_drop_radio=cfg.dynamic_drop_radio, text_blocks=cfg.text_blocks,
                                              img_blocks=cfg.img_blocks, num_sub=cfg.num_sub, n_text=cfg.n_text,
                                              n_views=cfg.n_views, box=cfg.box))),
            "out_counts": np.array([o.shape[0] for o in out]),
            "out_digest": np.array([digest(o) for o in out]),
            "out_sum": np.array([o.double().sum(0).numpy() for o in out]),
        }
        for k in ("centres", "idx2", "kept_centres", "kept_idx", "drop_idx", "point_proxy", "img_proxy", "translate",
                  "transform"):
            v = cap[k]
            rec[k] = v.numpy().astype(np.int32) if v.dtype == torch.int64 else v.numpy()
        P = torch.stack(pts, 0)
        for b in range(batch):     # rows the scatter (:495) changed, before removal (:467)
            ch = (cap["scattered"][b] != P[b]).any(-1).nonzero(as_tuple=True)[0]
            rec[f"changed_rows_{b}"] = ch.numpy().astype(np.int32)
            rec[f"changed_vals_{b}"] = cap["scattered"][b][ch].numpy()
        if full:
            for b, o in enumerate(out):
                rec[f"out_{b}"] = o.numpy()
        else:
            # 100k-point cases: keep only the rows the path changed plus strided samples of the survivors
            for b, o in enumerate(out):
                rec[f"out_head_{b}"] = o[:2048].numpy()
                rec[f"out_stride_{b}"] = o[::97].numpy()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: counts={rec['out_counts'].tolist()} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
