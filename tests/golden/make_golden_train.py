"""Generates tests/golden/c1_train.npz: the UNMODIFIED reference module (oracle/ref_shim.py shims) in train() mode with
every stochastic layer at rate 0 (drop_rate = attn_drop_rate = drop_path_rate = 0), one forward + backward on a seeded
scene batch — outputs, the BatchNorm running statistics after the step, and the parameter gradients of the fixed scalar
loss  L = sum_b <out_b, R_b>,  R_b = randn(N'_b, 3; seed 777 + b)  (norm of every gradient, full tensors for the small
parameters).  This pins the training-mode side of oracle/preshape_oracle.py.

Run in the build container only:    python tests/golden/make_golden_train.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from proxytransformation_b200 import synthetic as syn  # noqa: E402
from tests.golden_cases import TRAIN_CASE, train_loss_weights, FULL_GRAD_KEYS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    cfg, batch, first, wseed = TRAIN_CASE
    sd = syn.make_state_dict(cfg, wseed)
    pts, text_dict, img = syn.make_inputs(cfg, batch, first)
    kw = dict(cfg.module_kwargs(), drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0)
    net = ref_shim.build_module(kw, sd, pinned=True)
    net.train()
    nt = torch.get_num_threads()
    torch.set_num_threads(1)                       # pinned duplicate rule of the scatter (:495), see ref_shim.run_reference
    try:
        out = net(pts, text_dict, img)
        loss = sum((o * r).sum() for o, r in zip(out, train_loss_weights([o.shape[0] for o in out])))
        loss.backward()
    finally:
        torch.set_num_threads(nt)
    rec = {"loss": np.array(float(loss)), "out_counts": np.array([o.shape[0] for o in out])}
    for b, o in enumerate(out):
        rec[f"out_{b}"] = o.detach().numpy()
    for k, v in net.state_dict().items():
        if "running_" in k or "num_batches_tracked" in k:
            rec["bn/" + k] = v.numpy()
    names, norms = [], []
    for k, p in net.named_parameters():
        names.append(k)
        norms.append(float(p.grad.double().norm()) if p.grad is not None else -1.0)     # -1: parameter not on the path
        if k in FULL_GRAD_KEYS(cfg):
            rec["grad/" + k] = p.grad.numpy()
    rec["grad_names"] = np.array(names)
    rec["grad_norms"] = np.array(norms)
    path = os.path.join(HERE, "c1_train.npz")
    np.savez_compressed(path, **rec)
    print(f"c1_train: loss={float(loss):.6f} counts={rec['out_counts'].tolist()} "
          f"params with grad: {sum(n >= 0 for n in norms)}/{len(norms)} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
