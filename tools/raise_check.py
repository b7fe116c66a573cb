"""Drives the reference-maximum raise path of the tcgen05 pool kernel (late tokens scoring far above the first window) and
prints the error of both pool kernels against the oracle.  GPU box: python tools/raise_check.py [ramp]"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import preshape_oracle as po
from proxytransformation_b200 import ProxyTransformationNormReverse, synthetic as syn
ramp_v = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
cfg = syn.C2_WIDE.replace(n_views=40)
sd = syn.make_state_dict(cfg, 33, bf16_round=True)
_, _, img = syn.make_inputs(cfg.replace(n_points=8), 5, first_scene=310, img_dtype=torch.float32)
ramp = torch.ones(225); ramp[150:] = ramp_v
img = (img.reshape(5, 40, 512, 225) * ramp).reshape(5, 40, 512, 15, 15).to(torch.bfloat16)
want = po.image_proxies(sd, img.float(), cfg.num_heads)
for kernel in ("umma", "mma"):
    os.environ["PT_POOL_KERNEL"] = kernel
    m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval(); m.load_state_dict(sd); m = m.cuda()
    with torch.no_grad():
        got = m.get_img_proxy(img.cuda()).cpu()
    err = (got - want).abs().reshape(200, -1).max(-1)[0]
    print(kernel, "max err", float(err.max()), "views > 6e-5:", int((err > 6e-5).sum()), "nan:", int(torch.isnan(got).sum()), "worst view", int(err.argmax()))
