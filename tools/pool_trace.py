#!/usr/bin/env python
"""Phase timeline of the image-pool kernel (CTA 0): PT_POOL_DEBUG=8 python tools/pool_trace.py [batch]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PT_POOL_DEBUG"] = str(int(os.environ.get("PT_POOL_DEBUG", "0")) | 8)
import torch
from proxytransformation_b200 import ProxyTransformationNormReverse, _lib, synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = syn.C2_WIDE
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.cuda()
img = (torch.relu(torch.randn(B, cfg.n_views, 512, 15, 15, device="cuda")) * 1.5).bfloat16()
m.get_img_proxy(img); torch.cuda.synchronize()
L = _lib.load()
buf = (ctypes.c_ulonglong * 8)()
L.pt_debug_pool_trace(buf, 1)
m.get_img_proxy(img); torch.cuda.synchronize()
L.pt_debug_pool_trace(buf, 0)
views = -(-B * cfg.n_views // 148)
names = ["view barrier", "operand wait", "conversion+barrier", "score MMAs", "exchange+softmax", "weighted sums"]
tot = sum(buf[:6])
for n, v in zip(names, buf[:6]):
    print(f"{n:22s} {v / views:9.0f} cycles/view  {100.0 * v / tot:5.1f}%")
print(f"total {tot / views:.0f} cycles/view over {views} views")
