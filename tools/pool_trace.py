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
names = ["view barrier", "operand wait", "score MMAs", "exchange+softmax", "weighted sums"]
tot = sum(buf[:5]) or 1
for n, v in zip(names, buf[:5]):
    print(f"{n:22s} {v / views:9.0f} cycles/view  {100.0 * v / tot:5.1f}%")
print(f"total {tot / views:.0f} cycles/view over {views} views")
# device time of the whole image stage (mean + GEMMs + pool + GEMMs + LN), alternating two inputs larger than L2
img2 = img.flip(0).contiguous()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    m.get_img_proxy(img); m.get_img_proxy(img2)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    m.get_img_proxy(img); m.get_img_proxy(img2)
e1.record(); torch.cuda.synchronize()
print(f"image stage {e0.elapsed_time(e1) / 20:.4f} ms per {B} scenes  (PT_POOL_PF={os.environ.get('PT_POOL_PF', 'default')})")
