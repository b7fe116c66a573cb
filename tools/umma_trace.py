#!/usr/bin/env python
"""Per-role cycle trace of the tcgen05 image-pool kernel (CTA 0): python tools/umma_trace.py [batch]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PT_POOL_KERNEL"] = "umma"
os.environ["PT_UMMA_DEBUG"] = str(int(os.environ.get("PT_UMMA_DEBUG", "0")) | 16)
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
buf = (ctypes.c_ulonglong * 24)()
L.pt_debug_umma_trace(buf, 1)
m.get_img_proxy(img); torch.cuda.synchronize()
L.pt_debug_umma_trace(buf, 0)
views = -(-B * cfg.n_views // 148)
names = ["producer: wait empty", "issuer: wait full", "issuer: wait p_full", "issuer: wait wfull", "issuer: wait d2_empty", "issuer: total",
         "softmax: wait d1_full", "softmax: class exchange", "softmax: wait p_empty", "softmax: total", "softmax: bar_or/raise",
         "epilogue: s0", "epilogue: wait l_full", "epilogue: wait d2_full", "epilogue: store", "epilogue: total",
         "  softmax: tmem load", "  softmax: shuffles+stores", "  softmax: exchange barrier", "  softmax: gather", "  softmax: exp+split",
         "  softmax: P stores+arrive", "  softmax: end of view", "-"]
for n, v in zip(names, buf):
    print(f"{n:28s} {v / views:9.0f} cycles/view")
_lib.profile_enable(True)
img2 = img.flip(0).contiguous()
for _ in range(3):
    m.get_img_proxy(img); m.get_img_proxy(img2)
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
ms = prof["img_pool"][0] / prof["img_pool"][1]
print(f"pool kernel {ms:.4f} ms per {B} scenes = {ms * 1e-3 * 1.965e9 / views / 1e3:.1f} k cycles per view per CTA")
