#!/usr/bin/env python
"""Timeline of ONE forward of the benchmark workload with the image stage on its second stream: every kernel's start / end on a common
time axis (event timestamps), to see which kernels of the two streams actually ran side by side.  python tools/step_timeline.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proxytransformation_b200 import ProxyTransformationNormReverse, _lib, synthetic as syn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = syn.C2_WIDE
dev = torch.device("cuda", 0)
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.to(dev)
g = torch.Generator(device=dev).manual_seed(1)
P = torch.rand(B, cfg.n_points, 3, generator=g, device=dev) * torch.tensor(cfg.box, device=dev)
text = torch.randn(B, cfg.n_text, cfg.embed_dim, generator=g, device=dev)
mask = torch.ones(B, cfg.n_text, dtype=torch.uint8, device=dev)
hw = cfg.img_spacial_dim
img = (torch.relu(torch.randn(B, cfg.n_views, cfg.input_dim, hw, hw, generator=g, device=dev)) * 1.5).to(torch.bfloat16)
for _ in range(12):
    m.forward_packed(P, text, mask, img)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(3):
    m.forward_packed(P, text, mask, img)
torch.cuda.synchronize()
tl = _lib.profile_timeline()
_lib.profile_enable(False)
n = len(tl) // 3
step = tl[2 * n:]
t0 = min(s for _, s, _ in step)
print(f"{len(step)} launches, step span {max(e for _, _, e in step) - t0:.3f} ms")
for name, s, e in sorted(step, key=lambda r: r[1]):
    print(f"{s - t0:8.3f} {e - t0:8.3f}  {e - s:7.3f}  {name}")
