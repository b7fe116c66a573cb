#!/usr/bin/env python
"""Timeline of the image-pool kernel's load pipeline (CTA 0, views 40..43): PT_POOL_DEBUG=32 python tools/pool_events.py [batch]
Prints every logged event with its SM-clock offset: producer issue times of each ring load, and when warp 0 / warp 8
saw each slab pair land (scores) or started each weighted-sum step."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PT_POOL_DEBUG"] = str(int(os.environ.get("PT_POOL_DEBUG", "0")) | 32)
import torch
from proxytransformation_b200 import ProxyTransformationNormReverse, _lib, synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = syn.C2_WIDE
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.cuda()
img = (torch.relu(torch.randn(B, cfg.n_views, 512, 15, 15, device="cuda")) * 1.5).bfloat16()
L = _lib.load()
buf = (ctypes.c_longlong * 8192)()
m.get_img_proxy(img); torch.cuda.synchronize()
L.pt_debug_pool_events(buf, 4096)
m.get_img_proxy(img); torch.cuda.synchronize()
n = L.pt_debug_pool_events(buf, 4096)
ev = sorted((buf[2 * i + 1], buf[2 * i]) for i in range(n))      # 32-bit SM clock (a 0.9 ms kernel cannot wrap twice)
t0 = ev[0][0]
CODES = {0: "top of view", 1: "past view barrier", 2: "operands landed", 3: "score MMAs done", 4: "softmax done"}
FINE = {20: "41 after barrier (planes dead)", 21: "41 after store round + barrier", 22: "41 after add round + barrier", 23: "41 after row max + barrier",
        24: "41 after row sum + barrier", 25: "41 probabilities in registers"}
rows = {}
for t, i in ev:
    if i >= 100000:
        w, c = divmod(i - 100000, 1000)
        rows.setdefault((1, 100 + c), {})[w] = t - t0
        continue
    w, r = divmod(i, 1000); v, c = divmod(r, 10)
    rows.setdefault((v, c), {})[w] = t - t0
print("cycles since the first event; one column per consumer warp (lane 0)")
print(f"{'view event':34s} " + " ".join(f"w{w:<6d}" for w in range(16)))
for (v, c) in sorted(rows):
    label = FINE[c - 100] if c >= 100 else f"{40 + v:4d} {CODES.get(c, c)}"
    print(f"{label:34s} " + " ".join(f"{rows[(v, c)].get(w, -1):7d}" for w in range(16)))
