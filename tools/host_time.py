"""Host-side time of one forward_packed call (Python + ctypes + launches, no synchronisation): python tools/host_time.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proxytransformation_b200 import ProxyTransformationNormReverse, synthetic as syn
cfg = syn.C2_WIDE
B = int(os.environ.get("QB", "64"))
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.cuda()
pts, td, img = syn.make_inputs(cfg, 2, img_dtype=torch.bfloat16)
P = torch.stack(pts).cuda().repeat(B // 2, 1, 1).contiguous()
text = td["text_feats"].cuda().repeat(B // 2, 1, 1).contiguous()
mask = td["text_token_mask"].cuda().to(torch.uint8).repeat(B // 2, 1).contiguous()
im = img.cuda().repeat(B // 2, 1, 1, 1, 1).contiguous()
with torch.no_grad():
    for _ in range(3):
        m.forward_packed(P, text, mask, im)
    torch.cuda.synchronize()
    for n in (1, 4, 8):
        t0 = time.perf_counter()
        for _ in range(n):
            m.forward_packed(P, text, mask, im)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{n} forwards: host {1e3 * (t1 - t0) / n:.3f} ms per forward, with sync {1e3 * (t2 - t0) / n:.3f} ms", flush=True)
