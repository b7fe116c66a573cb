"""Quickest possible A/B of the BACK image stage: shipped pool kernel vs PT_POOL_SINGLE=1 (16 scenes x 196 views, 0.72 GB >> L2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proxytransformation_b200 import ProxyTransformationNormReverse, ops, synthetic as syn
B, V = int(os.environ.get("QB", "16")), 196
cfg = syn.C2_WIDE
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.cuda()
w = m._weights(torch.device("cuda"))
imgs = [torch.rand(B, V, 512, 15, 15, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
st = [ops.img_attnpool(imgs[k], w["img"], 8, params=w["img_struct"], stages=1) for k in range(2)]
for mode in ("0", "1", "0", "1"):
    os.environ["PT_POOL_SINGLE"] = mode
    for k in range(2):
        ops.img_attnpool(imgs[k], w["img"], 8, params=w["img_struct"], stages=2, out=st[k][0], ws=st[k][1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(10):
        ops.img_attnpool(imgs[k & 1], w["img"], 8, params=w["img_struct"], stages=2, out=st[k & 1][0], ws=st[k & 1][1])
    e1.record(); torch.cuda.synchronize()
    print(f"PT_POOL_SINGLE={mode}: BACK stage {e0.elapsed_time(e1) / 10:.4f} ms per {B} scenes", flush=True)
