"""A/B timing of the image stage's BACK half (pool kernel + value GEMMs + LayerNorm) and of the pool kernel alone (CUDA events
through pt_profile_*): tcgen05 kernel (PT_POOL_KERNEL=umma) vs mma.sync kernel, at bench size (QB scenes x 196 views, >> L2).
Usage (GPU box): [QB=64] python tools/pool_ab.py [kernels...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proxytransformation_b200 import ProxyTransformationNormReverse, ops, _lib, synthetic as syn
B, V = int(os.environ.get("QB", "64")), 196
cfg = syn.C2_WIDE
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.cuda()
imgs = [(torch.relu(torch.randn(B, V, 512, 15, 15, device="cuda")) * 1.5).bfloat16() for _ in range(2)]
kernels = [k for k in sys.argv[1:] if k in ("mma", "umma")] or ["mma", "umma", "mma", "umma"]
for kern in kernels:
    os.environ["PT_POOL_KERNEL"] = kern
    w = m._weights(torch.device("cuda"))
    st = [ops.img_attnpool(imgs[k], w["img"], 8, params=w["img_struct"], stages=1) for k in range(2)]
    for k in range(4):
        ops.img_attnpool(imgs[k & 1], w["img"], 8, params=w["img_struct"], stages=2, out=st[k & 1][0], ws=st[k & 1][1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(20):
        ops.img_attnpool(imgs[k & 1], w["img"], 8, params=w["img_struct"], stages=2, out=st[k & 1][0], ws=st[k & 1][1])
    e1.record(); torch.cuda.synchronize()
    _lib.profile_enable(True)
    for k in range(10):
        ops.img_attnpool(imgs[k & 1], w["img"], 8, params=w["img_struct"], stages=2, out=st[k & 1][0], ws=st[k & 1][1])
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    pool_ms = prof["img_pool"][0] / prof["img_pool"][1]
    gbs = B * V * 230400 / pool_ms / 1e6
    print(f"{kern}: BACK stage {e0.elapsed_time(e1) / 20:.4f} ms, pool kernel {pool_ms:.4f} ms = {gbs:.0f} GB/s algorithmic, per {B} scenes", flush=True)
