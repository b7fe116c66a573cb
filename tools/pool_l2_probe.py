"""Probe: how fast is the pool kernel (pass B) per view when its loads hit L2 instead of HBM, and how does it scale with the
number of persistent CTAs?  Times img_pool_mma_kernel alone (CUDA events through pt_profile_*).
  L2-resident: QB scenes x 196 views small enough for L2 (QB=1: 45 MB), the SAME input every launch.
  HBM: QB=64 (2.9 GB), two alternating inputs.
Usage (GPU box): python tools/pool_l2_probe.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proxytransformation_b200 import ProxyTransformationNormReverse, ops, _lib, synthetic as syn
V = 196
cfg = syn.C2_WIDE
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.cuda()
w = m._weights(torch.device("cuda"))
sms = torch.cuda.get_device_properties(0).multi_processor_count


def run(B, grid, nin, debug=0, reps=12):
    os.environ["PT_POOL_GRID"] = str(grid)
    os.environ["PT_POOL_DEBUG"] = str(debug)
    imgs = [(torch.relu(torch.randn(B, V, 512, 15, 15, device="cuda")) * 1.5).bfloat16() for _ in range(nin)]
    st = [ops.img_attnpool(imgs[k], w["img"], 8, params=w["img_struct"], stages=1) for k in range(nin)]
    for k in range(3):
        ops.img_attnpool(imgs[k % nin], w["img"], 8, params=w["img_struct"], stages=2, out=st[k % nin][0], ws=st[k % nin][1])
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for k in range(reps):
        ops.img_attnpool(imgs[k % nin], w["img"], 8, params=w["img_struct"], stages=2, out=st[k % nin][0], ws=st[k % nin][1])
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    ms = prof["img_pool"][0] / prof["img_pool"][1]
    views_per_cta = B * V / grid
    cyc = ms * 1e-3 * 1.965e9 / views_per_cta
    print(f"B={B:3d} grid={grid:3d} inputs={nin} debug={debug:2d}: pool {ms:.4f} ms, {views_per_cta:.1f} views/CTA, "
          f"{cyc / 1e3:.2f} k cycles per view per CTA, {B * V * 230400 / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    del imgs, st


for B, grid in ((1, 49), (1, 98), (2, 98), (2, 131), (3, 147)):
    run(B, grid, 1)                       # L2-resident (B=2: 90 MB, B=3: 135 MB - partially)
for grid in (49, 98, 116, 132, sms):
    run(64, grid, 2)                      # HBM
run(64, sms, 2, debug=2)                  # no slab traffic at all (16-byte loads): the compute floor of the schedule
run(64, 116, 2, debug=2)
