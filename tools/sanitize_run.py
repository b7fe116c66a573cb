#!/usr/bin/env python
"""One small forward through every sm_100a kernel (bf16 image features -> tensor-core pool kernel, tcgen05 GEMMs and
attention, geometry, scatter) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool synccheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proxytransformation_b200 import ProxyTransformationNormReverse, synthetic as syn

cfg = syn.C2_WIDE.replace(n_views=6, n_points=20000)
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.cuda()
pts, text_dict, img = syn.make_inputs(cfg, 2, first_scene=0, img_dtype=torch.bfloat16)
out = m([p.cuda() for p in pts], {k: v.cuda() for k, v in text_dict.items()}, img.cuda())
coords, feats = m.forward_sparse([p.cuda() for p in pts], {k: v.cuda() for k, v in text_dict.items()}, img.cuda(), 0.01)
torch.cuda.synchronize()
print("ok", [tuple(o.shape) for o in out], tuple(coords.shape))
# qkv_bias=True (bias inside the transposed-V epilogue of the QKV GEMM) and the N3 input-side kernel
cfgb = syn.C1.replace(qkv_bias=True)
mb = ProxyTransformationNormReverse(**cfgb.module_kwargs()).eval()
mb.load_state_dict(syn.make_state_dict(cfgb, 1))
mb = mb.cuda()
ptsb, tdb, imgb = syn.make_inputs(cfgb, 2, first_scene=3)
outb = mb([p.cuda() for p in ptsb], {k: v.cuda() for k, v in tdb.items()}, imgb.cuda())        # fp32 features: generic pool kernels
from proxytransformation_b200 import ops
views = [torch.rand(300 + 97 * v, 3) for v in range(5)]
ext = torch.eye(4).repeat(5, 1, 1)
ext[:, :3, 3] = torch.rand(5, 3)
agg = ops.aggregate_sample([v.cuda() for v in views], ext, torch.randint(0, sum(len(v) for v in views), (4096,)).cuda())
torch.cuda.synchronize()
print("ok", [tuple(o.shape) for o in outb], tuple(agg.shape))
# train() mode as a batch-statistics forward (bnstats.cu: fp64 statistics kernels + running-statistics update)
mt = ProxyTransformationNormReverse(**dict(syn.C1.module_kwargs(), drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0))
mt.load_state_dict(syn.make_state_dict(syn.C1, 2))
mt = mt.cuda().train()
ptst, tdt, imgt = syn.make_inputs(syn.C1, 2, first_scene=5)
with torch.no_grad():
    outt = mt([p.cuda() for p in ptst], {k: v.cuda() for k, v in tdt.items()}, imgt.cuda())
torch.cuda.synchronize()
print("ok", [tuple(o.shape) for o in outt], int(mt.text_trans_norm.num_batches_tracked))
