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
# round 2: fp16 features (half instantiation of the tcgen05 pool kernel), the mma.sync pool kernel, the shipped config's cluster count
# (691: streamed tcgen05 attention with per-scene padded V^T planes) and an odd point count (element paths of min/max / compaction)
pts16, td16, img16 = syn.make_inputs(cfg, 2, first_scene=9, img_dtype=torch.float16)
out16 = m([p.cuda() for p in pts16], {k: v.cuda() for k, v in td16.items()}, img16.cuda())
os.environ["PT_POOL_KERNEL"] = "mma"
m2 = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m2.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m2 = m2.cuda()
outm = m2([p.cuda() for p in pts], {k: v.cuda() for k, v in text_dict.items()}, img.cuda())
os.environ.pop("PT_POOL_KERNEL")
cfg3 = syn.C3.replace(n_views=3, n_points=30001)
m3 = ProxyTransformationNormReverse(**cfg3.module_kwargs()).eval()
m3.load_state_dict(syn.make_state_dict(cfg3, 4, bf16_round=True))
m3 = m3.cuda()
pts3, td3, img3 = syn.make_inputs(cfg3, 3, first_scene=11, img_dtype=torch.bfloat16)
out3 = m3([p.cuda() for p in pts3], {k: v.cuda() for k, v in td3.items()}, img3.cuda())
torch.cuda.synchronize()
print("ok", [tuple(o.shape) for o in out16], [tuple(o.shape) for o in outm], [tuple(o.shape) for o in out3])
# round 2 (late): the 8-warp form of the tcgen05 attention (two CTAs per SM; taken with more than 148 (scene, head) pairs), the pipelined
# GEMM epilogue with GELU / residual / transposed-V outputs (inside the blocks above) and the two-step ball-query loop (above)
g = torch.Generator().manual_seed(3)
q, k, v = (torch.randn(19, 200, 256, generator=g).cuda() for _ in range(3))
pt_ = torch.randn(19, 50, 256, generator=g).cuda()
mk = torch.ones(19, 50, dtype=torch.uint8)
mk[:, 40:] = 0
oa = ops.proxy_attention_tc(q, k, v, pt_, mk.cuda(), 8)
torch.cuda.synchronize()
print("ok", tuple(oa.shape))
