#!/usr/bin/env python
"""Step time of the benchmark workload (device-resident inputs, two alternating input sets) under the environment's overlap knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proxytransformation_b200 import ProxyTransformationNormReverse, synthetic as syn
B = 64
cfg = syn.C2_WIDE
dev = torch.device("cuda", 0)
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.to(dev)
sets = []
for sd in (1, 2):
    g = torch.Generator(device=dev).manual_seed(sd)
    P = torch.rand(B, cfg.n_points, 3, generator=g, device=dev) * torch.tensor(cfg.box, device=dev)
    text = torch.randn(B, cfg.n_text, cfg.embed_dim, generator=g, device=dev)
    mask = torch.ones(B, cfg.n_text, dtype=torch.uint8, device=dev)
    hw = cfg.img_spacial_dim
    img = (torch.relu(torch.randn(B, cfg.n_views, cfg.input_dim, hw, hw, generator=g, device=dev)) * 1.5).to(torch.bfloat16)
    sets.append((P, text, mask, img))
for i in range(30):
    m.forward_packed(*sets[i % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(40):
    m.forward_packed(*sets[i % 2])
e1.record()
torch.cuda.synchronize()
knobs = " ".join(f"{k}={os.environ[k]}" for k in ("PT_MEAN_CTAS", "PT_POOL_GRID", "PT_PARALLEL_BRANCH_ROWS", "PT_OVERLAP_IMG", "PT_ATTN_FORM") if k in os.environ)
print(f"{knobs or 'defaults'}: {e0.elapsed_time(e1) / 40:.4f} ms/step")
