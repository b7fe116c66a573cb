#!/usr/bin/env python
"""Stand-alone check of the tcgen05 attention core against an fp64 torch evaluation of :225-252."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proxytransformation_b200 import ops

def ref(q, k, v, pt, mask, heads):
    B, n, c = q.shape; l = pt.shape[1]; hd = c // heads; scale = hd ** -0.5
    f = lambda t, m: t.double().reshape(B, m, heads, hd).permute(0, 2, 1, 3)
    Q, K, V, P = f(q, n), f(k, n), f(v, n), f(pt, l)
    a1 = torch.softmax((P * scale) @ K.transpose(-1, -2), -1)
    pv = a1 @ V
    s2 = (Q * scale) @ P.transpose(-1, -2)
    if mask is not None:
        s2 = s2.masked_fill((mask == 0)[:, None, None, :], -1e9)
    o = torch.softmax(s2, -1) @ pv
    return o.permute(0, 2, 1, 3).reshape(B, n, c)

torch.manual_seed(0)
worst = 0.0
for (B, n, l, masked) in [(1, 16, 16, False), (2, 256, 64, True), (2, 256, 196, False), (3, 128, 50, True), (1, 64, 33, True), (2, 200, 256, False)]:
    c, heads = 256, 8
    q, k, v = (torch.randn(B, n, c, device="cuda") * 1.5 for _ in range(3))
    pt = torch.randn(B, l, c, device="cuda")
    mask = None
    if masked:
        mask = torch.ones(B, l, dtype=torch.uint8, device="cuda")
        for b in range(B):
            mask[b, l - 1 - 3 * b:] = 0
    o = ops.proxy_attention_tc(q, k, v, pt, mask, heads)
    torch.cuda.synchronize()
    err = (o.double() - ref(q, k, v, pt, mask, heads)).abs().max().item()
    worst = max(worst, err)
    print(f"B={B} n={n} l={l} masked={masked}: max |err| = {err:.3e}")
print("worst", worst)
sys.exit(0 if worst < 2e-5 else 1)
