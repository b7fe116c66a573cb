#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 3xBF16 GEMM on the ProxyBlock / image-pool shapes (CUDA events, warm L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proxytransformation_b200 import ops

from proxytransformation_b200 import _lib

def t(fn, n=20):
    """mean device time per launch (events around the kernel itself, via pt_profile_*)"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(n): fn()
    torch.cuda.synchronize()
    pr = _lib.profile_read()
    _lib.profile_enable(False)
    return sum(v[0] for v in pr.values()) / n * 1e3

dev = "cuda"
shapes = [("qkv", 16384, 768, 256, 0), ("proj", 16384, 256, 256, 0), ("fc1", 16384, 1024, 256, 1), ("fc2", 16384, 256, 1024, 0),
          ("pp_txt", 4096, 256, 256, 0), ("pp_img", 12544, 256, 256, 0), ("img_q", 12544, 256, 512, 0), ("bigK", 16384, 256, 4096, 0), ("bigK2", 16384, 1024, 4096, 0)]
for name, M, N, K, act in shapes:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5
    As, Ws = ops.split_bf16(A), ops.split_bf16(W)
    bias = torch.randn(N, device=dev); C = torch.empty(M, N, device=dev)
    res = []
    for bn in (64, 128, 256):
        if bn > N: continue
        us = t(lambda: ops.gemm_tc(As, Ws, M, N, K, bias=bias, act=act, C=C, ldc=N, bn=bn))
        res.append(f"bn{bn}: {us:7.1f} us ({3 * 2 * M * N * K / us / 1e6:6.1f} TF/s-equiv)")
    us_split = t(lambda: ops.split_bf16(A))
    print(f"{name:7s} M={M} N={N} K={K} act={act}  " + "  ".join(res) + f"   split(A) {us_split:.1f} us")
