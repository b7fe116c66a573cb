#!/usr/bin/env python
"""Per-CTA timeline of one tcgen05 GEMM launch (library built with PT_NVCC_DEFINES=-DPT_GEMM_DBG): kernel entry, set-up done, first
operand stage landed, first tile's MMAs committed, first accumulator seen by the epilogue, first epilogue done, last accumulator seen,
kernel exit — in microseconds after the earliest CTA entry (globaltimer)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from proxytransformation_b200 import ops, _lib

lib = _lib.load()
fn = lib.pt_debug_gemm_timeline
fn.restype = ctypes.c_int
names = ["entry", "setup", "1st full", "1st commit", "1st acc", "1st epi end", "last acc", "exit"]
shapes = [("qkv", 16384, 768, 256, 0, 256), ("proj", 16384, 256, 256, 0, 256), ("fc1", 16384, 1024, 256, 1, 256), ("fc2", 16384, 256, 1024, 0, 256), ("pp_txt", 4096, 256, 256, 0, 128)]
for name, M, N, K, act, bn in shapes:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    As, Ws = ops.split_bf16(A), ops.split_bf16(W)
    bias = torch.randn(N, device="cuda"); C = torch.empty(M, N, device="cuda")
    for _ in range(5): ops.gemm_tc(As, Ws, M, N, K, bias=bias, act=act, C=C, ldc=N, bn=bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # queue a few launches back to back so that the last one starts behind a busy GPU (no host gap in its timing)
    for _ in range(6): ops.gemm_tc(As, Ws, M, N, K, bias=bias, act=act, C=C, ldc=N, bn=bn)
    e0.record(); ops.gemm_tc(As, Ws, M, N, K, bias=bias, act=act, C=C, ldc=N, bn=bn); e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (160 * 8))()
    assert fn(buf) == 0
    tl = np.frombuffer(buf, dtype=np.uint64).reshape(160, 8).astype(np.int64)
    tiles = -(-M // 128) * -(-N // bn)
    ncta = min(tiles, 148)
    tl = tl[:ncta]
    t0 = tl[:, 0].min()
    rel = (tl - t0) / 1e3
    print(f"{name} M={M} N={N} K={K} bn={bn}: {tiles} tiles on {ncta} CTAs, event time {e0.elapsed_time(e1) * 1e3:.1f} us")
    for i, nm in enumerate(names):
        print(f"    {nm:12s} min {rel[:, i].min():7.2f}  median {np.median(rel[:, i]):7.2f}  max {rel[:, i].max():7.2f} us")
