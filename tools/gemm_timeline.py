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


def show(name, tiles):
    buf = (ctypes.c_ulonglong * (160 * 8))()
    assert fn(buf) == 0
    tl = np.frombuffer(buf, dtype=np.uint64).reshape(160, 8).astype(np.int64)[:min(tiles, 148)]
    rel = (tl - tl[:, 0].min()) / 1e3
    print(f"{name}: {tiles} tiles")
    for i, nm in enumerate(names):
        print(f"    {nm:12s} min {rel[:, i].min():7.2f}  median {np.median(rel[:, i]):7.2f}  max {rel[:, i].max():7.2f} us")


# the batched image-stage projections (12544 views = 64 scenes x 196): G2 (w_eff planes, K = 64), G4 (values, N = 32 per head)
BV, HEADS, C, YA, HD = 12544, 8, 512, 256, 32
q_s = ops.split_bf16(torch.randn(BV, 256, device="cuda"))
wk_s = ops.split_bf16(torch.randn(HEADS * C, 64, device="cuda") / 8)
wpl = torch.empty(BV, 2, HEADS * C, dtype=torch.bfloat16, device="cuda")
g2 = lambda: ops.gemm_tc(q_s, wk_s, BV, C, 64, batch=HEADS, a_koff_z=HD, w_row_z=C, c_split=wpl.view(-1)[:2 * HEADS * C].view(2, HEADS * C),
                         ldcs=2 * HEADS * C, cs_off_z=C)
ya_s = ops.split_bf16(torch.randn(BV, HEADS * YA, device="cuda"))
wv_s = ops.split_bf16(torch.randn(HEADS * HD, YA, device="cuda") / 16)
z_s = torch.empty(2, BV, HEADS * HD, dtype=torch.bfloat16, device="cuda")
g4 = lambda: ops.gemm_tc(ya_s, wv_s, BV, HD, YA, batch=HEADS, a_koff_z=YA, w_row_z=HD, c_split=z_s, ldcs=HEADS * HD, cs_off_z=HD)
for name, f, tiles in (("G2 w_eff 12544 x 512 x 64 x 8 heads", g2, 98 * 2 * 8), ("G4 values 12544 x 32 x 256 x 8 heads", g4, 98 * 8)):
    for _ in range(8):
        f()
    torch.cuda.synchronize()
    show(name, tiles)
