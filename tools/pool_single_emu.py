#!/usr/bin/env python
"""Executable specification (numpy, CPU) of the SINGLE-PASS image-pool schedule of DESIGN.md §7a — token windows in
u = token + class coordinates, per-class partial scores in a two-window ring, online softmax on the tokens a window
completes, lagging weighted sums with shifted probability fragments, final normalisation — checked against a float64
evaluation of the same algebra.  It follows the planned kernel phase by phase and uses its shared-memory layouts and
index formulas (class-major window rows, XOR-swizzled 16-byte chunks, two parity copies of the probabilities), so the
formulas can be validated without a GPU.  It is NOT on any product path.

    python tools/pool_single_emu.py [seed] [chunks per window: 4 (kernel as built) | 8 (128-byte pieces, the proposed next step)]
"""
import sys

import numpy as np

C, HW, HEADS, HD = 512, 225, 8, 32
NWIN, WCH = 8, 4                 # windows per view, 16-byte chunks (8 u-columns) per window: the kernel's shape (64-byte pieces)
NCHUNK = 29                      # aligned chunks per channel row (232 >= 225 + 7)
PP = 64                          # probability row pitch (elements): token slot i of the window sits at element i + 8 + copy


def set_shape(wch):
    """wch = 4: 64-byte pieces, 8 windows of 32 u-columns (img_pool_single_kernel); wch = 8: 128-byte pieces, 4 windows of 64."""
    global NWIN, WCH, PP
    WCH = wch
    NWIN = -(-32 // wch)             # 29 chunks -> 8 windows of 4 or 4 windows of 8
    PP = 8 * wch + 32                # slots + margins (8 + copy in front, >= 16 + 7 behind)


def bf16_round(x):
    """fp32 -> nearest-even bf16 (as fp32)."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def split(x):
    hi = bf16_round(x)
    return hi, bf16_round(np.asarray(x, np.float32) - hi)


def load_window(view_bytes, w):
    """cp.async stage: window w of a view -> shared-memory image [512 rows][4 chunks][8 elements] (physical order).
    Row of channel c = 8 r + s is s*64 + r; chunk ch of the window (u = 8 WCH w + 8 ch ..) lands at position
    ch ^ ((row >> 1) & (WCH - 1)); source address = 448 s + 3600 r + 16 (WCH w + ch), zero fill for chunks >= 29."""
    win = np.zeros((C, WCH, 8), np.float32)
    for row in range(C):
        s, r = row >> 6, row & 63
        for ch in range(WCH):
            gch = WCH * w + ch
            if gch >= NCHUNK:
                continue
            off = 448 * s + 3600 * r + 16 * gch
            win[row, ch ^ ((row >> 1) & (WCH - 1))] = view_bytes[off // 2: off // 2 + 8]
    return win


def chunk_of(win, row, ch):
    return win[row, ch ^ ((row >> 1) & (WCH - 1))]


def emulate_view(X, w_eff, cterm, xbar, scale):
    """X (512,225) bf16-valued fp32, w_eff (8,512) fp32, cterm (8,226), xbar (512,) -> (probabilities (8,226), Y (8,512))."""
    view = np.zeros(C * HW + 64, np.float32)
    view[:C * HW] = X.reshape(-1)                       # element index = 225 c + t ; byte address = 2 * index
    w_hi, w_lo = split(w_eff)
    m = np.zeros(HEADS, np.float32); l = np.zeros(HEADS, np.float32)
    sv0 = np.zeros(HEADS, np.float32)
    for h in range(HEADS):                              # mean token (attention token 0): opens the running softmax
        sv0[h] = scale * (np.float32((w_hi[h].astype(np.float64) + w_lo[h]) @ xbar.astype(np.float64)) + cterm[h, 0])
        m[h], l[h] = sv0[h], 1.0
    svbuf = np.full((HEADS, 232), -np.inf, np.float32)  # raw scaled scores of the spatial tokens
    Y = np.zeros((2, HEADS, C), np.float32)             # accumulators: [hi part | lo part of the probabilities][head][channel]
    W = 8 * WCH                                         # u-columns (= token slots) per window
    part = np.zeros((2, 8, HEADS, W), np.float32)       # [window parity][class][head][u_local]
    wins = {}
    for w in range(NWIN):
        wins[w] = load_window(view, w)
        win = wins[w]
        # ---- step 1: scores, warp = (class s, chunk pair cp); k-step j covers class rows r = 16 j .. 16 j + 15
        for s in range(8):
            for cp in range(WCH // 2):
                acc = np.zeros((16, 16), np.float64)                     # rows: 8 heads hi, 8 heads lo; cols: 2 chunks x 8
                for j in range(4):
                    rows = s * 64 + 16 * j + np.arange(16)
                    chan = 8 * (16 * j + np.arange(16)) + s
                    A = np.concatenate([w_hi[:, chan], w_lo[:, chan]], 0).astype(np.float64)      # (16, 16)
                    B = np.stack([np.concatenate([chunk_of(win, row, 2 * cp), chunk_of(win, row, 2 * cp + 1)]) for row in rows])
                    acc += A @ B.astype(np.float64)
                part[w & 1, s, :, 16 * cp:16 * cp + 16] = (acc[:8] + acc[8:]).astype(np.float32)
        # ---- step 2: softmax step, warp(s) = head, lane i = token slot: t = W w - 7 + i
        P = np.zeros((HEADS, W), np.float32)
        alpha = np.ones(HEADS, np.float32)
        for h in range(HEADS):
            sv = np.full(W, -np.inf, np.float32)
            for i in range(W):
                t = W * w - 7 + i
                if t < 0 or t >= HW:
                    continue
                acc = np.float32(0.0)
                for s in range(8):                                        # fixed order: bit-reproducible
                    ul = i - 7 + s                                        # u_local of class s for this token
                    acc = np.float32(acc + (part[(w - 1) & 1, s, h, ul + W] if ul < 0 else part[w & 1, s, h, ul]))
                sv[i] = scale * (acc + cterm[h, t + 1])
                svbuf[h, t] = sv[i]
            m_new = max(m[h], sv.max())
            alpha[h] = np.exp(np.float32(m[h] - m_new))
            P[h] = np.where(np.isfinite(sv), np.exp(sv - m_new), 0.0).astype(np.float32)
            l[h] = l[h] * alpha[h] + P[h].sum(dtype=np.float32)
            m[h] = m_new
        p_hi, p_lo = split(P)
        # shared-memory image of the probabilities: [copy][hi|lo][head][PP], slot i at element i + 8 + copy, zero margins
        pbuf = np.zeros((2, 2, HEADS, PP), np.float32)
        for cpy in range(2):
            pbuf[cpy, 0, :, 8 + cpy:8 + W + cpy] = p_hi
            pbuf[cpy, 1, :, 8 + cpy:8 + W + cpy] = p_lo
        # ---- step 3: weighted sums, warp = (class s, channel half hf); chunks -1 .. WCH-1 (+ one zero chunk), k-steps of 16 u
        Y *= alpha[None, :, None]
        for s in range(8):
            cpy = (s + 1) & 1
            for hf in range(2):
                for ks in range(WCH // 2 + 1):
                    # A fragment: u_local = 8 (2 ks - 1) + kk, token slot i = u_local + 7 - s, element = i + 8 + cpy (pairs aligned)
                    base = 16 * ks - 8 + 7 - s + 8 + cpy
                    assert base % 2 == 0 and base >= 0 and base + 16 <= PP
                    A = np.concatenate([pbuf[cpy, 0, :, base:base + 16], pbuf[cpy, 1, :, base:base + 16]], 0).astype(np.float64)   # (16 rows, 16 k)
                    for nt in range(4):
                        rows = s * 64 + 32 * hf + 8 * nt + np.arange(8)
                        Bt = []
                        for row in rows:                                   # B^T rows = channels, 16 k = two chunks
                            lo_ch, hi_ch = 2 * ks - 1, 2 * ks
                            c_lo = chunk_of(wins[w - 1], row, WCH - 1) if lo_ch < 0 and w > 0 else chunk_of(win, row, max(lo_ch, 0))
                            c_hi = chunk_of(win, row, min(hi_ch, WCH - 1))  # the chunk past the window: probabilities are zero there
                            Bt.append(np.concatenate([c_lo, c_hi]))
                        D = A @ np.stack(Bt).astype(np.float64).T          # (16, 8)
                        chan = 8 * (32 * hf + 8 * nt + np.arange(8)) + s
                        Y[0][:, chan] += D[:8].astype(np.float32)
                        Y[1][:, chan] += D[8:].astype(np.float32)
        wins.pop(w - 1, None)                                              # slot of window w-1 is free now
    # ---- end of view: normalise, mean-token term, final probabilities from the raw scores
    p0 = np.exp(sv0 - m) / l
    Yf = (Y[0] + Y[1]) / l[:, None] + p0[:, None] * xbar[None, :]
    probs = np.zeros((HEADS, HW + 1), np.float32)
    probs[:, 0] = p0
    probs[:, 1:] = np.exp(svbuf[:, :HW] - m[:, None]) / l[:, None]
    return probs, Yf


def reference_view(X, w_eff, cterm, xbar, scale):
    X, w_eff, xbar = X.astype(np.float64), w_eff.astype(np.float64), xbar.astype(np.float64)
    sc = np.concatenate([(w_eff @ xbar)[:, None], w_eff @ X], 1)
    sv = scale * (sc + cterm.astype(np.float64))
    P = np.exp(sv - sv.max(1, keepdims=True)); P /= P.sum(1, keepdims=True)
    return P, P[:, 1:] @ X.T + P[:, :1] * xbar[None, :]


def main():
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    set_shape(int(sys.argv[2]) if len(sys.argv) > 2 else 4)
    X = bf16_round(np.maximum(rng.standard_normal((C, HW)), 0) * 1.5)
    w_eff = (rng.standard_normal((HEADS, C)) * 0.08).astype(np.float32)
    cterm = (rng.standard_normal((HEADS, HW + 1)) * 0.5).astype(np.float32)
    xbar = X.mean(1).astype(np.float32)
    scale = np.float32(HD ** -0.5)
    probs, Y = emulate_view(X, w_eff, cterm, xbar, scale)
    P_ref, Y_ref = reference_view(X, w_eff, cterm, xbar, float(scale))
    print(f"probabilities: max abs err {np.abs(probs - P_ref).max():.3e} (max {P_ref.max():.3e})")
    print(f"weighted sums: max rel err {np.abs(Y - Y_ref).max() / np.abs(Y_ref).max():.3e}")
    ok = np.abs(probs - P_ref).max() < 1e-6 and np.abs(Y - Y_ref).max() / np.abs(Y_ref).max() < 2e-5
    print("OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
