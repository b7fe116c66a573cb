"""The executable specification of the single-pass image-pool schedule (tools/pool_single_emu.py: window layout, class shifts,
completed-token ranges, shifted probability fragments, online softmax, final normalisation) against a float64 evaluation of
the same algebra.  CPU only; the CUDA kernel transcribed from it (PT_POOL_SINGLE=1) is checked on the GPU by
tools/pool_check.py."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emu():
    spec = importlib.util.spec_from_file_location("pool_single_emu", os.path.join(ROOT, "tools", "pool_single_emu.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("wch", [4, 8], ids=["64B-windows(kernel)", "128B-windows(next)"])
@pytest.mark.parametrize("seed", [0, 7])
def test_single_pass_schedule_matches_float64(seed, wch):
    e = _emu()
    e.set_shape(wch)
    rng = np.random.default_rng(seed)
    X = e.bf16_round(np.maximum(rng.standard_normal((e.C, e.HW)), 0) * 1.5)
    w_eff = (rng.standard_normal((e.HEADS, e.C)) * 0.08).astype(np.float32)
    cterm = (rng.standard_normal((e.HEADS, e.HW + 1)) * 0.5).astype(np.float32)
    xbar = X.mean(1).astype(np.float32)
    scale = np.float32(e.HD ** -0.5)
    probs, Y = e.emulate_view(X, w_eff, cterm, xbar, scale)
    P_ref, Y_ref = e.reference_view(X, w_eff, cterm, xbar, float(scale))
    assert np.abs(probs - P_ref).max() < 1e-6
    assert np.abs(Y - Y_ref).max() / np.abs(Y_ref).max() < 2e-5
    np.testing.assert_allclose(probs.sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("wch", [4, 8])
def test_window_loader_covers_every_token_exactly_once(wch):
    """Every (channel, token) element appears at u = token + class in exactly one window chunk; chunks >= 29 are zero."""
    e = _emu()
    e.set_shape(wch)
    view = np.zeros(e.C * e.HW + 64, np.float32)
    view[:e.C * e.HW] = np.arange(1, e.C * e.HW + 1, dtype=np.float32)      # unique non-zero tags
    seen = np.zeros(e.C * e.HW, np.int32)
    for w in range(e.NWIN):
        win = e.load_window(view, w)
        for row in range(e.C):
            s, r = row >> 6, row & 63
            c = 8 * r + s
            for ch in range(e.WCH):
                vals = e.chunk_of(win, row, ch)
                for k in range(8):
                    t = 8 * e.WCH * w + 8 * ch + k - s
                    if 0 <= t < e.HW and e.WCH * w + ch < e.NCHUNK:
                        assert vals[k] == c * e.HW + t + 1
                        seen[c * e.HW + t] += 1
    assert (seen == 1).all()
