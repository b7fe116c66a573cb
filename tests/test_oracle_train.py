"""Training-mode side of the oracle (oracle first, CUDA later) against the fixture captured from the
unmodified reference in train() mode (tests/golden/make_golden_train.py): forward with BatchNorm batch statistics, the
running statistics after the step, and the parameter gradients of a fixed scalar loss through torch autograd of the
oracle's own forward.  Stochastic layers are at rate 0 on both sides.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import preshape_oracle as po
from oracle import ref_shim
from proxytransformation_b200 import synthetic as syn
from tests.golden_cases import FULL_GRAD_KEYS, GOLDEN_DIR, TRAIN_CASE, train_loss_weights


def _oracle_train_step():
    cfg, batch, first, wseed = TRAIN_CASE
    sd = syn.make_state_dict(cfg, wseed)
    leaves = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v) for k, v in sd.items()}
    pts, text_dict, img = syn.make_inputs(cfg, batch, first)
    train = {}
    nt = torch.get_num_threads()
    torch.set_num_threads(1)                # same accumulation order as the fixture (see tests/test_oracle_golden.py::_run)
    try:
        out = po.forward(leaves, pts, text_dict, img, grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio,
                         text_blocks=cfg.text_blocks, img_blocks=cfg.img_blocks, num_sub=cfg.num_sub, num_heads=cfg.num_heads,
                         train=train)
        loss = sum((o * r).sum() for o, r in zip(out, train_loss_weights([o.shape[0] for o in out])))
        loss.backward()
    finally:
        torch.set_num_threads(nt)
    return cfg, leaves, out, loss, train


def test_train_mode_forward_and_running_statistics_match_reference():
    g = dict(np.load(os.path.join(GOLDEN_DIR, "c1_train.npz"), allow_pickle=False))
    cfg, leaves, out, loss, train = _oracle_train_step()
    np.testing.assert_array_equal(np.array([o.shape[0] for o in out]), g["out_counts"])
    for b, o in enumerate(out):
        np.testing.assert_allclose(o.detach().numpy(), g[f"out_{b}"], rtol=0, atol=1e-4)      # the path's coordinate bar
    assert abs(float(loss) - float(g["loss"])) <= 1e-3 * max(1.0, abs(float(g["loss"])))
    bn_keys = [k[3:] for k in g if k.startswith("bn/")]
    assert len(bn_keys) == 4 * 3                                                             # four BatchNorm layers
    for k in bn_keys:
        got = train["bn"][k]
        if k.endswith("num_batches_tracked"):
            assert int(got) == int(g["bn/" + k])
        else:
            np.testing.assert_allclose(got.numpy(), g["bn/" + k], rtol=1e-5, atol=1e-6, err_msg=k)


def test_parameter_gradients_match_reference_autograd():
    g = dict(np.load(os.path.join(GOLDEN_DIR, "c1_train.npz"), allow_pickle=False))
    cfg, leaves, out, loss, train = _oracle_train_step()
    names, norms = [str(n) for n in g["grad_names"]], g["grad_norms"]
    checked = 0
    for name, want in zip(names, norms):
        grad = leaves[name].grad
        if want < 0:                                  # not on the path in the reference (blocks before the last one)
            assert grad is None or float(grad.abs().max()) == 0.0, name
            continue
        assert grad is not None, name
        got = float(grad.double().norm())
        # 2e-4 absolute: the bias of a conv that feeds a batch-statistics BatchNorm has an analytically zero gradient; both
        # sides hold rounding noise of ~1e-4 there
        assert abs(got - want) <= 2e-3 * want + 2e-4, (name, got, want)
        checked += 1
    assert checked == int((norms >= 0).sum()) and checked >= 60
    for k in sorted(FULL_GRAD_KEYS(cfg)):
        want = g["grad/" + k]
        np.testing.assert_allclose(leaves[k].grad.numpy(), want, rtol=0, atol=2e-3 * np.abs(want).max() + 1e-5, err_msg=k)   # floor: analytically zero gradients hold ~1e-6 of noise


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is only present in the build container")
def test_live_reference_train_mode_outputs():
    """Live cross-check where the reference tree exists: train() forward of the reference == oracle (same pins)."""
    cfg, batch, first, wseed = TRAIN_CASE
    sd = syn.make_state_dict(cfg, wseed)
    pts, text_dict, img = syn.make_inputs(cfg, batch, first)
    net = ref_shim.build_module(dict(cfg.module_kwargs(), drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0), sd, pinned=True)
    net.train()
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        with torch.no_grad():
            want = net(pts, text_dict, img)
            got = po.forward(sd, pts, text_dict, img, grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio,
                             text_blocks=cfg.text_blocks, img_blocks=cfg.img_blocks, num_sub=cfg.num_sub,
                             num_heads=cfg.num_heads, train={})
    finally:
        torch.set_num_threads(nt)
    assert [o.shape for o in got] == [o.shape for o in want]
    for a, b in zip(got, want):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=0, atol=1e-4)
