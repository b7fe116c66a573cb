"""Host-side boundary checks that need no GPU: C-ABI exports, state_dict layout, registry/config build, error behaviour."""
import ast
import ctypes
import json
import os
import re

import pytest
import torch

from oracle import ref_shim
from proxytransformation_b200 import MODELS, ProxyTransformationNormReverse, _lib, build_ext
from proxytransformation_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pt_preshape.h")
CFG_FIXTURE = os.path.join(ROOT, "tests", "golden", "preshape_cfg.json")
REF_CFG = "/root/reference/configs/grounding/proxy-tiblock33-gs12-wbias-ddr0.6-clip.py"


@pytest.fixture(scope="module")
def lib():
    build_ext.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pt_preshape.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared), "ctypes signature table out of sync with the header"


def test_library_loads_without_gpu_and_reports_version(lib):
    lib.pt_abi_version.restype = ctypes.c_int
    assert lib.pt_abi_version() == 1
    L = _lib.load()
    assert L.pt_minmax_ws_bytes(2, 100000) > 0
    assert L.pt_scatter_ws_bytes(2, 100000) >= 2 * 100000 * 4
    assert L.pt_proxy_block_ws_bytes(2, 256, 64, 256, 1024) > 0
    assert L.pt_img_attnpool_ws_bytes(8, 512, 225, 256, 8) > 0


def test_argument_errors_are_reported_without_touching_the_gpu():
    L = _lib.load()
    rc = L.pt_ball_query_firstk(None, None, 1, 1, 1, 1, 3.0, None, None, None)
    assert rc == -1 and b"null" in L.pt_last_error_string()
    rc = L.pt_cluster_dropout(None, None, 1, 64, 30, 45, 45, None, None, None, None, None, None)
    assert rc == -1 and b"n_drop" in L.pt_last_error_string()


@pytest.mark.parametrize("cfg", [syn.C1, syn.C2_WIDE, syn.C3], ids=lambda c: c.name)
def test_state_dict_layout_matches_spec_and_loads_strict(cfg):
    m = ProxyTransformationNormReverse(**cfg.module_kwargs())
    sd = m.state_dict()
    spec = syn.state_dict_spec(cfg)
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    m.load_state_dict(syn.make_state_dict(cfg, 3), strict=True)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_reference_checkpoint_loads_strict_both_ways():
    cfg = syn.C3
    ref = ref_shim.build_module(cfg.module_kwargs(), None)
    ours = ProxyTransformationNormReverse(**cfg.module_kwargs())
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in ref.parameters()) == 5792132


def test_constructor_defaults_match_reference_signature():
    import inspect
    sig = inspect.signature(ProxyTransformationNormReverse.__init__)
    want = dict(embed_dim=256, num_heads=8, n_points=100000, grid_size=4, text_blocks=1, img_blocks=1, dynamic_drop_radio=0.8,
                mlp_radio=4, qkv_bias=False, drop_rate=0.2, attn_drop_rate=0.2, drop_path_rate=0.2, num_sub=30,
                drop_radio=0.2, input_dim=512, img_spacial_dim=15)
    for k, v in want.items():
        assert sig.parameters[k].default == v, k
    assert list(sig.parameters)[1:19] == ["embed_dim", "num_heads", "n_points", "grid_size", "text_blocks", "img_blocks",
                                          "dynamic_drop_radio", "mlp_radio", "qkv_bias", "drop_rate", "attn_drop_rate",
                                          "drop_path_rate", "act_layer", "norm_layer", "num_sub", "drop_radio", "input_dim",
                                          "img_spacial_dim"]


def _load_py_config(path):
    """mmengine-free reader for the reference's python configs: executes `_base_` files first, then the file."""
    ns = {}
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "_base_":
            for rel in ast.literal_eval(node.value):
                ns.update(_load_py_config(os.path.normpath(os.path.join(os.path.dirname(path), rel))))
    exec(compile(src, path, "exec"), ns)
    return {k: v for k, v in ns.items() if not k.startswith("__")}


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not mounted (GPU box)")
def test_fixture_equals_shipped_config_preshape_dict():
    cfg = _load_py_config(REF_CFG)
    assert cfg["model"]["preshape"] == json.load(open(CFG_FIXTURE))


def test_registry_builds_module_from_shipped_config():
    """configs/grounding/proxy-tiblock33-gs12-wbias-ddr0.6-clip.py:41 (committed as a fixture) builds unchanged."""
    preshape = json.load(open(CFG_FIXTURE))
    assert preshape["type"] == "ProxyTransformationNormReverse"
    m = MODELS.build(preshape)
    assert isinstance(m, ProxyTransformationNormReverse)
    assert (m.num_cluster, m.real_cluster_num, m.keep1, m.num_sub) == (1728, 691, 1210, 30)
    assert len(m.textformer) == 3 and len(m.imgformer) == 3


def test_no_cpu_path_and_no_training_mode():
    cfg = syn.C1
    m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
    pts, td, img = syn.make_inputs(cfg, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(pts, td, img)
    # train() mode (batch-statistics kernels under no_grad, the autograd path otherwise) needs the device as well
    m.train()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(pts, td, img)
    m0 = ProxyTransformationNormReverse(**dict(cfg.module_kwargs(), drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0)).train()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m0(pts, td, img)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        m0(pts, td, img)


def test_oracle_is_not_imported_by_the_product_package():
    pkg = os.path.join(ROOT, "proxytransformation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "/root/reference" not in txt, f
