"""bench.py's output contract: ONE JSON line on stdout with the keys the driver reads, for both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    """--impl reference runs the reference's CPU path on the host cores (no GPU needed): the reference's own module when its
    bytecode is present under oracle/_ref (kind "reference"), else the oracle port (kind "port")."""
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--n-points", "20000")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "scenes/s" and d["higher_is_better"] is True and d["value"] > 0
    from oracle import ref_shim
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C2-wide")


@pytest.mark.gpu
def test_b200_arm_line_on_gpu():
    d = _run("--steps", "3", "--warmup", "3", "--batch", "16", "--cpu-scenes", "1")
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "core_region", "kernel_breakdown", "image_stage_roofline", "c4_strong",
                        "checks", "c1", "c3"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["value"] > 0 and d["gpu_launches"] > 0
    r = d["roofline"]
    # the dominant kernel is picked from the live profile: the tag with the largest total time
    top = max(d["kernel_breakdown"].items(), key=lambda kv: kv[1]["ms_per_step"])[0]
    assert r["kernel"] == top and r["bound"] in ("hbm", "tensor", "latency")
    if r["bound"] != "latency":
        assert 0 < r["frac"] < 1.2 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    st = d["image_stage_roofline"]
    assert st["kernels"] == ["img_mean", "gemm_img_3xbf16", "img_pool"] and 0 < st["frac"] < 1.0
    c = d["checks"]                                   # the timed inputs against the oracle: the parity bar of north_star
    assert c["idx_equal"] is True and c["count_equal"] is True and c["max_coord_err"] < 1e-4 and c["img_proxy_max_err"] < 6e-5
    assert d["c4_strong"]["scenes_total"] == 16 or d["c4_strong"]["scenes_total"] == 64
    assert d["c3"]["clusters"] == 691 and d["c3"]["ms_per_forward"] > 0 and d["c1"]["clusters"] == 16
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 4 * 100000 * 12 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] > 0
    assert d["config"]["workload"] == "C2-wide" and d["config"]["scenes_per_gpu_per_step"] == 16
