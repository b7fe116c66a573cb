#!/usr/bin/env python
"""Forward time of one BASELINE.json workload at a given batch (default: config 3, the shipped EmbodiedScan grounding config
gs=12 / ddr=0.6 / 3+3 blocks / num_sub=30 with batch 4, 50 views, 32 text tokens), device-resident inputs:
wall time per forward() as a caller sees it (host launch overhead included, one D2H of the counts), GPU time of the same
call (CUDA events) and the per-kernel breakdown; optional oracle-port CPU time beside it.  One JSON line.
    python tools/bench_config.py [--config c3|c3_wide|c2_wide|c2_room|c1] [--batch 4] [--iters 30] [--cpu-scenes 2]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proxytransformation_b200 import ProxyTransformationNormReverse, _lib, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--img-dtype", default="bf16", choices=["bf16", "f32"])
ap.add_argument("--cpu-scenes", type=int, default=0)
ap.add_argument("--graphs", action="store_true", help="replay the forward as one CUDA graph (module.cuda_graphs)")
a = ap.parse_args()
cfg = {"c3": syn.C3, "c3_wide": syn.C3_WIDE, "c2_wide": syn.C2_WIDE, "c2_room": syn.C2_ROOM, "c1": syn.C1}[a.config]
dev = torch.device("cuda:0")
dt = torch.bfloat16 if a.img_dtype == "bf16" else torch.float32
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True), strict=True)
m = m.to(dev)
m.cuda_graphs = a.graphs
sets = []
for k in range(2):
    pts, td, img = syn.make_inputs(cfg, a.batch, first_scene=100 * k, img_dtype=dt)
    sets.append(([p.to(dev) for p in pts], {n: v.to(dev) for n, v in td.items()}, img.to(dev)))
for k in range(4):
    m(*sets[k % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for k in range(a.iters):
    out = m(*sets[k % 2])
e1.record(); torch.cuda.synchronize()
wall_ms = (time.perf_counter() - t0) * 1e3 / a.iters
gpu_ms = e0.elapsed_time(e1) / a.iters
m.cuda_graphs = False
_lib.profile_enable(True)
for k in range(4):
    m(*sets[k % 2])
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
line = {"config": cfg.name, "batch": a.batch, "n_points": cfg.n_points, "clusters": cfg.real_cluster_num, "views": cfg.n_views,
        "text_tokens": cfg.n_text, "img_dtype": a.img_dtype, "cuda_graph": a.graphs, "wall_ms_per_forward": wall_ms, "gpu_ms_per_forward": gpu_ms,
        "scenes_per_s": a.batch / (wall_ms / 1e3), "kernel_ms": {k: v[0] / 4 for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
        "kernel_ms_total": sum(v[0] for v in prof.values()) / 4, "launches_per_forward": sum(v[1] for v in prof.values()) / 4}
if a.cpu_scenes:
    from oracle import preshape_oracle as po
    torch.set_num_threads(os.cpu_count() or 1)
    sd = syn.make_state_dict(cfg, 0, bf16_round=True)
    kw = dict(grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio, text_blocks=cfg.text_blocks, img_blocks=cfg.img_blocks,
              num_sub=cfg.num_sub, num_heads=cfg.num_heads, faithful_cost=True)
    data = [syn.make_inputs(cfg, 1, first_scene=500 + i, img_dtype=dt) for i in range(a.cpu_scenes + 1)]
    data = [(p, t, im.float()) for p, t, im in data]
    po.forward(sd, *data[0], **kw)
    t0 = time.perf_counter()
    for p, t, im in data[1:]:
        po.forward(sd, p, t, im, **kw)
    line["cpu_oracle_scenes_per_s"] = a.cpu_scenes / (time.perf_counter() - t0)
    line["cpu_cores"] = os.cpu_count()
print(json.dumps(line))
