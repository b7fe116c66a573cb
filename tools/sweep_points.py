#!/usr/bin/env python
"""Point-count sweep of the core region (image proxies precomputed) — BASELINE.json config 5 / SURVEY.md §8d:
N in 16k..512k at 1 GPU, scenes/s, algorithmic GB/s (24*N bytes per scene: points read once, packed output written once)
against the measured HBM peak, and the oracle port on the host cores for a bounded sample.  L2 is flushed between timed
iterations (256 MiB write).  Prints one JSON line per N."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proxytransformation_b200 import ProxyTransformationNormReverse, _lib, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--cpu-scenes", type=int, default=2)
ap.add_argument("--sizes", default="16384,32768,65536,131072,262144,524288")
a = ap.parse_args()
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for N in [int(x) for x in a.sizes.split(",")]:
    cfg = syn.C2_WIDE.replace(n_points=N, name=f"C2-wide-N{N}")
    sd = syn.make_state_dict(cfg, 0, bf16_round=True)
    m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    B = a.batch
    g = torch.Generator(device=dev).manual_seed(N)
    P = torch.rand(B, N, 3, generator=g, device=dev) * torch.tensor(cfg.box, device=dev)
    text = torch.randn(B, cfg.n_text, cfg.embed_dim, generator=g, device=dev)
    mask = torch.ones(B, cfg.n_text, dtype=torch.uint8, device=dev)
    img_proxy = torch.randn(B, cfg.n_views, cfg.embed_dim, generator=g, device=dev)
    for _ in range(5):                  # warm-up with the flush in place: the caching allocator reaches its steady state
        flush.fill_(1)
        m.forward_packed(P, text, mask, None, img_proxy=img_proxy)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(a.iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, counts = m.forward_packed(P, text, mask, None, img_proxy=img_proxy)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / a.iters
    sps = B / (ms / 1e3)
    # bandwidth kernels alone (minmax + scatter/compact), timed live through pt_profile_*
    _lib.profile_enable(True)
    m.forward_packed(P, text, mask, None, img_proxy=img_proxy)
    torch.cuda.synchronize()
    pr = _lib.profile_read()
    _lib.profile_enable(False)
    bw_ms = sum(pr[k][0] for k in ("minmax_partial", "scatter_mark", "scatter_count", "scatter_compact") if k in pr)
    bq_ms = pr.get("ball_query", (0.0, 0))[0]
    cpu = None
    if a.cpu_scenes > 0:
        from oracle import preshape_oracle as po
        torch.set_num_threads(os.cpu_count() or 1)
        pts, td, _ = syn.make_inputs(cfg.replace(n_views=1), a.cpu_scenes + 1, first_scene=500)
        ip = torch.randn(a.cpu_scenes + 1, cfg.n_views, cfg.embed_dim)
        kw = dict(grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio, text_blocks=cfg.text_blocks,
                  img_blocks=cfg.img_blocks, num_sub=cfg.num_sub, num_heads=cfg.num_heads, faithful_cost=True)
        po.forward(sd, pts[:1], {k: v[:1] for k, v in td.items()}, None, img_proxy=ip[:1], **kw)
        t0 = time.perf_counter()
        for i in range(1, a.cpu_scenes + 1):
            po.forward(sd, pts[i:i + 1], {k: v[i:i + 1] for k, v in td.items()}, None, img_proxy=ip[i:i + 1], **kw)
        cpu = a.cpu_scenes / (time.perf_counter() - t0)
    print(json.dumps({"n_points": N, "batch": B, "region": "core", "ms_per_step": ms, "scenes_per_s": sps,
                      "algorithmic_GBps": 24 * N * sps / 1e9, "hbm_frac_of_measured_peak": 24 * N * sps / 1e9 / peak,
                      "bandwidth_kernels_ms": bw_ms, "bandwidth_kernels_GBps": 24 * N * B / (bw_ms / 1e3) / 1e9 if bw_ms else None,
                      "ball_query_ms": bq_ms, "cpu_oracle_scenes_per_s": cpu, "cpu_cores": os.cpu_count()}), flush=True)
    del m, P, text, mask, img_proxy, out, counts
    torch.cuda.empty_cache()            # the next size starts from a clean allocator (no cudaMalloc inside a timed iteration)
