#!/usr/bin/env python
"""Benchmark of the preshape hot path (ProxyTransformationNormReverse.forward, eval) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch scenes-per-GPU]

Metric (BASELINE.json): ProxyBlock fwd scenes/sec at 100k points / 256 clusters / d=256 with 64 text + 196 image
proxies (config C2): one "step" = one forward of the module over a batch of synthetic scenes.
  value     whole-job scenes/s, inputs resident in HBM (CUDA events, max over ranks, barrier on both sides)
  e2e       same metric through the module's public forward() with HOST (pinned) inputs and host results: the H2D copy of
            points/text/mask/image features and the D2H read of the packed result are inside the timed region
  roofline  the dominant kernel (image-feature pooling pass, HBM-bound) timed live with CUDA events on its stream
  cpu_baseline / --impl reference   the reference's algorithm on the host cores: the oracle port (oracle/preshape_oracle.py
            with faithful_cost=True: all blocks, all 226 attention queries, exactly the work the reference's PyTorch path
            does).  The reference itself is pure Python that needs pytorch3d/timm/mmengine shims and lives only in the
            build container, so it cannot travel to the GPU box; kind = "port".
Multi-GPU: scenes are independent, so each rank processes its own shard with no data-path collective (weak scaling,
`--batch` scenes per GPU); one all_gather of a small per-rank metric tensor at the end.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "ProxyBlock fwd scenes/sec (100k pts, 256 clusters)"
UNIT = "scenes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="scenes per GPU per step")
    ap.add_argument("--img-dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--box", default="wide", choices=["wide", "room"])
    ap.add_argument("--n-points", type=int, default=100000)
    ap.add_argument("--cpu-scenes", type=int, default=96, help="scenes in the bounded CPU-baseline sample (about 10 s of host work)")
    ap.add_argument("--ref-scenes-per-step", type=int, default=8, help="--impl reference: scenes per timed step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--core", action="store_true", help="core region: image proxies precomputed (diagnostic)")
    return ap.parse_args()


def workload(args):
    from proxytransformation_b200 import synthetic as syn
    cfg = syn.C2_WIDE if args.box == "wide" else syn.C2_ROOM
    if args.n_points != cfg.n_points:
        cfg = cfg.replace(n_points=args.n_points, name=f"{cfg.name}-N{args.n_points}")
    return cfg


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md 'clocks' line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:      # timed region shorter than the sampling period: fall back to every sample taken
            for ts, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except (ValueError, IndexError):
                    pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU (reference arm)
def cpu_reference_rate(cfg, n_scenes: int, per_call: int = 8, warm: int = 1, img_dtype=torch.float32):
    """Oracle port, faithful cost, all host threads, `per_call` scenes per forward (the reference takes a batch as well);
    returns (scenes/s, cores, seconds)."""
    from oracle import preshape_oracle as po
    from proxytransformation_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.make_state_dict(cfg, 0, bf16_round=True)
    kw = dict(grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio, text_blocks=cfg.text_blocks,
              img_blocks=cfg.img_blocks, num_sub=cfg.num_sub, num_heads=cfg.num_heads, faithful_cost=True)
    per_call = max(1, min(per_call, n_scenes))
    calls = max(1, n_scenes // per_call)
    pool = [syn.make_inputs(cfg, per_call, first_scene=1000 + 16 * i, img_dtype=img_dtype) for i in range(2)]
    pool = [(p, t, im.float()) for p, t, im in pool]
    for i in range(warm):
        po.forward(sd, *pool[i % 2], **kw)
    t0 = time.perf_counter()
    for i in range(calls):                          # two distinct seeded batches, alternated: the oracle keeps no state
        po.forward(sd, *pool[i % 2], **kw)
    dt = time.perf_counter() - t0
    return calls * per_call / dt, cores, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload(args)
    per_step = max(1, args.ref_scenes_per_step)
    rates, t_all = [], 0.0
    from oracle import preshape_oracle as po
    from proxytransformation_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.make_state_dict(cfg, 0, bf16_round=True)
    kw = dict(grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio, text_blocks=cfg.text_blocks,
              img_blocks=cfg.img_blocks, num_sub=cfg.num_sub, num_heads=cfg.num_heads, faithful_cost=True)
    dt_img = torch.bfloat16 if args.img_dtype == "bf16" else torch.float32
    data = [syn.make_inputs(cfg, per_step, first_scene=2000 + i, img_dtype=dt_img) for i in range(2)]
    data = [(p, t, im.float()) for p, t, im in data]
    for i in range(args.warmup):
        p, t, im = data[i % 2]
        po.forward(sd, p, t, im, **kw)
    t0 = time.perf_counter()
    for i in range(args.steps):
        p, t, im = data[i % 2]
        po.forward(sd, p, t, im, **kw)
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg.name, "n_points": cfg.n_points, "clusters": cfg.real_cluster_num, "embed_dim": cfg.embed_dim,
                       "text_tokens": cfg.n_text, "image_views": cfg.n_views, "region": "full", "scenes_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x {per_step} scene(s), oracle port of the reference's PyTorch path "
                                       "(all blocks, full 226-token attention pool), fp32, all host threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measured_traffic(kernel: str, batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    (profiles/ncu_traffic.json, measured at a stated batch and scaled linearly in the batch); None if not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]
        return t["dram_bytes_per_launch"] * batch / t["batch"]
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    from proxytransformation_b200 import ProxyTransformationNormReverse, _lib, build_ext
    from proxytransformation_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build_ext.build()
    if world > 1:
        dist.barrier()
    _lib.load()

    cfg = workload(args)
    B = args.batch
    img_dtype = torch.bfloat16 if args.img_dtype == "bf16" else torch.float32
    sd = syn.make_state_dict(cfg, 0, bf16_round=True)           # "bf16" config: weights rounded once, fp32 math
    m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)

    # synthetic shard of this rank: scenes [rank*B, (rank+1)*B); two alternating input sets, each far larger than L2
    def make_set(seed_off):
        g = torch.Generator(device=dev).manual_seed(1234 + 7919 * rank + seed_off)
        box = torch.tensor(cfg.box, device=dev)
        P = torch.rand(B, cfg.n_points, 3, generator=g, device=dev) * box
        text = torch.randn(B, cfg.n_text, cfg.embed_dim, generator=g, device=dev)
        mask = torch.ones(B, cfg.n_text, dtype=torch.uint8, device=dev)
        for b in range(B):
            mask[b, cfg.n_text - (b % 8):] = 0
        hw = cfg.img_spacial_dim
        img = (torch.relu(torch.randn(B, cfg.n_views, cfg.input_dim, hw, hw, generator=g, device=dev)) * 1.5).to(img_dtype)
        return P, text, mask, img

    sets = [make_set(0), make_set(1)]
    img_proxy = None
    if args.core:
        img_proxy = [m.get_img_proxy(s[3]) for s in sets]

    def step(i):
        P, text, mask, img = sets[i % 2]
        return m.forward_packed(P, text, mask, img, img_proxy=img_proxy[i % 2] if args.core else None)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # ---- timed region: exactly K steps, device-resident inputs
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    w0 = time.time()
    ev0.record()
    for i in range(args.steps):
        out, counts = step(i)
    ev1.record()
    torch.cuda.synchronize()
    w1 = time.time()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(w0, w1) if rank == 0 else None

    # ---- per-kernel timing pass (same steps, events around every kernel on its stream)
    _lib.profile_enable(True)
    torch.cuda.synchronize()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for i in range(args.steps):
        step(i)
    pe1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    prof_ms = pe0.elapsed_time(pe1)
    _lib.profile_enable(False)

    # ---- core region: image proxies precomputed, i.e. everything but get_img_proxy; same steps, same timing
    core = None
    if not args.core:
        ip_pre = [m.get_img_proxy(s_[3]) for s_ in sets]
        for i in range(args.warmup):
            P, text, mask, img = sets[i % 2]
            m.forward_packed(P, text, mask, img, img_proxy=ip_pre[i % 2])
        sync_all()
        ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ce0.record()
        for i in range(args.steps):
            P, text, mask, img = sets[i % 2]
            m.forward_packed(P, text, mask, img, img_proxy=ip_pre[i % 2])
        ce1.record()
        torch.cuda.synchronize()
        c_ms = ce0.elapsed_time(ce1)
        if world > 1:
            t = torch.tensor([c_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c_ms = t.item()
        core = {"value": world * B * args.steps / (c_ms / 1e3), "unit": UNIT, "ms_per_step": c_ms / args.steps,
                "note": "image proxies (B,V,256) precomputed: 64 text + 196 image proxies as inputs, get_img_proxy excluded"}
        del ip_pre

    # ---- e2e: public forward() with pinned host inputs, host results
    e2e = None
    if not args.no_e2e:
        hsets = []
        for P, text, mask, img in sets:
            hsets.append(([p.cpu().pin_memory() for p in P], {"text_feats": text.cpu().pin_memory(),
                          "text_token_mask": mask.bool().cpu().pin_memory()}, img.cpu().pin_memory()))
        h2d = sum(p.numel() * 4 for p in hsets[0][0]) + hsets[0][1]["text_feats"].numel() * 4 + hsets[0][1]["text_token_mask"].numel() + \
            hsets[0][2].numel() * hsets[0][2].element_size()
        res = m(*hsets[0])
        d2h = B * cfg.n_points * 12 + 4 * B          # the packed (B,N,3) result block + B counts are copied back
        for i in range(max(1, args.warmup - 1)):
            m(*hsets[(i + 1) % 2])
        sync_all()
        e_steps = args.steps
        t0 = time.perf_counter()
        for i in range(e_steps):
            res = m(*hsets[i % 2])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e_ms = (t1 - t0) * 1e3
        if world > 1:
            t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = t.item()
        e2e = {"value": world * B * e_steps / (e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "steps": e_steps, "ms_per_step": e_ms / e_steps,
               "timer": "host perf_counter around forward() incl. copies, max over ranks",
               "pipeline": f"{m.host_chunk_scenes}-scene chunks over H2D / compute / D2H streams, one host sync per call"}

    # ---- the one collective of the path: all_gather of per-rank metric tensors (sharding.py)
    from proxytransformation_b200 import sharding
    cnt_host = counts.cpu().tolist()
    mine = sharding.scene_metrics([out[b, :cnt_host[b]] for b in range(B)], elapsed_ms=ms, launches=launches, device=dev)
    allm = sharding.gather_metrics(mine).cpu()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_max = allm[:, 3].max().item()
    scenes = allm[:, 0].sum().item() * args.steps
    value = scenes / (ms_max / 1e3)

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    top = max(prof.items(), key=lambda kv: kv[1][0]) if prof else None
    esz = 2 if img_dtype == torch.bfloat16 else 4
    tile_bytes = cfg.input_dim * cfg.img_spacial_dim ** 2 * esz
    alg_bytes = {   # algorithmic bytes per launch of the HBM-bound kernels
        "img_pool": B * cfg.n_views * tile_bytes, "img_mean": B * cfg.n_views * tile_bytes,
        "scatter_compact": B * cfg.n_points * 24, "minmax_partial": B * cfg.n_points * 12,
    }
    roof = None
    dom = "img_pool" if ("img_pool" in prof and not args.core) else (top[0] if top else None)
    if dom in prof and dom in alg_bytes:
        t_ms, n = prof[dom]
        ach = alg_bytes[dom] / (t_ms / n / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": measured_traffic(dom, B), "peak_source": peak_src, "avg_launch_ms": t_ms / n,
                "algorithmic_bytes_per_launch": alg_bytes[dom]}
    elif dom in prof:
        t_ms, n = prof[dom]
        roof = {"bound": "latency", "kernel": dom, "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                "traffic": None, "avg_launch_ms": t_ms / n}
    # whole-path HBM roofline: algorithmic bytes per scene / measured copy bandwidth
    path_bytes = 24 * cfg.n_points + 2 * cfg.embed_dim * (cfg.n_text + cfg.n_views) + cfg.n_text + (0 if args.core else cfg.n_views * tile_bytes)
    path_bound = hbm_peak * 1e9 / path_bytes
    breakdown = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in
                 sorted(prof.items(), key=lambda kv: -kv[1][0])}

    cpu = None
    if not args.no_cpu_baseline:
        v, cores, dt = cpu_reference_rate(cfg, args.cpu_scenes, per_call=args.ref_scenes_per_step, img_dtype=img_dtype)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_scenes} scenes of the same workload in batches of {min(args.ref_scenes_per_step, args.cpu_scenes)} ({dt:.1f} s), "
                         "oracle port of the reference's PyTorch path (all blocks, full 226-token attention pool), fp32, all host threads"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (3xBF16 tensor-core dense layers, bf16 image features)" if img_dtype == torch.bfloat16 else "f32",
            "data": "synthetic",
            "config": {"workload": cfg.name, "n_points": cfg.n_points, "clusters": cfg.real_cluster_num, "embed_dim": cfg.embed_dim,
                       "text_tokens": cfg.n_text, "image_views": cfg.n_views, "img_feat_dtype": args.img_dtype,
                       "region": "core" if args.core else "full", "scenes_per_gpu_per_step": B, "box_m": list(cfg.box),
                       "l2": "two alternating input sets per rank, each %.1f GB >> 126 MB L2" % (B * cfg.n_views * tile_bytes / 1e9),
                       "sharding": f"{world} x {B} independent scenes, no data-path collective"},
            "e2e": e2e, "gpu_launches": int(allm[:, 4].sum().item()), "clocks": clocks, "roofline": roof,
            "path_roofline": {"algorithmic_bytes_per_scene": path_bytes, "hbm_bound_scenes_per_s_per_gpu": path_bound,
                              "frac": value / world / path_bound},
            "core_region": core, "kernel_breakdown": breakdown, "profiled_pass_ms_per_step": prof_ms / args.steps, "cpu_baseline": cpu,
            "checks": {"survivors_per_scene": allm[0, 1].item() / B}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _guard_stdout():
    """Libraries (NCCL's version banner, torchrun) write to fd 1; the contract is ONE JSON line on stdout.  Route fd 1 to
    stderr for the duration of the run and give print() a private handle on the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    a = parse()
    _guard_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
