#!/usr/bin/env python
"""Benchmark of the preshape hot path (ProxyTransformationNormReverse.forward, eval) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch scenes-per-GPU]

Metric (BASELINE.json): ProxyBlock fwd scenes/sec at 100k points / 256 clusters / d=256 with 64 text + 196 image
proxies (config C2): one "step" = one forward of the module over a batch of synthetic scenes.
  value     whole-job scenes/s, inputs resident in HBM (CUDA events, max over ranks, barrier on both sides)
  e2e       same metric through the module's public forward() with HOST (pinned) inputs and host results: the H2D copy of
            points/text/mask/image features and the D2H read of the packed result are inside the timed region
  roofline  the dominant kernel (largest total time in the live per-kernel profile; today the image-feature pooling pass,
            HBM-bound) timed with CUDA events on its stream; image_stage_roofline: the whole image stage against one read
  cpu_baseline / --impl reference   the reference's CPU path on the host cores: the reference's OWN module (kind "reference":
            its bytecode is compiled from /root/reference into oracle/_ref/ by oracle/build.py and travels with the repo;
            pytorch3d / timm / registry come from the oracle's shims) or, when that file is absent, the oracle port with
            faithful_cost=True (all blocks, all 226 attention queries; kind "port").
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
    ap.add_argument("--no-checks", action="store_true", help="skip the oracle check of the timed batch")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1 / C3 / strong-scaling side measurements")
    return ap.parse_args()


def config_dict(cfg, args, world: int, batch: int) -> dict:
    """The `config` object of the JSON line — identical for both arms (the driver compares them key by key)."""
    esz = 2 if args.img_dtype == "bf16" else 4
    tile_bytes = cfg.input_dim * cfg.img_spacial_dim ** 2 * esz
    return {"workload": cfg.name, "n_points": cfg.n_points, "clusters": cfg.real_cluster_num, "embed_dim": cfg.embed_dim,
            "text_tokens": cfg.n_text, "image_views": cfg.n_views, "img_feat_dtype": args.img_dtype,
            "region": "core" if args.core else "full", "scenes_per_gpu_per_step": batch, "box_m": list(cfg.box),
            "l2": "two alternating input sets per rank, each %.1f GB >> 126 MB L2" % (batch * cfg.n_views * tile_bytes / 1e9),
            "sharding": f"{world} x {batch} independent scenes, no data-path collective"}


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

    def wait_first(self, timeout: float = 10.0):
        """block until nvidia-smi has initialised NVML and delivered its first sample, so that its start-up (which takes driver
        locks for up to seconds on a multi-GPU box) never falls into the timed region"""
        t_end = time.time() + timeout
        while self.proc is not None and not self.rows and time.time() < t_end and self.proc.poll() is None:
            time.sleep(0.02)

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
def cpu_forward(cfg):
    """-> (forward(points, text_dict, img_feat), kind, description): the reference's OWN module when its bytecode travelled
    with the repo (oracle/_ref, built by oracle/build.py::build_ref from /root/reference; pytorch3d's two CPU loops come from
    the C restatement, timm / registry from shims), else the oracle port with faithful cost."""
    from proxytransformation_b200 import synthetic as syn
    sd = syn.make_state_dict(cfg, 0, bf16_round=True)
    try:
        from oracle import ref_shim
        if ref_shim.available():
            net = ref_shim.build_module(cfg.module_kwargs(), sd, pinned=False)
            ref_shim.use_native_fps(True)

            def fwd(p, t, im):
                with torch.no_grad():
                    return net(p, t, im)
            return fwd, "reference", ("the reference's own ProxyTransformationNormReverse.forward (unmodified module, eval, all blocks, "
                                      "full 226-token attention pool; pytorch3d ball query / FPS = C restatement of its CPU loops)")
        else:
            sys.stderr.write("reference bytecode oracle/_ref/ not found (built by __graft_entry__.build() where /root/reference exists); timing the oracle port\n")
    except Exception as e:      # pragma: no cover - fall back to the port, say why
        sys.stderr.write(f"reference module unavailable ({e!r}); timing the oracle port\n")
    from oracle import preshape_oracle as po
    kw = dict(grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio, text_blocks=cfg.text_blocks,
              img_blocks=cfg.img_blocks, num_sub=cfg.num_sub, num_heads=cfg.num_heads, faithful_cost=True)
    return (lambda p, t, im: po.forward(sd, p, t, im, **kw)), "port", ("oracle port of the reference's PyTorch path (all blocks, "
                                                                        "full 226-token attention pool)")


def cpu_reference_rate(cfg, n_scenes: int, per_call: int = 8, warm: int = 1, img_dtype=torch.float32):
    """The reference's CPU path on all host threads, `per_call` scenes per forward (the reference takes a batch as well);
    returns (scenes/s, cores, seconds, kind, description)."""
    from proxytransformation_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind, desc = cpu_forward(cfg)
    per_call = max(1, min(per_call, n_scenes))
    calls = max(1, n_scenes // per_call)
    pool = [syn.make_inputs(cfg, per_call, first_scene=1000 + 16 * i, img_dtype=img_dtype) for i in range(2)]
    pool = [(p, t, im.float()) for p, t, im in pool]
    for i in range(warm):
        fwd(*pool[i % 2])
    t0 = time.perf_counter()
    for i in range(calls):                          # two distinct seeded batches, alternated: no state is kept
        fwd(*pool[i % 2])
    dt = time.perf_counter() - t0
    return calls * per_call / dt, cores, dt, kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload(args)
    per_step = max(1, args.ref_scenes_per_step)
    from proxytransformation_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind, desc = cpu_forward(cfg)
    dt_img = torch.bfloat16 if args.img_dtype == "bf16" else torch.float32
    data = [syn.make_inputs(cfg, per_step, first_scene=2000 + i, img_dtype=dt_img) for i in range(2)]
    data = [(p, t, im.float()) for p, t, im in data]
    for i in range(args.warmup):
        fwd(*data[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        fwd(*data[i % 2])
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # same workload description as the B200 arm (the CPU arm's step is a bounded sample of it: see cpu_baseline.sample)
            "config": config_dict(cfg, args, args.gpus, args.batch),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{args.steps} steps x {per_step} scene(s) of the workload, {desc}, fp32, all host threads"},
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


def forward_latency(cfg, batch: int, dev, img_dtype, iters: int = 20, graph: bool = False):
    """GPU ms per forward() (device-resident inputs, CUDA events, one D2H of the counts inside) of another BASELINE.json
    workload: two alternating input sets, warm-up calls for at least 0.2 s."""
    from proxytransformation_b200 import ProxyTransformationNormReverse
    from proxytransformation_b200 import synthetic as syn
    m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
    m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True), strict=True)
    m = m.to(dev)
    m.cuda_graphs = graph
    sets = []
    for k in range(2):
        pts, td, img = syn.make_inputs(cfg, batch, first_scene=100 * k, img_dtype=img_dtype)
        sets.append(([p.to(dev) for p in pts], {n: v.to(dev) for n, v in td.items()}, img.to(dev)))
    with torch.no_grad():
        t_w, k = time.time(), 0
        while k < 4 or time.time() - t_w < 0.2:         # (these run after CPU-side phases of the bench: 0.2 s under load first)
            m(*sets[k % 2])
            k += 1
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(iters):
            m(*sets[k % 2])
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def oracle_checks(m, cfg, sets, B: int) -> dict:
    """Parity of the TIMED inputs: the first and the last scene of input set 0 through the same forward_packed call the timed region
    makes, against the oracle (CPU) on identical inputs — cluster indices bit-exact, image proxies, transformed coordinates.  (The
    last scene's views are the LAST views the persistent image-pool CTAs handle: ring wrap-arounds, barrier parities and buffer
    rotations after ~85 views per CTA.)"""
    import numpy as np
    from oracle import preshape_oracle as po
    from proxytransformation_b200 import synthetic as syn
    P, text, mask, img = sets[0]
    tr = {}
    out, counts = m.forward_packed(P, text, mask, img, trace=tr)
    torch.cuda.synchronize()
    sd = syn.make_state_dict(cfg, 0, bf16_round=True)
    res = {"scene": "first and last scene of the timed input set 0 vs the oracle on identical inputs", "scenes_checked": [],
           "idx_equal": True, "count_equal": True, "max_coord_err": 0.0, "img_proxy_max_err": 0.0, "img_proxy_views_checked": 0,
           "coord_tolerance": 1e-4, "img_proxy_tolerance": 6e-5}
    for s_ in sorted({0, P.shape[0] - 1}):
        otr = {}
        want = po.forward(sd, [P[s_].cpu()], {"text_feats": text[s_:s_ + 1].cpu(), "text_token_mask": mask[s_:s_ + 1].bool().cpu()},
                          img[s_:s_ + 1].float().cpu(), grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio,
                          text_blocks=cfg.text_blocks, img_blocks=cfg.img_blocks, num_sub=cfg.num_sub, num_heads=cfg.num_heads, trace=otr)[0]
        n0 = int(counts[s_].item())
        got = out[s_, :n0].cpu()
        res["scenes_checked"].append(int(s_))
        res["idx_equal"] &= bool(np.array_equal(tr["kept_idx"][s_].cpu().numpy(), otr["kept_idx"][0].numpy()) and
                                 np.array_equal(tr["drop_idx"][s_].cpu().numpy(), otr["drop_idx"][0].numpy()) and
                                 np.array_equal(tr["idx2"][s_].cpu().numpy(), otr["idx2"][0].numpy()))
        res["count_equal"] &= bool(n0 == want.shape[0])
        res["max_coord_err"] = max(res["max_coord_err"], float((got - want).abs().max())) if n0 == want.shape[0] else None
        res["img_proxy_max_err"] = max(res["img_proxy_max_err"], float((tr["img_proxy"][s_].cpu() - otr["img_proxy"][0]).abs().max()))
        res["img_proxy_views_checked"] += int(cfg.n_views)
        if res["max_coord_err"] is None:
            break
    return res


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

    # set-up, not part of the contract's W warm-up steps: the first calls pack the weights, set function attributes and grow the
    # caching allocator's pools of BOTH streams the forward uses (a cudaMalloc inside the timed region would stall the device)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # started before the set-up steps: NVML start-up overlaps them, not the timed region
    t_setup = time.time()
    i = 0
    while i < 4 or time.time() - t_setup < 0.3:      # ... and at least 0.3 s under load: the clocks are up before the warm-up steps
        step(i)
        i += 1
        if i % 4 == 0:
            torch.cuda.synchronize()
    sync_all()
    if rank == 0:
        sampler.wait_first()

    mallocs = [0]

    def timed_region():
        for i in range(args.warmup):
            step(i)
        sync_all()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        a0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        t0 = time.time()
        e0.record()
        for i in range(args.steps):
            res = step(i)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        mallocs[0] = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - a0      # cudaMalloc calls of the caching allocator inside the K steps
        if world > 1:
            dist.barrier()
        return e0.elapsed_time(e1), _lib.launch_count() - l0, t0, t1, res

    # ---- timed region: W warm-up steps, then exactly K steps, device-resident inputs
    ms, launches, w0, w1, (out, counts) = timed_region()

    # ---- per-kernel timing pass (same steps, events around every kernel on its stream).  The timed region above runs the image
    # stage on a second stream next to the geometric stages; for the per-kernel durations (roofline, kernel_breakdown) the kernels
    # run one after the other on one stream, otherwise every duration would include the kernels it shared the GPU with.
    ov_saved = m.overlap_img_stage
    m.overlap_img_stage = "0"
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
    m.overlap_img_stage = ov_saved

    # One re-measurement when the host could not keep the device fed.  The K steps are enqueued asynchronously, so the event time
    # of the timed region is device time unless something stalled the launching thread (a driver lock held by another process,
    # a page-in): the per-kernel pass just above ran the SAME kernels serialised on one stream with an event pair around each,
    # which bounds a healthy timed region from above.  A first measurement more than 1.25x that bound is kept under "remeasured"
    # and replaced by a second timed region (W warm-up steps + K steps again).
    remeasured = None
    stalled = torch.tensor([1.0 if ms > 1.25 * prof_ms else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(stalled, op=dist.ReduceOp.MAX)
    if stalled.item() > 0:
        first = ms / args.steps
        ms, launches, w0, w1, (out, counts) = timed_region()
        remeasured = {"first_ms_per_step": first, "serialised_pass_ms_per_step": prof_ms / args.steps,
                      "reason": "first timed region exceeded 1.25x the serialised per-kernel pass (host-side stall); second measurement reported"}
    clocks = sampler.stop(w0, w1) if rank == 0 else None

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

    # ---- BASELINE config 4 as worded: 64 scenes in total, sharded over the ranks (strong scaling: 64 / N scenes per rank).
    # Latency-bound at 8 scenes per rank — reported beside the weak-scaling headline, same timing rules.
    c4 = None
    if not args.core and not args.no_extra:
        b4 = max(1, min(B, 64 // world))
        sub = [tuple(t[:b4] for t in s_) for s_ in sets]
        t_w, i = time.time(), 0
        while i < 3 or time.time() - t_w < 0.3:      # the PCIe-bound e2e loop above leaves the GPU mostly idle: 0.3 s under load first
            m.forward_packed(*sub[i % 2])
            i += 1
            if i % 4 == 0:
                torch.cuda.synchronize()
        sync_all()
        s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0_.record()
        for i in range(args.steps):
            m.forward_packed(*sub[i % 2])
        s1_.record()
        torch.cuda.synchronize()
        c4_ms = s0_.elapsed_time(s1_)
        if world > 1:
            t = torch.tensor([c4_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c4_ms = t.item()
        c4 = {"scenes_total": b4 * world, "scenes_per_rank": b4, "ms_per_step": c4_ms / args.steps, "unit": UNIT,
              "value": b4 * world * args.steps / (c4_ms / 1e3), "scaling": "strong", "l2": "%.2f GB of image features per rank per step" %
              (b4 * cfg.n_views * cfg.input_dim * cfg.img_spacial_dim ** 2 * (2 if img_dtype == torch.bfloat16 else 4) / 1e9)}
        del sub

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
    ms_ranks = [round(float(x) / args.steps, 4) for x in allm[:, 3].tolist()]
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
    # dominant kernel = the profile tag with the largest total time in the timed steps (picked live, not hard-coded)
    bf16_peak = peaks.get("bf16_tflops_sustained") or 1400.0
    c, hid, n_, Lt, V_ = cfg.embed_dim, 4 * cfg.embed_dim, cfg.real_cluster_num, cfg.n_text, cfg.n_views
    blk_flops = lambda l: 2.0 * B * (n_ * c * 3 * c + l * c * c + n_ * c * c + 2 * n_ * c * hid)      # useful FLOPs of one live block's GEMMs
    tensor_flops = {"gemm_tc_3xbf16": 3.0 * (blk_flops(Lt) + blk_flops(V_))}                            # x3: hi*hi + lo*hi + hi*lo products
    roof = None
    dom = top[0] if top else None
    if dom in prof and dom in alg_bytes:
        t_ms, n = prof[dom]
        ach = alg_bytes[dom] / (t_ms / n / 1e3) / 1e9
        traffic = measured_traffic(dom, B)
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": t_ms / n,
                "algorithmic_bytes_per_launch": alg_bytes[dom]}
        if traffic:     # what the kernel really moves (ncu dram bytes: features + per-view operand planes + outputs) over the same duration
            roof["traffic_gbs"] = traffic / (t_ms / n / 1e3) / 1e9
            roof["traffic_frac_of_peak"] = roof["traffic_gbs"] / hbm_peak
    elif dom in prof and dom in tensor_flops:
        t_ms, n = prof[dom]
        ach = tensor_flops[dom] / (t_ms / args.steps / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak,
                "traffic": None, "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)", "avg_launch_ms": t_ms / n,
                "note": "all launches of the tag per step; issued FLOPs = 3 x useful (3xBF16 split)"}
    elif dom in prof:
        t_ms, n = prof[dom]
        roof = {"bound": "latency", "kernel": dom, "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                "traffic": None, "avg_launch_ms": t_ms / n}
    # the image stage as a whole (mean pass + query-side projections + pooling pass + value-side projections) against ONE
    # algorithmic read of the feature maps
    stage = None
    if not args.core and "img_pool" in prof:
        st_ms = sum(prof[k][0] for k in ("img_mean", "img_pool", "gemm_img_3xbf16") if k in prof) / args.steps
        st_alg = B * cfg.n_views * tile_bytes
        st_traffic = None
        if all(measured_traffic(k, B) is not None for k in ("img_mean", "img_pool")):
            st_traffic = measured_traffic("img_mean", B) + measured_traffic("img_pool", B) + (measured_traffic("gemm_img_3xbf16", B) or 0.0)
        stage = {"kernels": [k for k in ("img_mean", "gemm_img_3xbf16", "img_pool") if k in prof], "ms_per_step": st_ms,
                 "algorithmic_bytes_per_step": st_alg, "achieved": st_alg / (st_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                 "frac": st_alg / (st_ms / 1e3) / 1e9 / hbm_peak, "traffic": st_traffic,
                 "note": "the query of the attention pool depends on the spatial mean of the whole view, so the stage reads the features twice"}
    # whole-path HBM roofline: algorithmic bytes per scene / measured copy bandwidth
    path_bytes = 24 * cfg.n_points + 2 * cfg.embed_dim * (cfg.n_text + cfg.n_views) + cfg.n_text + (0 if args.core else cfg.n_views * tile_bytes)
    path_bound = hbm_peak * 1e9 / path_bytes
    breakdown = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in
                 sorted(prof.items(), key=lambda kv: -kv[1][0])}

    cpu = None
    if not args.no_cpu_baseline:
        v, cores, dt, kind, desc = cpu_reference_rate(cfg, args.cpu_scenes, per_call=args.ref_scenes_per_step, img_dtype=img_dtype)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{args.cpu_scenes} scenes of the same workload in batches of {min(args.ref_scenes_per_step, args.cpu_scenes)} ({dt:.1f} s), "
                         f"{desc}, fp32, all host threads"}

    checks = {"survivors_per_scene": allm[0, 1].item() / B}
    if not args.core and not args.no_checks:
        checks.update(oracle_checks(m, cfg, sets, B))
    extra = {}
    if not args.core and not args.no_extra:
        # the other BASELINE.json configurations, single GPU forward latency (rank 0): C1 (4 096 points / 16 clusters, one scene),
        # C3 (the shipped grounding config gs=12 / ddr=0.6 -> 691 clusters, batch 4, 50 views, 32 text tokens), eager and CUDA graph
        del sets
        torch.cuda.empty_cache()
        for key, c_, b_ in (("c1", syn.C1, 1), ("c3", syn.C3, 4), ("c3_wide", syn.C3_WIDE, 4)):
            ms_e = forward_latency(c_, b_, dev, img_dtype)
            ms_g = forward_latency(c_, b_, dev, img_dtype, graph=True)
            extra[key] = {"workload": c_.name, "batch": b_, "clusters": c_.real_cluster_num, "views": c_.n_views, "ms_per_forward": ms_e,
                          "ms_per_forward_cuda_graph": ms_g, "scenes_per_s": b_ / (min(ms_e, ms_g) / 1e3)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (3xBF16 tensor-core dense layers, bf16 image features)" if img_dtype == torch.bfloat16 else "f32",
            "data": "synthetic",
            "config": config_dict(cfg, args, world, B),
            "e2e": e2e, "gpu_launches": int(allm[:, 4].sum().item()), "clocks": clocks, "roofline": roof,
            "path_roofline": {"algorithmic_bytes_per_scene": path_bytes, "hbm_bound_scenes_per_s_per_gpu": path_bound,
                              "frac": value / world / path_bound},
            "image_stage_roofline": stage, "c4_strong": c4,
            "ms_per_step_by_rank": ms_ranks, "core_region": core, "kernel_breakdown": breakdown, "profiled_pass_ms_per_step": prof_ms / args.steps,
            "profiled_pass": "one stream, kernels back to back (the timed region overlaps the image stage with the geometric stages on two streams)",
            "cpu_baseline": cpu,
            "checks": checks}
    line["device_mallocs_in_timed_region"] = int(mallocs[0])
    if remeasured is not None:
        line["remeasured"] = remeasured
    line.update(extra)
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
