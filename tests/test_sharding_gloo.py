"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path — contiguous scene shards with no
data-path collective and one all_gather of per-rank metric tensors.  Each rank runs the ORACLE on its shard (the CUDA
path needs a GPU); the union of the shards must equal the single-process result scene by scene."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from proxytransformation_b200 import sharding
from proxytransformation_b200 import synthetic as syn


def test_shard_ranges_cover_every_scene_once():
    for n in (0, 1, 3, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) <= -(-n // world) if n else True
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_scene_outputs(first, count):
    from oracle import preshape_oracle as po
    cfg = syn.C1
    sd = syn.make_state_dict(cfg, 0)
    if count == 0:
        return []
    pts, text_dict, img = syn.make_inputs(cfg, count, first_scene=first)
    return po.forward(sd, pts, text_dict, img, grid_size=cfg.grid_size, dynamic_drop_radio=cfg.dynamic_drop_radio,
                      text_blocks=cfg.text_blocks, img_blocks=cfg.img_blocks, num_sub=cfg.num_sub)


def _worker(rank, world, port, n_scenes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        a, b = sharding.shard_range(n_scenes, rank, world)
        outs = _oracle_scene_outputs(a, b - a)
        table = sharding.gather_metrics(sharding.scene_metrics(outs, elapsed_ms=10.0 + rank, launches=7))
        q.put((rank, table.numpy().tolist(), [tuple(o.shape) for o in outs]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_match_single_process():
    n_scenes, world = 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scenes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    whole = _oracle_scene_outputs(0, n_scenes)
    ref = sharding.scene_metrics(whole)
    tables = [torch.tensor(t, dtype=torch.float64) for _, t, _ in got]
    assert torch.equal(tables[0], tables[1])                       # every rank sees the same gathered table
    job = sharding.reduce_job(tables[0])
    assert job["n_scenes"] == n_scenes and job["survivors"] == int(ref[1].item())
    assert abs(job["coord_checksum"] - ref[2].item()) <= 1e-6 * max(1.0, abs(ref[2].item()))
    assert job["elapsed_ms"] == 11.0 and job["launches"] == 14      # time is max over ranks, launches add up
    shapes = [s for _, _, ss in got for s in ss]
    assert shapes == [tuple(o.shape) for o in whole]                # shard order == scene order
