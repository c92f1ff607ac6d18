"""world_size-2 gloo tests of the multi-GPU bookkeeping: image sharding and the per-image metric gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from image_restoration_and_enhancement_b200 import metrics, sweep
    idx = sweep.shard(n_items, rank, world)
    rng = lambda i: np.random.default_rng(i)
    vals = {"psnr": [float(rng(i).uniform(5, 40)) for i in idx], "ssim": [float(rng(i + 1000).uniform(0, 1)) for i in idx],
            "lpips": [float(rng(i + 2000).uniform(0, 1)) for i in idx]}            # the (psnr, ssim, lpips) triple of north_star
    full = metrics.gather_per_image(idx, vals)
    if rank == 0:
        q.put(metrics.summarize("denoise", full, n_items))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gather_is_bit_identical_to_single_process():
    from image_restoration_and_enhancement_b200 import metrics, sweep
    n = 13                                          # ragged: ranks hold 7 and 6 items
    assert sweep.shard(n, 0, 2) == [0, 2, 4, 6, 8, 10, 12] and sweep.shard(n, 1, 2) == [1, 3, 5, 7, 9, 11]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = lambda i: np.random.default_rng(i)
    single = metrics.summarize("denoise", {"psnr": [float(rng(i).uniform(5, 40)) for i in range(n)],
                                            "ssim": [float(rng(i + 1000).uniform(0, 1)) for i in range(n)],
                                            "lpips": [float(rng(i + 2000).uniform(0, 1)) for i in range(n)]}, n)
    assert set(got["metrics"]) == {"psnr", "ssim", "lpips"}
    for k in ("psnr", "ssim", "lpips"):
        for stat in ("mean", "std", "min", "max", "median"):
            assert got["metrics"][k][stat] == single["metrics"][k][stat]        # bit-exact, not approx


def test_gather_single_process_orders_by_index():
    from image_restoration_and_enhancement_b200 import metrics
    out = metrics.gather_per_image([2, 0, 1], {"psnr": [30.0, 10.0, 20.0]})
    assert out == {"psnr": [10.0, 20.0, 30.0]}


def test_synthetic_pairs_are_seeded():
    from image_restoration_and_enhancement_b200 import synth
    for task in ("denoise", "sr", "colorize", "inpaint"):
        a, b = synth.make_pair(task, 5, 64, 64), synth.make_pair(task, 5, 64, 64)
        assert all((a[k] == b[k]).all() for k in a) and a["input"].shape == (64, 64, 3) and a["input"].dtype == np.uint8
    it = synth.make_pair("inpaint", 1, 64, 64)
    assert set(np.unique(it["mask"])) <= {0, 255} and (it["input"][it["mask"] == 255] == 0).all()
    g = synth.make_pair("colorize", 2, 64, 64)["input"]
    assert (g[..., 0] == g[..., 1]).all() and (g[..., 1] == g[..., 2]).all()
