"""world_size-2 gloo test of the N>1 host logic: unit sharding + the single metric all-reduce."""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from blindshadowremoval_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, frame, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.shard_images(n_images, frame, rank, world)
    # per-image "metric" = image index; each rank scores only its own shard
    sums = {"psnr": float(sum(range(b, e))), "auc": float(sum(i * 0.5 for i in range(b, e)))}
    out = sharding.reduce_metrics(sums, e - b)
    q.put((rank, b, e, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_metric_reduce():
    world, n, frame = 2, 50, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, frame, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 30), (30, 50)]          # 3 + 2 chunks of 10
    for r in res:
        assert r[3]["count"] == n
        assert abs(r[3]["psnr"] - (n - 1) / 2) < 1e-12               # identical global mean on every rank
        assert abs(r[3]["auc"] - (n - 1) / 4) < 1e-12


def test_reduce_metrics_single_process():
    out = sharding.reduce_metrics({"ssim": 3.0}, 4)
    assert out == {"ssim": 0.75, "count": 4.0}
