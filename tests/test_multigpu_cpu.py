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


class _FakeGen:
    """Stands in for the CUDA generator in the CPU test of the evaluation loop: a deterministic function of the chunk."""

    def forward_chunk(self, chunk, frame=None, share=True):
        import numpy as np
        img, mask, face = chunk[..., 0:3], chunk[..., 6:7], chunk[..., 16:17]
        pred = (0.8 * (mask == 2) + 0.1 * img[..., 0:1]) * face
        return np.clip(img, 0, 1), pred.astype(np.float32)


def _chunk(i):
    import numpy as np
    rng = np.random.default_rng(100 + i)
    c = rng.random((2, 256, 256, 17), dtype=np.float32)
    c[..., 6] = rng.integers(0, 3, (2, 256, 256))                  # SFW label map {0,1,2}
    c[..., 16] = (rng.random((2, 256, 256)) > 0.2)
    return c


def _eval_worker(rank, world, port, n_chunks, q):
    from blindshadowremoval_b200.evaluate import evaluate_sfw
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q.put((rank, evaluate_sfw(_FakeGen(), _chunk, n_chunks, frame=2, rank=rank, world=world)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sfw_evaluation_equals_single_process():
    """Config 3 host logic: chunks of 2 frames dealt to 2 ranks, one all-reduce -> every rank holds the means a single
    process computes over all chunks (train_with_TSM.py:633-637 + utils.py:136-171)."""
    from blindshadowremoval_b200.evaluate import evaluate_sfw
    n_chunks = 5
    ref = evaluate_sfw(_FakeGen(), _chunk, n_chunks, frame=2)
    assert ref["count"] == n_chunks and 0.5 < ref["auc"] <= 1.0 and ref["psnr"] > 0 and -1 <= ref["ssim"] <= 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, n_chunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, out in res:
        assert out["count"] == n_chunks
        for k in ("auc", "psnr", "ssim"):
            assert abs(out[k] - ref[k]) < 1e-9, (k, out[k], ref[k])
