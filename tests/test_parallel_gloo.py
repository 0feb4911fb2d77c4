"""N > 1 host logic on CPU: world_size 2 over gloo (rendezvous on 127.0.0.1)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crowdsam_b200 import parallel


def test_shard_range_matches_reference_slicing():
    # tools/batch_eval.py:82-89: batch = n // world, last rank takes the remainder
    assert [parallel.shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 10)]
    assert [parallel.shard_range(64, r, 8) for r in range(8)] == [(8 * r, 8 * r + 8) for r in range(8)]
    assert parallel.shard_range(3, 0, 1) == (0, 3)
    covered = sorted(i for r in range(3) for i in range(*parallel.shard_range(7, r, 3)))
    assert covered == list(range(7))


def _fake_dets(rank, n_images):
    rng = np.random.default_rng(100 + rank)
    out = []
    for i in range(n_images):
        n = int(rng.integers(0, 5))
        if i == 1:
            out.append(None)            # an image with no detection at all
            continue
        out.append({"boxes": rng.uniform(0, 1000, (n, 4)).astype(np.float32), "scores": rng.uniform(0, 1, n).astype(np.float32),
                    "categories": np.zeros(n, dtype=np.int64)})
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, end = parallel.shard_range(7, rank, world)       # ragged: rank 1 gets the remainder
    dets = _fake_dets(rank, end - start)
    gathered = parallel.gather_detections(dets, nmax=8, device="cpu")
    ok = len(gathered) == world
    for r in range(world):
        s, e = parallel.shard_range(7, r, world)
        ref = _fake_dets(r, e - s)
        ok &= len(gathered[r]) == e - s
        for g, d in zip(gathered[r], ref):
            if d is None:
                ok &= len(g["boxes"]) == 0
            else:
                ok &= np.array_equal(g["boxes"], d["boxes"]) and np.array_equal(g["scores"], d["scores"])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_gather_single_process():
    out = parallel.gather_detections(_fake_dets(0, 3), nmax=8)
    assert len(out) == 1 and len(out[0]) == 3
