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
        n = int(rng.integers(0, 5)) if not (rank == 1 and i == 2) else 300     # one crowded image (CrowdHuman-like)
        if i == 1:
            out.append(None)            # an image with no detection at all
            continue
        out.append({"boxes": rng.uniform(0, 1000, (n, 4)).astype(np.float32), "scores": rng.uniform(0, 1, n).astype(np.float32),
                    "categories": np.zeros(n, dtype=np.int64)})
    return out


def _load_item(i):
    """(image, gt_boxes, image_id) as crowdsam.utils.load_img_and_annotation returns them."""
    img = np.full((8, 8, 3), i, dtype=np.uint8)
    return img, np.zeros((i + 2, 4)), f"img{i:03d}"


class _StubModel:
    """generate() with the MaskData surface tools/test.py:65-72 consumes; the count depends on the image."""

    def generate(self, image):
        i = int(image[0, 0, 0])
        n = 150 if i == 5 else i % 4
        rng = np.random.default_rng(i)
        return {"boxes": rng.uniform(0, 500, (n, 4)).astype(np.float32), "scores": rng.uniform(0, 1, n).astype(np.float32),
                "categories": np.zeros(n, dtype=np.int64), "points": np.zeros((n, 2)),
                "rles": [{"size": [8, 8], "counts": f"r{i}_{k}"} for k in range(n)]}


def _reference_item(i):
    """What tools/test.py:62-72 appends for image i (and batch_eval.py:17-28 concatenates in rank order)."""
    image, gt, image_id = _load_item(i)
    result = _StubModel().generate(image)
    d = {"image_id": image_id, "num_gt": len(gt) - 1}
    d.update({k: v.tolist() for k, v in result.items() if k in ["boxes", "scores", "categories"]})
    d.update({k: v for k, v in result.items() if k in ["rles"]})
    return d


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, end = parallel.shard_range(7, rank, world)       # ragged: rank 1 gets the remainder
    dets = _fake_dets(rank, end - start)
    gathered = parallel.gather_detections(dets, device="cpu")          # buffer width = exchanged global maximum
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
    # the launcher that replaces tools/batch_eval.py: per-rank slices, one gather, merged list on rank 0
    from crowdsam_b200 import batch_eval

    merged = batch_eval.run_sharded(_StubModel(), _load_item, 7, rank, world, device="cpu", seed=42)
    if rank == 0:
        ref = [_reference_item(i) for i in range(7)]
        ok &= merged == ref
    else:
        ok &= merged is None
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


def test_gather_never_truncates():
    """More detections than an explicit buffer width is an error, not a silently shortened list (round-1 ADVICE);
    the default sizes the buffer from the data."""
    import pytest

    dets = _fake_dets(1, 4)                      # image 2 of rank 1 holds 300 detections
    with pytest.raises(ValueError):
        parallel.gather_detections(dets, nmax=64)
    with pytest.raises(ValueError):
        parallel.pack_detections(dets, 64, "cpu")
    out = parallel.gather_detections(dets)
    assert len(out[0][2]["boxes"]) == 300 and np.array_equal(out[0][2]["boxes"], dets[2]["boxes"])


def test_run_sharded_single_process_and_coco():
    from crowdsam_b200 import batch_eval

    merged = batch_eval.run_sharded(_StubModel(), _load_item, 7, 0, 1)
    assert merged == [_reference_item(i) for i in range(7)]
    gt = {"images": [{"file_name": f"img{i:03d}.jpg", "id": i} for i in range(7)], "categories": [{"id": 1, "name": "person"}]}
    coco = batch_eval.convert_to_coco(merged, gt)
    n = sum(len(m["boxes"]) for m in merged)
    assert len(coco["annotations"]) == n and coco["annotations"][-1]["id"] == n - 1
    a, b = coco["annotations"][0], merged[1]["boxes"][0]
    assert a["image_id"] == "img001" and a["bbox"] == [b[0], b[1], b[2] - b[0], b[3] - b[1]]
