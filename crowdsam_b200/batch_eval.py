"""`torchrun` launcher that replaces the reference's multi-GPU fan-out.

Reference: tools/batch_eval.py:60-103 starts one `python tools/test.py --start_idx a --end_idx b --local_rank r`
subprocess per GPU over contiguous image slices, each dumping `temp_result_{rank}.json`, then merges the files
(:17-28) and converts to COCO (:29-58).  Here ONE job does it:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        -m crowdsam_b200.batch_eval -c configs/crowdhuman.yaml [test.max_prompts 500 ...]

rank r builds one CrowdSAM on cuda:r, processes images `shard_range(n, r, W)` (the reference's slicing, last rank
takes the remainder) with the exact per-image loop of tools/test.py:62-72, the detection lists are exchanged with
one all-gather (`parallel.gather_detections`, NCCL; nothing truncated) and rank 0 writes `result.json` with the
schema of tools/test.py:66-72,84-89 — a list, in image order, of
{image_id, num_gt, boxes, scores, categories[, rles]}.  `--coco` additionally writes the COCO-format detection
file batch_eval.py:29-58 hands to the evaluation script (the evaluation itself is outside the hot path).
"""
from __future__ import annotations

import argparse
import json
import os
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import parallel


def run_sharded(model, load_item: Callable[[int], Tuple[np.ndarray, np.ndarray, object]], n_items: int,
                rank: int, world: int, device="cpu", seed: Optional[int] = None, group=None) -> Optional[List[Dict]]:
    """Process this rank's slice and gather.  Returns the merged per-image list on rank 0, None elsewhere.
    `model.generate(image)` must return the MaskData-like mapping of crowdsam.model.CrowdSAM.generate."""
    start, end = parallel.shard_range(n_items, rank, world)
    if seed is not None:
        # tools/test.py:29-30 seeds once per process; every subprocess of the reference starts from the same seed
        np.random.seed(seed)
        torch.random.manual_seed(seed)
    dets, metas = [], []
    for id_ in range(start, end):
        image, gt_boxes, image_id = load_item(id_)
        result = model.generate(image)                                          # tools/test.py:65
        keys = dict(result.items())
        dets.append({"boxes": np.asarray(keys["boxes"], dtype=np.float32).reshape(-1, 4),
                     "scores": np.asarray(keys["scores"], dtype=np.float32).reshape(-1)[: len(keys["boxes"])],
                     "categories": np.asarray(keys.get("categories", np.zeros(len(keys["boxes"]))))})
        meta = {"image_id": image_id, "num_gt": len(gt_boxes) - 1}             # tools/test.py:68
        if "rles" in keys:
            meta["rles"] = keys["rles"]
        metas.append(meta)
    gathered = parallel.gather_detections(dets, nmax=None, device=device, group=group)   # the one tensor exchange
    all_meta = parallel.gather_objects(metas, group=group)
    if rank != 0:
        return None
    flat = [d for per_rank in gathered for d in per_rank]
    assert len(flat) == len(all_meta) == n_items
    out = []
    for d, m in zip(flat, all_meta):
        item = {"image_id": m["image_id"], "num_gt": m["num_gt"], "boxes": d["boxes"].tolist(),
                "scores": d["scores"].tolist(), "categories": d["categories"].tolist()}
        if "rles" in m:
            item["rles"] = m["rles"]
        out.append(item)
    return out


def convert_to_coco(det_result: List[Dict], gt_js: Dict) -> Dict:
    """batch_eval.py:29-58: XYXY -> XYWH annotations with running ids, image ids = file name stems."""
    images = gt_js["images"]
    for im in images:
        im["id"] = im["file_name"][:-4]
    annotations, next_id = [], 0
    for k, item in enumerate(det_result):
        image_id = images[k]["id"] if images != [] else item["image_id"]
        for score, box in zip(item["scores"], item["boxes"]):
            x0, y0, x1, y1 = box
            annotations.append({"category_id": 1, "bbox": [x0, y0, x1 - x0, y1 - y0], "image_id": image_id,
                                "iscrowd": False, "area": (y1 - y0) * (x1 - x0), "id": next_id, "score": score})
            next_id += 1
    return {"images": images, "annotations": annotations, "categories": gt_js["categories"]}


def main(argv=None) -> int:
    from .dropin.crowdsam import utils
    from .pipeline import CrowdSAM

    ap = argparse.ArgumentParser(description="CrowdSAM multi-GPU evaluation (one torchrun job)")
    ap.add_argument("-c", "--config_file", default="./configs/crowdhuman.yaml")
    ap.add_argument("-s", "--save_path", default="")
    ap.add_argument("--coco", default="", help="also write the COCO-format detection file here")
    ap.add_argument("options", nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)
    config = utils.modify_config(utils.load_config(args.config_file), args.options)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("crowdsam_b200.batch_eval needs CUDA devices (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    config["environ"]["device"] = f"cuda:{local_rank}"                           # tools/test.py:44-47
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    out_dir = config["environ"]["output_dir"]
    os.makedirs(os.path.join(out_dir, "log"), exist_ok=True)
    logger = utils.setup_logger(os.path.join(out_dir, "log"), quiet=rank != 0)
    model = CrowdSAM(config, logger)
    annots = json.load(open(config["data"]["json_file"]))
    root, dataset = config["data"]["dataset_root"], config["data"]["dataset"]
    merged = run_sharded(model, lambda i: utils.load_img_and_annotation(root, annots, dataset, i),
                         len(annots["images"]), rank, world, device=f"cuda:{local_rank}",
                         seed=config["environ"]["seed"])
    if rank == 0:
        path = args.save_path or os.path.join(out_dir, "result.json")
        json.dump(merged, open(path, "w"), ensure_ascii=True)                    # tools/test.py:84-89
        print(f"dump json file to {path}")
        if args.coco:
            json.dump(convert_to_coco(merged, annots), open(args.coco, "w"), ensure_ascii=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
