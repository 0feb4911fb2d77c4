"""The drop-in claim, executed: `CrowdSAM(config, logger)` built from checkpoint FILES through
`sam_model_registry['vit_l'](checkpoint=..., n_class=1)` + adapter `load_state_dict(strict=False)` + DINOv2 checkpoint
(reference model.py:33-42,88-115 <-> pipeline.py), driven by the statements of the reference's tools/test.py
(:14-35 environment, :37-58 setup, :62-72 per-image loop, :84-89 json.dump) with `crowdsam` / `segment_anything_cs`
resolved from crowdsam_b200/dropin.  The result must equal what the same pipeline gives with an injected predictor
holding the same weights."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import weights  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "crowdsam_b200", "dropin")

YAML = """
environ:
  seed: 42
  device: "cuda"
  output_dir: "{out}"
data:
  dataset: "crowdhuman"
  dataset_root: "{root}"
  json_file: "{root}/val_visible.json"
  odgt_file: "{root}/annotation_val.odgt"
model:
  dino_repo: "./dinov2"
  dino_checkpoint: "{ckpt}/dinov2_vitl14_pretrain.pth"
  dino_model: "dinov2_vitl14"
  sam_checkpoint: "{ckpt}/sam_vit_l_0b3195.pth"
  sam_model: "vit_l"
  sam_arch: "crowdsam"
  sam_adapter_checkpoint: "{ckpt}/10_shot.pth"
  n_class: 1
  max_size: 1024
  trainfree: False
test:
  output_rles: True
  crop_n_layers: 0
  crop_nms_thresh: 0.7
  crop_overlap_ratio: 0.341
  pos_sim_thresh: 0.5
  apply_box_offsets: False
  grid_size: 192
  max_prompts: 500
  filter_thresh: 0.7
  points_per_batch: 32
  mask_selection: "max_iou"
  max_size: 1024
  fuse_simmap: False
  min_mask_region_area: 100
  box_nms_thresh: 0.65
  stability_score_thresh: 0.8
  stability_score_offset: 1
  pred_iou_thresh: 0.1
vis:
  vis_thresh: 0.6
"""
ADAPTER = ("mask_decoder.dino_proj.", "mask_decoder.point_classifier.", "mask_decoder.parallel_iou_head.")
OPTIONS = ["test.grid_size", "8", "test.pos_sim_thresh", "-1", "test.max_prompts", "64", "test.filter_thresh", "2.0"]


def _write_fixture_files(tmp):
    import cv2

    ckpt, root, out = os.path.join(tmp, "weights"), os.path.join(tmp, "dataset"), os.path.join(tmp, "outputs")
    os.makedirs(ckpt); os.makedirs(os.path.join(root, "Images"))
    sam_sd, dino_sd = weights.make_sam_state("vit_l"), weights.make_dino_state("dinov2_vitl14")
    torch.save({k: v for k, v in sam_sd.items() if not k.startswith(ADAPTER)}, os.path.join(ckpt, "sam_vit_l_0b3195.pth"))
    torch.save({k[len("mask_decoder."):]: v for k, v in sam_sd.items() if k.startswith(ADAPTER)}, os.path.join(ckpt, "10_shot.pth"))
    torch.save(dino_sd, os.path.join(ckpt, "dinov2_vitl14_pretrain.pth"))
    images, annotations = [], []
    for i, hw in enumerate([(1024, 1024), (600, 900), (1024, 1024)]):
        img = weights.synthetic_image(3 + i, *hw)
        name = f"im{i}.png"
        cv2.imwrite(os.path.join(root, "Images", name), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
        images.append({"file_name": name, "id": 100 + i, "height": hw[0], "width": hw[1]})
        for j in range(i + 2):
            annotations.append({"image_id": 100 + i, "bbox": [10.0 * j, 5.0, 50.0, 80.0], "id": len(annotations)})
    json.dump({"images": images, "annotations": annotations, "categories": [{"id": 1, "name": "person"}]},
              open(os.path.join(root, "val_visible.json"), "w"))
    cfg_path = os.path.join(tmp, "crowdhuman.yaml")
    open(cfg_path, "w").write(YAML.format(out=out, root=root, ckpt=ckpt))
    return cfg_path


def test_tools_test_py_statements_on_the_dropin(tmp_path):
    cfg_path = _write_fixture_files(str(tmp_path))
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    for k in [k for k in sys.modules if k == "crowdsam" or k.startswith("crowdsam.")]:
        del sys.modules[k]
    # ---- tools/test.py:7-10
    from crowdsam.model import CrowdSAM
    from crowdsam.utils import (load_img_and_annotation, setup_logger, data_meta, load_config, modify_config,
                                visualize_result, evaluate_boxes)  # noqa: F401
    import crowdsam.model as cm

    assert cm.__file__.startswith(DROPIN)
    # ---- tools/test.py:26-35 (envrion_init)
    configs = load_config(cfg_path)
    configs = modify_config(configs, OPTIONS)
    np.random.seed(configs["environ"]["seed"])
    torch.random.manual_seed(configs["environ"]["seed"])
    os.makedirs(configs["environ"]["output_dir"], exist_ok=True)
    os.makedirs(configs["environ"]["output_dir"] + "/log", exist_ok=True)
    logger = setup_logger(configs["environ"]["output_dir"] + "/log")
    config = configs
    # ---- tools/test.py:41-58
    dataset_path = config["data"]["dataset_root"]
    n_class, class_names = data_meta[config["data"]["dataset"]][1:]
    if "cuda" in config["environ"]["device"]:
        torch.cuda.set_device(0)
        config["environ"]["device"] = "cuda:0"
    model = CrowdSAM(config, logger)
    annots = json.load(open(config["data"]["json_file"]))
    image_ids = list(range(0, len(annots["images"])))
    # ---- tools/test.py:60-72
    output_content = []
    for id_ in image_ids:
        image, gt_boxes, image_id = load_img_and_annotation(dataset_path, annots, config["data"]["dataset"], id_)
        result = model.generate(image)
        instance_dict = {"image_id": image_id, "num_gt": len(gt_boxes) - 1}
        instance_dict.update({k: v.tolist() for k, v in result.items() if k in ["boxes", "scores", "categories"]})
        instance_dict.update({k: v for k, v in result.items() if k in ["rles"]})
        output_content.append(instance_dict)
        del result
    # ---- tools/test.py:84-89
    file_path = os.path.join(config["environ"]["output_dir"], "result.json")
    json.dump(output_content, open(file_path, "w"), ensure_ascii=True)
    back = json.load(open(file_path))
    assert [b["image_id"] for b in back] == [100, 101, 102] and [b["num_gt"] for b in back] == [1, 2, 3]
    assert all(set(b) == {"image_id", "num_gt", "boxes", "scores", "categories", "rles"} for b in back)
    assert any(len(b["boxes"]) > 0 for b in back)
    for b in back:
        for r in b["rles"]:
            assert isinstance(r["counts"], str) and len(r["size"]) == 2
    # ---- the checkpoint-built model holds exactly the weights of the files, on the CUDA engines
    from crowdsam_b200.pipeline import CrowdSAM as B200CrowdSAM
    from test_gpu_model import make_predictor

    assert CrowdSAM is B200CrowdSAM and model.predictor.model.image_encoder.img_size == 1024
    assert model.predictor.model.mask_threshold == 0.0 and model.predictor.device.type == "cuda"
    pred, *_ = make_predictor("vit_l", "dinov2_vitl14")
    ref_model = B200CrowdSAM(config, logger, predictor=pred)
    np.random.seed(config["environ"]["seed"])
    for id_, got in zip(image_ids, output_content):
        image, _, _ = load_img_and_annotation(dataset_path, annots, config["data"]["dataset"], id_)
        ref = ref_model.generate(image)
        assert got["boxes"] == ref["boxes"].tolist() and got["scores"] == ref["scores"].tolist()
        assert [r["counts"] for r in got["rles"]] == [r["counts"] for r in ref["rles"]]
    # PIL input as tools/demo.py:48-49 passes it
    from PIL import Image

    np.random.seed(7)
    a = model.generate(Image.open(os.path.join(dataset_path, "Images", "im1.png")))
    np.random.seed(7)
    b = model.generate(load_img_and_annotation(dataset_path, annots, "crowdhuman", 1)[0])
    assert a["boxes"].tolist() == b["boxes"].tolist()
