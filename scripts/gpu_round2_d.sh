#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python scripts/debug_determinism.py tiny_l 3 2>&1 | tail -8
python scripts/debug_determinism.py vit_l 2 2>&1 | tail -6
timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool memcheck python scripts/debug_determinism.py tiny_l 2 2>&1 | grep -E "rep |ERROR SUMMARY" | tail -8
python -m pytest tests -m gpu -q -k "mask_iou_nms_as_selection or mask_post_vs_oracle or pair" 2>&1 | tail -8
