#!/bin/bash
export PYTHONPATH=$PWD
timeout 600 python scripts/run_config.py vit_l 64 2>&1 | tail -4
timeout 600 python scripts/run_config.py vit_b 8 2>&1 | tail -2
timeout 600 python scripts/run_config.py vit_h 32 2>&1 | tail -2
