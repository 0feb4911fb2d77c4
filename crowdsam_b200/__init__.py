"""crowdsam_b200 — B200-native (sm_100a) implementation of the Crowd-SAM inference hot path.

Public surface:
  crowdsam_b200.lib        ctypes binding of libcsam_sm100.so (include/csam.h)
  crowdsam_b200.ops        torch-tensor front end of the kernels
  crowdsam_b200.engine     encoder / DINOv2 / decoder orchestration
  crowdsam_b200.predictor  SamPredictor   (segment_anything_cs.SamPredictor surface)
  crowdsam_b200.pipeline   CrowdSAM       (crowdsam.model.CrowdSAM surface)
  crowdsam_b200.dropin     directory holding `segment_anything_cs` and `crowdsam` packages with the
                           reference's import names: put it on PYTHONPATH (or call install_dropin())
                           and tools/test.py / tools/batch_eval.py run unchanged.
"""
import os
import sys

__version__ = "0.1.0"
DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def install_dropin() -> str:
    """Make `import segment_anything_cs` / `import crowdsam` resolve to the B200 implementation."""
    if DROPIN_DIR not in sys.path:
        sys.path.insert(0, DROPIN_DIR)
    return DROPIN_DIR
