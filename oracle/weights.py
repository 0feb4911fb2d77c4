"""Synthetic weights / images shared by every side of a parity run.  TEST INFRASTRUCTURE.

The recipe itself lives in crowdsam_b200/synthetic.py (pure data generation: the product's bench needs it too and
must not import oracle/); this module re-exports it so that the real reference (tests/golden/make_golden.py), the
oracle restatement and the CUDA path load the same state_dict."""
from crowdsam_b200.synthetic import (DINO_ARCHS, SAM_ARCHS, make_dino_state, make_sam_state,  # noqa: F401
                                     synthetic_image)
