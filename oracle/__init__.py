"""CPU oracle for the Crowd-SAM inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product (``crowdsam_b200``) never imports this
package and fails loudly when its CUDA library is missing.

Contents
--------
``weights.py``    synthetic "recipe v1" state_dicts shared by the
                  oracle, the real reference (when importable) and the CUDA path.
``restate.py``    plain PyTorch fp32 restatement of the reference algorithm,
                  every function citing the reference file:line it follows.
``ref_import.py`` imports the *real* reference from /root/reference with the
                  three shims of SURVEY.md §8c (only where that tree exists).

Parity pin: ``tests/golden/*.npz`` were produced by running the real reference
(``tests/golden/make_golden.py``, committed) and ``tests/test_oracle_golden.py``
checks ``restate.py`` against them, so the oracle is pinned to the reference.
The COCO compressed-RLE string (pycocotools, absent from this image and from
the reference tree) is restated from the published COCO API algorithm and is
"parity unpinned" for that one field.
"""
