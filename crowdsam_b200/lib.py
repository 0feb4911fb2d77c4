"""ctypes binding of libcsam_sm100.so (include/csam.h).

The library is the product's only compute path.  There is no fallback: if the shared
object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSAM_LIB_PATH", os.path.join(_HERE, "_C", "libcsam_sm100.so"))
_lib = None

vp, ci, cf, cll = C.c_void_p, C.c_int, C.c_float, C.c_longlong


class GemmArgs(C.Structure):
    _fields_ = [("a_hi", vp), ("a_lo", vp), ("w_hi", vp), ("w_lo", vp),
                ("M", ci), ("N", ci), ("K", ci), ("lda", ci), ("ldw", ci),
                ("bias", vp), ("row_scale", vp), ("col_scale", vp), ("act", ci),
                ("residual", vp), ("ldr", ci), ("res_mod", ci),
                ("row_map", vp),
                ("out_f32", vp), ("ldo", ci),
                ("out_hi", vp), ("out_lo", vp), ("ldh", ci),
                ("impl", ci), ("b_mn_major", ci),
                ("epi", ci), ("gamma", vp), ("beta", vp), ("eps", cf),
                ("pe", vp), ("ldpe", ci), ("pe_mod", ci), ("out2_hi", vp), ("out2_lo", vp),
                ("hyper", vp), ("masks", vp),
                ("res_hi", vp), ("res_lo", vp), ("ldrh", ci)]


class LnArgs(C.Structure):
    _fields_ = [("x", vp), ("ldx", ci), ("rows_in", ci),
                ("add", vp), ("ldadd", ci), ("add_mod", ci),
                ("row_map", vp), ("rows_out", ci), ("cols", ci),
                ("gamma", vp), ("beta", vp), ("eps", cf), ("normalize", ci),
                ("out_f32", vp), ("ldo", ci),
                ("out_hi", vp), ("out_lo", vp), ("ldh", ci),
                ("pe", vp), ("ldpe", ci), ("pe_mod", ci), ("out2_hi", vp), ("out2_lo", vp),
                ("act", ci)]


class AttnArgs(C.Structure):
    _fields_ = [("qkv_hi", vp), ("qkv_lo", vp), ("ld_qkv", ci),
                ("groups", ci), ("tokens", ci), ("heads", ci), ("hd", ci), ("scale", cf),
                ("rel_h", vp), ("rel_w", vp), ("S", ci),
                ("out_hi", vp), ("out_lo", vp), ("ld_out", ci),
                ("scratch", vp), ("scratch_bytes", cll),
                ("impl", ci), ("p_split", ci)]


class DecAttnArgs(C.Structure):
    _fields_ = [("q", vp), ("Bq", ci), ("k", vp), ("v", vp), ("Bk", ci),
                ("B", ci), ("nq", ci), ("nk", ci), ("heads", ci), ("hd", ci),
                ("out_f32", vp), ("out_hi", vp), ("out_lo", vp),
                ("ldq", ci), ("ldk", ci), ("ldv", ci)]


class I2TArgs(C.Structure):
    _fields_ = [("x_hi", vp), ("x_lo", vp), ("x_shared", ci),
                ("peq_hi", vp), ("peq_lo", vp), ("b1_hi", vp), ("b1_lo", vp), ("b2_hi", vp), ("b2_lo", vp),
                ("P", ci), ("bias", vp), ("gamma", vp), ("beta", vp), ("eps", cf),
                ("out_hi", vp), ("out_lo", vp)]


class T2IArgs(C.Structure):
    _fields_ = [("x_hi", vp), ("x_lo", vp), ("x_shared", ci), ("pek_hi", vp), ("pek_lo", vp),
                ("b1_hi", vp), ("b1_lo", vp), ("P", ci), ("xbar", vp), ("wv_t", vp), ("bv", vp),
                ("out_f32", vp), ("out_hi", vp), ("out_lo", vp)]


class PostArgs(C.Structure):
    _fields_ = [("low", vp), ("P", ci), ("sel", vp), ("planes", ci),
                ("in_h", ci), ("in_w", ci), ("out_h", ci), ("out_w", ci),
                ("thr", cf), ("off", cf),
                ("counts", vp), ("boxes", vp),
                ("keep", vp), ("n_keep", ci),
                ("masks", vp), ("logits", vp)]


_SIGS = {
    "csam_last_error": (C.c_char_p, []),
    "csam_abi_version": (ci, []),
    "csam_launch_count": (cll, []),
    "csam_gemm": (ci, [C.POINTER(GemmArgs), vp]),
    "csam_patchify": (ci, [vp, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp]),
    "csam_layernorm": (ci, [C.POINTER(LnArgs), vp]),
    "csam_vit_attention_scratch_bytes": (cll, [ci, ci, ci, ci, ci]),
    "csam_vit_attention": (ci, [C.POINTER(AttnArgs), vp]),
    "csam_im2col3x3": (ci, [vp, vp, ci, ci, vp, vp, vp]),
    "csam_transpose_f32": (ci, [vp, ci, ci, vp, vp]),
    "csam_bilinear": (ci, [vp, ci, ci, ci, vp, ci, ci, ci, vp]),
    "csam_prompt_tokens": (ci, [vp, vp, ci, vp, vp, vp, vp, vp, vp]),
    "csam_attn_few_keys": (ci, [C.POINTER(DecAttnArgs), vp]),
    "csam_attn_few_queries": (ci, [C.POINTER(DecAttnArgs), vp]),
    "csam_dec_fold_i2t": (ci, [vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp]),
    "csam_dec_i2t_layer": (ci, [C.POINTER(I2TArgs), vp]),
    "csam_dec_fold_t2i": (ci, [vp, ci, vp, vp, vp, vp]),
    "csam_dec_t2i": (ci, [C.POINTER(T2IArgs), vp]),
    "csam_upscale_shuffle_ln_gelu": (ci, [vp, ci, vp, vp, cf, vp, vp, vp]),
    "csam_upscale_hyper_masks": (ci, [vp, ci, vp, vp, vp]),
    "csam_softmax_weights": (ci, [vp, ci, ci, vp, vp, vp, vp]),
    "csam_select_candidates": (ci, [vp, vp, ci, ci, vp, vp, vp, vp]),
    "csam_mask_post_stats": (ci, [C.POINTER(PostArgs), vp]),
    "csam_mask_post_write": (ci, [C.POINTER(PostArgs), vp]),
    "csam_box_nms_scratch_bytes": (cll, [ci]),
    "csam_box_nms": (ci, [vp, vp, ci, cf, vp, vp, vp, cll, vp]),
    "csam_mask_overlap_scratch_bytes": (cll, [ci]),
    "csam_mask_overlap": (ci, [vp, ci, ci, ci, vp, vp, vp, cll, vp]),
    "csam_points_occupied": (ci, [vp, ci, ci, ci, vp, vp, ci, vp, vp]),
    "csam_small_regions_scratch_bytes": (cll, [ci, ci, ci]),
    "csam_remove_small_regions": (ci, [vp, ci, ci, ci, ci, ci, vp, vp, cll, vp]),
    "csam_rle_scratch_bytes": (cll, [ci, ci, ci]),
    "csam_rle_count": (ci, [vp, ci, ci, ci, vp, vp, cll, vp]),
    "csam_rle_fill": (ci, [vp, ci, ci, ci, vp, vp, ci, vp, vp, vp, vp]),
    "csam_coco_rle_strings": (ci, [vp, vp, ci, vp, cll, vp]),
}

EXPORTS = tuple(_SIGS.keys())


def build(verbose: bool = False) -> str:
    """Compile libcsam_sm100.so in-tree with nvcc for sm_100a (works without a GPU)."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    res = subprocess.run(["bash", script], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libcsam_sm100.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def load():
    """Load the shared library (no CUDA call is made here, so this works on a CPU-only host)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(crowdsam_b200 has no fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)       # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.csam_abi_version() != 10:
        raise RuntimeError("libcsam_sm100.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().csam_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libcsam {what}: {msg}")


def launch_count() -> int:
    return int(load().csam_launch_count())
