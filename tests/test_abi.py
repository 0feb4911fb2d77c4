"""CPU-side checks of the C-ABI boundary: the shared library builds/loads without a GPU and exports every
symbol that include/csam.h declares; ctypes struct layouts match the header field order."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "csam.h")).read()


def test_library_exports_every_declared_symbol():
    from crowdsam_b200 import lib

    if not os.path.exists(lib.LIB_PATH):
        lib.build()
    l = lib.load()
    declared = set(re.findall(r"CSAM_API\s+[\w\s\*]+?\b(csam_\w+)\s*\(", _header()))
    assert len(declared) >= 25
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    for name in declared:
        assert hasattr(l, name), name
    assert l.csam_abi_version() == int(re.search(r"#define CSAM_ABI_VERSION (\d+)", _header()).group(1))
    assert l.csam_launch_count() == 0 or l.csam_launch_count() > 0     # callable without a device


@pytest.mark.parametrize("cname,pyname", [("csam_gemm_args", "GemmArgs"), ("csam_ln_args", "LnArgs"),
                                          ("csam_attn_args", "AttnArgs"), ("csam_dec_attn_args", "DecAttnArgs"),
                                          ("csam_post_args", "PostArgs")])
def test_struct_field_order_matches_header(cname, pyname):
    from crowdsam_b200 import lib

    body = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\} " + cname + ";", _header(), re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.replace("*", " ").split(","):
            fields.append(re.findall(r"(\w+)\s*$", part.strip())[0])
    py = [f[0] for f in getattr(lib, pyname)._fields_]
    assert fields == py, (fields, py)


def test_no_product_import_of_oracle():
    """The product must never route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "crowdsam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dp, f)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from crowdsam_b200 import lib

    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError):
        lib.load()


def test_cpu_device_is_rejected():
    """No CPU fallback: engines refuse to run on a CPU model."""
    from crowdsam_b200.build import _build_sam

    sam = _build_sam(128, 2, 2, 1, (1,))
    with pytest.raises(RuntimeError):
        sam.image_encoder.engine()
