"""CPU-side checks of the C ABI: the library loads, exports every symbol that
include/ngx.h declares, and refuses to compute without a GPU (no fallback)."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ngx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ngx_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from ngmix_b200 import _lib
    _lib.build()
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.ngx_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ngmix_b200 import _lib
    with pytest.raises(_lib.NgxError):
        _lib.gmix_fill("exp", [[0, 0, 0, 0, 1.0, 1.0]])


def test_product_does_not_import_oracle():
    """the product package must never import, link or call the oracle"""
    pkg = os.path.join(ROOT, "ngmix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src, f
                assert "liboracle" not in src, f
                assert "oracle.h" not in src, f


def test_exp_table_matches_reference():
    """the __constant__ exp(i) table is the reference's _EXP_LOOKUP"""
    import numpy as np
    from conftest import load_golden
    src = open(os.path.join(ROOT, "ngmix_b200", "csrc", "ngx_core.cuh")).read()
    body = src.split("c_exp_lookup[16] = {")[1].split("}")[0]
    vals = np.array([float(x) for x in body.replace("\n", " ").split(",")])
    assert np.array_equal(vals, load_golden("fastexp")["lookup"])
