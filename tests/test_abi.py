"""The C-ABI library loads and exports exactly what include/pdsb.h declares (no compute calls:
this runs without a GPU)."""
import ctypes
import os
import re

import pytest

from pdspy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pdsb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdsb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for needed in ("pdsb_sample_image", "pdsb_loglike", "pdsb_loglike_batch", "pdsb_chi2", "pdsb_chisq",
                   "pdsb_grid", "pdsb_freqcorrect", "pdsb_dataset_create"):
        assert needed in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_every_declared_symbol():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_version_and_error_string_without_gpu():
    L = _lib.load()
    assert L.pdsb_version() == 1
    assert isinstance(L.pdsb_last_error(), bytes)


def test_bad_arguments_are_reported_not_crashed():
    L = _lib.load()
    assert L.pdsb_device_count(None) == 1            # PDSB_ERR_ARG
    assert b"count" in L.pdsb_last_error()


def test_no_cpu_fallback():
    """Without a device every compute entry point must fail loudly."""
    L = _lib.load()
    n = ctypes.c_int(-1)
    rc = L.pdsb_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    assert L.pdsb_init(0) != 0
    with pytest.raises(_lib.PdsbError):
        _lib.check(L.pdsb_synchronize())
    import numpy as np
    from pdspy_b200.interferometry import chisq, Visibilities
    z = np.zeros((4, 1))
    d = Visibilities(np.zeros(4), np.zeros(4), np.ones(1), z, z, z + 1)
    with pytest.raises(_lib.PdsbError):
        chisq(d, d)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under pdspy_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pdspy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "liboracle" not in txt, f
