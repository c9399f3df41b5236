"""libcl2.so loads on a CPU-only host and exports every symbol include/cl2.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from cloops2_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cl2.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cl2_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SIGNATURES.keys())


def test_library_exports_every_declared_symbol():
    _lib.build()
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    version = ctypes.cast(_lib.load().cl2_version(), ctypes.c_char_p).value
    assert b"sm_100a" in version
    assert ("src:" + _lib.source_hash()).encode() in version      # the shipped library was compiled from these sources


def test_context_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.Cl2Error):
        _lib.Context(0)
