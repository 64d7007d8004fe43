"""The C-ABI boundary: libmcl.so loads and exports every symbol include/mcl.h declares; the ctypes table covers the
header; without a GPU the product path fails loudly (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

import pymcl_b200  # noqa: F401
from pymcl_b200 import _lib, build
from conftest import HAS_GPU, REPO


def header_symbols():
    src = open(os.path.join(REPO, "include", "mcl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcl_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    build.build()  # no-op when libmcl.so is newer than its sources
    assert os.path.isfile(_lib.LIB_PATH)
    ctypes.CDLL(_lib.LIB_PATH)


def test_every_header_symbol_is_exported_and_typed():
    syms = header_symbols()
    assert len(syms) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libmcl.so does not export %s" % s
    assert set(syms) == set(_lib.SIGNATURES), (set(syms) ^ set(_lib.SIGNATURES))


def test_header_cites_the_reference_interface_it_replaces():
    src = open(os.path.join(REPO, "include", "mcl.h")).read()
    for needle in ("Matcher.py:147-163", "Matcher.py:262-291", "Matcher.py:169-194", "Matcher.py:196-230",
                   "Matcher.py:232-259", "analyze.py:441-445", "analyze.py:312-336"):
        assert needle in src


@pytest.mark.skipif(HAS_GPU, reason="a GPU is present")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(_lib.MclError) as e:
        _lib.Context(0)
    assert "no CPU fallback" in str(e.value) or "failed" in str(e.value)
    from pymcl_b200.Matcher import Matcher
    m = Matcher("SIFT", width=160, height=120)
    m.setIndex({"map/0/angle000.png": (None, None)})
    m.setDirectory("map/0")
    m.setQueryDescriptors(None)
    with pytest.raises(_lib.MclError):
        m.run()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "py-mcl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(oracle|pipeline|ref_shims)\b", text, flags=re.M), f
