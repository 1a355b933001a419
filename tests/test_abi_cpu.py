"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200seg.h declares
(no compute calls without a GPU); descriptor structs mirror the header."""
import ctypes
import os
import re

import pytest

from ct_image_segmentation_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    header = open(os.path.join(ROOT, "include", "b200seg.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(b200seg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_error_string(lib):
    assert lib.b200seg_version() == 100
    assert isinstance(lib.b200seg_last_error(), bytes)
    assert lib.b200seg_launch_count() == 0


def test_descriptor_layout():
    assert ctypes.sizeof(_lib.ConvDesc) == 23 * 4
    assert ctypes.sizeof(_lib.NormDesc) == 40 and _lib.NormDesc.spatial.offset == 8
    assert ctypes.sizeof(_lib.DiceDesc) == 32 and _lib.DiceDesc.ld.offset == 16


def test_argument_validation_without_gpu(lib):
    d = _lib.ConvDesc(1, 16, 16, 4, 4, 4, 4, 4, 5, 3, 3, 3, 1, 1, 1, 1, 1, 1, 16, 16, 0, 0, 0)
    rc = lib.b200seg_conv_fprop(ctypes.byref(d), 1, 1, None, None, 1, None)
    assert rc == -1 and b"inconsistent" in lib.b200seg_last_error()
    assert lib.b200seg_packed_weight_bytes(ctypes.byref(d), 0) == 27 * 16 * 16 * 4
    assert lib.b200seg_conv_wgrad_workspace_bytes(ctypes.byref(d)) > 0


def test_product_has_no_oracle_import():
    pkg = os.path.join(ROOT, "ct_image_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"
