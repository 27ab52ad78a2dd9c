"""The C-ABI shared library loads and exports every symbol include/sspslam_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    with open(os.path.join(ROOT, "include", "sspslam_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_python_binding_covers_the_header(lib):
    from sspslam_b200 import cabi
    assert sorted(cabi.EXPORTS) == _declared()


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.ssb_version()
    assert isinstance(lib.ssb_last_error(), bytes)


def test_bad_arguments_are_reported_not_crashed(lib):
    from sspslam_b200 import cabi
    rc = lib.ssb_set_scalar(None, b"dt", 0.001)
    assert rc < 0 and b"ssb_set_scalar" in lib.ssb_last_error()
    assert lib.ssb_n_steps(None) == -1
    rc = lib.ssb_ssp_encode(0, None, None, None, 1, 2, 7)
    assert rc < 0
    try:
        cabi.check(rc, "ssb_ssp_encode")
    except cabi.SsbError as e:
        assert "bad arguments" in str(e)
    else:
        raise AssertionError("SsbError expected")


def test_library_is_in_tree_and_self_contained():
    from sspslam_b200 import cabi
    assert os.path.dirname(cabi.LIB_PATH) == os.path.join(ROOT, "semantic-spiking-neural-slam-2023_b200")
    ctypes.CDLL(cabi.LIB_PATH)   # loads without torch / python symbols


def test_builder_library_exports_its_header(lib):
    """libssb_builder.so (include/sspslam_b200_builder.h) loads without a GPU and exports what its header declares."""
    from sspslam_b200 import cabi
    with open(os.path.join(ROOT, "include", "sspslam_b200_builder.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", text)))
    blib = cabi.load_builder()
    assert declared == sorted(cabi.BUILDER_EXPORTS)
    for name in declared:
        assert hasattr(blib, name)
    assert isinstance(blib.ssb_builder_last_error(), bytes)
