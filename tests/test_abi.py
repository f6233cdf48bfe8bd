"""The C-ABI library loads and exports every function include/c2b200.h declares (no compute, no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    with open(os.path.join(ROOT, "include", "c2b200.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(c2b_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    from crispresso2_b200 import _lib
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_cuda_library_exports_every_symbol():
    import __graft_entry__ as g
    g.build()                                  # nvcc cross-compiles for sm_100a without a GPU
    from crispresso2_b200 import _lib
    lib = ctypes.CDLL(_lib.DEFAULT_LIB)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_missing_library_fails_loudly(tmp_path):
    from crispresso2_b200 import _lib
    with pytest.raises(_lib.LibraryMissing):
        _lib.load(str(tmp_path / "nope.so"))


def test_struct_sizes_match_header():
    from crispresso2_b200 import _lib
    assert _lib.ALN_DTYPE.itemsize == 32 and _lib.REC_DTYPE.itemsize == 16 and _lib.EDIT_DTYPE.itemsize == 8
    assert _lib.NVEC == 42 and _lib.NSCAL == 29 and _lib.NHIST == 6


def test_sass_is_sm100a_with_dpx_ops():
    """The shipped kernel is native sm_100a code: DPX integer ops (packed 16-bit and 32-bit), warp shuffles, and the
    TMA bulk copy (UBLKCP) + mbarrier (SYNCS) that stage the reference tile in shared memory."""
    import subprocess
    from crispresso2_b200 import _lib
    out = subprocess.run(["cuobjdump", "-sass", _lib.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "VIMNMX3" in out and "VIADDMNMX" in out and "SHFL" in out
    assert "VIMNMX3.S16x2" in out and "VIADDMNMX.S16x2" in out
    assert "UBLKCP" in out and "SYNCS" in out
