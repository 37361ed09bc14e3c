"""The C-ABI library loads without a GPU and exports every symbol include/vsrb200.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from vsrlab_b200 import build, _lib
    build.build()
    return _lib.load()


def header_symbols():
    text = (ROOT / "include" / "vsrb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsrb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from vsrlab_b200 import _lib
    names = header_symbols()
    assert len(names) >= 14
    assert sorted(_lib.SYMBOLS.keys()) == names
    for n in names:
        assert getattr(lib, n) is not None


def test_version_and_geometry_helpers_run_on_cpu(lib):
    from vsrlab_b200 import _lib
    assert lib.vsrb_version() == 100
    g = _lib.ConvGeom()
    g.kh = g.kw = 3
    g.n_seg = 1
    g.seg_c[0] = 64
    g.cout = 64
    g.groups = 1
    g.dtype = _lib.BF16
    assert lib.vsrb_packed_weight_bytes(C.byref(g)) == 1024 + 9 * 64 * 64 * 2
    g.dtype = _lib.F32
    assert lib.vsrb_packed_weight_bytes(C.byref(g)) == 1024 + 9 * 64 * 64 * 4
    g.kh = 4                                            # even kernels are rejected with a message, not a crash
    assert lib.vsrb_packed_weight_bytes(C.byref(g)) == 0
    assert b"unsupported" in lib.vsrb_last_error()


def test_struct_layout_matches_header(lib):
    """ctypes mirrors of the C structs: sizes follow the C layout rules of the header."""
    from vsrlab_b200 import _lib
    assert C.sizeof(_lib.ConvGeom) == 12 * 4
    assert C.sizeof(_lib.ConvArgs) % 8 == 0
    assert _lib.ConvArgs.packed.offset % 8 == 0 and _lib.ConvArgs.out_img_stride.offset % 8 == 0
