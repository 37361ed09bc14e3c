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
    # bias header + classic image (9 taps x 64 x 64 bf16) + the ring-walk image: 2 CTA ranks x (3 filter columns x 96 rows x 128 B
    # + the 32 x 64 B im2col tile of a 3-channel segment)
    assert lib.vsrb_packed_weight_bytes(C.byref(g)) == 1024 + 9 * 64 * 64 * 2 + 2 * (3 * 96 * 128 + 32 * 64)
    g.dtype = _lib.F32
    assert lib.vsrb_packed_weight_bytes(C.byref(g)) == 1024 + 9 * 64 * 64 * 4
    g.kh = 4                                            # even kernels are rejected with a message, not a crash
    assert lib.vsrb_packed_weight_bytes(C.byref(g)) == 0
    assert b"unsupported" in lib.vsrb_last_error()


def test_struct_layout_matches_header(lib):
    """ctypes mirrors of the C structs: sizes follow the C layout rules of the header."""
    from vsrlab_b200 import _lib
    assert C.sizeof(_lib.ConvGeom) == 16 * 4          # kh kw n_seg seg_c[4] seg_off[4] cout pixshuf groups dtype transpose
    assert C.sizeof(_lib.ConvArgs) % 8 == 0
    assert _lib.ConvArgs.packed.offset % 8 == 0 and _lib.ConvArgs.out_img_stride.offset % 8 == 0


def test_conv_plan_selection(lib):
    """The tiling the library picks per geometry (DESIGN.md 4.1 cost model): the hot 3x3 64->64 conv and the
    SPyNet 7x7 convs use the stacked layout, stems with 3 input channels and 1x1 / pixel-shuffle convs the classic one."""
    from vsrlab_b200 import _lib

    def plan(k, segs, cout, pixshuf=0):
        g = _lib.ConvGeom()
        g.kh = g.kw = k
        g.n_seg = len(segs)
        for i, c in enumerate(segs):
            g.seg_c[i] = c
        g.cout, g.groups, g.dtype, g.pixshuf = cout, 1, _lib.BF16, pixshuf
        info = (C.c_int32 * 8)()
        assert lib.vsrb_conv_plan_info(C.byref(g), info) == 0
        return list(info)
    stacked, n_tile, n_blocks, mma_n = plan(3, [64], 64)[:4]
    assert (stacked, n_tile, n_blocks, mma_n) == (1, 64, 1, 192)
    assert plan(3, [64, 3], 64)[:4] == [1, 64, 1, 192]          # cat([lr_i, feat]) stem: two K segments
    assert plan(3, [3], 64)[0] == 0                              # 16-channel (32-byte-row) stem stays classic
    assert plan(1, [64, 64], 64)[:4] == [0, 64, 1, 64]           # 1x1 fusion conv
    assert plan(3, [64], 256, pixshuf=2)[:4] == [0, 128, 2, 128]  # upsampling conv: N = 128, two blocks
    for segs, cout in (([8], 32), ([32], 64), ([64], 32), ([32], 16), ([16], 2)):
        info = plan(7, segs, cout)
        # stacked, N <= 256; a block's weights either stream (<= 128 KiB) or half of them stays resident per CTA of a pair
        assert info[0] == 1 and info[3] <= 256 and info[7] <= 200
    assert plan(7, [64], 32)[4] == 32 and plan(7, [32], 64)[4] == 32      # CTA-pair plans keep 64-byte K rows for 7x7
