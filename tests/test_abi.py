"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/vangan_b200.h declares;
the ctypes signature table covers exactly those symbols.  No compute call is made (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "vangan_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(vg_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from van_gan_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "missing export: %s" % n
    assert lib.vg_abi_version() == 4


def test_ctypes_table_matches_header():
    from van_gan_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_no_cpu_fallback_paths():
    """the product package never imports the oracle and raises when the library is missing"""
    pkg = os.path.join(ROOT, "van-gan_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            with open(os.path.join(pkg, fn)) as f:
                src = f.read()
            assert "import oracle" not in src and "from oracle" not in src, fn
    from van_gan_b200 import _lib
    saved, _lib._lib = _lib._lib, None
    real = _lib.LIB_PATH
    _lib.LIB_PATH = real + ".missing"
    try:
        try:
            _lib.lib()
            raised = False
        except _lib.VgError:
            raised = True
        assert raised
    finally:
        _lib.LIB_PATH, _lib._lib = real, saved


def test_descriptor_struct_layouts():
    from van_gan_b200 import _lib
    assert ctypes.sizeof(_lib.ConvDesc) == 13 * 4
    assert ctypes.sizeof(_lib.InDesc) == 64     # 12 x 4 bytes + 8-byte seed (aligned) + 8-byte seed_dev pointer
    lib = _lib.lib()
    d = _lib.ConvDesc(1, 10, 10, 10, 16, 16, 3, 1, _lib.VG_BF16, _lib.VG_BF16, 0)
    assert lib.vg_conv3d_packed_bytes(d, 0) > 27 * 16 * 16 * 2
    assert lib.vg_conv3d_packed_bytes(d, 1) > 27 * 16 * 16 * 2
    bad = _lib.ConvDesc(1, 10, 10, 10, 16, 16, 5, 1, _lib.VG_BF16, _lib.VG_BF16, 0)
    assert lib.vg_conv3d_packed_bytes(bad, 0) == 0
    assert lib.vg_soft_skel_bwd_workspace_bytes(1, 8, 8, 8) == 6 * 512 * 4 + 2 * 64 * 4   # six volumes + the per-level maxima
