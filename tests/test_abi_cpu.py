"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/adipose_b200.h declares; no compute call is made (no GPU here)."""
import os
import re
import sys

import numpy as np
import pytest

import adipose_unet_b200 as A
from adipose_unet_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _lib.load()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "adipose_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adp_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    syms = _header_symbols()
    assert len(syms) >= 25
    assert sorted(_lib.SIGNATURES) == syms
    for s in syms:
        assert hasattr(lib, s), f"library does not export {s}"


def test_abi_version_and_no_gpu_behaviour(lib):
    assert lib.adp_abi_version() == 3
    import ctypes as C
    if lib.adp_device_count() == 0:
        h = C.c_void_p()
        rc = lib.adp_create(0, 1, 44, 16, C.byref(h))
        assert rc == -3, "adp_create must fail loudly without an sm_100 device (no CPU fallback)"
        assert b"no CUDA device" in lib.adp_last_error() or b"sm_100" in lib.adp_last_error()


def test_tta_op_tables(lib):
    import ctypes as C
    from oracle import geometry as G
    ops = (C.c_int * 8)()
    for mode, name in ((1, "minimal"), (2, "basic"), (3, "full")):
        n = lib.adp_tta_ops(mode, ops)
        assert list(ops[:n]) == G.TTA_OPCODES[name]
    from adipose_unet_b200 import api
    assert api.TTA_OPCODES == G.TTA_OPCODES
    assert api._INV == G.D4_INVERSE
    x = np.arange(25, dtype=np.float32).reshape(5, 5)
    for op in range(8):
        np.testing.assert_array_equal(api._aug_fn(op)(x), G.d4_apply(op, x))


def test_sass_is_blackwell_native():
    """The conv kernel must contain tcgen05 MMA, TMEM loads and TMA loads (SASS mnemonics of
    /opt/skills/guides/B200_PROFILING.md)."""
    import shutil, subprocess
    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cu):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cu, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP"):
        assert mnem in sass, mnem


def test_host_geometry_matches_oracle():
    from adipose_unet_b200 import api
    from oracle import geometry as G
    for (h, w, t, ov) in [(1024, 1024, 1024, 0.5), (1500, 2000, 1024, 0.5), (2560, 3072, 1024, 0.75),
                          (300, 520, 128, 0.9), (1000, 1000, 1024, 0.5), (32768, 32768, 1024, 0.5)]:
        sw = api.SlidingWindowInference(t, ov, "none", verbose=False)
        assert sw.stride == G.stride_for(t, ov)
        assert sw.extract_tile_positions((h, w)) == G.tile_positions(h, w, t, sw.stride)
    np.testing.assert_array_equal(api.GaussianBlender.__new__(api.GaussianBlender).__class__(64).weight_map,
                                  G.gaussian_window(64))
    m = api.metrics_from_counts(10, 3, 2, 85)
    assert m == G.metrics_from_counts(10, 3, 2, 85)
    assert api.metrics_from_counts(0, 0, 0, 7) == G.metrics_from_counts(0, 0, 0, 7)


def test_bench_help_renders():
    """argparse formats help strings with %: an unescaped per-cent sign makes `bench.py --help` raise."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-400:]
    assert "--max-forwards" in r.stdout and "--impl" in r.stdout
