"""C-ABI hygiene (no GPU needed): the library builds for sm_100a, loads, and exports exactly
the symbols include/srfdet_b200.h declares; the ctypes prototypes cover all of them."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from srfdet_b200 import build
    build.build()
    from srfdet_b200 import _lib
    return _lib.load()


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'srfdet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return set(re.findall(r'\b(srf_[a-z0-9_]+)\s*\(', src))


def test_header_matches_prototypes(lib):
    from srfdet_b200 import _lib
    assert _header_symbols() == set(_lib.PROTOTYPES)


def test_library_exports_every_symbol(lib):
    from srfdet_b200 import _lib
    out = subprocess.check_output(['nm', '-D', '--defined-only', _lib.SO_PATH], text=True)
    exported = set(re.findall(r'\bT (srf_[a-z0-9_]+)', out))
    assert _header_symbols() <= exported, _header_symbols() - exported


def test_host_only_entry_points(lib):
    """Calls that never touch a device."""
    from srfdet_b200 import _lib as L
    assert lib.srf_version() >= 100
    g = L.make_geom([0.075, 0.075, 0.2], [-55.2, -55.2, -5.0, 55.2, 55.2, 3.0])
    assert list(g.grid) == [1472, 1472, 40]
    g = L.make_geom([0.05, 0.05, 0.1], [0, -40, -3, 70.4, 40, 1])
    assert list(g.grid) == [1408, 1600, 40]
    for vs, pc in [([0.075, 0.075, 0.2], [-55.2, -55.2, -5.0, 55.2, 55.2, 3.0]), ([0.05, 0.05, 0.1], [0, -40, -3, 70.4, 40, 1]),
                   ([0.1, 0.1, 0.15], [-76.8, -76.8, -2, 76.8, 76.8, 4]), ([0.3, 0.7, 0.25], [-1.05, 0, 0, 1.05, 4.55, 1.125])]:
        a, b = L.make_geom(vs, pc), L.make_geom_c(vs, pc)       # host arithmetic == srf_geom_init
        assert list(a.grid) == list(b.grid) and list(a.vs) == list(b.vs) and list(a.lo) == list(b.lo) and list(a.hi) == list(b.hi)
    assert lib.srf_index_bytes(41 * 1472 * 1472) > 2 * 41 * 1472 * 1472 // 8
    assert lib.srf_hard_voxelize_ws_bytes(300000, 10, 160000) > 0
    assert lib.srf_linear_tile_k(6272) == 128 and lib.srf_linear_tile_n(32) == 32
    # argument validation reports through srf_last_error, no exception crosses the boundary
    rc = lib.srf_index_clear(None, 0, None)
    assert rc == -1 and b'srf_index_clear' in lib.srf_last_error()


def test_sass_contains_tcgen05():
    """The sparse-conv / GEMM kernel really is a tcgen05 kernel (UTCHMMA + TMEM loads)."""
    from srfdet_b200 import _lib
    sass = subprocess.check_output(['cuobjdump', '-sass', _lib.SO_PATH], text=True)
    assert 'UTCHMMA' in sass and 'LDTM' in sass and 'LDGSTS' in sass


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under srfdet_b200/ may reference it."""
    for dp, _, files in os.walk(os.path.join(ROOT, 'srfdet_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt and 'srf_oracle' not in txt, f
