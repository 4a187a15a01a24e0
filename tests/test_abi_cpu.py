"""CPU-side checks of the drop-in boundary: the C-ABI library builds in-tree, loads, exports every
symbol include/gokalman_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gokalman_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gkb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_surface():
    syms = _declared_symbols()
    for must in ("gkb_create_lti", "gkb_create_information_from_state", "gkb_create_hybrid", "gkb_create_srif",
                 "gkb_update", "gkb_nl_run", "gkb_mc_chisquare", "gkb_set_measurement_matrix", "gkb_set_noise",
                 "gkb_reset", "gkb_destroy", "gkb_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from gokalman_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = C.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert set(declared) == bound, (set(declared) ^ bound)


def test_shape_table_and_version():
    import gokalman_b200 as gk
    lib = gk.load()
    assert b"sm_100a" in lib.gkb_version()
    for n, m in ((2, 1), (3, 1), (4, 1), (4, 2), (6, 2)):  # every BASELINE config shape of the register kernels
        for kind in range(6):
            assert lib.gkb_shape_supported(kind, n, m) == 1
    for n in (16, 24, 32, 40, 48, 56, 64):  # large-state Vanilla (kernels_tile.cu); no other kind, no other n
        assert lib.gkb_shape_supported(0, n, 8) == 1 and lib.gkb_shape_supported(0, n, 1) == 1
        assert lib.gkb_shape_supported(2, n, 8) == 0 and lib.gkb_shape_supported(0, n, 9) == 0
    assert lib.gkb_shape_supported(0, 44, 8) == 0 and lib.gkb_shape_supported(0, 72, 8) == 0 and lib.gkb_shape_supported(0, 128, 8) == 0
    for kind in range(6):  # every kind reaches the north star's n <= 8
        assert lib.gkb_shape_supported(kind, 8, 3) == 1 and lib.gkb_shape_supported(kind, 7, 1) == 1
        assert lib.gkb_shape_supported(kind, 9, 1) == 0
    assert lib.gkb_shape_supported(4, 8, 2) == 1 and lib.gkb_shape_supported(5, 7, 2) == 1 and lib.gkb_shape_supported(4, 9, 2) == 0


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_gpu():
    """The product path must fail loudly when no CUDA device is usable."""
    import gokalman_b200 as gk
    import fixtures as fx
    f = fx.jerk3()
    with pytest.raises(gk.GkbError) as ei:
        gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(f["Q"], f["R"]))
    assert ei.value.code == -9
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under gokalman_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gokalman_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "gko_" not in txt, (fn, "references the oracle")
                assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_python_mirror_validates_dimensions_without_touching_the_gpu():
    """checkMatDims failures (vanilla.go:23-31) are raised before any device call."""
    import gokalman_b200 as gk
    import fixtures as fx
    f = fx.jerk3()
    noise = gk.NewNoiseless(f["Q"], f["R"])
    with pytest.raises(gk.GkbError) as ei:
        gk.NewVanilla(np.zeros(2), f["P0"], f["F"], f["G"], f["H"], noise)
    assert ei.value.code == -1
    with pytest.raises(ValueError):
        gk.NewNoiseless(None, f["R"])  # noise.go:30-32 panics
    with pytest.raises(ValueError):
        gk.NewAWGN(-np.eye(2), np.eye(1))  # noise.go:149-151 panics ("process noise invalid")
