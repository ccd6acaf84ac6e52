"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/bann.h declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "bann.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bann_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import rs_bann_b200 as rb
    lib = ctypes.CDLL(rb.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bann.h but not exported"
    # and the Python binding declares a prototype for each of them
    assert set(syms) == set(rb.PROTOTYPES), set(syms) ^ set(rb.PROTOTYPES)


def test_no_cpu_fallback():
    import rs_bann_b200 as rb
    if rb.cuda_available():
        pytest.skip("CUDA device present")
    with pytest.raises(rb.BannError, match="no CPU fallback"):
        rb.Context()


def test_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rs-bann_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


def test_selector_constants_match_the_header():
    """The Python mirror of the test / profiling selectors (Net.K1_*, Net.HMC_*, Net.TC_*) carries the header's enum values."""
    import rs_bann_b200 as rb
    src = open(os.path.join(ROOT, "include", "bann.h")).read()
    enums = dict((k, int(v)) for k, v in re.findall(r"\b(BANN_[A-Z0-9_]+)\s*=\s*(\d+)", src))
    for py, c in (("K1_AUTO", "BANN_K1_AUTO"), ("K1_TENSOR", "BANN_K1_TENSOR"), ("K1_FFMA", "BANN_K1_FFMA"), ("K1_GENERIC", "BANN_K1_GENERIC"),
                  ("HMC_AUTO", "BANN_HMC_AUTO"), ("HMC_LAUNCHES", "BANN_HMC_LAUNCHES"), ("HMC_PERSISTENT", "BANN_HMC_PERSISTENT"),
                  ("TC_FOUR_WARPS", "BANN_TC_FOUR_WARPS"), ("TC_FIVE_WARPS", "BANN_TC_FIVE_WARPS"),
                  ("TC_FIVE_WARPS_PLAIN", "BANN_TC_FIVE_WARPS_PLAIN")):
        assert getattr(rb.Net, py) == enums[c], (py, c)
