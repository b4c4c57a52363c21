"""CPU-side checks of the C ABI: the library loads, exports every symbol include/fac.h declares, and
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fac.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fac_[a-z_0-9]+)\s*\(", src)) - {"fac_read_fn", "fac_write_fn", "fac_match_fn",
                                                                     "fac_replace_fn"})


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build_gpu_library()
    from fac_b200 import _abi
    return _abi.load_library()


def test_every_declared_symbol_is_exported(lib):
    from fac_b200 import _abi
    typed = {name for name, _, _ in _abi.SYMBOLS}
    for name in _declared():
        assert hasattr(lib, name), name
        assert name in typed, "ctypes mirror misses %s" % name


def test_abi_version(lib):
    assert lib.fac_abi_version() == 2


def test_struct_layouts():
    from fac_b200 import _abi
    assert C.sizeof(_abi.fac_match) == 32
    assert C.sizeof(_abi.fac_limits) == 10
    assert _abi.fac_match.similarity.offset == 20 and _abi.fac_match.edits.offset == 28


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from fac_b200 import FuzzyAhoCorasickBuilder, SearchError
    with pytest.raises(SearchError) as ei:
        FuzzyAhoCorasickBuilder.new().build(["abc"])
    assert ei.value.status == 3  # FAC_CUDA_ERROR
    assert "no CPU fallback" in str(ei.value)
