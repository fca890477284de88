"""The C-ABI library loads, exports every symbol include/rf_b200.h declares, and fails loudly
(no CPU fallback) when no GPU is visible.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from rag_foundation_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rf_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rf_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 18
    L = _capi.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in rf_b200.h but not exported by librf_b200.so"
    assert sorted(_capi.SIGNATURES) == names, "ctypes binding and header disagree"


def test_struct_layout_matches_header():
    assert C.sizeof(_capi.rf_config) == 32
    assert C.sizeof(_capi.rf_stats) == 80


def test_library_holds_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_error_strings_and_build_info():
    L = _capi.lib()
    assert L.rf_strerror(0) == b"ok"
    assert b"busy" in L.rf_strerror(_capi.RF_EBUSY)
    buf = C.create_string_buffer(256)
    assert L.rf_build_info(buf, 256) == 0 and b"sm_100a" in buf.value


def test_error_code_mapping():
    with pytest.raises(TimeoutError):   # retryable in the reference: gemini_rag.py:22-27
        _capi.check(_capi.RF_EBUSY)
    with pytest.raises(RuntimeError):
        _capi.check(_capi.RF_ECUDA)
    _capi.check(_capi.RF_OK)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rag_foundation_b200 import Engine
    with pytest.raises(RuntimeError) as ei:
        Engine(capacity_rows=1024)
    assert "no CUDA device" in str(ei.value) or "code -4" in str(ei.value)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: the package and the development tools never touch it
    (measurement scripts that need it as a checker live under tests/perf/)."""
    for top in ("rag_foundation_b200", "tools", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src, f
                    assert "librf1_oracle" not in src, f


def test_rf1w_host_functions_match_the_oracle():
    """rf_idf_weights / rf_weight_query are pure host functions of the C-ABI: callable without a GPU."""
    import ctypes as C

    import numpy as np

    from oracle import c_oracle as co
    from rag_foundation_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(11)
    for n in (0, 1, 77, 10_000, 10 ** 8):
        df = rng.integers(0, n + 1, 256).astype(np.uint64)
        w = np.zeros(256, np.uint8)
        assert L.rf_idf_weights(df.ctypes.data, n, 256, w.ctypes.data) == 0
        assert (w == co.idf_weights(df, n)).all()
        q = rng.integers(0, 9, 256).astype(np.int8)
        qw = np.zeros(256, np.int8)
        assert L.rf_weight_query(q.ctypes.data, w.ctypes.data, 256, qw.ctypes.data) == 0
        assert (qw == co.weight_query(q, w)).all()
    bad = np.full(256, 5, np.uint64)
    assert L.rf_idf_weights(bad.ctypes.data, 4, 256, w.ctypes.data) == _capi.RF_EINVAL   # df > n
