"""Run the reference's OWN adapter tests (backend/tests/test_gemini_rag.py) with B200Rag standing
in for GeminiRag.  Only where /root/reference exists (the build container); the GPU box skips it.

The reference module is imported with the same 3-line sqlalchemy stub as
tests/golden/make_reference_golden.py; the tests that exercise the real Gemini HTTP client
(retry/REST fallback) are not about this path and are left out."""
import importlib.util
import os
import sys
import types

import pytest

REF = "/root/reference"
REF_TEST = os.path.join(REF, "backend", "tests", "test_gemini_rag.py")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_TEST), reason="reference checkout not present")

WANTED = [
    "test_extract_citations_handles_empty_response",
    "test_extract_citations_handles_missing_metadata",
    "test_extract_citations_handles_missing_chunks",
    "test_extract_citations_returns_valid_structure",
    "test_extract_citations_logs_warning_on_parsing_error",
    "test_new_stream_ids_returns_unique_ids",
]


@pytest.fixture(scope="module")
def ref_test_class():
    os.environ.setdefault("ENVIRONMENT", "test")
    os.environ.setdefault("GEMINI_MOCK_MODE", "true")
    os.environ.setdefault("JWT_SECRET", "x" * 64)
    os.environ.setdefault("GEMINI_API_KEY", "fake-key-for-tests")
    if "sqlalchemy" not in sys.modules:
        sa = types.ModuleType("sqlalchemy"); eng = types.ModuleType("sqlalchemy.engine")
        url = types.ModuleType("sqlalchemy.engine.url"); url.make_url = lambda s: s
        sys.modules.update({"sqlalchemy": sa, "sqlalchemy.engine": eng, "sqlalchemy.engine.url": url})
    sys.path.insert(0, os.path.join(REF, "backend"))
    try:
        spec = importlib.util.spec_from_file_location("ref_test_gemini_rag", REF_TEST)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception as exc:   # a missing optional dependency of the reference
        pytest.skip(f"reference test module does not import here: {exc}")
    finally:
        sys.path.remove(os.path.join(REF, "backend"))
    from rag_foundation_b200.adapter import B200Rag
    mod.GeminiRag = B200Rag          # the drop-in: the reference's tests now exercise our adapter
    return mod.TestGeminiRag


@pytest.mark.parametrize("name", WANTED)
def test_reference_adapter_test_passes_against_b200rag(ref_test_class, name, caplog):
    inst = ref_test_class()
    fn = getattr(inst, name)
    if "caplog" in fn.__code__.co_varnames[:fn.__code__.co_argcount]:
        fn(caplog)
    else:
        fn()
