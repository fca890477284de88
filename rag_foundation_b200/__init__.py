"""rag_foundation_b200 -- B200-native store-scoped chunk retrieval for rag-foundation.

(The repository task names this package `rag-foundation_b200/`; Python cannot import a hyphenated
name, so the directory is `rag_foundation_b200/`.)

Only the reference's one data-parallel hot path lives here: chunk featurisation at ingest and
query-vs-chunk scoring + per-store top-k behind the GeminiRag-shaped adapter.  Compute is
hand-written sm_100a CUDA in librf_b200.so (csrc/), reached through the C-ABI in include/rf_b200.h.
There is no CPU fallback and nothing here imports oracle/.
"""
from .adapter import B200Rag, UploadResult, get_rag_client  # noqa: F401
from .engine import Engine, EngineGroup, unpack_keys  # noqa: F401

__all__ = ["B200Rag", "UploadResult", "get_rag_client", "Engine", "EngineGroup", "unpack_keys"]
