"""The two Prometheus series the reference's adapter feeds (backend/app/metrics.py:6-8, observed at
services/gemini_rag.py:330-352, 509-515, 544-551, 626-629, 668-671, 686-694): calls by
(operation, status) and latency by operation, with the reference's own operation labels
("upload", "generate", "generate_stream").

Inside the backend process `app.metrics` is loaded and B200Rag feeds the reference's own collector objects,
so /metrics needs no change.  Standalone (the engine daemon, tests) the same series names live in this
module's own CollectorRegistry (`REGISTRY`; never the default one, where they would collide with the
reference's as soon as `app.metrics` is imported); without prometheus_client the hooks are no-ops.
"""
from __future__ import annotations

import sys


class _NoOp:
    def labels(self, *a, **k):
        return self

    def inc(self, *a, **k):
        pass

    def observe(self, *a, **k):
        pass


try:
    from prometheus_client import CollectorRegistry, Counter, Histogram
    REGISTRY = CollectorRegistry()
    _own = {"gemini_calls_total": Counter("gemini_api_calls_total", "Gemini API calls", ["operation", "status"], registry=REGISTRY),
            "gemini_latency": Histogram("gemini_api_latency_seconds", "Gemini API latency", ["operation"], registry=REGISTRY)}
    SOURCE = "standalone"
except Exception:   # noqa: BLE001
    REGISTRY = None
    _own = {"gemini_calls_total": _NoOp(), "gemini_latency": _NoOp()}
    SOURCE = "disabled"


class _Series:
    """Resolves at each use: the reference's collector when `app.metrics` is loaded in this process, else ours."""

    def __init__(self, name: str):
        self._name = name

    def labels(self, *a, **k):
        ref = sys.modules.get("app.metrics")
        target = getattr(ref, self._name, None) if ref is not None else None
        return (target if target is not None else _own[self._name]).labels(*a, **k)


gemini_calls_total = _Series("gemini_calls_total")
gemini_latency = _Series("gemini_latency")
