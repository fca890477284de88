"""bench.py host-side pieces that need no GPU: the query generator the GPU arm uses equals the
oracle's, and the CPU (reference) arm prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_queries_equal_oracle_queries():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import c_oracle as co, rf1
    zb = rf1.zipf_bucket_table()
    Q = bench.make_queries(16)
    for i in range(16):
        assert (Q[i] == co.synth_query(bench.SEED, i, zb)).all()


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "chunks/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("configs[1]") and line["higher_is_better"] is True


def test_b200_arm_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
