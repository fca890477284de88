#!/bin/bash
# One GPU-box pass: the -m gpu tests, the default bench line (with its configs legs) and the CPU arm.
# usage (under gpurun): bash tools/gpu_check.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_pytest_$tag.log
tail -8 gpurun_out/r2_pytest_$tag.log
( time python bench.py --steps 200 --warmup 20 > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err )
tail -3 gpurun_out/r2_bench_$tag.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref_$tag.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_$tag.json"))
print(json.dumps({k: v for k, v in d.items() if k != "configs"})[:2600])
print(json.dumps(d.get("configs"), indent=1)[:7000])
r = json.load(open("gpurun_out/r2_ref_$tag.json"))
print("reference arm:", r["value"], "cores", r["cpu_baseline"]["cores"], "same config:", r["config"] == d["config"])
PY
