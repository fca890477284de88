#!/bin/bash
# One GPU-box pass of the profiler (under gpurun, one GPU): the launch list of the bench command and
# full captures of the three kernels the bench legs spend their time in.  The plain run comes first and
# must exit 0; numbers printed under ncu are never bench values.
# usage: bash tools/ncu_r02.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/ncu_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches_$tag.csv $CMD > gpurun_out/ncu_l_$tag.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:score_topk_scan -s 10 -c 2 -f -o gpurun_out/ncu_scan_$tag $CMD --no-configs > gpurun_out/ncu_s_$tag.log 2>&1
echo "scan capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:score_topk_gemm_pair -s 2 -c 2 -f -o gpurun_out/ncu_gemm_pair_$tag $CMD > gpurun_out/ncu_g_$tag.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:tokenize|rows_from_tokens" -s 2 -c 2 -f -o gpurun_out/ncu_ingest_$tag $CMD > gpurun_out/ncu_i_$tag.log 2>&1
echo "ingest capture rc=$?"
ls -la gpurun_out | tail -12
