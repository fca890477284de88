#!/bin/bash
# Second profiler pass of round 2 (under gpurun, one GPU): the batched kernels after this round's work and the
# 1024-feature instantiations.  Plain run first (must exit 0); numbers printed under ncu are never bench values.
# usage: bash tools/ncu_r02b.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity --legs cfg2,wide"
$CMD > gpurun_out/ncu_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/ncu_launches_$tag.csv $CMD > gpurun_out/ncu_l_$tag.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:score_topk_gemm_pair -s 2 -c 2 -f -o gpurun_out/ncu_gemm_pair_$tag $CMD > gpurun_out/ncu_g_$tag.log 2>&1
echo "pair gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:score_topk_gemm_wide -s 2 -c 2 -f -o gpurun_out/ncu_gemm_wide_$tag $CMD > gpurun_out/ncu_w_$tag.log 2>&1
echo "wide gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:score_topk_scan_tma_kernel<6, 12, 4>" -s 10 -c 2 -f -o gpurun_out/ncu_scan_wide_$tag $CMD > gpurun_out/ncu_sw_$tag.log 2>&1
echo "wide scan capture rc=$?"
ls -la gpurun_out | grep $tag
