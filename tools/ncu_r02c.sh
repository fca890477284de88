#!/bin/bash
# Third profiler pass of round 2 (under gpurun, one GPU): DRAM traffic of the scan kernel at the shard sizes of the
# multi-GPU bench lines (100 M chunks over 1 / 2 / 4 / 8 GPUs) and in the configs[4] leg, for roofline.traffic at
# N > 1 (profiles/roofline_traffic.json).  Plain run first (must exit 0); numbers printed under ncu are never bench values.
# usage: bash tools/ncu_r02c.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for n in 12500000 25000000 50000000 100000000; do
  CMD="python bench.py --chunks $n --steps 8 --warmup 3 --no-cpu-baseline --no-parity --no-configs --streams 1"
  $CMD > gpurun_out/ncu_plain_${tag}_$n.log 2>&1 || { echo "plain run failed ($n)"; tail -5 gpurun_out/ncu_plain_${tag}_$n.log; exit 1; }
  ncu $M --clock-control none -k regex:score_topk_scan_tma -s 70 -c 2 --csv --log-file gpurun_out/ncu_traffic_${tag}_$n.csv $CMD > gpurun_out/ncu_t_${tag}_$n.log 2>&1
  echo "shard $n rc=$?"
done
CMD="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --legs cfg4"
$CMD > gpurun_out/ncu_plain_${tag}_cfg4.log 2>&1 || { echo "plain run failed (cfg4)"; exit 1; }
ncu $M --clock-control none -k regex:score_topk_scan_tma -s 200 -c 400 --csv --log-file gpurun_out/ncu_traffic_${tag}_cfg4.csv $CMD > gpurun_out/ncu_t_${tag}_cfg4.log 2>&1
echo "cfg4 rc=$?"
ls -la gpurun_out | grep ncu_traffic_$tag
