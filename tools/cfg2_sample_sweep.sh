#!/bin/bash
# configs[2] (1024 queries x 1 M chunks) against the size of the floor pass's sample (RF_GEMM_SAMPLE: rows for the single-CTA
# kernel; the pair kernel takes twice as many).  usage (under gpurun): bash tools/cfg2_sample_sweep.sh
for S in ${SWEEP:-32768 65536 131072}; do
RF_GEMM_SAMPLE=$S python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --legs cfg2 > gpurun_out/cfg2_s$S.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/cfg2_s$S.json')); c=d['configs']['cfg2']; print('RF_GEMM_SAMPLE $S', c['ms_per_batch'], c['roofline']['frac'])"
done
