#!/bin/bash
# sustained-load experiment: 40 back-to-back table-mode MSMs with per-stage device times, clocks / power sampled every 20 ms
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1f}
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 20 > $OUT/clocks_$TAG.csv &
SMI=$!
python tests/run_msm.py 24 40 0 0 0 2 > $OUT/sustained_$TAG.log 2>&1
kill $SMI
grep -o "accum[^]]*]=\[[^]]*\]" $OUT/sustained_$TAG.log | awk -F'[][,]' '{print $(NF-4)}' | tr '\n' ' '; echo
awk -F', ' '{print $2, $4, $5, $6}' $OUT/clocks_$TAG.csv | sort | uniq -c | sort -rn | head -30
