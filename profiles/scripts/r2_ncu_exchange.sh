#!/bin/bash
# ncu --set full of the exchange kernel on the closing code: one plain transpose launch and one twiddle + transpose launch (2^12 x 2^13, one GPU)
set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python tests/run_exchange.py 12 13"
$CMD > $OUT/r2_plain_exchange.log 2>&1; tail -3 $OUT/r2_plain_exchange.log
ncu --set full --clock-control none --import-source on -k regex:'k_ntt_exchange' -s 12 -c 2 -o $OUT/r2_prof_exchange -f $CMD > $OUT/r2_ncu_exchange.log 2>&1
ls -la $OUT/r2_prof_exchange.ncu-rep
