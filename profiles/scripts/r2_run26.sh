#!/bin/bash
set -u
echo "== e2e 2^24: scatter cap x digits cap"
for sc in 148 296; do for dc in 148 296 592 1184; do echo "scatter $sc digits $dc"; PANDA_MSM_SIDE_CTAS=$sc PANDA_MSM_SIDE_DIGITS=$dc python profiles/scripts/streamed_times.py 24 3,4; done; done
PANDA_MSM_SIDE_CTAS=148 PANDA_MSM_SIDE_DIGITS=148 PANDA_MSM_TRACE=1 python profiles/scripts/streamed_times.py 24 4 2>&1 | tail -29
