#!/bin/bash
set -u
PANDA_MSM_TRACE=1 python profiles/scripts/streamed_times.py 24 4 2>&1 | tail -60
