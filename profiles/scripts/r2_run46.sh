#!/bin/bash
set -u
PANDA_MSM_CHUNK_GROWTH=3.2 PANDA_MSM_TRACE=1 python profiles/scripts/streamed_times.py 24 3 2>&1 | tail -42 | grep -v "uploaded"
