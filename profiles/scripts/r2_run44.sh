#!/bin/bash
set -u
for g in 3.0 3.5 4.0 5.0; do echo "2 chunks growth $g"; PANDA_MSM_CHUNK_GROWTH=$g python profiles/scripts/streamed_times.py 24 2; done
echo "3 chunks default"; python profiles/scripts/streamed_times.py 24 3
for g in 3.0 4.0; do echo "3 chunks growth $g"; PANDA_MSM_CHUNK_GROWTH=$g python profiles/scripts/streamed_times.py 24 3; done
