#!/bin/bash
# round 2, call 9: BLS12-381 + new host API tests, reference GPU kernels diagnostic, resident chunking at 2^25 / 2^26
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py tests/test_gpu_field_curve.py -m gpu -q -k "bls12_381 or field_ops or curve_ops or host_api" ) > $OUT/r2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/r2_pytest9.log
python - <<'PY' 2>&1 | tail -12
import sys, subprocess, os
sys.path.insert(0, '.')
import numpy as np, oracle as O
for k in (16, 20):
    n = 1 << k
    b = O.gen_bases(0, O.seed_for(k), n); s = O.gen_scalars(1, O.seed_for(k) + 1, n)
    try:
        rj, best, allms = O.ref_gpu_msm(b, s, k, reps=3)
        exp = O.jac_to_affine(0, O.expected_progression_msm(0, O.seed_for(k), s, n))
        print("ref gpu", k, best, allms, bool((O.jac_to_affine(0, rj) == exp).all()))
    except Exception as e:
        print("ref gpu", k, "ERR", e)
PY
for sp in 1 2; do echo "== 2^25 split $sp"; PANDA_MSM_SPLIT=$sp python profiles/scripts/stage_times.py 25 2>&1 | tail -1 | cut -c1-160; done
for sp in 2 8; do echo "== 2^26 split $sp"; PANDA_MSM_SPLIT=$sp python profiles/scripts/stage_times.py 26 2>&1 | tail -1 | cut -c1-160; done
for sp in 2; do echo "== 2^24 split $sp"; PANDA_MSM_SPLIT=$sp python profiles/scripts/stage_times.py 24 2>&1 | tail -1 | cut -c1-160; done
