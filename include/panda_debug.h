/*
 * panda_debug.h -- diagnostic entry points of libpanda-cuda (B200 implementation).  Not part of the reference's
 * ABI; used by tests/ (to check the device field / curve arithmetic against the oracle one operation at a time)
 * and by bench.py (per-stage timings, integer-pipe peak).  All pointers are DEVICE pointers unless noted.
 */
#ifndef PANDA_DEBUG_H
#define PANDA_DEBUG_H

#include "panda_interface.h"

#ifdef __cplusplus
extern "C" {
#endif

/* field ids: 0 BN254 Fq, 1 BN254 Fr, 2 BLS12-377 Fq, 3 BLS12-377 Fr, 4 BLS12-381 Fq, 5 BLS12-381 Fr (same numbering as oracle/panda_oracle.c;
 * field 5 is 255 bits wide: canonical operands only, no PANDA_FOP_INV -- see Bls381Fr in csrc/field.cuh) */
enum panda_debug_field_opcode {
    PANDA_FOP_MUL = 0, PANDA_FOP_ADD = 1, PANDA_FOP_SUB = 2, PANDA_FOP_SQR = 3, PANDA_FOP_FROM_MONT = 4,
    PANDA_FOP_TO_MONT = 5, PANDA_FOP_INV = 6, PANDA_FOP_NEG = 7
};
/* out[i] = a[i] (op) b[i], canonical Montgomery limbs; b is ignored by unary ops */
panda_error panda_debug_field_op(int field_id, int op, const void *a, const void *b, void *out, size_t count, panda_stream stream);

/* curve ids: 0 BN254, 1 BLS12-377, 2 BLS12-381.  Points are Jacobian triples (x||y||z) except q of MADD (affine x||y). */
enum panda_debug_curve_opcode {
    PANDA_COP_MADD = 0,        /* out = p + q(affine), through the XYZZ mixed addition the MSM uses */
    PANDA_COP_ADD = 1,         /* out = p + q, through the XYZZ addition */
    PANDA_COP_DBL_XYZZ = 2,    /* out = 2p, XYZZ doubling */
    PANDA_COP_DBL_JAC = 3,     /* out = 2p, Jacobian dbl-2009-l */
    PANDA_COP_TO_HOMOGENEOUS = 4
};
panda_error panda_debug_curve_op(int curve_id, int op, const void *p, const void *q, void *out, size_t count, panda_stream stream);

typedef struct panda_debug_msm_plan_info {
    unsigned window_bits, windows, buckets_per_window, segment_len, segments_per_window, reduce_chunk;
    size_t workspace_bytes;
    unsigned folded, bucket_sets, groups, phases;   /* folded = 1: plan for a precomputed 2^(c*j)*P table (one bucket set) */
    size_t table_bytes;
} panda_debug_msm_plan_info;
/* the plan msm_execute would use for n points (c_override / seg_override = 0: automatic); folded selects the table plan */
panda_error panda_debug_msm_plan(int curve_id, size_t n, int folded, unsigned c_override, unsigned seg_override, panda_debug_msm_plan_info *out);

/* MSM with explicit window width / segment length / table mode and per-stage device times.
 * table_mode: -1 default (env PANDA_MSM_PRECOMPUTE, else 3), 0 never use tables, 1 auto (unannounced bases get a table after their second
 * sighting with the same fingerprint), 2 eager (table at first sight), 3 tables only for registered base sets.
 * stage_ms (HOST float[7], may be NULL): digits, scan, scatter, accumulate, bucket_reduce, window_reduce, final.
 * info (HOST unsigned[3], may be NULL): folded, window bits, windows actually used.
 * Synchronises the stream when stage_ms or info is given. */
panda_error panda_debug_msm_timed(int curve_id, const panda_msm_configuration cfg, size_t n, unsigned c_override, unsigned seg_override,
                                  int table_mode, float *stage_ms, unsigned *info);

/* panda_msm_execute_*_host_scalars with an explicit table mode and chunk count (0 = automatic). */
panda_error panda_debug_msm_streamed(int curve_id, const panda_msm_configuration cfg, size_t n, int table_mode, unsigned chunks);

/* Integer-pipe microbenchmarks on the current device (synchronous).
 * kind 0: independent IMAD (32-bit) chains, 1: independent IMAD.WIDE chains, 2: Montgomery products (BN254 Fq) in
 * 4 independent chains per thread.  Fills *ms (device time of the timed launch) and *ops (instructions of that kind /
 * modular products executed). */
panda_error panda_debug_int_peak(int kind, unsigned iters, float *ms, unsigned long long *ops);

/* panda_ntt_execute_bn254_v1 / panda_intt_execute_bn254_v1 with the device time of every pass (HOST float[4], unused entries 0).
 * Synchronises the stream. */
panda_error panda_debug_ntt_timed(const panda_ntt_configuration_v1 cfg, int inverse, float *pass_ms);

/* out = omega^(2^k) for a BN254 Fr element, HOST pointers, Montgomery in / out: the host-side helper the multi-GPU NTT derives its
 * sub-roots with (no device involved) */
panda_error panda_debug_fr_pow2k_host(const void *omega, unsigned k, void *out);

#ifdef __cplusplus
}
#endif
#endif /* PANDA_DEBUG_H */
