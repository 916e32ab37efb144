/*
 * panda_interface.h -- C ABI of libpanda-cuda (B200 / sm_100a implementation).
 *
 * Drop-in for the reference's src/cuda/core/panda_interface.cuh:10-110, i.e. exactly the symbols the Rust side
 * binds in src/gpu_ffi/binding.rs:3-115 with the #[repr(C)] types of src/gpu_ffi/common.rs:40-208.
 * Plain pointers and sizes only; every function returns panda_error: 0 on success, otherwise the raw
 * cudaError_t value (reference convention, panda_interface.cu:11-191; Rust only tests `!= 0`).
 *
 * Device pointers are owned by the caller.  MSM inputs (bases, scalars) are never written: the reference converts
 * scalars in place (msm_cuda.cuh:155, msm_host.cuh:293-296) and thereby corrupts cached scalars on their second
 * use -- this implementation does not.  The NTT entry points use d_src and d_dst as ping-pong storage exactly like the
 * reference's pass loop (fft.cu:193-211): after the call the buffer *flag names holds the result and the other one is
 * scratch (d_src keeps its contents only for single-pass sizes, log_n <= 8).
 * MSM execution is asynchronous on cfg.stream (the Rust caller records + syncs an event on it afterwards,
 * gpu_manager/unit.rs:60-62) and never blocks the host; temporaries are allocated stream-ordered from cfg.mem_pool.
 * All calls act on the calling thread's current device (the reference hard-codes device 0,
 * msm_cuda.cuh:554-555).
 */
#ifndef PANDA_INTERFACE_H
#define PANDA_INTERFACE_H

#include <stddef.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* panda_interface.cuh:10-16 */
typedef enum panda_error {
    panda_success = 0,
    panda_error_invalid_value = 1,
    panda_error_memory_allocation = 2,
    panda_error_not_ready = 600
} panda_error;

/* panda_interface.cuh:18-31; gpu_ffi/common.rs:40-44, 89-93, 134-138 -- raw cudaStream_t / cudaEvent_t / cudaMemPool_t, by value */
typedef struct panda_stream { void *handle; } panda_stream;
typedef struct panda_event { void *handle; } panda_event;
typedef struct panda_mem_pool { void *handle; } panda_mem_pool;

/* panda_interface.cuh:33-37; gpu_ffi/common.rs:168-173 */
typedef enum panda_msm_result_coordinate_type {
    JACOBIAN = 0,   /* (X, Y, Z), x = X/Z^2, y = Y/Z^3 -- what arkworks' G1Projective::new expects (tests/test.rs:87-102) */
    PROJECTIVE      /* homogeneous (X*Z, Y, Z^3), projective.cuh:66-77 */
} panda_msm_result_coordinate_type;

typedef void (*panda_host_fn)(void *user_data);

/* ---- device / stream / event / memory plumbing: panda_interface.cuh:40-68, panda_interface.cu:11-150 ---- */
panda_error panda_get_device_number(int *count);
panda_error panda_get_device(int *device_id);
panda_error panda_set_device(int device_id);
panda_error panda_stream_create(panda_stream *stream, bool blocking_sync);
panda_error panda_stream_wait_event(panda_stream stream, panda_event event);
panda_error panda_stream_sync(panda_stream stream);          /* spelling defined by the reference C side (panda_interface.cu:36-39) */
panda_error panda_stream_synchronize(panda_stream stream);   /* spelling the Rust side links against (binding.rs:14); undefined in the reference */
panda_error panda_stream_query(panda_stream stream);         /* binding.rs:16; declared but undefined in the reference */
panda_error panda_stream_destroy(panda_stream stream);
panda_error panda_launch_host_fn(panda_stream stream, panda_host_fn fn, void *user_data);
panda_error panda_event_create(panda_event *event, bool blocking_sync, bool disable_timing);
panda_error panda_event_record(panda_event event, panda_stream stream);
panda_error panda_event_sync(panda_event event);
panda_error panda_event_query(panda_event event);
panda_error panda_event_destroy(panda_event event);
panda_error panda_mem_get_info(size_t *free, size_t *total);
panda_error panda_malloc(void **ptr, size_t size);
panda_error panda_malloc_host(void **ptr, size_t size);
panda_error panda_free(void *ptr);
panda_error panda_free_host(void *ptr);
panda_error panda_host_register(void *ptr, size_t size);
panda_error panda_host_unregister(void *ptr);
panda_error panda_device_disable_peer_access(int device_id); /* binding.rs:54; undefined in the reference */
panda_error panda_device_enable_peer_access(int device_id);  /* binding.rs:56; undefined in the reference */
panda_error panda_memcpy(void *dst, const void *src, size_t count);
panda_error panda_memcpy_async(void *dst, const void *src, size_t count, panda_stream stream);
panda_error panda_memset(void *ptr, int value, size_t count);
panda_error panda_memset_async(void *ptr, int value, size_t count, panda_stream stream);
panda_error panda_mem_pool_create(panda_mem_pool *pool, int device_id);
panda_error panda_mem_pool_destroy(panda_mem_pool pool);
panda_error panda_malloc_from_pool_async(void **ptr, size_t size, panda_mem_pool pool, panda_stream stream);
panda_error panda_free_async(void *ptr, panda_stream stream);

/* ---- MSM: panda_interface.cuh:70-84; gpu_ffi/common.rs:175-185 (48 bytes, passed by value) ---- */
typedef struct panda_msm_configuration {
    panda_mem_pool mem_pool;        /* pool for the execution's temporaries (NULL handle: the device's default pool) */
    panda_stream stream;            /* stream the execution is ordered on */
    void *bases;                    /* device: 2^log_scalars_count affine points, x||y, Fq Montgomery LE; x == 0 <=> identity */
    void *scalars;                  /* device: 2^log_scalars_count x 32 B, Fr Montgomery LE */
    void *results;                  /* device: 3 Fq elements (96 B BN254, 144 B BLS12-377), Montgomery, canonical */
    unsigned log_scalars_count;
    panda_msm_result_coordinate_type msm_result_coordinate_type;
} panda_msm_configuration;
typedef panda_msm_configuration msm_configuration;

panda_error panda_msm_setup_bn254(void);
panda_error panda_msm_execute_bn254(const panda_msm_configuration exec_cfg);
/* Reference: CPU debug path taking HOST pointers for bases / scalars / results (msm_host.cuh:267-370).
 * Here the same contract (host pointers in, 96-byte result out, synchronous) is served by staging through the
 * device; the host buffers are not modified.  Like the reference (msm_host.cuh:372-383 never reads the coordinate
 * flag) the result is always Jacobian. */
panda_error panda_msm_execute_bn254_host(const panda_msm_configuration exec_cfg);
panda_error panda_msm_tear_down(void);

/* ---- NTT: panda_interface.cuh:86-110; gpu_ffi/common.rs:187-208 ---- */
typedef struct panda_ntt_configuration {
    panda_mem_pool mem_pool;
    panda_stream stream;
    void *d_src;                    /* device: 2^log_n Fr elements, Montgomery LE, natural order */
    void *d_dst;                    /* device: scratch / output buffer of the same size */
    unsigned log_n;
    void *flag;                     /* HOST unsigned*: written before return; 0 -> result in d_src, 1 -> result in d_dst */
} panda_ntt_configuration;
typedef panda_ntt_configuration ntt_configuration;

typedef struct panda_ntt_configuration_v1 {
    panda_mem_pool mem_pool;
    panda_stream stream;
    void *d_src;
    void *d_dst;
    void *d_omega;                  /* HOST pointer to the 2^log_n-th root of unity (32 B, Montgomery) -- unit.rs:507-518 */
    unsigned log_n;
    void *flag;
} panda_ntt_configuration_v1;
typedef panda_ntt_configuration_v1 ntt_configuration_v1;

/* input_omega: HOST pointer to omega for the transform size the caller is going to use (wrapper.rs:199-210). */
panda_error panda_ntt_setup_bn254(void *input_omega);
/* Forward DFT y[j] = sum_i x[i] * omega^(i*j), natural order in and out, no scaling (the function the reference's
 * disabled radix_fft text describes, fft.cu:107-169).  *flag = ceil(log_n / 8) & 1 as in fft.cu:193-211. */
panda_error panda_ntt_execute_bn254(panda_ntt_configuration exec_cfg);
panda_error panda_ntt_execute_bn254_v1(const panda_ntt_configuration_v1 exec_cfg);
panda_error panda_ntt_tear_down(void);

/* ---- additions the configs in BASELINE.json need; no reference precedent, same naming pattern ---- */

/* BLS12-377 G1: bases 96 B (x||y, 12 x u32 each), scalars 32 B (Fr, 253 bit), result 144 B. */
panda_error panda_msm_setup_bls12_377(void);
panda_error panda_msm_execute_bls12_377(const panda_msm_configuration exec_cfg);
panda_error panda_msm_execute_bls12_377_host(const panda_msm_configuration exec_cfg);   /* HOST pointers, 144-byte Jacobian result; see panda_msm_execute_bn254_host */

/* BLS12-381 G1 (third curve; the reference's README.md:36 lists it as planned, it ships no parameters for it): same shapes as BLS12-377 --
 * bases 96 B (x||y, 12 x u32 Montgomery), scalars 32 B (Fr, 255 bit, Montgomery), result 144 B.  Every entry point below exists for it under the
 * same name with the curve replaced. */
panda_error panda_msm_setup_bls12_381(void);
panda_error panda_msm_execute_bls12_381(const panda_msm_configuration exec_cfg);
panda_error panda_msm_execute_bls12_381_n(const panda_msm_configuration exec_cfg, size_t n);
panda_error panda_msm_execute_bls12_381_host(const panda_msm_configuration exec_cfg);
panda_error panda_msm_execute_bls12_381_host_scalars(const panda_msm_configuration exec_cfg, size_t n);
panda_error panda_msm_register_bases_bls12_381(const void *d_bases, size_t n, panda_stream stream);
panda_error panda_msm_combine_bls12_381(const void *partials, unsigned count, void *result,
                                        panda_msm_result_coordinate_type coord, panda_stream stream);

/* MSM over an arbitrary point count (a shard of a larger MSM): like panda_msm_execute_* but n need not be a
 * power of two; cfg.log_scalars_count is ignored. */
panda_error panda_msm_execute_bn254_n(const panda_msm_configuration exec_cfg, size_t n);
panda_error panda_msm_execute_bls12_377_n(const panda_msm_configuration exec_cfg, size_t n);

/* One bucket class of an MSM sharded over class_count GPUs (a power of two <= 64) that all hold ALL n points and scalars: the call adds up only
 * the digits whose bucket index is congruent to class_index modulo class_count and returns their weighted sum as a Jacobian point; the
 * class_count partials add up to the MSM (panda_msm_combine_*).  Every GPU does 1 / class_count of the additions AND of the bucket reduction at the
 * window width of the whole job, where a shard by point range (panda_msm_execute_*_n on n / class_count points) repeats the reduction per GPU and
 * narrows its windows.  Adjacent buckets belong to different classes, so skewed digit distributions stay balanced.  No reference precedent. */
panda_error panda_msm_execute_bn254_class(const panda_msm_configuration exec_cfg, size_t n, unsigned class_count, unsigned class_index);
panda_error panda_msm_execute_bls12_377_class(const panda_msm_configuration exec_cfg, size_t n, unsigned class_count, unsigned class_index);
panda_error panda_msm_execute_bls12_381_class(const panda_msm_configuration exec_cfg, size_t n, unsigned class_count, unsigned class_index);

/* init_msm for cached bases (wrapper.rs:122-152 keeps the device pointer and reuses it across calls): announces that the n
 * affine points at d_bases stay unchanged until panda_msm_unregister_bases / panda_msm_tear_down.  The library builds its
 * table of 2^(c*j) * P multiples right away (asynchronous on `stream`; W * n * 64 bytes of HBM, capped by PANDA_MSM_TABLE_BUDGET
 * [GiB], default half of the free memory beyond a 4 GiB reserve) and every later MSM on d_bases -- or on a prefix of it -- runs the
 * one-bucket-set pipeline.  Pointers that were never announced run the windowed pipeline: an execute call never allocates a
 * table or synchronises the host on its own (PANDA_MSM_PRECOMPUTE=1 opts unannounced pointers into automatic tables at the price
 * of a content fingerprint + 8-byte read-back per call).  Re-registering a pointer replaces the old entry. */
panda_error panda_msm_register_bases_bn254(const void *d_bases, size_t n, panda_stream stream);
panda_error panda_msm_register_bases_bls12_377(const void *d_bases, size_t n, panda_stream stream);
panda_error panda_msm_unregister_bases(const void *d_bases);

/* MSM whose scalars are still in HOST memory (cfg.scalars: host pointer, pinned for full overlap; cfg.bases / cfg.results:
 * device pointers as in panda_msm_execute_bn254).  Replaces the copy-then-execute sequence of unit.rs:103-188
 * (panda_msm_bn254_gpu_with_cached_bases): the library uploads the scalars in chunks on its own copy stream and sorts /
 * accumulates chunk q while chunk q+1 is still crossing PCIe.  n need not be a power of two;
 * cfg.log_scalars_count is ignored.  Asynchronous on cfg.stream (pinned memory); the host buffer must stay valid until the
 * stream reaches the end of the call. */
panda_error panda_msm_execute_bn254_host_scalars(const panda_msm_configuration exec_cfg, size_t n);
panda_error panda_msm_execute_bls12_377_host_scalars(const panda_msm_configuration exec_cfg, size_t n);

/* Sum `count` Jacobian partial results (device, count x 96 B / 144 B; e.g. the all-gathered per-GPU results of an
 * MSM sharded by point range) into `result` (device) in the requested coordinates.  Asynchronous on `stream`. */
panda_error panda_msm_combine_bn254(const void *partials, unsigned count, void *result,
                                    panda_msm_result_coordinate_type coord, panda_stream stream);
panda_error panda_msm_combine_bls12_377(const void *partials, unsigned count, void *result,
                                        panda_msm_result_coordinate_type coord, panda_stream stream);

/* Inverse transform: x = (1/n) * DFT_{omega^-1}(y); omega is the FORWARD root (host pointer), same flag contract. */
panda_error panda_intt_execute_bn254_v1(const panda_ntt_configuration_v1 exec_cfg);

/* Coset transforms (what a PLONK / halo2 prover calls around its quotient polynomial; no reference precedent): with coset
 * generator g (HOST pointer, 32 B Montgomery) forward computes y = NTT(x_i * g^i), inverse != 0 computes
 * x_i = g^-i * INTT(y)_i; same omega / flag contract as panda_ntt_execute_bn254_v1.  The forward pre-scaling is done in place
 * in d_src (both buffers are work storage, see the note at the top). */
panda_error panda_ntt_coset_execute_bn254_v1(const panda_ntt_configuration_v1 exec_cfg, const void *coset_gen, int inverse);

/* Bit-reversal permutation of 2^log_n Fr elements, out of place: d_dst[bitrev(i)] = d_src[i].  Together with the natural-order
 * transforms it gives the natural-in / reversed-out (and reversed-in / natural-out) variants provers ask for. */
panda_error panda_ntt_bit_reverse_bn254(const void *d_src, void *d_dst, unsigned log_n, panda_stream stream);

/* Batched transform: `batch` independent 2^log_n-point (I)NTTs stored back to back in d_src (d_dst: same size); same omega /
 * flag contract as panda_ntt_execute_bn254_v1 (flag tells which buffer holds all `batch` results).  inverse != 0: omega^-1 and
 * the 1/2^log_n scale.  Building block of polynomial-batch provers and of the multi-GPU four-step transform. */
panda_error panda_ntt_batch_execute_bn254_v1(const panda_ntt_configuration_v1 exec_cfg, unsigned batch, int inverse);

/* Exchange step of the four-step transform n = n1 * n2 sharded over the GPUs of one box (no reference
 * precedent).  Reads the 2^log_rows x 2^log_cols row-major matrix d_src, multiplies element (r, c) by
 * omega^((row_offset + r) * c)  (omega: HOST pointer, primitive 2^log_n-th root, NULL = no twiddle; inverse != 0: omega^-1),
 * and writes it transposed: column block h (of `parts` equal blocks, parts a power of two <= 16) goes to
 * dst[h][(c mod block) * ld + col_offset + r].  dst is a HOST array of `parts` DEVICE pointers: per-rank staging chunks for an
 * all-to-all, or peer-mapped buffers of the other GPUs (the kernel then stores over NVLink straight into the consumer's row
 * layout).  parts = 1 without omega is a plain transpose.  Asynchronous on `stream`. */
typedef struct panda_ntt_exchange_configuration {
    panda_stream stream;
    const void *d_src;
    unsigned log_rows, log_cols;
    unsigned row_offset;
    unsigned log_n;
    const void *omega;
    int inverse;
    unsigned parts;
    void *const *dst;
    size_t ld;
    size_t col_offset;
} panda_ntt_exchange_configuration;
panda_error panda_ntt_exchange_bn254(const panda_ntt_exchange_configuration *cfg);

/* ---- multi-GPU entry points: ONE process drives several GPUs of one box (the reference is
 * single-device -- cudaSetDevice(0) in msm_cuda.cuh:554-555, "Supports one GPU by default" wrapper.rs:38 -- but its bindings
 * already declare the peer-access calls such a caller needs, binding.rs:54-56).  The device of each shard is taken from its
 * pointers; the calls are asynchronous on the per-device streams, ordered across devices by events, and leave the caller's current
 * device unchanged. ---- */

/* MSM sharded by contiguous point range: per_device[d] describes shard d exactly like panda_msm_execute_bn254 (bases / scalars /
 * results on GPU d, a stream and a pool of that GPU, 2^log_scalars_count points; the _n variants take the point counts instead).
 * Every GPU runs the whole pipeline on its shard (one host thread per GPU queues the work); the 96-byte Jacobian partials are copied
 * peer-to-peer to the first shard's GPU and summed there.  On completion per_device[0].results holds the total in
 * per_device[0].msm_result_coordinate_type; per_device[d > 0].results hold shard d's Jacobian partial.  Completion is ordered on
 * per_device[0].stream. */
panda_error panda_msm_execute_bn254_multi(const panda_msm_configuration *per_device, int n_dev);
panda_error panda_msm_execute_bn254_multi_n(const panda_msm_configuration *per_device, const size_t *counts, int n_dev);
panda_error panda_msm_execute_bls12_377_multi(const panda_msm_configuration *per_device, int n_dev);
panda_error panda_msm_execute_bls12_377_multi_n(const panda_msm_configuration *per_device, const size_t *counts, int n_dev);
panda_error panda_msm_execute_bls12_381_multi(const panda_msm_configuration *per_device, int n_dev);
panda_error panda_msm_execute_bls12_381_multi_n(const panda_msm_configuration *per_device, const size_t *counts, int n_dev);

/* Four-step NTT n = n1 * n2 (n1 = 2^floor(log_n / 2)) sharded over n_dev GPUs (a power of two <= 16, n_dev <= n1) with ONE exchange:
 * local transpose, batched n1-point transforms, panda_ntt_exchange_bn254 storing over NVLink straight into the peers' d_dst
 * (peer access is enabled by the call), batched n2-point transforms.
 *   forward: d_src[g] = column block g of the natural-order n1 x n2 matrix x[i1 * n2 + i2] (columns g * n2/G ..), row-major, not written;
 *            d_dst[g] = row block g of the result: element [j1l][j2] = X[(g * n1/G + j1l) + n1 * j2].
 *   inverse (inverse != 0): takes that row-block layout in d_src and returns the column blocks of (1/n) * DFT_{omega^-1} in d_dst,
 *            so forward -> pointwise work -> inverse needs no re-shuffle.
 * omega: HOST pointer, primitive 2^log_n-th root.  Each shard is n / n_dev elements; scratch (2 x shard) comes from each GPU's default pool. */
typedef struct panda_ntt_multi_configuration {
    unsigned n_dev;
    const panda_stream *streams;    /* HOST array: one stream per GPU */
    void *const *d_src;             /* HOST array of n_dev DEVICE pointers */
    void *const *d_dst;             /* HOST array of n_dev DEVICE pointers (plain cudaMalloc / panda_malloc memory: written by the peers) */
    const void *omega;
    unsigned log_n;
    int inverse;
} panda_ntt_multi_configuration;
panda_error panda_ntt_execute_bn254_multi(const panda_ntt_multi_configuration *cfg);

/* Library identification: "panda-b200 <version> sm_100a". */
const char *panda_version(void);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* PANDA_INTERFACE_H */
