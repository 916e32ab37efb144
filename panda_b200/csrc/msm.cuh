// msm.cuh -- Pippenger multi-scalar multiplication pipeline for sm_100a (internal interface).
//
// Replaces the reference's msm_execute_cuda and its kernels (src/cuda/core/unit/msm/msm_cuda.cuh:552-769,
// kernels :148-282, :373-497; host Horner :59-77).  See DESIGN.md for the data layout and the per-kernel
// rooflines.  Everything runs on the caller's stream, temporaries come from the caller's memory pool
// (stream-ordered), inputs are never written (the reference converts scalars in place, msm_cuda.cuh:155).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace pb {

enum CurveId { CURVE_BN254 = 0, CURVE_BLS12_377 = 1, CURVE_BLS12_381 = 2 };
inline size_t curve_fq_bytes(CurveId c) { return c == CURVE_BN254 ? 32 : 48; }                                  // base-field element (8 or 12 limbs)
inline uint32_t curve_scalar_bits(CurveId c) { return c == CURVE_BLS12_377 ? 253 : c == CURVE_BLS12_381 ? 255 : 254; }
inline CurveId curve_from_id(int id) { return id == 1 ? CURVE_BLS12_377 : id == 2 ? CURVE_BLS12_381 : CURVE_BN254; }
enum CoordType { COORD_JACOBIAN = 0, COORD_PROJECTIVE = 1 };   // curve.cuh:23-27

static constexpr uint32_t MSM_MAX_CHUNKS = 8;

struct MsmPlan {
    uint32_t n;            // points
    uint32_t table_n;      // folded: row length of the table (>= n: an MSM may use a prefix of a registered base set)
    uint32_t c;            // window width in bits (signed digits)
    uint32_t windows;      // W digit windows per scalar
    uint32_t wide;         // windows 0 .. wide-1 are c bits wide, the others c-1 (table plan: balanced windows); W for uniform windows
    uint32_t nb;           // buckets per bucket set = 2^(c-1)
    uint32_t folded;       // 1: precomputed 2^(c*j)*P table, every window feeds ONE bucket set; 0: one bucket set per window
    uint32_t sets;         // bucket sets: 1 (folded) or W
    uint32_t stride;       // sorted-entry capacity per set: W*n (folded) or n
    uint32_t seg_len;      // L: sorted entries handled by one accumulation thread
    uint32_t segs_ps;      // ceil(stride / L) segments per set
    uint32_t chunk;        // m: buckets folded serially by one reduction thread
    uint32_t chunks_ps;    // nb / m
    uint32_t groups;       // CTAs per set in the group-reduce stage (<= 256, divides chunks_ps)
    uint32_t chunks;       // Q: point-range chunks that are sorted and accumulated separately and share the bucket reduction
                           //    (Q > 1 only for streamed scalars: chunk q computes while chunk q+1 is still being uploaded)
    uint32_t chunk_begin[MSM_MAX_CHUNKS + 1];   // chunk q covers points [chunk_begin[q], chunk_begin[q + 1]): sizes grow geometrically
    uint32_t phases;       // folded scatter: bucket ranges (a power of two <= 32, each at most 2^18 buckets): the codes are grouped by range tile by tile
                           //    and scattered one range at a time (L2-resident output slice; pipelined with the accumulation of the range before)
    uint32_t codes_stride; // folded: code capacity per chunk (whole 256-scalar tiles of W codes)
    uint32_t heads_stride; // folded: u16 header entries per chunk (tiles * (phases + 1))
    uint32_t class_log2;   // bucket-class shard (multi-GPU): this run only takes the digits whose bucket index is congruent to class_index
    uint32_t class_index;  //    modulo 2^class_log2; nb counts the buckets of that class.  0 / 0: the whole MSM
    // workspace layout (byte offsets into one arena)
    size_t off_counts, off_offsets, off_cursor, off_biglist, off_tiles, off_digits, off_heads, off_sorted, off_slots, off_chunks, off_gsums, bytes;
    size_t table_bytes;    // folded: size of the precomputed table (W * n affine points)
};

// window width / segment length selection; c_override = 0 -> cost model.  table_budget: bytes available for a table (folded only).
MsmPlan msm_make_plan(CurveId curve, uint32_t n, bool folded, uint32_t c_override, uint32_t seg_override, size_t table_budget = ~(size_t)0,
                      uint32_t chunks = 1, uint32_t class_log2 = 0);

// The library's per-device side streams and, for a streamed MSM, where the scalars come from (msm_impl.cuh: "streams by role").
//   host_scalars / dev_scalars / copy_stream   scalars still in host memory: uploaded chunk by chunk (sub-chunk by sub-chunk) on copy_stream
//   aux_stream   (high priority) the scatters: bucket range by bucket range (table plan), window by window (windowed plan)
//   dig_stream   (high priority) digits + scan of the chunks of a streamed MSM, so that chunk q+1 is recoded as it arrives while chunk q's
//                ranges are still being scattered
//   aux2 .. aux4_stream   with the caller's stream, the four streams the accumulation launches rotate over
struct MsmFeed { const void *host_scalars; void *dev_scalars; cudaStream_t copy_stream; cudaStream_t aux_stream; cudaStream_t aux2_stream; cudaStream_t aux3_stream; cudaStream_t aux4_stream; cudaStream_t dig_stream; };

// Per-stage device timings (ms) filled when msm_run is called with timings != nullptr (adds event syncs;
// the benchmark harness uses it to attribute time to kernels -- never set on the product path).
struct MsmStageTimes { float digits, scan, scatter, accumulate, bucket_reduce, window_reduce, final; int folded; unsigned c, windows; };

// How reused bases are handled (PANDA_MSM_PRECOMPUTE = 0 | 1 | 2 | 3 overrides the default REGISTERED):
//   OFF         never build tables, not even for registered sets
//   AUTO        opt-in: an UNANNOUNCED (device pointer, n) pair seen a second time with an identical 64-bit content fingerprint gets a
//               table (costs a fingerprint pass + an 8-byte read-back, i.e. a host synchronisation, on every call)
//   EAGER       opt-in: unannounced pointers get a table at first sight
//   REGISTERED  default: tables only for base sets announced with msm_register_bases (init_msm); an unannounced pointer runs the
//               windowed plan -- execute never blocks the host and never allocates a table behind the caller's back
enum MsmTableMode { MSM_TABLE_OFF = 0, MSM_TABLE_AUTO = 1, MSM_TABLE_EAGER = 2, MSM_TABLE_REGISTERED = 3, MSM_TABLE_DEFAULT = -1 };

// bases: n affine points (x||y Montgomery), scalars: n x 32 B (Montgomery), result: 3 field elements.
// pool may be nullptr (default pool).  Asynchronous on `stream` (the opt-in table modes AUTO / EAGER add an 8-byte fingerprint
// read-back, see MsmTableMode).
cudaError_t msm_run(CurveId curve, const void *bases, const void *scalars, uint32_t n, void *result,
                    CoordType coord, cudaMemPool_t pool, cudaStream_t stream,
                    uint32_t c_override = 0, uint32_t seg_override = 0, MsmStageTimes *timings = nullptr,
                    int table_mode = MSM_TABLE_DEFAULT, uint32_t class_count = 1, uint32_t class_index = 0);

// Bucket-class shards (class_count > 1, a power of two <= 64): the run adds up only the digits whose bucket index is congruent to
// class_index modulo class_count and returns  sum_{b = class_index (mod class_count)} (b + 1) * B_b  as a Jacobian point; the class_count
// partials add up to the MSM (msm_combine).  Every GPU of a sharded MSM sees ALL points and ALL scalars but does 1 / class_count of the
// additions AND of the bucket reduction with the window width of the whole job -- a shard by point range repeats the reduction on every
// GPU and has to narrow its windows.  Adjacent buckets belong to different classes, so skewed digit distributions stay balanced.

// Same MSM with the scalars in HOST memory (pinned for full overlap; pageable works).  With a table for `bases` the points are
// processed in chunks that share the bucket reduction, so the upload of chunk q+1 overlaps the sort / accumulation of chunk q;
// without one the scalars are uploaded in one piece.  Asynchronous on `stream` for pinned memory.
cudaError_t msm_run_streamed(CurveId curve, const void *bases, const void *host_scalars, uint32_t n, void *result, CoordType coord,
                             cudaMemPool_t pool, cudaStream_t stream, int table_mode = MSM_TABLE_DEFAULT, uint32_t chunks_override = 0);

// result = sum of `count` Jacobian partials (the per-GPU results of a sharded MSM), then coordinate conversion.
cudaError_t msm_combine(CurveId curve, const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream);

// init_msm for a base set that stays put until it is unregistered: builds its table right away (asynchronous on `stream`) and
// lets every later MSM on `bases` (or a prefix of it) skip the content fingerprint and its host synchronisation.
cudaError_t msm_register_bases(CurveId curve, const void *bases, uint32_t n, cudaStream_t stream);
cudaError_t msm_unregister_bases(const void *bases);

// drops every table of the current device (registered or not); synchronises
cudaError_t msm_release_tables();

}  // namespace pb
