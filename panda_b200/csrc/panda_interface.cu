// panda_interface.cu -- the C ABI (include/panda_interface.h) over the sm_100a MSM / NTT implementation.
//
// Same symbols, argument meaning and error convention as the reference's src/cuda/core/panda_interface.cu:11-191
// (every function returns the cudaError_t value, 0 = success), plus the spellings the Rust bindings expect but the
// reference never defined (src/gpu_ffi/binding.rs:14,16,54-56) and the BLS12-377 / sharding / inverse-NTT additions.
#include "panda_interface.h"
#include "panda_debug.h"
#include "msm.cuh"
#include "ntt.cuh"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <cuda_runtime.h>

static inline panda_error perr(cudaError_t e) { return static_cast<panda_error>(e); }
static inline cudaStream_t cu(panda_stream s) { return static_cast<cudaStream_t>(s.handle); }
static inline cudaEvent_t cu(panda_event e) { return static_cast<cudaEvent_t>(e.handle); }
static inline cudaMemPool_t cu(panda_mem_pool p) { return static_cast<cudaMemPool_t>(p.handle); }

extern "C" {

const char *panda_version(void) { return "panda-b200 0.1.0 sm_100a"; }

// ---- plumbing: one CUDA runtime call each (panda_interface.cu:11-150) ---------------------------------------

panda_error panda_get_device_number(int *count) { return perr(cudaGetDeviceCount(count)); }
panda_error panda_get_device(int *device_id) { return perr(cudaGetDevice(device_id)); }
panda_error panda_set_device(int device_id) { return perr(cudaSetDevice(device_id)); }

panda_error panda_stream_create(panda_stream *stream, bool blocking_sync) {
    // common.cu:11-21: plain (blocking) stream; optionally the blocking-sync wait policy
    cudaStream_t s = nullptr;
    cudaError_t e = cudaStreamCreate(&s);
    if (e != cudaSuccess) return perr(e);
    stream->handle = s;
    if (!blocking_sync) return panda_success;
    cudaStreamAttrValue policy{};
    policy.syncPolicy = cudaSyncPolicyBlockingSync;
    return perr(cudaStreamSetAttribute(s, cudaStreamAttributeSynchronizationPolicy, &policy));
}
panda_error panda_stream_wait_event(panda_stream stream, panda_event event) { return perr(cudaStreamWaitEvent(cu(stream), cu(event), 0)); }
panda_error panda_stream_sync(panda_stream stream) { return perr(cudaStreamSynchronize(cu(stream))); }
panda_error panda_stream_synchronize(panda_stream stream) { return perr(cudaStreamSynchronize(cu(stream))); }
panda_error panda_stream_query(panda_stream stream) { return perr(cudaStreamQuery(cu(stream))); }
panda_error panda_stream_destroy(panda_stream stream) { return perr(cudaStreamDestroy(cu(stream))); }
panda_error panda_launch_host_fn(panda_stream stream, panda_host_fn fn, void *user_data) { return perr(cudaLaunchHostFunc(cu(stream), fn, user_data)); }

panda_error panda_event_create(panda_event *event, bool blocking_sync, bool disable_timing) {
    unsigned flags = (blocking_sync ? cudaEventBlockingSync : cudaEventDefault) | (disable_timing ? cudaEventDisableTiming : cudaEventDefault);
    return perr(cudaEventCreateWithFlags(reinterpret_cast<cudaEvent_t *>(&event->handle), flags));
}
panda_error panda_event_record(panda_event event, panda_stream stream) { return perr(cudaEventRecord(cu(event), cu(stream))); }
panda_error panda_event_sync(panda_event event) { return perr(cudaEventSynchronize(cu(event))); }
panda_error panda_event_query(panda_event event) { return perr(cudaEventQuery(cu(event))); }
panda_error panda_event_destroy(panda_event event) { return perr(cudaEventDestroy(cu(event))); }

panda_error panda_mem_get_info(size_t *free, size_t *total) { return perr(cudaMemGetInfo(free, total)); }
panda_error panda_malloc(void **ptr, size_t size) { return perr(cudaMalloc(ptr, size)); }
panda_error panda_malloc_host(void **ptr, size_t size) { return perr(cudaMallocHost(ptr, size)); }
panda_error panda_free(void *ptr) { return perr(cudaFree(ptr)); }
panda_error panda_free_host(void *ptr) { return perr(cudaFreeHost(ptr)); }
panda_error panda_host_register(void *ptr, size_t size) { return perr(cudaHostRegister(ptr, size, cudaHostRegisterDefault)); }
panda_error panda_host_unregister(void *ptr) { return perr(cudaHostUnregister(ptr)); }
panda_error panda_device_disable_peer_access(int device_id) { return perr(cudaDeviceDisablePeerAccess(device_id)); }
panda_error panda_device_enable_peer_access(int device_id) { return perr(cudaDeviceEnablePeerAccess(device_id, 0)); }
panda_error panda_memcpy(void *dst, const void *src, size_t count) { return perr(cudaMemcpy(dst, src, count, cudaMemcpyDefault)); }
panda_error panda_memcpy_async(void *dst, const void *src, size_t count, panda_stream stream) {
    return perr(cudaMemcpyAsync(dst, src, count, cudaMemcpyDefault, cu(stream)));
}
panda_error panda_memset(void *ptr, int value, size_t count) { return perr(cudaMemset(ptr, value, count)); }
panda_error panda_memset_async(void *ptr, int value, size_t count, panda_stream stream) { return perr(cudaMemsetAsync(ptr, value, count, cu(stream))); }

panda_error panda_mem_pool_create(panda_mem_pool *pool, int device_id) {
    // common.cu:23-29: device-local pool that never trims (release threshold = max)
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device_id;
    cudaMemPool_t p = nullptr;
    cudaError_t e = cudaMemPoolCreate(&p, &props);
    if (e != cudaSuccess) return perr(e);
    pool->handle = p;
    unsigned long long threshold = ~0ull;
    return perr(cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &threshold));
}
panda_error panda_mem_pool_destroy(panda_mem_pool pool) { return perr(cudaMemPoolDestroy(cu(pool))); }
panda_error panda_malloc_from_pool_async(void **ptr, size_t size, panda_mem_pool pool, panda_stream stream) {
    return perr(cudaMallocFromPoolAsync(ptr, size, cu(pool), cu(stream)));
}
panda_error panda_free_async(void *ptr, panda_stream stream) { return perr(cudaFreeAsync(ptr, cu(stream))); }

// ---- MSM (panda_interface.cu:152-170) -----------------------------------------------------------------------

static panda_error msm_execute(pb::CurveId curve, const panda_msm_configuration &cfg, size_t n) {
    if (!cfg.results || (n && (!cfg.bases || !cfg.scalars))) return perr(cudaErrorInvalidValue);
    if (n > (size_t)1 << 30) return perr(cudaErrorInvalidValue);
    pb::CoordType coord = cfg.msm_result_coordinate_type == PROJECTIVE ? pb::COORD_PROJECTIVE : pb::COORD_JACOBIAN;
    return perr(pb::msm_run(curve, cfg.bases, cfg.scalars, (uint32_t)n, cfg.results, coord, cu(cfg.mem_pool), cu(cfg.stream)));
}

// One bucket class of a sharded MSM (msm.cuh): the Jacobian partial of the buckets congruent to class_index modulo class_count.
static panda_error msm_execute_class(pb::CurveId curve, const panda_msm_configuration &cfg, size_t n, unsigned class_count, unsigned class_index) {
    if (!cfg.results || (n && (!cfg.bases || !cfg.scalars))) return perr(cudaErrorInvalidValue);
    if (n > (size_t)1 << 30) return perr(cudaErrorInvalidValue);
    return perr(pb::msm_run(curve, cfg.bases, cfg.scalars, (uint32_t)n, cfg.results, pb::COORD_JACOBIAN, cu(cfg.mem_pool), cu(cfg.stream), 0, 0, nullptr,
                            pb::MSM_TABLE_DEFAULT, class_count, class_index));
}

// Host-pointer variant: stage through the device on cfg.stream, synchronous like the reference's CPU path.
static panda_error msm_execute_host(pb::CurveId curve, const panda_msm_configuration &cfg) {
    if (!cfg.results || !cfg.bases || !cfg.scalars || cfg.log_scalars_count > 30) return perr(cudaErrorInvalidValue);   // before anything is queued
    const size_t n = (size_t)1 << cfg.log_scalars_count;
    const size_t fq = pb::curve_fq_bytes(curve);
    cudaStream_t s = cu(cfg.stream);
    uint8_t *d = nullptr;
    const size_t bases_bytes = n * 2 * fq, scalars_bytes = n * 32, off_s = (bases_bytes + 255) & ~(size_t)255,
                 off_r = off_s + ((scalars_bytes + 255) & ~(size_t)255);
    cudaError_t e = cu(cfg.mem_pool) ? cudaMallocFromPoolAsync((void **)&d, off_r + 256, cu(cfg.mem_pool), s) : cudaMallocAsync((void **)&d, off_r + 256, s);
    if (e != cudaSuccess) return perr(e);
    do {
        if ((e = cudaMemcpyAsync(d, cfg.bases, bases_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(d + off_s, cfg.scalars, scalars_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess) break;
        panda_msm_configuration dc = cfg;
        dc.bases = d; dc.scalars = d + off_s; dc.results = d + off_r;
        dc.msm_result_coordinate_type = JACOBIAN;   // the reference host path never reads the flag (msm_host.cuh:372-383)
        if ((e = (cudaError_t)msm_execute(curve, dc, n)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(cfg.results, d + off_r, 3 * fq, cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
        e = cudaStreamSynchronize(s);
    } while (0);
    cudaError_t f = cudaFreeAsync(d, s);
    return perr(e != cudaSuccess ? e : f);
}

static panda_error msm_execute_host_scalars(pb::CurveId curve, const panda_msm_configuration &cfg, size_t n) {
    if (!cfg.results || (n && (!cfg.bases || !cfg.scalars))) return perr(cudaErrorInvalidValue);
    if (n > (size_t)1 << 30) return perr(cudaErrorInvalidValue);
    pb::CoordType coord = cfg.msm_result_coordinate_type == PROJECTIVE ? pb::COORD_PROJECTIVE : pb::COORD_JACOBIAN;
    return perr(pb::msm_run_streamed(curve, cfg.bases, cfg.scalars, (uint32_t)n, cfg.results, coord, cu(cfg.mem_pool), cu(cfg.stream)));
}
panda_error panda_msm_execute_bn254_host_scalars(const panda_msm_configuration cfg, size_t n) { return msm_execute_host_scalars(pb::CURVE_BN254, cfg, n); }
panda_error panda_msm_execute_bls12_377_host_scalars(const panda_msm_configuration cfg, size_t n) { return msm_execute_host_scalars(pb::CURVE_BLS12_377, cfg, n); }

panda_error panda_msm_register_bases_bn254(const void *d_bases, size_t n, panda_stream stream) {
    if (n > (size_t)1 << 30) return perr(cudaErrorInvalidValue);
    return perr(pb::msm_register_bases(pb::CURVE_BN254, d_bases, (uint32_t)n, cu(stream)));
}
panda_error panda_msm_register_bases_bls12_377(const void *d_bases, size_t n, panda_stream stream) {
    if (n > (size_t)1 << 30) return perr(cudaErrorInvalidValue);
    return perr(pb::msm_register_bases(pb::CURVE_BLS12_377, d_bases, (uint32_t)n, cu(stream)));
}
panda_error panda_msm_unregister_bases(const void *d_bases) { return perr(pb::msm_unregister_bases(d_bases)); }

// BLS12-381 G1 (third curve; same shapes as BLS12-377: 96-byte points, 144-byte results, 32-byte scalars of 255 bits)
panda_error panda_msm_setup_bls12_381(void) { return panda_success; }
panda_error panda_msm_execute_bls12_381(const panda_msm_configuration cfg) {
    if (cfg.log_scalars_count > 30) return perr(cudaErrorInvalidValue);
    return msm_execute(pb::CURVE_BLS12_381, cfg, (size_t)1 << cfg.log_scalars_count);
}
panda_error panda_msm_execute_bls12_381_n(const panda_msm_configuration cfg, size_t n) { return msm_execute(pb::CURVE_BLS12_381, cfg, n); }
panda_error panda_msm_execute_bls12_381_host(const panda_msm_configuration cfg) { return msm_execute_host(pb::CURVE_BLS12_381, cfg); }
panda_error panda_msm_execute_bls12_381_host_scalars(const panda_msm_configuration cfg, size_t n) { return msm_execute_host_scalars(pb::CURVE_BLS12_381, cfg, n); }
panda_error panda_msm_register_bases_bls12_381(const void *d_bases, size_t n, panda_stream stream) {
    if (n > (size_t)1 << 30) return perr(cudaErrorInvalidValue);
    return perr(pb::msm_register_bases(pb::CURVE_BLS12_381, d_bases, (uint32_t)n, cu(stream)));
}
panda_error panda_msm_combine_bls12_381(const void *partials, unsigned count, void *result, panda_msm_result_coordinate_type coord, panda_stream stream) {
    return perr(pb::msm_combine(pb::CURVE_BLS12_381, partials, count, result, coord == PROJECTIVE ? pb::COORD_PROJECTIVE : pb::COORD_JACOBIAN, cu(stream)));
}

panda_error panda_msm_setup_bn254(void) { return panda_success; }          // nothing to prepare: msm_cuda.cuh:786-795 is empty too
panda_error panda_msm_setup_bls12_377(void) { return panda_success; }
panda_error panda_msm_tear_down(void) { return perr(pb::msm_release_tables()); }   // idempotent (wrapper.rs:297-312 calls it once per base set); drops cached tables

panda_error panda_msm_execute_bn254(const panda_msm_configuration cfg) {
    if (cfg.log_scalars_count > 30) return perr(cudaErrorInvalidValue);
    return msm_execute(pb::CURVE_BN254, cfg, (size_t)1 << cfg.log_scalars_count);
}
panda_error panda_msm_execute_bn254_n(const panda_msm_configuration cfg, size_t n) { return msm_execute(pb::CURVE_BN254, cfg, n); }
panda_error panda_msm_execute_bn254_class(const panda_msm_configuration cfg, size_t n, unsigned class_count, unsigned class_index) {
    return msm_execute_class(pb::CURVE_BN254, cfg, n, class_count, class_index);
}
panda_error panda_msm_execute_bls12_377_class(const panda_msm_configuration cfg, size_t n, unsigned class_count, unsigned class_index) {
    return msm_execute_class(pb::CURVE_BLS12_377, cfg, n, class_count, class_index);
}
panda_error panda_msm_execute_bls12_381_class(const panda_msm_configuration cfg, size_t n, unsigned class_count, unsigned class_index) {
    return msm_execute_class(pb::CURVE_BLS12_381, cfg, n, class_count, class_index);
}
panda_error panda_msm_execute_bn254_host(const panda_msm_configuration cfg) { return msm_execute_host(pb::CURVE_BN254, cfg); }
panda_error panda_msm_execute_bls12_377_host(const panda_msm_configuration cfg) { return msm_execute_host(pb::CURVE_BLS12_377, cfg); }
panda_error panda_msm_execute_bls12_377(const panda_msm_configuration cfg) {
    if (cfg.log_scalars_count > 30) return perr(cudaErrorInvalidValue);
    return msm_execute(pb::CURVE_BLS12_377, cfg, (size_t)1 << cfg.log_scalars_count);
}
panda_error panda_msm_execute_bls12_377_n(const panda_msm_configuration cfg, size_t n) { return msm_execute(pb::CURVE_BLS12_377, cfg, n); }

panda_error panda_msm_combine_bn254(const void *partials, unsigned count, void *result, panda_msm_result_coordinate_type coord, panda_stream stream) {
    return perr(pb::msm_combine(pb::CURVE_BN254, partials, count, result, coord == PROJECTIVE ? pb::COORD_PROJECTIVE : pb::COORD_JACOBIAN, cu(stream)));
}
panda_error panda_msm_combine_bls12_377(const void *partials, unsigned count, void *result, panda_msm_result_coordinate_type coord, panda_stream stream) {
    return perr(pb::msm_combine(pb::CURVE_BLS12_377, partials, count, result, coord == PROJECTIVE ? pb::COORD_PROJECTIVE : pb::COORD_JACOBIAN, cu(stream)));
}

// ---- NTT (panda_interface.cu:172-191) -------------------------------------------------------------------------

// setup state is per device (the reference keeps ONE process-global table, fft.cu:223, so two managers on two GPUs race)
static constexpr int MAX_DEVICES = 64;
static std::mutex g_omega_mutex;
static unsigned char g_setup_omega[MAX_DEVICES][32];
static bool g_setup_done[MAX_DEVICES] = {};

panda_error panda_ntt_setup_bn254(void *input_omega) {
    if (!input_omega) return perr(cudaErrorInvalidValue);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return perr(e);
    if (dev < 0 || dev >= MAX_DEVICES) return perr(cudaErrorInvalidDevice);
    std::lock_guard<std::mutex> lock(g_omega_mutex);
    memcpy(g_setup_omega[dev], input_omega, 32);      // fft.cu:62-73 copies omega to the device here; tables are built at first execute
    g_setup_done[dev] = true;
    return panda_success;
}

static panda_error ntt_execute(void *d_src, void *d_dst, unsigned log_n, const void *omega, bool inverse, panda_stream stream, void *flag) {
    if (!flag || (log_n && (!d_src || !d_dst))) return perr(cudaErrorInvalidValue);
    unsigned in_dst = 0;
    cudaError_t e = pb::ntt_run(pb::NTT_BN254_FR, d_src, d_dst, log_n, omega, inverse, cu(stream), &in_dst);
    *static_cast<unsigned *>(flag) = in_dst;      // host write, before returning (unit.rs:458 reads it immediately)
    return perr(e);
}

panda_error panda_ntt_execute_bn254(panda_ntt_configuration cfg) {
    unsigned char omega[32];
    int dev = 0;
    cudaError_t de = cudaGetDevice(&dev);
    if (de != cudaSuccess) return perr(de);
    {
        std::lock_guard<std::mutex> lock(g_omega_mutex);
        if (dev < 0 || dev >= MAX_DEVICES || !g_setup_done[dev]) return perr(cudaErrorNotReady);
        memcpy(omega, g_setup_omega[dev], 32);
    }
    return ntt_execute(cfg.d_src, cfg.d_dst, cfg.log_n, omega, false, cfg.stream, cfg.flag);
}
panda_error panda_ntt_execute_bn254_v1(const panda_ntt_configuration_v1 cfg) {
    if (!cfg.d_omega) return perr(cudaErrorInvalidValue);
    return ntt_execute(cfg.d_src, cfg.d_dst, cfg.log_n, cfg.d_omega, false, cfg.stream, cfg.flag);
}
panda_error panda_intt_execute_bn254_v1(const panda_ntt_configuration_v1 cfg) {
    if (!cfg.d_omega) return perr(cudaErrorInvalidValue);
    return ntt_execute(cfg.d_src, cfg.d_dst, cfg.log_n, cfg.d_omega, true, cfg.stream, cfg.flag);
}
panda_error panda_ntt_batch_execute_bn254_v1(const panda_ntt_configuration_v1 cfg, unsigned batch, int inverse) {
    if (!cfg.d_omega || !cfg.flag || !cfg.d_src || !cfg.d_dst) return perr(cudaErrorInvalidValue);
    unsigned in_dst = 0;
    cudaError_t e = pb::ntt_run(pb::NTT_BN254_FR, cfg.d_src, cfg.d_dst, cfg.log_n, cfg.d_omega, inverse != 0, cu(cfg.stream), &in_dst, batch);
    *static_cast<unsigned *>(cfg.flag) = in_dst;
    return perr(e);
}
panda_error panda_ntt_coset_execute_bn254_v1(const panda_ntt_configuration_v1 cfg, const void *coset_gen, int inverse) {
    if (!cfg.d_omega || !cfg.flag || !coset_gen || !cfg.d_src || !cfg.d_dst) return perr(cudaErrorInvalidValue);
    cudaError_t e = cudaSuccess;
    if (!inverse) e = pb::ntt_coset_scale(pb::NTT_BN254_FR, cfg.d_src, cfg.log_n, coset_gen, false, cu(cfg.stream));
    unsigned in_dst = 0;
    if (e == cudaSuccess) e = pb::ntt_run(pb::NTT_BN254_FR, cfg.d_src, cfg.d_dst, cfg.log_n, cfg.d_omega, inverse != 0, cu(cfg.stream), &in_dst);
    *static_cast<unsigned *>(cfg.flag) = in_dst;
    if (e == cudaSuccess && inverse) e = pb::ntt_coset_scale(pb::NTT_BN254_FR, in_dst ? cfg.d_dst : cfg.d_src, cfg.log_n, coset_gen, true, cu(cfg.stream));
    return perr(e);
}
panda_error panda_ntt_bit_reverse_bn254(const void *d_src, void *d_dst, unsigned log_n, panda_stream stream) {
    return perr(pb::ntt_bit_reverse(pb::NTT_BN254_FR, d_src, d_dst, log_n, cu(stream)));
}
panda_error panda_ntt_exchange_bn254(const panda_ntt_exchange_configuration *cfg) {
    if (!cfg) return perr(cudaErrorInvalidValue);
    return perr(pb::ntt_exchange(pb::NTT_BN254_FR, cfg->d_src, cfg->log_rows, cfg->log_cols, cfg->row_offset, cfg->omega, cfg->log_n, cfg->inverse != 0,
                                 cfg->parts, cfg->dst, cfg->ld, cfg->col_offset, cu(cfg->stream)));
}
panda_error panda_ntt_tear_down(void) {       // the current device's NTT unit: its setup omega and its cached twiddle tables
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < MAX_DEVICES) {
        std::lock_guard<std::mutex> lock(g_omega_mutex);
        g_setup_done[dev] = false;
    } else cudaGetLastError();
    return perr(pb::ntt_release_tables());
}

// ---- diagnostics (include/panda_debug.h) --------------------------------------------------------------------------

panda_error panda_debug_ntt_timed(const panda_ntt_configuration_v1 cfg, int inverse, float *pass_ms) {
    if (!cfg.d_omega || !cfg.flag || !cfg.d_src || !cfg.d_dst || !pass_ms) return perr(cudaErrorInvalidValue);
    unsigned in_dst = 0;
    cudaError_t e = pb::ntt_run(pb::NTT_BN254_FR, cfg.d_src, cfg.d_dst, cfg.log_n, cfg.d_omega, inverse != 0, cu(cfg.stream), &in_dst, 1, pass_ms);
    *static_cast<unsigned *>(cfg.flag) = in_dst;
    return perr(e);
}

panda_error panda_debug_fr_pow2k_host(const void *omega, unsigned k, void *out) {
    if (!omega || !out) return perr(cudaErrorInvalidValue);
    pb::ntt_pow2k_host(omega, k, out);
    return panda_success;
}

panda_error panda_debug_msm_plan(int curve, size_t n, int folded, unsigned c_override, unsigned seg_override, panda_debug_msm_plan_info *out) {
    if (!out) return perr(cudaErrorInvalidValue);
    pb::MsmPlan p = pb::msm_make_plan(pb::curve_from_id(curve), (uint32_t)n, folded != 0, c_override, seg_override);
    out->window_bits = p.c; out->windows = p.windows; out->buckets_per_window = p.nb; out->segment_len = p.seg_len;
    out->segments_per_window = p.segs_ps; out->reduce_chunk = p.chunk; out->workspace_bytes = p.bytes;
    out->folded = p.folded; out->bucket_sets = p.sets; out->groups = p.groups; out->phases = p.phases; out->table_bytes = p.table_bytes;
    return panda_success;
}

panda_error panda_debug_msm_timed(int curve, const panda_msm_configuration cfg, size_t n, unsigned c_override, unsigned seg_override, int table_mode,
                                  float *stage_ms, unsigned *info) {
    pb::MsmStageTimes t{};
    pb::CoordType coord = cfg.msm_result_coordinate_type == PROJECTIVE ? pb::COORD_PROJECTIVE : pb::COORD_JACOBIAN;
    cudaError_t e = pb::msm_run(pb::curve_from_id(curve), cfg.bases, cfg.scalars, (uint32_t)n, cfg.results, coord,
                                cu(cfg.mem_pool), cu(cfg.stream), c_override, seg_override, (stage_ms || info) ? &t : nullptr, table_mode);
    if (stage_ms) {
        stage_ms[0] = t.digits; stage_ms[1] = t.scan; stage_ms[2] = t.scatter; stage_ms[3] = t.accumulate;
        stage_ms[4] = t.bucket_reduce; stage_ms[5] = t.window_reduce; stage_ms[6] = t.final;
    }
    if (info) { info[0] = (unsigned)t.folded; info[1] = t.c; info[2] = t.windows; }
    return perr(e);
}

panda_error panda_debug_msm_timed_class(int curve, const panda_msm_configuration cfg, size_t n, unsigned class_count, unsigned class_index,
                                        float *stage_ms, unsigned *info) {
    pb::MsmStageTimes t{};
    cudaError_t e = pb::msm_run(pb::curve_from_id(curve), cfg.bases, cfg.scalars, (uint32_t)n, cfg.results, pb::COORD_JACOBIAN,
                                cu(cfg.mem_pool), cu(cfg.stream), 0, 0, (stage_ms || info) ? &t : nullptr, pb::MSM_TABLE_DEFAULT, class_count, class_index);
    if (stage_ms) {
        stage_ms[0] = t.digits; stage_ms[1] = t.scan; stage_ms[2] = t.scatter; stage_ms[3] = t.accumulate;
        stage_ms[4] = t.bucket_reduce; stage_ms[5] = t.window_reduce; stage_ms[6] = t.final;
    }
    if (info) { info[0] = (unsigned)t.folded; info[1] = t.c; info[2] = t.windows; }
    return perr(e);
}

panda_error panda_debug_msm_streamed(int curve, const panda_msm_configuration cfg, size_t n, int table_mode, unsigned chunks) {
    pb::CoordType coord = cfg.msm_result_coordinate_type == PROJECTIVE ? pb::COORD_PROJECTIVE : pb::COORD_JACOBIAN;
    return perr(pb::msm_run_streamed(pb::curve_from_id(curve), cfg.bases, cfg.scalars, (uint32_t)n, cfg.results, coord,
                                     cu(cfg.mem_pool), cu(cfg.stream), table_mode, chunks));
}

}  // extern "C"
