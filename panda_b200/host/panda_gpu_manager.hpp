// panda_gpu_manager.hpp -- C++ host API over the C ABI (include/panda_interface.h).
//
// The reference's host side is Rust (no Rust toolchain in this image), so this header mirrors it in C++, name for name:
//   src/gpu_ffi/common.rs:5-38      PandaGpuError
//   src/gpu_ffi/common.rs:40-157    PandaStream / PandaEvent / PandaMemPool helpers
//   src/gpu_manager/common.rs:5-76  log_2, malloc_from_pool_async, memcpy_async, free_async, memory_alloc_and_copy, memory_copy_and_free
//   src/gpu_manager/wrapper.rs:8-347  PandaGpuManager, PandaGpuManagerInitUnitType, get_device_number, device_info, set_device
//   src/gpu_manager/unit.rs:10-543  panda_msm_bn254_gpu{,_with_cached_bases,_with_cached_scalars,_with_cached_input,_host},
//                                   panda_ntt_bn254_gpu{,_v1}
// Same argument meaning and error behaviour (a Rust Err(PandaGpuError::X) is a thrown PandaGpuException{X}).  Deliberate
// fixes, each marked "fix:" below: no pinned result staging buffer per call (unit.rs:67-74 allocates and leaks one), the result buffer is
// allocated on the stream that writes it (unit.rs:33-40 allocates on h2d_stream), the NTT waits for
// its upload before executing (unit.rs:426-453 does not).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/panda_interface.h"

namespace panda {

enum class PandaGpuError {
    GetDeviceCountError, SetDeviceError, DeviceGetDeviceMemoryInfoError, CreateContextError, InitUnitTypeError, MSMBasesAddrError,
    NTTOmegaAddrError, SetBasesErr, SchedulingErr, GetExponentAddressErr, GetResultAddressesErr, StartProcessingErr,
    FinishProcessingErr, DestroyContextErr, BasesIndexErr, MemPoolCreateErr, AsyncPoolMallocErr, AsyncMemcopyErr, NttExecErr,
    StremCreateErr, StreamDestroyErr, StreamWaitEventErr, StreamSyncErr, EventCreateErr, EventRecordErr, EventDestroyErr, EventSyncErr
};

inline const char *to_string(PandaGpuError e) {
    static const char *names[] = {"GetDeviceCountError", "SetDeviceError", "DeviceGetDeviceMemoryInfoError", "CreateContextError",
        "InitUnitTypeError", "MSMBasesAddrError", "NTTOmegaAddrError", "SetBasesErr", "SchedulingErr", "GetExponentAddressErr",
        "GetResultAddressesErr", "StartProcessingErr", "FinishProcessingErr", "DestroyContextErr", "BasesIndexErr", "MemPoolCreateErr",
        "AsyncPoolMallocErr", "AsyncMemcopyErr", "NttExecErr", "StremCreateErr", "StreamDestroyErr", "StreamWaitEventErr", "StreamSyncErr",
        "EventCreateErr", "EventRecordErr", "EventDestroyErr", "EventSyncErr"};
    return names[static_cast<int>(e)];
}

struct PandaGpuException : std::runtime_error {
    PandaGpuError kind;
    explicit PandaGpuException(PandaGpuError k) : std::runtime_error(to_string(k)), kind(k) {}
};

inline void check(panda_error rc, PandaGpuError kind) { if (rc != panda_success) throw PandaGpuException(kind); }

constexpr size_t BN254_SCALAR_WIDTH_BITS = 254;   // gpu_manager/mod.rs:12-16
constexpr size_t BN254_POINT_WIDTH_BITS = 254;
constexpr size_t FIELD_ELEMENT_LEN = 32;

// ---- gpu_ffi/common.rs handle helpers --------------------------------------------------------------------------------------
struct PandaStream {
    panda_stream raw{nullptr};
    static PandaStream create() { PandaStream s; check(panda_stream_create(&s.raw, true), PandaGpuError::StremCreateErr); return s; }
    static PandaStream null() { return PandaStream{}; }
    void destroy() const { check(panda_stream_destroy(raw), PandaGpuError::StreamDestroyErr); }
    void wait(panda_event e) const { check(panda_stream_wait_event(raw, e), PandaGpuError::StreamWaitEventErr); }
    void sync() const { check(panda_stream_synchronize(raw), PandaGpuError::StreamSyncErr); }
};

struct PandaEvent {
    panda_event raw{nullptr};
    static PandaEvent create() { PandaEvent e; check(panda_event_create(&e.raw, true, true), PandaGpuError::EventCreateErr); return e; }
    void record(const PandaStream &s) const { check(panda_event_record(raw, s.raw), PandaGpuError::EventRecordErr); }
    void sync() const { check(panda_event_sync(raw), PandaGpuError::EventSyncErr); }
    void destroy() const { check(panda_event_destroy(raw), PandaGpuError::EventDestroyErr); }
};

struct PandaMemPool {
    panda_mem_pool raw{nullptr};
    static PandaMemPool create(size_t device_id) {
        PandaMemPool p; check(panda_mem_pool_create(&p.raw, static_cast<int>(device_id)), PandaGpuError::MemPoolCreateErr); return p;
    }
};

struct PandaDeviceInfo { uint64_t free = 0, total = 0; };
enum class PandaMSMResultCoordinateType { Jacobian = 0, Projective = 1 };
// curves of the MSM entry points (no counterpart in the reference's Rust API, which is BN254 only; README.md:36 announces the others)
enum class PandaCurve { Bn254 = 0, Bls12_377 = 1 };
inline size_t fq_len(PandaCurve c) { return c == PandaCurve::Bls12_377 ? 48 : 32; }   // bytes of a base-field element; scalars are 32 bytes on both

// ---- gpu_manager/common.rs ---------------------------------------------------------------------------------------------------
inline uint32_t log_2(size_t num) {
    if (num == 0) throw std::invalid_argument("log_2(0)");
    uint32_t pow = 0;
    while ((size_t(1) << (pow + 1)) <= num) pow++;
    return pow;
}
inline void malloc_from_pool_async(void **ptr, size_t size, const PandaMemPool &pool, const PandaStream &stream) {
    check(panda_malloc_from_pool_async(ptr, size, pool.raw, stream.raw), PandaGpuError::AsyncPoolMallocErr);
}
inline void memcpy_async(void *dst, const void *src, size_t size, const PandaStream &stream) {
    check(panda_memcpy_async(dst, src, size, stream.raw), PandaGpuError::AsyncMemcopyErr);
}
inline void free_async(void *ptr, const PandaStream &stream) { check(panda_free_async(ptr, stream.raw), PandaGpuError::AsyncMemcopyErr); }

enum class PandaGpuManagerInitUnitType {   // wrapper.rs:23-29 (the long spellings are the Rust variant names)
    None, MSM, NTT, ALL,
    PandaGpuManagerInitUnitTypeNone = None, PandaGpuManagerInitUnitTypeMSM = MSM, PandaGpuManagerInitUnitTypeNTT = NTT,
    PandaGpuManagerInitUnitTypeALL = ALL
};

struct ByteSlice { const uint8_t *data = nullptr; size_t len = 0; };

// ---- gpu_manager/wrapper.rs ----------------------------------------------------------------------------------------------------
inline int get_device_number() {
    int count = 0;
    check(panda_get_device_number(&count), PandaGpuError::GetDeviceCountError);
    return count;
}
inline void set_device(size_t device_id) { check(panda_set_device(static_cast<int>(device_id)), PandaGpuError::SetDeviceError); }
inline PandaDeviceInfo device_info(int device_id) {
    check(panda_set_device(device_id), PandaGpuError::SetDeviceError);
    size_t free = 0, total = 0;
    check(panda_mem_get_info(&free, &total), PandaGpuError::DeviceGetDeviceMemoryInfoError);
    return PandaDeviceInfo{free, total};
}

class PandaGpuManager {
  public:
    std::vector<void *> d_bases;       // wrapper.rs:15-17: public, the caller pushes cached device pointers here
    std::vector<void *> d_scalars;
    std::vector<size_t> scalars_len;   // bytes

    // wrapper.rs:32-53 (`new`)
    static PandaGpuManager create(size_t device_id) {
        if (get_device_number() == 0) throw PandaGpuException(PandaGpuError::GetDeviceCountError);
        PandaGpuManager gm;
        gm.device_id_ = device_id;
        gm.mem_pool_ = init_hardware(device_id);
        gm.make_streams();
        return gm;
    }
    // wrapper.rs:55-113
    static PandaGpuManager init_all(size_t device_id, PandaGpuManagerInitUnitType unit, const std::vector<ByteSlice> *bases, const ByteSlice *omega) {
        if (get_device_number() == 0) throw PandaGpuException(PandaGpuError::GetDeviceCountError);
        PandaGpuManager gm;
        gm.device_id_ = device_id;
        gm.mem_pool_ = init_hardware(device_id);
        switch (unit) {
            case PandaGpuManagerInitUnitType::None: throw PandaGpuException(PandaGpuError::MSMBasesAddrError);
            case PandaGpuManagerInitUnitType::MSM:
                if (!bases) throw PandaGpuException(PandaGpuError::MSMBasesAddrError);
                gm.d_bases = init_msm(*bases);
                break;
            case PandaGpuManagerInitUnitType::NTT:
                if (!omega) throw PandaGpuException(PandaGpuError::NTTOmegaAddrError);
                init_ntt(*omega);
                break;
            case PandaGpuManagerInitUnitType::ALL:
                if (!bases) throw PandaGpuException(PandaGpuError::MSMBasesAddrError);
                gm.d_bases = init_msm(*bases);
                if (!omega) throw PandaGpuException(PandaGpuError::NTTOmegaAddrError);
                init_ntt(*omega);
                break;
        }
        gm.make_streams();
        return gm;
    }
    static PandaMemPool init_hardware(size_t device_id) {   // wrapper.rs:115-120
        try { set_device(device_id); } catch (const PandaGpuException &) {}
        return PandaMemPool::create(device_id);
    }
    static void *upload(const ByteSlice &b) {   // panda_malloc + blocking panda_memcpy, wrapper.rs:131-146 / 154-170
        void *d = nullptr;
        check(panda_malloc(&d, b.len), PandaGpuError::CreateContextError);
        check(panda_memcpy(d, b.data, b.len), PandaGpuError::CreateContextError);
        return d;
    }
    // cached bases are announced to the library (panda_msm_register_bases_bn254): it builds its table of precomputed multiples
    // once, here, and every MSM on the cached pointer runs without a content check.  No counterpart in wrapper.rs: the
    // reference's panda_msm_setup_bn254 is an empty hook (msm_cuda.cuh:786-795).
    static void *upload_bases(const ByteSlice &b, PandaCurve curve = PandaCurve::Bn254) {
        void *d = upload(b);
        panda_stream null_stream{nullptr};
        const size_t points = b.len / (2 * fq_len(curve));
        if (points) check(curve == PandaCurve::Bls12_377 ? panda_msm_register_bases_bls12_377(d, points, null_stream)
                                                         : panda_msm_register_bases_bn254(d, points, null_stream), PandaGpuError::SetBasesErr);
        return d;
    }
    static std::vector<void *> init_msm(const std::vector<ByteSlice> &bases) {   // wrapper.rs:122-152
        std::vector<void *> out;
        for (const auto &b : bases) out.push_back(upload_bases(b));
        check(panda_msm_setup_bn254(), PandaGpuError::CreateContextError);
        return out;
    }
    static void *init_msm_cached_bases(const ByteSlice &bases, PandaCurve curve = PandaCurve::Bn254) { return upload_bases(bases, curve); } // wrapper.rs:154-170
    static void *init_msm_cached_scalars(const ByteSlice &scalars) { return upload(scalars); }   // wrapper.rs:172-188
    static std::pair<void *, void *> init_msm_cached(const ByteSlice &scalars, const ByteSlice &bases) {   // wrapper.rs:190-197
        void *s = init_msm_cached_scalars(scalars);
        void *b = init_msm_cached_bases(bases);
        return {s, b};
    }
    static void init_ntt(const ByteSlice &omega) {   // wrapper.rs:199-210
        check(panda_ntt_setup_bn254(const_cast<uint8_t *>(omega.data)), PandaGpuError::CreateContextError);
    }

    void set_config(PandaMSMResultCoordinateType t) { coord_ = t; }   // wrapper.rs:212-214
    PandaMemPool get_mem_pool() const { return mem_pool_; }
    PandaStream get_stream() const { return default_stream_; }
    PandaStream get_h2d_stream() const { return h2d_stream_; }
    PandaStream get_d2h_stream() const { return d2h_stream_; }
    PandaStream get_exec_stream() const { return exec_stream_; }
    PandaMSMResultCoordinateType get_msm_result_coordinate_type() const { return coord_; }
    void *get_params_bases_ptr_mut(size_t i) const { return i < d_bases.size() ? d_bases[i] : nullptr; }
    void *get_params_scalars_ptr_mut(size_t i) const { return i < d_scalars.size() ? d_scalars[i] : nullptr; }
    size_t get_params_scalars_len(size_t i) const { return i < scalars_len.size() ? scalars_len[i] : 0; }

    void wait_h2d() const {   // wrapper.rs:256-262
        PandaEvent e = PandaEvent::create();
        e.record(h2d_stream_);
        exec_stream_.wait(e.raw);
        e.destroy();          // fix: the reference never destroys its events
    }
    void wait_exec() const {   // wrapper.rs:264-269
        PandaEvent e = PandaEvent::create();
        e.record(exec_stream_);
        d2h_stream_.wait(e.raw);
        e.destroy();
    }
    void destroy() const { check(panda_mem_pool_destroy(mem_pool_.raw), PandaGpuError::DestroyContextErr); }   // wrapper.rs:271-279
    void sync() const { h2d_stream_.sync(); exec_stream_.sync(); d2h_stream_.sync(); }                      // wrapper.rs:281-287
    size_t device_id() const { return device_id_; }
    void deinit() {   // wrapper.rs:297-312 (tear_down is idempotent here; streams are destroyed too)
        sync();
        for (void *p : d_bases) { check(panda_msm_unregister_bases(p), PandaGpuError::DestroyContextErr); check(panda_free(p), PandaGpuError::DestroyContextErr); check(panda_msm_tear_down(), PandaGpuError::DestroyContextErr); }
        for (void *p : d_scalars) check(panda_free(p), PandaGpuError::DestroyContextErr);
        d_bases.clear(); d_scalars.clear(); scalars_len.clear();
        check(panda_mem_pool_destroy(mem_pool_.raw), PandaGpuError::DestroyContextErr);
        default_stream_.destroy(); h2d_stream_.destroy(); d2h_stream_.destroy(); exec_stream_.destroy();
    }

  private:
    void make_streams() {
        default_stream_ = PandaStream::create(); h2d_stream_ = PandaStream::create();
        d2h_stream_ = PandaStream::create(); exec_stream_ = PandaStream::create();
    }
    size_t device_id_ = 0;
    PandaMemPool mem_pool_{};
    PandaStream default_stream_{}, h2d_stream_{}, d2h_stream_{}, exec_stream_{};
    PandaMSMResultCoordinateType coord_ = PandaMSMResultCoordinateType::Jacobian;
};

// gpu_manager/common.rs:54-76
inline void *memory_alloc_and_copy(const PandaGpuManager &gm, const ByteSlice &h, const PandaStream &stream) {
    void *d = nullptr;
    malloc_from_pool_async(&d, h.len, gm.get_mem_pool(), stream);
    memcpy_async(d, h.data, h.len, stream);
    return d;
}
inline void memory_copy_and_free(uint8_t *h, size_t len, void *d, const PandaStream &stream) {
    memcpy_async(h, d, len, stream);
    free_async(d, stream);
}

// ---- gpu_manager/unit.rs -----------------------------------------------------------------------------------------------------
namespace detail {

// steps 3-7 of unit.rs:31-100, shared by the four device MSM variants (of either curve)
inline std::vector<uint8_t> msm_execute_and_fetch(const PandaGpuManager &gm, void *d_scalars, void *d_bases, size_t scalars_len,
                                                  bool free_scalars, bool free_bases, bool scalars_on_host = false, PandaCurve curve = PandaCurve::Bn254) {
    const uint32_t log_scalars_count = log_2(scalars_len / FIELD_ELEMENT_LEN);
    const size_t result_buf_len = fq_len(curve) * 3;
    const bool bls = curve == PandaCurve::Bls12_377;
    void *d_result = nullptr;
    malloc_from_pool_async(&d_result, result_buf_len, gm.get_mem_pool(), gm.get_exec_stream());   // fix: exec stream (A12)
    panda_msm_configuration cfg{};
    cfg.mem_pool = gm.get_mem_pool().raw;
    cfg.stream = gm.get_exec_stream().raw;
    cfg.bases = d_bases;
    cfg.scalars = d_scalars;
    cfg.results = d_result;
    cfg.log_scalars_count = log_scalars_count;
    cfg.msm_result_coordinate_type = static_cast<panda_msm_result_coordinate_type>(gm.get_msm_result_coordinate_type());
    const size_t n = (size_t)1 << log_scalars_count;
    if (scalars_on_host) check(bls ? panda_msm_execute_bls12_377_host_scalars(cfg, n) : panda_msm_execute_bn254_host_scalars(cfg, n), PandaGpuError::SchedulingErr);
    else check(bls ? panda_msm_execute_bls12_377(cfg) : panda_msm_execute_bn254(cfg), PandaGpuError::SchedulingErr);
    std::vector<uint8_t> out(result_buf_len);
    // fix: unit.rs:67-74 allocates (and leaks) a pinned staging buffer per call; the result bytes go straight into the result vector
    panda_error rc = panda_memcpy_async(out.data(), d_result, result_buf_len, gm.get_exec_stream().raw);
    if (rc == panda_success) rc = panda_stream_synchronize(gm.get_exec_stream().raw);               // unit.rs:60-62 + :76
    if (rc != panda_success) throw PandaGpuException(PandaGpuError::CreateContextError);
    if (free_scalars) free_async(d_scalars, gm.get_exec_stream());
    if (free_bases) free_async(d_bases, gm.get_exec_stream());
    free_async(d_result, gm.get_exec_stream());
    return out;
}

}  // namespace detail

// The five MSM shapes of unit.rs, with the curve as a parameter; the reference's names (BN254) and the BLS12-377 names follow.
// unit.rs:10-101
inline std::vector<uint8_t> panda_msm_gpu(const PandaGpuManager &gm, PandaCurve curve, const ByteSlice &scalars, const ByteSlice &bases) {
    if (bases.len / (2 * fq_len(curve)) < ((size_t)1 << log_2(scalars.len / FIELD_ELEMENT_LEN))) throw PandaGpuException(PandaGpuError::MSMBasesAddrError);
    void *d_scalars = memory_alloc_and_copy(gm, scalars, gm.get_h2d_stream());
    void *d_bases = memory_alloc_and_copy(gm, bases, gm.get_h2d_stream());
    gm.wait_h2d();
    return detail::msm_execute_and_fetch(gm, d_scalars, d_bases, scalars.len, true, true, false, curve);
}
// unit.rs:103-188
inline std::vector<uint8_t> panda_msm_gpu_with_cached_bases(const PandaGpuManager &gm, PandaCurve curve, const ByteSlice &scalars, size_t bases_index) {
    void *d_bases = gm.get_params_bases_ptr_mut(bases_index);
    if (!d_bases) throw PandaGpuException(PandaGpuError::BasesIndexErr);
    // fix: unit.rs:113-131 uploads every scalar before the MSM starts; panda_msm_execute_*_host_scalars streams them in chunks
    // on the library's copy stream and overlaps the upload with the sort / accumulation of the chunks already on the device
    return detail::msm_execute_and_fetch(gm, const_cast<uint8_t *>(scalars.data), d_bases, scalars.len, false, false, true, curve);
}
// unit.rs:190-275
inline std::vector<uint8_t> panda_msm_gpu_with_cached_scalars(const PandaGpuManager &gm, PandaCurve curve, size_t scalars_index, const ByteSlice &bases) {
    const size_t len = gm.get_params_scalars_len(scalars_index);
    void *d_scalars = gm.get_params_scalars_ptr_mut(scalars_index);
    if (len == 0 || !d_scalars) throw PandaGpuException(PandaGpuError::BasesIndexErr);
    void *d_bases = memory_alloc_and_copy(gm, bases, gm.get_h2d_stream());
    gm.wait_h2d();
    return detail::msm_execute_and_fetch(gm, d_scalars, d_bases, len, false, true, false, curve);
}
// unit.rs:277-361
inline std::vector<uint8_t> panda_msm_gpu_with_cached_input(const PandaGpuManager &gm, PandaCurve curve, size_t scalars_index, size_t bases_index) {
    const size_t len = gm.get_params_scalars_len(scalars_index);
    if (len == 0) throw PandaGpuException(PandaGpuError::BasesIndexErr);
    void *d_scalars = gm.get_params_scalars_ptr_mut(scalars_index);
    void *d_bases = gm.get_params_bases_ptr_mut(bases_index);
    if (!d_scalars || !d_bases) throw PandaGpuException(PandaGpuError::BasesIndexErr);
    return detail::msm_execute_and_fetch(gm, d_scalars, d_bases, len, false, false, false, curve);
}
// unit.rs:363-416: host pointers straight into panda_msm_execute_*_host
inline std::vector<uint8_t> panda_msm_gpu_host(const PandaGpuManager &gm, PandaCurve curve, const ByteSlice &scalars, const ByteSlice &bases) {
    std::vector<uint8_t> out(fq_len(curve) * 3);
    panda_msm_configuration cfg{};
    cfg.mem_pool = gm.get_mem_pool().raw;
    cfg.stream = gm.get_exec_stream().raw;
    cfg.bases = const_cast<uint8_t *>(bases.data);
    cfg.scalars = const_cast<uint8_t *>(scalars.data);
    cfg.results = out.data();
    cfg.log_scalars_count = log_2(scalars.len / FIELD_ELEMENT_LEN);
    cfg.msm_result_coordinate_type = static_cast<panda_msm_result_coordinate_type>(gm.get_msm_result_coordinate_type());
    check(curve == PandaCurve::Bls12_377 ? panda_msm_execute_bls12_377_host(cfg) : panda_msm_execute_bn254_host(cfg), PandaGpuError::SchedulingErr);
    return out;
}

// the reference's API (BN254)
inline std::vector<uint8_t> panda_msm_bn254_gpu(const PandaGpuManager &gm, const ByteSlice &scalars, const ByteSlice &bases) {
    return panda_msm_gpu(gm, PandaCurve::Bn254, scalars, bases);
}
inline std::vector<uint8_t> panda_msm_bn254_gpu_with_cached_bases(const PandaGpuManager &gm, const ByteSlice &scalars, size_t bases_index) {
    return panda_msm_gpu_with_cached_bases(gm, PandaCurve::Bn254, scalars, bases_index);
}
inline std::vector<uint8_t> panda_msm_bn254_gpu_with_cached_scalars(const PandaGpuManager &gm, size_t scalars_index, const ByteSlice &bases) {
    return panda_msm_gpu_with_cached_scalars(gm, PandaCurve::Bn254, scalars_index, bases);
}
inline std::vector<uint8_t> panda_msm_bn254_gpu_with_cached_input(const PandaGpuManager &gm, size_t scalars_index, size_t bases_index) {
    return panda_msm_gpu_with_cached_input(gm, PandaCurve::Bn254, scalars_index, bases_index);
}
inline std::vector<uint8_t> panda_msm_bn254_gpu_host(const PandaGpuManager &gm, const ByteSlice &scalars, const ByteSlice &bases) {
    return panda_msm_gpu_host(gm, PandaCurve::Bn254, scalars, bases);
}

namespace detail {
inline void ntt_fetch(const PandaGpuManager &gm, uint8_t *scalars, size_t len, void *d_src, void *d_dst, unsigned flag) {
    // unit.rs:457-476: copy back from the buffer the flag names, free both
    panda_error rc = panda_memcpy(scalars, flag == 0 ? d_src : d_dst, len);
    free_async(d_src, gm.get_exec_stream());
    free_async(d_dst, gm.get_exec_stream());
    if (rc != panda_success) throw PandaGpuException(PandaGpuError::CreateContextError);
}
}  // namespace detail

// unit.rs:418-479
inline void panda_ntt_bn254_gpu(const PandaGpuManager &gm, uint8_t *scalars, size_t len, uint32_t log_n) {
    if (len != (size_t(1) << log_n) * 32) throw std::invalid_argument("scalars.len() != (1 << log_n) * 32");   // unit.rs:423 assert_eq!
    void *d_src = memory_alloc_and_copy(gm, ByteSlice{scalars, len}, gm.get_h2d_stream());
    void *d_dst = nullptr;
    malloc_from_pool_async(&d_dst, len, gm.get_mem_pool(), gm.get_h2d_stream());
    gm.wait_h2d();                                                     // fix: the reference executes without waiting for the upload
    unsigned flag = 0;
    panda_ntt_configuration cfg{gm.get_mem_pool().raw, gm.get_exec_stream().raw, d_src, d_dst, log_n, &flag};
    check(panda_ntt_execute_bn254(cfg), PandaGpuError::SchedulingErr);
    gm.get_exec_stream().sync();
    detail::ntt_fetch(gm, scalars, len, d_src, d_dst, flag);
}
// unit.rs:481-543
inline void panda_ntt_bn254_gpu_v1(const PandaGpuManager &gm, uint8_t *scalars, size_t len, const ByteSlice &omega, uint32_t log_n) {
    void *d_src = memory_alloc_and_copy(gm, ByteSlice{scalars, len}, gm.get_h2d_stream());
    void *d_dst = nullptr;
    malloc_from_pool_async(&d_dst, len, gm.get_mem_pool(), gm.get_h2d_stream());
    gm.wait_h2d();
    unsigned flag = 0;
    panda_ntt_configuration_v1 cfg{gm.get_mem_pool().raw, gm.get_exec_stream().raw, d_src, d_dst, const_cast<uint8_t *>(omega.data), log_n, &flag};
    check(panda_ntt_execute_bn254_v1(cfg), PandaGpuError::SchedulingErr);
    gm.get_exec_stream().sync();
    detail::ntt_fetch(gm, scalars, len, d_src, d_dst, flag);
}

// ---- additions: the entry points README.md:36 promises ("BLS12-377 ... will be easy") and the inverse transform, in the same shape as the
// functions above (no counterpart in unit.rs) ------------------------------------------------------------------------------------------

// panda_ntt_bn254_gpu_v1's inverse: scalars <- (1/n) DFT_{omega^-1}(scalars), omega = the FORWARD root
inline void panda_intt_bn254_gpu_v1(const PandaGpuManager &gm, uint8_t *scalars, size_t len, const ByteSlice &omega, uint32_t log_n) {
    if (len != (size_t(1) << log_n) * 32) throw std::invalid_argument("scalars.len() != (1 << log_n) * 32");
    void *d_src = memory_alloc_and_copy(gm, ByteSlice{scalars, len}, gm.get_h2d_stream());
    void *d_dst = nullptr;
    malloc_from_pool_async(&d_dst, len, gm.get_mem_pool(), gm.get_h2d_stream());
    gm.wait_h2d();
    unsigned flag = 0;
    panda_ntt_configuration_v1 cfg{gm.get_mem_pool().raw, gm.get_exec_stream().raw, d_src, d_dst, const_cast<uint8_t *>(omega.data), log_n, &flag};
    check(panda_intt_execute_bn254_v1(cfg), PandaGpuError::SchedulingErr);
    gm.get_exec_stream().sync();
    detail::ntt_fetch(gm, scalars, len, d_src, d_dst, flag);
}

// BLS12-377 G1 (bases 96 B per point: x || y, 12 x u32 Montgomery; scalars 32 B; result 144 B, Jacobian or Projective per set_config):
// the same five shapes.  Cached bases of this curve are uploaded with init_msm_cached_bases(bases, PandaCurve::Bls12_377).
inline std::vector<uint8_t> panda_msm_bls12_377_gpu(const PandaGpuManager &gm, const ByteSlice &scalars, const ByteSlice &bases) {
    return panda_msm_gpu(gm, PandaCurve::Bls12_377, scalars, bases);
}
inline std::vector<uint8_t> panda_msm_bls12_377_gpu_with_cached_bases(const PandaGpuManager &gm, const ByteSlice &scalars, size_t bases_index) {
    return panda_msm_gpu_with_cached_bases(gm, PandaCurve::Bls12_377, scalars, bases_index);
}
inline std::vector<uint8_t> panda_msm_bls12_377_gpu_with_cached_scalars(const PandaGpuManager &gm, size_t scalars_index, const ByteSlice &bases) {
    return panda_msm_gpu_with_cached_scalars(gm, PandaCurve::Bls12_377, scalars_index, bases);
}
inline std::vector<uint8_t> panda_msm_bls12_377_gpu_with_cached_input(const PandaGpuManager &gm, size_t scalars_index, size_t bases_index) {
    return panda_msm_gpu_with_cached_input(gm, PandaCurve::Bls12_377, scalars_index, bases_index);
}
inline std::vector<uint8_t> panda_msm_bls12_377_gpu_host(const PandaGpuManager &gm, const ByteSlice &scalars, const ByteSlice &bases) {
    return panda_msm_gpu_host(gm, PandaCurve::Bls12_377, scalars, bases);
}

}  // namespace panda
