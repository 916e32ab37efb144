// panda_host_capi.cpp -- flat C entry points over panda_gpu_manager.hpp so that the host API can be driven from
// Python (tests/, bench.py) without a Rust toolchain.  One function per Rust API item; errors come back as
// -(1 + PandaGpuError index), 0 = Ok.  Built into libpanda-host.so, which links libpanda-cuda.so.
#include "panda_gpu_manager.hpp"

#include <new>

using namespace panda;

namespace {
template <class F>
int guarded(F &&f) {
    try { f(); return 0; }
    catch (const PandaGpuException &e) { return -(1 + static_cast<int>(e.kind)); }
    catch (const std::bad_alloc &) { return -1000; }
    catch (const std::exception &) { return -1001; }
}
inline ByteSlice bs(const void *p, size_t n) { return ByteSlice{static_cast<const uint8_t *>(p), n}; }
}  // namespace

extern "C" {

const char *panda_host_error_name(int code) { return code < 0 && code > -1000 ? to_string(static_cast<PandaGpuError>(-code - 1)) : (code == 0 ? "Ok" : "HostError"); }

int panda_host_get_device_number(int *count) { return guarded([&] { *count = get_device_number(); }); }
int panda_host_device_info(int device_id, unsigned long long *free_bytes, unsigned long long *total_bytes) {
    return guarded([&] { PandaDeviceInfo i = device_info(device_id); *free_bytes = i.free; *total_bytes = i.total; });
}
int panda_host_set_device(size_t device_id) { return guarded([&] { set_device(device_id); }); }

int panda_host_manager_new(size_t device_id, void **out) {
    return guarded([&] { *out = new PandaGpuManager(PandaGpuManager::create(device_id)); });
}
// unit: 0 None, 1 MSM, 2 NTT, 3 ALL; bases: `bases_count` (ptr,len) pairs; omega may be NULL
int panda_host_manager_init_all(size_t device_id, int unit, const void *const *bases_ptrs, const size_t *bases_lens, size_t bases_count,
                                const void *omega, void **out) {
    return guarded([&] {
        std::vector<ByteSlice> b;
        for (size_t i = 0; i < bases_count; i++) b.push_back(bs(bases_ptrs[i], bases_lens[i]));
        ByteSlice om = bs(omega, 32);
        *out = new PandaGpuManager(PandaGpuManager::init_all(device_id, static_cast<PandaGpuManagerInitUnitType>(unit),
                                                             bases_ptrs ? &b : nullptr, omega ? &om : nullptr));
    });
}
int panda_host_manager_deinit(void *gm) { return guarded([&] { auto *m = static_cast<PandaGpuManager *>(gm); m->deinit(); delete m; }); }
int panda_host_manager_set_config(void *gm, int coord) {
    return guarded([&] { static_cast<PandaGpuManager *>(gm)->set_config(static_cast<PandaMSMResultCoordinateType>(coord)); });
}
int panda_host_manager_sync(void *gm) { return guarded([&] { static_cast<PandaGpuManager *>(gm)->sync(); }); }
size_t panda_host_manager_device_id(void *gm) { return static_cast<PandaGpuManager *>(gm)->device_id(); }
int panda_host_init_ntt(const void *omega) { return guarded([&] { PandaGpuManager::init_ntt(bs(omega, 32)); }); }

// init_msm_cached_bases / _scalars + pushing the pointer into the manager's public vectors (what a Rust caller does by hand)
int panda_host_manager_cache_bases(void *gm, const void *bases, size_t len, size_t *index) {
    return guarded([&] {
        auto *m = static_cast<PandaGpuManager *>(gm);
        m->d_bases.push_back(PandaGpuManager::init_msm_cached_bases(bs(bases, len)));
        *index = m->d_bases.size() - 1;
    });
}
int panda_host_manager_cache_bases_curve(void *gm, int curve, const void *bases, size_t len, size_t *index) {
    return guarded([&] {
        auto *m = static_cast<PandaGpuManager *>(gm);
        m->d_bases.push_back(PandaGpuManager::init_msm_cached_bases(bs(bases, len), static_cast<PandaCurve>(curve)));
        *index = m->d_bases.size() - 1;
    });
}
int panda_host_manager_cache_scalars(void *gm, const void *scalars, size_t len, size_t *index) {
    return guarded([&] {
        auto *m = static_cast<PandaGpuManager *>(gm);
        m->d_scalars.push_back(PandaGpuManager::init_msm_cached_scalars(bs(scalars, len)));
        m->scalars_len.push_back(len);
        *index = m->d_scalars.size() - 1;
    });
}
void *panda_host_manager_bases_ptr(void *gm, size_t index) { return static_cast<PandaGpuManager *>(gm)->get_params_bases_ptr_mut(index); }
void *panda_host_manager_scalars_ptr(void *gm, size_t index) { return static_cast<PandaGpuManager *>(gm)->get_params_scalars_ptr_mut(index); }
void *panda_host_manager_exec_stream(void *gm) { return static_cast<PandaGpuManager *>(gm)->get_exec_stream().raw.handle; }
void *panda_host_manager_mem_pool(void *gm) { return static_cast<PandaGpuManager *>(gm)->get_mem_pool().raw.handle; }

// MSM variants; result96 receives FIELD_ELEMENT_LEN * 3 bytes
int panda_host_msm_bn254_gpu(void *gm, const void *scalars, size_t scalars_len, const void *bases, size_t bases_len, void *result96) {
    return guarded([&] { auto r = panda_msm_bn254_gpu(*static_cast<PandaGpuManager *>(gm), bs(scalars, scalars_len), bs(bases, bases_len)); memcpy(result96, r.data(), r.size()); });
}
int panda_host_msm_bn254_gpu_with_cached_bases(void *gm, const void *scalars, size_t scalars_len, size_t bases_index, void *result96) {
    return guarded([&] { auto r = panda_msm_bn254_gpu_with_cached_bases(*static_cast<PandaGpuManager *>(gm), bs(scalars, scalars_len), bases_index); memcpy(result96, r.data(), r.size()); });
}
int panda_host_msm_bn254_gpu_with_cached_scalars(void *gm, size_t scalars_index, const void *bases, size_t bases_len, void *result96) {
    return guarded([&] { auto r = panda_msm_bn254_gpu_with_cached_scalars(*static_cast<PandaGpuManager *>(gm), scalars_index, bs(bases, bases_len)); memcpy(result96, r.data(), r.size()); });
}
int panda_host_msm_bn254_gpu_with_cached_input(void *gm, size_t scalars_index, size_t bases_index, void *result96) {
    return guarded([&] { auto r = panda_msm_bn254_gpu_with_cached_input(*static_cast<PandaGpuManager *>(gm), scalars_index, bases_index); memcpy(result96, r.data(), r.size()); });
}
int panda_host_msm_bn254_gpu_host(void *gm, const void *scalars, size_t scalars_len, const void *bases, size_t bases_len, void *result96) {
    return guarded([&] { auto r = panda_msm_bn254_gpu_host(*static_cast<PandaGpuManager *>(gm), bs(scalars, scalars_len), bs(bases, bases_len)); memcpy(result96, r.data(), r.size()); });
}
int panda_host_ntt_bn254_gpu(void *gm, void *scalars, size_t len, unsigned log_n) {
    return guarded([&] { panda_ntt_bn254_gpu(*static_cast<PandaGpuManager *>(gm), static_cast<uint8_t *>(scalars), len, log_n); });
}
int panda_host_ntt_bn254_gpu_v1(void *gm, void *scalars, size_t len, const void *omega, unsigned log_n) {
    return guarded([&] { panda_ntt_bn254_gpu_v1(*static_cast<PandaGpuManager *>(gm), static_cast<uint8_t *>(scalars), len, bs(omega, 32), log_n); });
}

int panda_host_intt_bn254_gpu_v1(void *gm, void *scalars, size_t len, const void *omega, unsigned log_n) {
    return guarded([&] { panda_intt_bn254_gpu_v1(*static_cast<PandaGpuManager *>(gm), static_cast<uint8_t *>(scalars), len, bs(omega, 32), log_n); });
}
// the five MSM shapes with the curve as a parameter (0 BN254, 1 BLS12-377); result receives 3 base-field elements (96 / 144 bytes)
int panda_host_msm_gpu(void *gm, int curve, const void *scalars, size_t scalars_len, const void *bases, size_t bases_len, void *result) {
    return guarded([&] { auto r = panda_msm_gpu(*static_cast<PandaGpuManager *>(gm), static_cast<PandaCurve>(curve), bs(scalars, scalars_len), bs(bases, bases_len)); memcpy(result, r.data(), r.size()); });
}
int panda_host_msm_gpu_with_cached_bases(void *gm, int curve, const void *scalars, size_t scalars_len, size_t bases_index, void *result) {
    return guarded([&] { auto r = panda_msm_gpu_with_cached_bases(*static_cast<PandaGpuManager *>(gm), static_cast<PandaCurve>(curve), bs(scalars, scalars_len), bases_index); memcpy(result, r.data(), r.size()); });
}
int panda_host_msm_gpu_with_cached_scalars(void *gm, int curve, size_t scalars_index, const void *bases, size_t bases_len, void *result) {
    return guarded([&] { auto r = panda_msm_gpu_with_cached_scalars(*static_cast<PandaGpuManager *>(gm), static_cast<PandaCurve>(curve), scalars_index, bs(bases, bases_len)); memcpy(result, r.data(), r.size()); });
}
int panda_host_msm_gpu_with_cached_input(void *gm, int curve, size_t scalars_index, size_t bases_index, void *result) {
    return guarded([&] { auto r = panda_msm_gpu_with_cached_input(*static_cast<PandaGpuManager *>(gm), static_cast<PandaCurve>(curve), scalars_index, bases_index); memcpy(result, r.data(), r.size()); });
}
int panda_host_msm_gpu_host(void *gm, int curve, const void *scalars, size_t scalars_len, const void *bases, size_t bases_len, void *result) {
    return guarded([&] { auto r = panda_msm_gpu_host(*static_cast<PandaGpuManager *>(gm), static_cast<PandaCurve>(curve), bs(scalars, scalars_len), bs(bases, bases_len)); memcpy(result, r.data(), r.size()); });
}
int panda_host_msm_bls12_377_gpu(void *gm, const void *scalars, size_t scalars_len, const void *bases, size_t bases_len, void *result144) {
    return panda_host_msm_gpu(gm, 1, scalars, scalars_len, bases, bases_len, result144);
}

}  // extern "C"
