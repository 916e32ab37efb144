// multi.cu -- the multi-GPU entry points of the C ABI: ONE process drives the GPUs of one box (SURVEY.md section 8b "new symbols",
// section 8e).  The reference is single-device (cudaSetDevice(0) inside execute, msm_cuda.cuh:554-555; "Supports one GPU by
// default", src/gpu_manager/wrapper.rs:38) but already declares the peer-access calls a multi-GPU caller needs
// (src/gpu_ffi/binding.rs:54-56); these are the functions such a caller would bind.
//
//   MSM  sharded by contiguous point range: every GPU runs the whole single-GPU pipeline on its shard (its slice of the cached
//        bases and its table live only there), the 96-byte Jacobian partials are copied peer-to-peer to the first GPU and summed.
//        One host thread per GPU queues the work, so the launch sequences of the shards overlap.
//   NTT  four-step transform n = n1 * n2 with ONE exchange: local transpose, batched n1-point transforms, the fused
//        twiddle + transpose + all-to-all kernel storing over NVLink straight into the peers' output buffers (k_ntt_exchange), batched
//        n2-point transforms.  Same sharded layout as panda_b200/sharded_ntt.py: column blocks in, row blocks out.
// Ordering across devices is by CUDA events only (no host synchronisation); everything is asynchronous on the per-device streams.
#include "panda_interface.h"
#include "msm.cuh"
#include "ntt.cuh"

#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

namespace {

inline panda_error perr(cudaError_t e) { return static_cast<panda_error>(e); }
inline cudaStream_t cu(panda_stream s) { return static_cast<cudaStream_t>(s.handle); }
inline cudaMemPool_t cu(panda_mem_pool p) { return static_cast<cudaMemPool_t>(p.handle); }

struct DeviceGuard {            // the entry points leave the caller's current device as they found it
    int saved = 0;
    DeviceGuard() { cudaGetDevice(&saved); }
    ~DeviceGuard() { cudaSetDevice(saved); }
};

cudaError_t device_of(const void *ptr, int *dev) {
    cudaPointerAttributes a{};
    cudaError_t e = cudaPointerGetAttributes(&a, ptr);
    if (e != cudaSuccess) return e;
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return cudaErrorInvalidValue;
    *dev = a.device;
    return cudaSuccess;
}

// peer access both ways between all listed devices; already enabled is fine
cudaError_t enable_peer_access(const std::vector<int> &devs) {
    for (int a : devs) {
        cudaError_t e = cudaSetDevice(a);
        if (e != cudaSuccess) return e;
        for (int b : devs) {
            if (a == b) continue;
            int can = 0;
            e = cudaDeviceCanAccessPeer(&can, a, b);
            if (e != cudaSuccess) return e;
            if (!can) return cudaErrorPeerAccessUnsupported;
            e = cudaDeviceEnablePeerAccess(b, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

panda_error msm_multi(pb::CurveId curve, const panda_msm_configuration *cfgs, const size_t *counts, int n_dev) {
    if (!cfgs || n_dev <= 0 || n_dev > 64) return perr(cudaErrorInvalidValue);
    const size_t rbytes = 3 * pb::curve_fq_bytes(curve);
    DeviceGuard guard;
    std::vector<int> dev(n_dev);
    std::vector<size_t> n(n_dev);
    for (int d = 0; d < n_dev; d++) {
        if (!cfgs[d].results) return perr(cudaErrorInvalidValue);
        if (!counts && cfgs[d].log_scalars_count > 30) return perr(cudaErrorInvalidValue);
        n[d] = counts ? counts[d] : (size_t)1 << cfgs[d].log_scalars_count;
        if (n[d] > ((size_t)1 << 30) || (n[d] && (!cfgs[d].bases || !cfgs[d].scalars))) return perr(cudaErrorInvalidValue);
        cudaError_t e = device_of(cfgs[d].results, &dev[d]);
        if (e != cudaSuccess) return perr(e);
    }
    // 1. every shard on its own GPU, queued by its own host thread; partial d lands in cfgs[d].results (Jacobian)
    std::vector<cudaError_t> rc(n_dev, cudaSuccess);
    std::vector<cudaEvent_t> done(n_dev, nullptr);
    auto shard = [&](int d) {
        cudaError_t e = cudaSetDevice(dev[d]);
        if (e == cudaSuccess)
            e = pb::msm_run(curve, cfgs[d].bases, cfgs[d].scalars, (uint32_t)n[d], cfgs[d].results, pb::COORD_JACOBIAN, cu(cfgs[d].mem_pool), cu(cfgs[d].stream));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[d], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(done[d], cu(cfgs[d].stream));
        rc[d] = e;
    };
    {
        std::vector<std::thread> workers;
        for (int d = 1; d < n_dev; d++) workers.emplace_back(shard, d);
        shard(0);
        for (auto &w : workers) w.join();
    }
    cudaError_t err = cudaSuccess;
    for (int d = 0; d < n_dev; d++) if (rc[d] != cudaSuccess && err == cudaSuccess) err = rc[d];
    // 2. the tiny exchange: n_dev x 96 bytes to the first GPU (peer copies ordered behind the shards' events), then the sum
    if (err == cudaSuccess && (n_dev > 1 || cfgs[0].msm_result_coordinate_type == PROJECTIVE)) {
        cudaStream_t s0 = cu(cfgs[0].stream);
        uint8_t *gather = nullptr;
        err = cudaSetDevice(dev[0]);
        if (err == cudaSuccess) err = cudaMallocAsync((void **)&gather, (size_t)n_dev * rbytes, s0);
        if (err == cudaSuccess) {
            for (int d = 0; d < n_dev && err == cudaSuccess; d++) {
                if (d) err = cudaStreamWaitEvent(s0, done[d], 0);
                if (err == cudaSuccess) err = cudaMemcpyPeerAsync(gather + (size_t)d * rbytes, dev[0], cfgs[d].results, dev[d], rbytes, s0);
            }
            if (err == cudaSuccess)
                err = pb::msm_combine(curve, gather, (uint32_t)n_dev, cfgs[0].results,
                                      cfgs[0].msm_result_coordinate_type == PROJECTIVE ? pb::COORD_PROJECTIVE : pb::COORD_JACOBIAN, s0);
            cudaError_t f = cudaFreeAsync(gather, s0);
            if (err == cudaSuccess) err = f;
        }
    }
    for (int d = 0; d < n_dev; d++) if (done[d]) cudaEventDestroy(done[d]);      // a destroyed event still orders the waits already queued on it
    if (err != cudaSuccess) fprintf(stderr, "[panda-b200] multi-GPU MSM failed: %s\n", cudaGetErrorString(err));
    return perr(err);
}

}  // namespace

namespace {
// sub-roots of the four-step transform: omega^(n2) has order n1 (the column transforms), omega^(n1) order n2 (the row transforms)
void sub_roots(const void *omega_host, unsigned l1, unsigned l2, unsigned char *rows32, unsigned char *cols32) {
    pb::ntt_pow2k_host(omega_host, l2, rows32);
    pb::ntt_pow2k_host(omega_host, l1, cols32);
}
}  // namespace

extern "C" {

panda_error panda_msm_execute_bn254_multi(const panda_msm_configuration *per_device, int n_dev) { return msm_multi(pb::CURVE_BN254, per_device, nullptr, n_dev); }
panda_error panda_msm_execute_bn254_multi_n(const panda_msm_configuration *per_device, const size_t *counts, int n_dev) {
    return counts ? msm_multi(pb::CURVE_BN254, per_device, counts, n_dev) : perr(cudaErrorInvalidValue);
}
panda_error panda_msm_execute_bls12_377_multi(const panda_msm_configuration *per_device, int n_dev) {
    return msm_multi(pb::CURVE_BLS12_377, per_device, nullptr, n_dev);
}
panda_error panda_msm_execute_bls12_377_multi_n(const panda_msm_configuration *per_device, const size_t *counts, int n_dev) {
    return counts ? msm_multi(pb::CURVE_BLS12_377, per_device, counts, n_dev) : perr(cudaErrorInvalidValue);
}

panda_error panda_msm_execute_bls12_381_multi(const panda_msm_configuration *per_device, int n_dev) {
    return msm_multi(pb::CURVE_BLS12_381, per_device, nullptr, n_dev);
}
panda_error panda_msm_execute_bls12_381_multi_n(const panda_msm_configuration *per_device, const size_t *counts, int n_dev) {
    return counts ? msm_multi(pb::CURVE_BLS12_381, per_device, counts, n_dev) : perr(cudaErrorInvalidValue);
}

panda_error panda_ntt_execute_bn254_multi(const panda_ntt_multi_configuration *cfg) {
    if (!cfg || !cfg->streams || !cfg->d_src || !cfg->d_dst || !cfg->omega) return perr(cudaErrorInvalidValue);
    const unsigned G = cfg->n_dev, log_n = cfg->log_n;
    if (G == 0 || G > 16 || (G & (G - 1)) || log_n > 28) return perr(cudaErrorInvalidValue);
    unsigned lg = 0; while ((1u << lg) < G) lg++;
    const unsigned l1 = log_n / 2, l2 = log_n - l1;
    if (l1 < lg) return perr(cudaErrorInvalidValue);               // every GPU needs at least one row of the n1 x n2 matrix
    const bool inverse = cfg->inverse != 0;
    const size_t local = ((size_t)1 << log_n) >> lg, bytes = local * 32;
    DeviceGuard guard;
    std::vector<int> dev(G);
    std::vector<cudaStream_t> st(G);
    for (unsigned g = 0; g < G; g++) {
        if (!cfg->d_src[g] || !cfg->d_dst[g] || cfg->d_src[g] == cfg->d_dst[g]) return perr(cudaErrorInvalidValue);
        cudaError_t e = device_of(cfg->d_dst[g], &dev[g]);
        if (e != cudaSuccess) return perr(e);
        st[g] = cu(cfg->streams[g]);
    }
    cudaError_t err = G > 1 ? enable_peer_access(dev) : cudaSuccess;
    if (err != cudaSuccess) return perr(err);

    unsigned char om_rows[32], om_cols[32];
    sub_roots(cfg->omega, l1, l2, om_rows, om_cols);

    std::vector<uint8_t *> A(G, nullptr), B(G, nullptr);
    std::vector<cudaEvent_t> started(G, nullptr), exchanged(G, nullptr);
    std::vector<void *> held(G, nullptr);            // where device g's data is after the first batched transform
    auto fail = [&](cudaError_t e) { if (err == cudaSuccess) err = e; };
    // phase 1: scratch, "my output buffer may be written from now on", local steps up to the exchange
    for (unsigned g = 0; g < G && err == cudaSuccess; g++) {
        cudaError_t e = cudaSetDevice(dev[g]);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&A[g], bytes, st[g]);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&B[g], bytes, st[g]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&started[g], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&exchanged[g], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(started[g], st[g]);
        unsigned in_dst = 0;
        if (e == cudaSuccess && !inverse) {
            void *one[1] = {A[g]};                                                                       // 1. At[i2l][i1]
            e = pb::ntt_exchange(pb::NTT_BN254_FR, cfg->d_src[g], l1, l2 - lg, 0, nullptr, 0, false, 1, one, (size_t)1 << l1, 0, st[g]);
            if (e == cudaSuccess) e = pb::ntt_run(pb::NTT_BN254_FR, A[g], B[g], l1, om_rows, false, st[g], &in_dst, 1u << (l2 - lg));   // 2. Yt[i2l][j1]
        } else if (e == cudaSuccess) {
            e = cudaMemcpyAsync(A[g], cfg->d_src[g], bytes, cudaMemcpyDeviceToDevice, st[g]);                // the transforms ping-pong: keep the caller's input intact
            if (e == cudaSuccess) e = pb::ntt_run(pb::NTT_BN254_FR, A[g], B[g], l2, om_cols, true, st[g], &in_dst, 1u << (l1 - lg));    // Z[j1l][i2]
        }
        held[g] = in_dst ? B[g] : A[g];
        if (e != cudaSuccess) fail(e);
    }
    // phase 2: the exchange -- twiddle, transpose and all-to-all in one kernel, storing into every GPU's d_dst over NVLink
    const unsigned x_rows_log = inverse ? l1 - lg : l2 - lg, x_cols_log = inverse ? l2 : l1;
    for (unsigned g = 0; g < G && err == cudaSuccess; g++) {
        cudaError_t e = cudaSetDevice(dev[g]);
        for (unsigned h = 0; h < G && e == cudaSuccess; h++) if (h != g) e = cudaStreamWaitEvent(st[g], started[h], 0);
        const size_t rows = (size_t)1 << x_rows_log;
        if (e == cudaSuccess)
            e = pb::ntt_exchange(pb::NTT_BN254_FR, held[g], x_rows_log, x_cols_log, (unsigned)(g * rows), cfg->omega, log_n, inverse, G, cfg->d_dst, rows * G,
                                 g * rows, st[g]);
        if (e == cudaSuccess) e = cudaEventRecord(exchanged[g], st[g]);
        if (e != cudaSuccess) fail(e);
    }
    // phase 3: every GPU waits for all stores into its buffer, then finishes locally; the result ends in d_dst
    for (unsigned g = 0; g < G && err == cudaSuccess; g++) {
        cudaError_t e = cudaSetDevice(dev[g]);
        for (unsigned h = 0; h < G && e == cudaSuccess; h++) if (h != g) e = cudaStreamWaitEvent(st[g], exchanged[h], 0);
        unsigned in_dst = 0;
        if (e == cudaSuccess && !inverse) {
            e = pb::ntt_run(pb::NTT_BN254_FR, cfg->d_dst[g], A[g], l2, om_cols, false, st[g], &in_dst, 1u << (l1 - lg));               // 4. X[j1l][j2]
            if (e == cudaSuccess && in_dst) e = cudaMemcpyAsync(cfg->d_dst[g], A[g], bytes, cudaMemcpyDeviceToDevice, st[g]);
        } else if (e == cudaSuccess) {
            e = pb::ntt_run(pb::NTT_BN254_FR, cfg->d_dst[g], A[g], l1, om_rows, true, st[g], &in_dst, 1u << (l2 - lg));                // At[i2l][i1]
            void *src = in_dst ? (void *)A[g] : cfg->d_dst[g];
            void *tgt = in_dst ? cfg->d_dst[g] : (void *)B[g];
            void *one[1] = {tgt};
            if (e == cudaSuccess) e = pb::ntt_exchange(pb::NTT_BN254_FR, src, l2 - lg, l1, 0, nullptr, 0, false, 1, one, (size_t)1 << (l2 - lg), 0, st[g]);   // A[i1][i2l]
            if (e == cudaSuccess && !in_dst) e = cudaMemcpyAsync(cfg->d_dst[g], B[g], bytes, cudaMemcpyDeviceToDevice, st[g]);
        }
        if (e != cudaSuccess) fail(e);
    }
    for (unsigned g = 0; g < G; g++) {
        cudaSetDevice(dev[g]);
        if (err != cudaSuccess && exchanged[g]) {
            // error path: peers may still be storing into / reading from buffers of this call; drain before the scratch is released
            for (unsigned h = 0; h < G; h++) cudaStreamSynchronize(st[h]);
        }
        if (A[g]) cudaFreeAsync(A[g], st[g]);
        if (B[g]) cudaFreeAsync(B[g], st[g]);
        if (started[g]) cudaEventDestroy(started[g]);
        if (exchanged[g]) cudaEventDestroy(exchanged[g]);
    }
    if (err != cudaSuccess) fprintf(stderr, "[panda-b200] multi-GPU NTT failed: %s\n", cudaGetErrorString(err));
    return perr(err);
}

}  // extern "C"
