// debug.cu -- diagnostic kernels behind include/panda_debug.h: single-operation field / curve checks for the parity
// tests and integer-pipe microbenchmarks for the roofline denominators.
#include "panda_debug.h"
#include "ec.cuh"

#include <cuda_runtime.h>

using namespace pb;

namespace {

template <class F>
__global__ void k_field_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        F x = F::load(a + i * F::N), y = F::zero(), r;
        if (op == PANDA_FOP_MUL || op == PANDA_FOP_ADD || op == PANDA_FOP_SUB) y = F::load(b + i * F::N);
        switch (op) {
            case PANDA_FOP_MUL: r = x * y; break;
            case PANDA_FOP_ADD: r = x + y; break;
            case PANDA_FOP_SUB: r = x - y; break;
            case PANDA_FOP_SQR: r = x.sqr(); break;
            case PANDA_FOP_FROM_MONT: r = x.from_mont(); break;
            case PANDA_FOP_TO_MONT: r = x.to_mont(); break;
            case PANDA_FOP_INV: r = fe_inverse(x); break;
            default: r = x.neg(); break;
        }
        r.canon().store(out + i * F::N);
    }
}

template <class F>
__global__ void k_curve_op(int op, const uint32_t *p, const uint32_t *q, uint32_t *out, size_t count) {
    using J = Jacobian<F>;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        J a;
        a.x = F::load(p + i * 3 * F::N); a.y = F::load(p + i * 3 * F::N + F::N); a.z = F::load(p + i * 3 * F::N + 2 * F::N);
        J r;
        if (op == PANDA_COP_MADD) {
            Affine<F> b = Affine<F>::load(q + i * 2 * F::N);
            Xyzz<F> acc = a.to_xyzz();
            if (!b.is_identity()) acc.madd(b.x, b.y);
            r = J::from_xyzz(acc);
        } else if (op == PANDA_COP_ADD) {
            J b;
            b.x = F::load(q + i * 3 * F::N); b.y = F::load(q + i * 3 * F::N + F::N); b.z = F::load(q + i * 3 * F::N + 2 * F::N);
            Xyzz<F> acc = a.to_xyzz();
            acc.add(b.to_xyzz());
            r = J::from_xyzz(acc);
        } else if (op == PANDA_COP_DBL_XYZZ) {
            r = J::from_xyzz(a.to_xyzz().dbl());
        } else if (op == PANDA_COP_DBL_JAC) {
            r = a.dbl();
        } else {
            r = a.to_homogeneous();
        }
        r.store_canonical(out + i * 3 * F::N);
    }
}

__global__ void k_imad_peak(unsigned iters, uint32_t seed, uint32_t *sink) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    const uint32_t m = seed | 1;
#pragma unroll 1
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a0 = a0 * m + a1; a1 = a1 * m + a2; a2 = a2 * m + a3; a3 = a3 * m + a0;
            a4 = a4 * m + a5; a5 = a5 * m + a6; a6 = a6 * m + a7; a7 = a7 * m + a4;
        }
    }
    if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x12345678u) sink[0] = a0;
}

__global__ void k_imad_wide_peak(unsigned iters, uint32_t seed, unsigned long long *sink) {
    // 16 independent accumulators per thread, each step one 32 x 32 + 64 -> 64 multiply-add (IMAD.WIDE.U32)
    unsigned long long a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = (unsigned long long)(seed + threadIdx.x) * (2 * k + 3);
    const uint32_t m = seed | 1;
#pragma unroll 1
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = (unsigned long long)(uint32_t)a[k] * m + a[k];
        }
    }
    unsigned long long x = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= a[k];
    if (x == 0x12345678ull) sink[0] = x;
}

__global__ void __launch_bounds__(256) k_modmul_peak(unsigned iters, uint32_t seed, uint32_t *sink) {
    using F = FqBn254;
    F a = F::one(), b = F::r2(), c = F::one(), d = F::r2();
    a.l[0] += threadIdx.x + seed; b.l[1] ^= threadIdx.x; c.l[2] += seed; d.l[3] ^= seed + threadIdx.x;
#pragma unroll 1
    for (unsigned i = 0; i < iters; i++) {
        a = a * b; b = b * c; c = c * d; d = d * a;
    }
    F s = a + b + c + d;
    if (s.l[0] == 0x12345678u && s.l[7] == 0x9abcdef0u) s.store(sink);
}

template <class F>
panda_error launch_field(int op, const void *a, const void *b, void *out, size_t count, cudaStream_t s) {
    if (!count) return panda_success;
    unsigned blocks = (unsigned)((count + 127) / 128);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_field_op<F><<<blocks, 128, 0, s>>>(op, (const uint32_t *)a, (const uint32_t *)b, (uint32_t *)out, count);
    return (panda_error)cudaGetLastError();
}
template <class F>
panda_error launch_curve(int op, const void *p, const void *q, void *out, size_t count, cudaStream_t s) {
    if (!count) return panda_success;
    unsigned blocks = (unsigned)((count + 127) / 128);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_curve_op<F><<<blocks, 128, 0, s>>>(op, (const uint32_t *)p, (const uint32_t *)q, (uint32_t *)out, count);
    return (panda_error)cudaGetLastError();
}

}  // namespace

extern "C" {

panda_error panda_debug_field_op(int field_id, int op, const void *a, const void *b, void *out, size_t count, panda_stream stream) {
    cudaStream_t s = (cudaStream_t)stream.handle;
    switch (field_id) {
        case 0: return launch_field<FqBn254>(op, a, b, out, count, s);
        case 1: return launch_field<FrBn254>(op, a, b, out, count, s);
        case 2: return launch_field<FqBls377>(op, a, b, out, count, s);
        case 3: return launch_field<FrBls377>(op, a, b, out, count, s);
        case 4: return launch_field<Fe<Bls381Fq>>(op, a, b, out, count, s);
        case 5: return launch_field<Fe<Bls381Fr>>(op, a, b, out, count, s);      // canonical operands only: see Bls381Fr in field.cuh
    }
    return (panda_error)cudaErrorInvalidValue;
}

panda_error panda_debug_curve_op(int curve_id, int op, const void *p, const void *q, void *out, size_t count, panda_stream stream) {
    cudaStream_t s = (cudaStream_t)stream.handle;
    if (curve_id == 0) return launch_curve<FqBn254>(op, p, q, out, count, s);
    if (curve_id == 1) return launch_curve<FqBls377>(op, p, q, out, count, s);
    if (curve_id == 2) return launch_curve<Fe<Bls381Fq>>(op, p, q, out, count, s);
    return (panda_error)cudaErrorInvalidValue;
}

panda_error panda_debug_int_peak(int kind, unsigned iters, float *ms, unsigned long long *ops) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (panda_error)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    void *sink = nullptr;
    if ((e = cudaMalloc(&sink, 256)) != cudaSuccess) return (panda_error)e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    const unsigned blocks = (unsigned)sms * 8, threads = 256;
    for (int rep = 0; rep < 2; rep++) {     // first launch warms up, second is timed
        cudaEventRecord(t0);
        if (kind == 0) k_imad_peak<<<blocks, threads>>>(iters, 12345u, (uint32_t *)sink);
        else if (kind == 1) k_imad_wide_peak<<<blocks, threads>>>(iters, 12345u, (unsigned long long *)sink);
        else k_modmul_peak<<<blocks, threads>>>(iters, 12345u, (uint32_t *)sink);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
    }
    if (e == cudaSuccess) {
        cudaEventElapsedTime(ms, t0, t1);
        const unsigned long long per_thread = kind == 2 ? 4ull * iters : 64ull * iters;
        *ops = per_thread * blocks * threads;
        e = cudaGetLastError();
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(sink);
    return (panda_error)e;
}

}  // extern "C"
