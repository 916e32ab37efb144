// fp64_probe2.cu (same field code as fp64_probe.cu) -- round-2 feasibility probe: can the B200's FP64 pipe (64 DFMA/clk/SM) carry part of the
// Montgomery products that today all sit on the half-rate IMAD.WIDE path?  Measures the raw pipe rates, their
// overlap, and a 5 x 52-bit-limb Montgomery product built from DFMA hi/lo splits, alone and next to integer warps.
// Build (from this directory):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -I../../panda_b200/csrc -I../../include fp64_probe.cu -o _bin/fp64_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "field.cuh"

using namespace pb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void k_dfma(unsigned iters, double seed, double *sink) {
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = seed + threadIdx.x * 0.001 + k;
    const double m = seed * 0.5, c = seed * 0.25;
#pragma unroll 1
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = __fma_rz(a[k], m, c);
        }
    }
    double x = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) x += a[k];
    if (x == 0.12345) sink[0] = x;
}

__global__ void k_wide(unsigned iters, uint32_t seed, unsigned long long *sink) {
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

// NF dfma + NW imad.wide per inner step, same thread
template <int NF, int NW>
__global__ void k_mix(unsigned iters, uint32_t seed, unsigned long long *sink) {
    unsigned long long a[8];
    double f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = (unsigned long long)(seed + threadIdx.x) * (2 * k + 3); f[k] = seed + threadIdx.x * 0.001 + k; }
    const uint32_t m = seed | 1;
    const double fm = seed * 0.5, fc = seed * 0.25;
#pragma unroll 1
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k < NW) a[k] = (unsigned long long)(uint32_t)a[k] * m + a[k];
                if (k < NF) f[k] = __fma_rz(f[k], fm, fc);
            }
        }
    }
    unsigned long long x = 0; double y = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { x ^= a[k]; y += f[k]; }
    if (x == 0x12345678ull || y == 0.12345) sink[0] = x;
}

// 64-bit three-input adds (IADD3 + IADD3.X)
__global__ void k_iadd3(unsigned iters, uint32_t seed, unsigned long long *sink) {
    unsigned long long a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = (unsigned long long)(seed + threadIdx.x) * (2 * k + 3);
#pragma unroll 1
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = a[k] + a[(k + 1) & 15] + a[(k + 5) & 15];
        }
    }
    unsigned long long x = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= a[k];
    if (x == 0x12345678ull) sink[0] = x;
}

// ------------------------------------------------------------------------------------------------
// 5 x 52-bit limb Montgomery product on the FP64 pipe, R' = 2^260, BN254 Fq.
struct Bn254Fq52 {
    static constexpr uint64_t PINV = 0x20782e4866389ull;      // -p^-1 mod 2^52
    __host__ __device__ static constexpr uint64_t mod(int i) {
        constexpr uint64_t v[5] = {0x8c16d87cfd47ull, 0x916871ca8d3c2ull, 0x181585d97816aull, 0xa029b85045b68ull, 0x30644e72e131ull};
        return v[i];
    }
};

static constexpr uint64_t M52 = (1ull << 52) - 1;
static constexpr uint64_t E52 = 0x4330000000000000ull;       // bits of 2^52
static constexpr uint64_t E104 = 0x4670000000000000ull;      // bits of 2^104

struct F52 { uint64_t l[5]; };

__device__ __forceinline__ double u2d(uint64_t x) { return __longlong_as_double((long long)(x | E52)) - 4503599627370496.0; }

// hi / lo halves of a * b (a, b integers < 2^52 held in doubles) as raw bit patterns: hi = E104 | floor(ab / 2^52), lo = E52 | (ab mod 2^52)
__device__ __forceinline__ void prod(double a, double b, uint64_t &hi, uint64_t &lo) {
    const double c1 = __longlong_as_double((long long)E104);
    const double c2 = __longlong_as_double((long long)(E104 + 1));   // 2^104 + 2^52
    const double h = __fma_rz(a, b, c1);
    const double s = c2 - h;
    const double l = __fma_rz(a, b, s);
    hi = (uint64_t)__double_as_longlong(h);
    lo = (uint64_t)__double_as_longlong(l);
}

template <class P>
__device__ __forceinline__ F52 mul52(const F52 &a, const F52 &b) {
    double ad[5], bd[5];
#pragma unroll
    for (int i = 0; i < 5; i++) { ad[i] = u2d(a.l[i]); bd[i] = u2d(b.l[i]); }
    // column k receives the low halves of the products with i + j == k and the high halves of those with i + j == k - 1,
    // from the plain product (25) and from the reduction (25).  Start every column at minus the exponent patterns it will collect.
    uint64_t col[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        int nlo = 0, nhi = 0;
        for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) { if (i + j == k) nlo += 2; if (i + j + 1 == k) nhi += 2; }
        col[k] = 0ull - ((uint64_t)nlo * E52 + (uint64_t)nhi * E104);
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
#pragma unroll
        for (int j = 0; j < 5; j++) {
            uint64_t hi, lo;
            prod(ad[i], bd[j], hi, lo);
            col[i + j] += lo;
            col[i + j + 1] += hi;
        }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint64_t m = ((col[k] & M52) * P::PINV) & M52;
        const double md = u2d(m);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            uint64_t hi, lo;
            prod(md, (double)P::mod(j), hi, lo);
            col[k + j] += lo;
            col[k + j + 1] += hi;
        }
        col[k + 1] += col[k] >> 52;
    }
    F52 r;
#pragma unroll
    for (int i = 0; i < 4; i++) { r.l[i] = col[5 + i] & M52; col[6 + i] += col[5 + i] >> 52; }
    r.l[4] = col[9];
    return r;
}

__global__ void __launch_bounds__(256) k_mul52_check(const uint64_t *a, const uint64_t *b, uint64_t *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F52 x, y;
    for (int k = 0; k < 5; k++) { x.l[k] = a[i * 5 + k]; y.l[k] = b[i * 5 + k]; }
    F52 r = mul52<Bn254Fq52>(x, y);
    for (int k = 0; k < 5; k++) out[i * 5 + k] = r.l[k];
}


// FP64 product with ILP independent chains per thread
template <int ILP, int MINB>
__global__ void __launch_bounds__(128, MINB) k_fp(unsigned iters, uint32_t seed, uint64_t *sink) {
    F52 x[ILP], y[ILP];
#pragma unroll
    for (int c = 0; c < ILP; c++)
        for (int k = 0; k < 5; k++) { x[c].l[k] = (Bn254Fq52::mod(k) >> 1) + threadIdx.x + seed + c; y[c].l[k] = (Bn254Fq52::mod(k) >> 2) ^ (threadIdx.x + 3 * c); }
#pragma unroll 1
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < ILP; c++) x[c] = mul52<Bn254Fq52>(x[c], y[c]);
#pragma unroll
        for (int c = 0; c < ILP; c++) y[c] = mul52<Bn254Fq52>(y[c], x[c]);
    }
    uint64_t s = 0;
#pragma unroll
    for (int c = 0; c < ILP; c++) s ^= x[c].l[0] ^ y[c].l[1];
    if (s == 0x12345678ull) sink[0] = s;
}

template <int ILP, int MINB>
__global__ void __launch_bounds__(128, MINB) k_int(unsigned iters, uint32_t seed, uint64_t *sink) {
    using F = FqBn254;
    extern __shared__ uint32_t pad[];
    F x[ILP], y[ILP];
#pragma unroll
    for (int c = 0; c < ILP; c++) { x[c] = F::one(); y[c] = F::r2(); x[c].l[0] += threadIdx.x + seed + c; y[c].l[1] ^= threadIdx.x + c; }
#pragma unroll 1
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < ILP; c++) x[c] = x[c] * y[c];
#pragma unroll
        for (int c = 0; c < ILP; c++) y[c] = y[c] * x[c];
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < ILP; c++) s ^= x[c].l[0] ^ y[c].l[7];
    if (s == 0x12345678u) { sink[0] = s; pad[0] = s; }
}

template <class K>
static float time_launch(K launch) {
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(t0)); launch(); CK(cudaEventRecord(t1)); CK(cudaEventSynchronize(t1)); CK(cudaGetLastError());
    }
    CK(cudaEventElapsedTime(&ms, t0, t1));
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    return ms;
}

static void *sink;
static int sms;

template <int ILP, int MINB>
static void run_fp(unsigned ctas_per_sm) {
    const unsigned it = 1500;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fp<ILP, MINB>, 128, 0));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_fp<ILP, MINB>));
    float ms = time_launch([&] { k_fp<ILP, MINB><<<sms * ctas_per_sm, 128>>>(it, 1u, (uint64_t *)sink); });
    printf("fp64 ILP=%d regs=%3d occ=%d CTAs/SM, %u CTAs/SM launched: %8.3f ms %7.2f G modmul/s\n", ILP, fa.numRegs, occ, ctas_per_sm, ms,
           (double)sms * ctas_per_sm * 128 * it * 2 * ILP / ms / 1e6);
}
template <int ILP, int MINB>
static void run_int(unsigned ctas_per_sm) {
    const unsigned it = 1500;
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_int<ILP, MINB>));
    float ms = time_launch([&] { k_int<ILP, MINB><<<sms * ctas_per_sm, 128>>>(it, 1u, (uint64_t *)sink); });
    printf("int  ILP=%d regs=%3d, %u CTAs/SM launched: %8.3f ms %7.2f G modmul/s\n", ILP, fa.numRegs, ctas_per_sm, ms,
           (double)sms * ctas_per_sm * 128 * it * 2 * ILP / ms / 1e6);
}

// two kernels side by side: the integer one capped to int_ctas CTAs per SM by dynamic shared memory, the FP64 one with fp_ctas per SM
template <int ILPF, int MINBF, int ILPI, int MINBI>
static void run_pair(unsigned int_ctas, unsigned fp_ctas, unsigned it_int, unsigned it_fp) {
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t t0, t1, e2;
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1)); CK(cudaEventCreate(&e2));
    // shared memory so that exactly (int_ctas + fp_ctas) CTAs fit: each integer CTA takes 200 KB / (int_ctas + fp_ctas) ... the FP64 kernel takes none,
    // so cap the integer kernel only: smem per CTA = 220 KB / int_ctas leaves no room for one more integer CTA
    const size_t smem = (size_t)(220 * 1024 / int_ctas) & ~(size_t)1023;
    CK(cudaFuncSetAttribute(k_int<ILPI, MINBI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(t0, s1));
        CK(cudaStreamWaitEvent(s2, t0, 0));
        k_int<ILPI, MINBI><<<sms * int_ctas, 128, smem, s1>>>(it_int, 1u, (uint64_t *)sink);
        k_fp<ILPF, MINBF><<<sms * fp_ctas, 128, 0, s2>>>(it_fp, 1u, (uint64_t *)sink);
        CK(cudaEventRecord(e2, s2));
        CK(cudaStreamWaitEvent(s1, e2, 0));
        CK(cudaEventRecord(t1, s1));
        CK(cudaEventSynchronize(t1)); CK(cudaGetLastError());
    }
    CK(cudaEventElapsedTime(&ms, t0, t1));
    const double mm = (double)sms * 128 * 2 * ((double)int_ctas * it_int * ILPI + (double)fp_ctas * it_fp * ILPF);
    printf("pair int %u CTAs/SM x %u it (ILP %d) + fp64 %u CTAs/SM x %u it (ILP %d): %8.3f ms %7.2f G modmul/s\n", int_ctas, it_int, ILPI, fp_ctas, it_fp, ILPF, ms, mm / ms / 1e6);
    cudaStreamDestroy(s1); cudaStreamDestroy(s2);
}

int main() {
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaMalloc(&sink, 4096));
    run_int<2, 3>(3); run_int<2, 3>(2); run_int<2, 3>(1); run_int<1, 3>(3);
    run_fp<1, 3>(3); run_fp<1, 3>(2); run_fp<1, 3>(1);
    run_fp<2, 2>(2); run_fp<2, 2>(1);
    run_fp<3, 1>(1); run_fp<3, 2>(2);
    run_fp<4, 1>(1);
    // pairs: 2 integer CTAs + 1 FP64 CTA per SM
    for (unsigned f = 600; f <= 1800; f += 300) run_pair<2, 2, 2, 3>(2, 1, 1500, f);
    for (unsigned f = 600; f <= 1800; f += 300) run_pair<3, 1, 2, 3>(2, 1, 1500, f);
    for (unsigned f = 600; f <= 1800; f += 300) run_pair<2, 2, 2, 3>(1, 2, 1500, f);
    for (unsigned f = 600; f <= 1800; f += 300) run_pair<1, 3, 2, 3>(2, 2, 1500, f);
    return 0;
}
