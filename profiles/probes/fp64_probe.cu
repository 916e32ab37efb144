// fp64_probe.cu -- round-2 feasibility probe: can the B200's FP64 pipe (64 DFMA/clk/SM) carry part of the
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

// role split: warp w of a CTA runs the FP64 product when ((w / 4) % period) < fp_of, else the integer product
__global__ void __launch_bounds__(256) k_modmul_hybrid(unsigned iters_fp, unsigned iters_int, int period, int fp_of, uint32_t seed, uint64_t *sink) {
    const int w = threadIdx.x >> 5;
    if (((w >> 2) % period) < fp_of) {
        F52 a, b, c, d;
        for (int k = 0; k < 5; k++) { a.l[k] = (Bn254Fq52::mod(k) >> 1) + threadIdx.x + seed; b.l[k] = (Bn254Fq52::mod(k) >> 2) ^ threadIdx.x; c.l[k] = a.l[k] ^ 0x55; d.l[k] = b.l[k] + 77; }
#pragma unroll 1
        for (unsigned i = 0; i < iters_fp; i++) {
            a = mul52<Bn254Fq52>(a, b); b = mul52<Bn254Fq52>(b, c); c = mul52<Bn254Fq52>(c, d); d = mul52<Bn254Fq52>(d, a);
        }
        uint64_t s = a.l[0] ^ b.l[1] ^ c.l[2] ^ d.l[3];
        if (s == 0x12345678ull) sink[0] = s;
    } else {
        using F = FqBn254;
        F a = F::one(), b = F::r2(), c = F::one(), d = F::r2();
        a.l[0] += threadIdx.x + seed; b.l[1] ^= threadIdx.x; c.l[2] += seed; d.l[3] ^= seed + threadIdx.x;
#pragma unroll 1
        for (unsigned i = 0; i < iters_int; i++) {
            a = a * b; b = b * c; c = c * d; d = d * a;
        }
        F s = a + b + c + d;
        if (s.l[0] == 0x12345678u && s.l[7] == 0x9abcdef0u) sink[0] = s.l[0];
    }
}

template <class K>
static float time_launch(K launch) {
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(t0));
        launch();
        CK(cudaEventRecord(t1));
        CK(cudaEventSynchronize(t1));
        CK(cudaGetLastError());
    }
    CK(cudaEventElapsedTime(&ms, t0, t1));
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    return ms;
}

int main(int argc, char **argv) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SMs %d, clock attr %d kHz\n", sms, clk);
    void *sink;
    CK(cudaMalloc(&sink, 4096));
    const unsigned blocks = sms * 8, threads = 256;
    const double T = (double)blocks * threads;
    const unsigned it = 4000;
    float ms;
    ms = time_launch([&] { k_dfma<<<blocks, threads>>>(it, 1.0000001, (double *)sink); });
    printf("DFMA alone         : %8.3f ms  %7.2f T DFMA/s\n", ms, T * it * 64 / ms / 1e9);
    ms = time_launch([&] { k_wide<<<blocks, threads>>>(it, 12345u, (unsigned long long *)sink); });
    printf("IMAD.WIDE alone    : %8.3f ms  %7.2f T/s\n", ms, T * it * 64 / ms / 1e9);
    ms = time_launch([&] { k_iadd3<<<blocks, threads>>>(it, 12345u, (unsigned long long *)sink); });
    printf("64-bit 3-input add : %8.3f ms  %7.2f T adds/s\n", ms, T * it * 64 / ms / 1e9);
    ms = time_launch([&] { k_mix<8, 8><<<blocks, threads>>>(it, 12345u, (unsigned long long *)sink); });
    printf("mix 8 DFMA + 8 WIDE: %8.3f ms  %7.2f T DFMA/s + %7.2f T WIDE/s\n", ms, T * it * 64 / ms / 1e9, T * it * 64 / ms / 1e9);
    ms = time_launch([&] { k_mix<8, 4><<<blocks, threads>>>(it, 12345u, (unsigned long long *)sink); });
    printf("mix 8 DFMA + 4 WIDE: %8.3f ms  %7.2f T DFMA/s + %7.2f T WIDE/s\n", ms, T * it * 64 / ms / 1e9, T * it * 32 / ms / 1e9);
    ms = time_launch([&] { k_mix<8, 0><<<blocks, threads>>>(it, 12345u, (unsigned long long *)sink); });
    printf("mix 8 DFMA + 0 WIDE: %8.3f ms  %7.2f T DFMA/s\n", ms, T * it * 64 / ms / 1e9);
    ms = time_launch([&] { k_mix<0, 8><<<blocks, threads>>>(it, 12345u, (unsigned long long *)sink); });
    printf("mix 0 DFMA + 8 WIDE: %8.3f ms  %7.2f T WIDE/s\n", ms, T * it * 64 / ms / 1e9);

    // correctness data for the host-side check (python reads the binary dump)
    if (argc > 1) {
        const int n = 4096;
        uint64_t *ha = (uint64_t *)malloc(n * 40), *hb = (uint64_t *)malloc(n * 40), *ho = (uint64_t *)malloc(n * 40);
        uint64_t s = 88172645463325252ull;
        auto rnd = [&] { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
        for (int i = 0; i < n * 5; i++) { ha[i] = rnd() & M52; hb[i] = rnd() & M52; }
        for (int i = 0; i < n; i++) { ha[i * 5 + 4] &= (1ull << 48) - 1; hb[i * 5 + 4] &= (1ull << 48) - 1; }   // values < 2^256
        for (int k = 0; k < 5; k++) { ha[k] = M52; hb[k] = M52; ha[5 + k] = 0; hb[10 + k] = k == 0; }
        ha[4] = hb[4] = (1ull << 49) - 1;
        uint64_t *da, *db, *dout;
        CK(cudaMalloc(&da, n * 40)); CK(cudaMalloc(&db, n * 40)); CK(cudaMalloc(&dout, n * 40));
        CK(cudaMemcpy(da, ha, n * 40, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb, n * 40, cudaMemcpyHostToDevice));
        k_mul52_check<<<(n + 255) / 256, 256>>>(da, db, dout, n);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ho, dout, n * 40, cudaMemcpyDeviceToHost));
        FILE *f = fopen(argv[1], "wb");
        fwrite(ha, 40, n, f); fwrite(hb, 40, n, f); fwrite(ho, 40, n, f);
        fclose(f);
        printf("wrote %s\n", argv[1]);
    }

    const unsigned mi = 2000;
    ms = time_launch([&] { k_modmul_hybrid<<<blocks, threads>>>(mi, mi, 1, 0, 1u, (uint64_t *)sink); });
    const double int_rate = T * mi * 4 / ms / 1e6;
    printf("modmul int only    : %8.3f ms  %7.2f G modmul/s\n", ms, int_rate);
    ms = time_launch([&] { k_modmul_hybrid<<<blocks, threads>>>(mi, mi, 1, 1, 1u, (uint64_t *)sink); });
    const double fp_rate = T * mi * 4 / ms / 1e6;
    printf("modmul fp64 only   : %8.3f ms  %7.2f G modmul/s\n", ms, fp_rate);
    // half the warps each; sweep the iteration ratio to find the balance
    for (int pct = 40; pct <= 160; pct += 20) {
        const unsigned ifp = mi * pct / 100;
        ms = time_launch([&] { k_modmul_hybrid<<<blocks, threads>>>(ifp, mi, 2, 1, 1u, (uint64_t *)sink); });
        printf("hybrid 1:1 warps, fp iters %3d%% : %8.3f ms  %7.2f G modmul/s\n", pct, ms, (T / 2) * (ifp + mi) * 4 / ms / 1e6);
    }
    return 0;
}
