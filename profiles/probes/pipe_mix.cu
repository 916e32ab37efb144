// pipe_mix.cu -- which B200 issue resources does DFMA share?  In-thread mixes of NF DFMA chains with NX chains of another class.
// KIND 0: IMAD.WIDE (32x32+64), 1: IMAD (32-bit), 2: IADD3 (32-bit three-input add), 3: FFMA, 4: LOP3, 5: 64-bit add pair (IADD3 + IADD3.X), 6: DADD
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int NF, int NX, int KIND>
__global__ void __launch_bounds__(256) k_mix(unsigned iters, uint32_t seed, unsigned long long *sink) {
    unsigned long long a[8]; uint32_t u[8]; float g[8]; double f[8], h[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        a[k] = (unsigned long long)(seed + threadIdx.x) * (2 * k + 3); u[k] = seed * (k + 7) + threadIdx.x;
        g[k] = seed * 0.001f + k + threadIdx.x; f[k] = seed + threadIdx.x * 0.001 + k; h[k] = f[k] * 3;
    }
    const uint32_t m = seed | 1, m2 = seed * 77 + 5;
    const double fm = seed * 0.5, fc = seed * 0.25;
    const float gm = seed * 0.5f, gc = seed * 0.25f;
#pragma unroll 1
    for (unsigned i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k < NX) {
                    if (KIND == 0) a[k] = (unsigned long long)(uint32_t)a[k] * m + a[k];
                    else if (KIND == 1) u[k] = u[k] * m + m2;
                    else if (KIND == 2) u[k] = u[k] + u[(k + 1) & 7] + m2;
                    else if (KIND == 3) g[k] = fmaf(g[k], gm, gc);
                    else if (KIND == 4) u[k] = (u[k] & u[(k + 1) & 7]) ^ m2;
                    else if (KIND == 5) a[k] = a[k] + a[(k + 1) & 7] + a[(k + 3) & 7];
                    else if (KIND == 6) h[k] = __dadd_rn(h[k], fc);
                }
                if (k < NF) f[k] = __fma_rz(f[k], fm, fc);
            }
        }
    }
    unsigned long long x = 0; double y = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { x ^= a[k] ^ u[k] ^ __float_as_uint(g[k]); y += f[k] + h[k]; }
    if (x == 0x12345678ull || y == 0.12345) sink[0] = x;
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
static unsigned blocks;
template <int NF, int NX, int KIND>
static void run(const char *name) {
    const unsigned it = 2000;
    float ms = time_launch([&] { k_mix<NF, NX, KIND><<<blocks, 256>>>(it, 12345u, (unsigned long long *)sink); });
    // warp-instructions per SMSP: blocks*8 warps / (148*4) SMSPs
    const double warps_per_smsp = blocks * 8.0 / (148 * 4);
    const double clk = ms * 1e-3 * 1.965e9;
    const double steps = (double)it * 8 * warps_per_smsp;      // inner steps per SMSP
    printf("%-10s NF=%d NX=%d : %8.3f ms  %6.2f clk per inner step (per SMSP, at 1965 MHz)\n", name, NF, NX, ms, clk / steps);
}
template <int KIND>
static void sweep(const char *name) {
    run<0, 8, KIND>(name); run<0, 4, KIND>(name); run<8, 8, KIND>(name); run<8, 4, KIND>(name); run<4, 8, KIND>(name); run<4, 4, KIND>(name);
}
int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaMalloc(&sink, 4096));
    blocks = sms * 8;
    run<8, 0, 0>("DFMA"); run<4, 0, 0>("DFMA");
    sweep<0>("IMAD.WIDE"); sweep<1>("IMAD"); sweep<2>("IADD3"); sweep<3>("FFMA"); sweep<4>("LOP3"); sweep<5>("ADD64x3"); sweep<6>("DADD");
    return 0;
}
