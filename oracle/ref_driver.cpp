// TEST INFRASTRUCTURE ONLY -- driver for the UNMODIFIED reference host (CPU debug) MSM path.
//
// Built by oracle/Makefile from the reference sources where they lie under /root/reference
// (panda_interface.cu, common/common.cu, unit/ntt/fft.cu) into oracle/_ref/ref_host_msm.
// Nothing under panda_b200/ links or executes this; only tests/ and bench.py's cpu_baseline /
// --impl reference legs run it.
//
// The reference entry point is panda_msm_execute_bn254_host (src/cuda/core/panda_interface.cu:162-165
// -> unit/msm/msm_host.cuh:372-383 -> :267-370).  It only needs cudaMallocHost/cudaFreeHost from the
// CUDA runtime (msm_host.cuh:124,128,305,308,357-365) and relies on that memory being zero (its bucket
// array is never initialised, msm_host.cuh:114-132), so both are interposed here with calloc/free.
// That also lets it run on a box without a GPU.
//
// usage: ref_host_msm <bases.bin> <scalars.bin> <log_n> <out96.bin> [coord=0]
//   bases  : 2^log_n * 64 B (x||y, Fq Montgomery, LE)      scalars: 2^log_n * 32 B (Fr Montgomery, LE)
//   out    : 96 B Jacobian (x||y||z, Montgomery) exactly as the reference memcpy's it (msm_host.cuh:352)
// prints "ref_time_ms <t>" (wall time of the reference call only).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <cuda_runtime_api.h>

extern "C" cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
extern "C" cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }

extern "C" {
struct ref_handle { void *handle; };
struct ref_msm_configuration {          // panda_interface.cuh:70-79 (48 bytes, passed by value)
    ref_handle mem_pool;
    ref_handle stream;
    void *bases;
    void *scalars;
    void *results;
    unsigned log_scalars_count;
    int msm_result_coordinate_type;
};
unsigned panda_msm_execute_bn254_host(const ref_msm_configuration cfg);
}

static bool slurp(const char *path, std::vector<unsigned char> &buf, size_t want) {
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); return false; }
    buf.resize(want);
    size_t got = fread(buf.data(), 1, want, f);
    fclose(f);
    if (got != want) { fprintf(stderr, "%s: short read %zu of %zu\n", path, got, want); return false; }
    return true;
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s bases.bin scalars.bin log_n out.bin [coord]\n", argv[0]); return 2; }
    unsigned log_n = (unsigned)atoi(argv[3]);
    int coord = argc > 5 ? atoi(argv[5]) : 0;
    size_t n = (size_t)1 << log_n;
    std::vector<unsigned char> bases, scalars;
    if (!slurp(argv[1], bases, n * 64) || !slurp(argv[2], scalars, n * 32)) return 1;
    unsigned char out[96];
    memset(out, 0, sizeof out);
    ref_msm_configuration cfg{};
    cfg.bases = bases.data();
    cfg.scalars = scalars.data();          // NB: the reference converts these in place (msm_host.cuh:293-296)
    cfg.results = out;
    cfg.log_scalars_count = log_n;
    cfg.msm_result_coordinate_type = coord;
    auto t0 = std::chrono::steady_clock::now();
    unsigned rc = panda_msm_execute_bn254_host(cfg);
    auto t1 = std::chrono::steady_clock::now();
    if (rc) { fprintf(stderr, "reference returned %u\n", rc); return 1; }
    FILE *f = fopen(argv[4], "wb");
    if (!f) { perror(argv[4]); return 1; }
    fwrite(out, 1, 96, f);
    fclose(f);
    printf("ref_time_ms %.3f\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
    return 0;
}
