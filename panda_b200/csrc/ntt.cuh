// ntt.cuh -- radix-2^r multi-pass NTT over a 256-bit prime field for sm_100a (internal interface).
//
// Replaces the reference's src/cuda/core/unit/ntt/fft.cu:62-73 (set_up), :171-216 (execute) and the
// radix_fft kernel whose body is compiled out there (:107-169).  Semantics = the disabled text's (bellperson
// radix_fft): forward DFT y[j] = sum_i x[i] * omega^(i*j), natural order in and out, no scaling, ping-pong
// between d_src and d_dst once per pass with ceil(log_n / 8) passes, so *flag = passes & 1 exactly as in
// fft.cu:193-211.  See DESIGN.md for the pass structure and rooflines.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pb {

enum NttField { NTT_BN254_FR = 0 };

// omega_host: primitive 2^log_n-th root of unity, Montgomery form, HOST pointer (32 bytes).
// inverse: transform with omega^-1 and scale by 1/n.
// *result_in_dst: 1 if the output is in d_dst, 0 if in d_src (both buffers are used as ping-pong storage).
// Asynchronous on `stream` (twiddle tables are built on first use of an (omega, log_n) pair and cached per device).
cudaError_t ntt_run(NttField field, void *d_src, void *d_dst, unsigned log_n, const void *omega_host, bool inverse,
                    cudaStream_t stream, unsigned *result_in_dst);

// frees the cached twiddle tables of every device
cudaError_t ntt_release_tables();

}  // namespace pb
