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
// batch: number of independent 2^log_n-point transforms stored back to back in d_src (d_dst has the same size).
cudaError_t ntt_run(NttField field, void *d_src, void *d_dst, unsigned log_n, const void *omega_host, bool inverse,
                    cudaStream_t stream, unsigned *result_in_dst, unsigned batch = 1, float *pass_ms = nullptr);   // pass_ms: HOST float[4], per-pass device time (diagnostics; synchronises)

// Exchange step of the four-step (multi-GPU) transform: tiled transpose of the 2^log_rows x 2^log_cols matrix d_src fused
// with the twiddle omega^((row_offset + r) * c) (omega_host = nullptr: none; omega of order 2^log_n, inverse: omega^-1);
// column block h of `parts` equal blocks is written to dst[h][(c mod block) * ld + col_offset + r].  dst is a HOST array of
// device pointers (local staging chunks or peer-mapped buffers).  Asynchronous on `stream`.
cudaError_t ntt_exchange(NttField field, const void *d_src, unsigned log_rows, unsigned log_cols, unsigned row_offset, const void *omega_host,
                         unsigned log_n, bool inverse, unsigned parts, void *const *dst, size_t ld, size_t col_offset, cudaStream_t stream);

// d_data[i] *= g^i (inverse: g^-i) for i < 2^log_n; gen_host: HOST pointer to the coset generator g (Montgomery).
cudaError_t ntt_coset_scale(NttField field, void *d_data, unsigned log_n, const void *gen_host, bool inverse, cudaStream_t stream);

// d_dst[bitrev(i)] = d_src[i] (out of place): turns natural-order data into bit-reversed order and back
cudaError_t ntt_bit_reverse(NttField field, const void *d_src, void *d_dst, unsigned log_n, cudaStream_t stream);

// omega^(2^k) on the HOST (32 bytes, Montgomery in / out): sub-roots of the multi-GPU four-step transform
void ntt_pow2k_host(const void *omega_host, unsigned k, void *out_host);

// frees the cached twiddle tables of the current device
cudaError_t ntt_release_tables();

}  // namespace pb
