// msm.cu -- Pippenger MSM for sm_100a: signed-digit windows, counting sort of point indices by bucket,
// load-balanced bucket accumulation (XYZZ mixed additions), parallel running-sum bucket reduction with
// warp-shuffle stitching, on-device window combination.  See msm.cuh / DESIGN.md.
//
// Replaces src/cuda/core/unit/msm/msm_cuda.cuh:552-769 of the reference (kernels :148-282, :373-497 and
// the host-side Horner :59-77).  Written from scratch for B200; nothing here is derived from that code.
// This file holds the plan (window width, segment length, workspace layout) and the per-curve dispatch;
// the kernels live in msm_impl.cuh.
#include "msm.cuh"

#include <algorithm>

namespace pb {

cudaError_t msm_run_bn254(const void *bases, const void *scalars, uint32_t n, void *result, CoordType coord, cudaMemPool_t pool,
                          cudaStream_t stream, uint32_t c_override, uint32_t seg_override, MsmStageTimes *timings);
cudaError_t msm_run_bls12_377(const void *bases, const void *scalars, uint32_t n, void *result, CoordType coord, cudaMemPool_t pool,
                              cudaStream_t stream, uint32_t c_override, uint32_t seg_override, MsmStageTimes *timings);
cudaError_t msm_combine_bn254(const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream);
cudaError_t msm_combine_bls12_377(const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream);

// ----------------------------------------------------------------------------------------------------
// plan

static uint32_t windows_for(uint32_t bits, uint32_t c) {
    // signed digits need the top window to stay <= 2^(c-1) after the incoming carry: its bit width t must be <= c-1
    uint32_t W = (bits + c - 1) / c;
    int t = (int)bits - (int)(W - 1) * (int)c;
    if (t > (int)c - 1) W++;
    return W;
}

static uint32_t pow2_floor(uint64_t v) { uint32_t r = 1; while ((uint64_t)r * 2 <= v) r *= 2; return r; }

MsmPlan msm_make_plan(CurveId curve, uint32_t n, uint32_t c_override, uint32_t seg_override) {
    const uint32_t bits = curve == CURVE_BLS12_377 ? 253 : 254;
    const size_t fq_bytes = curve == CURVE_BLS12_377 ? 48 : 32;
    MsmPlan p{};
    p.n = n;
    uint32_t best_c = 8;
    if (c_override >= 8 && c_override <= 16) best_c = c_override;
    else {
        double best = 1e300;
        for (uint32_t c = 8; c <= 16; c++) {
            double W = windows_for(bits, c), nb = (double)(1u << (c - 1));
            // modmul counts: mixed add 10, bucket combine + running sums ~ (2 + 1.5) full adds of 14
            double cost = (double)n * W * 10.0 + W * nb * 3.5 * 14.0;
            if (cost < best) { best = cost; best_c = c; }
        }
    }
    p.c = best_c;
    p.windows = windows_for(bits, p.c);
    p.nb = 1u << (p.c - 1);
    // segment length: about one average bucket, so that most buckets end up with one or two partial sums, but
    // never so long that the accumulation kernel has fewer than ~4 waves of threads (148 SMs x 384 threads)
    uint64_t entries = (uint64_t)n * p.windows;
    uint32_t L = pow2_floor(std::max<uint64_t>(1, n / p.nb));
    L = std::min<uint32_t>(L, pow2_floor(std::max<uint64_t>(1, entries / (148ull * 384 * 4))));
    L = std::min<uint32_t>(std::max<uint32_t>(L, 8), 512);
    if (seg_override) L = seg_override;
    p.seg_len = L;
    p.segs_pw = n ? (n + L - 1) / L : 0;
    uint32_t m = pow2_floor(std::max<uint64_t>(1, ((uint64_t)p.windows * p.nb) / 65536));
    m = std::min<uint32_t>(std::min<uint32_t>(m, 32), p.nb);
    p.chunk = m;
    p.chunks_pw = p.nb / m;

    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    p.off_counts = off;  off = align(off + (size_t)p.windows * p.nb * 4 + 4);      // + the big-bucket counter
    p.off_offsets = off; off = align(off + (size_t)p.windows * (p.nb + 1) * 4);
    p.off_cursor = off;  off = align(off + (size_t)p.windows * p.nb * 4);
    p.off_biglist = off; off = align(off + (size_t)p.windows * p.nb * 4);
    p.off_digits = off;  off = align(off + (size_t)p.windows * n * 2);
    p.off_sorted = off;  off = align(off + (size_t)p.windows * n * 4);
    p.off_slots = off;   off = align(off + (size_t)p.windows * ((size_t)p.segs_pw + p.nb) * 4 * fq_bytes);
    p.off_chunks = off;  off = align(off + (size_t)p.windows * p.chunks_pw * 2 * 4 * fq_bytes);
    p.off_wsums = off;   off = align(off + (size_t)p.windows * 4 * fq_bytes);
    p.bytes = off;
    return p;
}


cudaError_t msm_run(CurveId curve, const void *bases, const void *scalars, uint32_t n, void *result, CoordType coord,
                    cudaMemPool_t pool, cudaStream_t stream, uint32_t c_override, uint32_t seg_override, MsmStageTimes *timings) {
    if (curve == CURVE_BLS12_377) return msm_run_bls12_377(bases, scalars, n, result, coord, pool, stream, c_override, seg_override, timings);
    return msm_run_bn254(bases, scalars, n, result, coord, pool, stream, c_override, seg_override, timings);
}

cudaError_t msm_combine(CurveId curve, const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream) {
    if (curve == CURVE_BLS12_377) return msm_combine_bls12_377(partials, count, result, coord, stream);
    return msm_combine_bn254(partials, count, result, coord, stream);
}

}  // namespace pb
