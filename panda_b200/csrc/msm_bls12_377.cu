// msm_bls12_377.cu -- instantiates the MSM pipeline (msm_impl.cuh) for one curve.
#include "msm_impl.cuh"

namespace pb {

cudaError_t msm_run_bls12_377(const void *bases, const void *scalars, uint32_t n, void *result, CoordType coord, cudaMemPool_t pool,
                         cudaStream_t stream, uint32_t c_override, uint32_t seg_override, MsmStageTimes *timings) {
    return msm_run_t<Bls377>(CURVE_BLS12_377, bases, scalars, n, result, coord, pool, stream, c_override, seg_override, timings);
}
cudaError_t msm_combine_bls12_377(const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream) {
    return msm_combine_t<Bls377>(partials, count, result, coord, stream);
}

}  // namespace pb
