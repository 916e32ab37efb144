// msm_bls12_381.cu -- instantiates the MSM pipeline (msm_impl.cuh) for one curve.
#include "msm_impl.cuh"

namespace pb {

cudaError_t msm_pipeline_bls12_381(const MsmPlan &p, const void *points, const void *scalars, void *result, CoordType coord, cudaMemPool_t pool,
                            cudaStream_t stream, MsmStageTimes *timings, const MsmFeed *feed) {
    return msm_pipeline_t<Bls381>(p, points, scalars, result, coord, pool, stream, timings, feed);
}
cudaError_t msm_build_table_bls12_381(const void *bases, uint32_t n, uint32_t c, uint32_t W, uint32_t wide, void *table, cudaStream_t stream) {
    return msm_build_table_t<Bls381>(bases, n, c, W, wide, table, stream);
}
cudaError_t msm_combine_bls12_381(const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream) {
    return msm_combine_t<Bls381>(partials, count, result, coord, stream);
}

}  // namespace pb
