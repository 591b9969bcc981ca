// Pencil-marching flux kernel, option "arithmetic" = 1: the same kernel template compiled WITH FMA contraction, Float32
// HLL combination and Green-Gauss difference, approximate reciprocals (march_kernel.cuh, physics.cuh: hll_flux_f32).
// Not bit-identical to the reference's roundings; DESIGN.md 4.1 states the measured error under both normalisations.
#include "march_kernel.cuh"

namespace ibx {

int march_flux_fast(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, int hyb, ibx_fluid f, int flux_kind, const float* P,
                    const float* S, float* R, float* cfl, const double* GF, const float* GC, cudaStream_t st) {
  return march_flux_impl<true>(c, D, blocks, n, hyb, f, flux_kind, P, S, R, cfl, GF, GC, st);
}

}  // namespace ibx
