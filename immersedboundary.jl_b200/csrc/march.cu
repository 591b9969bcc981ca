// Pencil-marching flux kernel, reference-exact arithmetic (compiled with -fmad=false): see march_kernel.cuh.
#include "march_kernel.cuh"

namespace ibx {

bool march_supported(const ibx_domain& D) { return D.nd == 3 && D.block_size == 8 && D.all_pow2; }

int march_flux(ibx_ctx* c, const ibx_domain& D, const int32_t* blocks, int n, int hyb, ibx_fluid f, int flux_kind, const float* P,
               const float* S, float* R, float* cfl, const double* GF, const float* GC, cudaStream_t st) {
  if (c->opt_arith == 1) return march_flux_fast(c, D, blocks, n, hyb, f, flux_kind, P, S, R, cfl, GF, GC, st);
  return march_flux_impl<false>(c, D, blocks, n, hyb, f, flux_kind, P, S, R, cfl, GF, GC, st);
}

}  // namespace ibx
