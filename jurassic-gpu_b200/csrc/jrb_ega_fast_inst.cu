// jrb_ega_fast_inst.cu -- compiled once per continuum mask (-DJRB_MASK=0..15): instantiates the specialised EGA
// kernel for that continuum combination.  (The reference stamps out its 16 variants with the X-macro header
// src/jr_multiversion4gases.h; here they are template instantiations spread over translation units so that they
// build in parallel.)
#include "jrb_ega_fast.cuh"
#include "jrb_ega_tiled.cuh"

#ifndef JRB_MASK
#error "compile with -DJRB_MASK=<0..15>"
#endif

namespace jrb {

template <>
cudaError_t launch_ega_fast_mask<JRB_MASK>(const EgaArgs &a, cudaStream_t stream, int sm_count, int *ngb_out) {
  if (ngb_out) *ngb_out = a.ng;
  return launch_ega_fast_t<JRB_MASK>(a, stream, sm_count);
}

template <>
cudaError_t launch_ega_tiled_mask<JRB_MASK>(const EgaArgs &a, cudaStream_t stream, int sm_count, int *n_launched) {
  return a.unsorted_columns ? launch_ega_tiled_tm<JRB_MASK, true>(a, stream, sm_count, n_launched) : launch_ega_tiled_tm<JRB_MASK, false>(a, stream, sm_count, n_launched);
}

} // namespace jrb
