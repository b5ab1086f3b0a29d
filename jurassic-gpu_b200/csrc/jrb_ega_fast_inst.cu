// jrb_ega_fast_inst.cu -- compiled once per continuum mask (-DJRB_MASK=0..15): instantiates the specialised EGA
// kernels for 1..8 gases.  (The reference stamps out its 16 variants with the X-macro header
// src/jr_multiversion4gases.h; here they are template instantiations spread over translation units so that they
// build in parallel.)
#include "jrb_ega_fast.cuh"

#ifndef JRB_MASK
#error "compile with -DJRB_MASK=<0..15>"
#endif

namespace jrb {

template <>
cudaError_t launch_ega_fast_mask<JRB_MASK>(const EgaArgs &a, cudaStream_t stream, int sm_count, int *ngb_out) {
  if (ngb_out) *ngb_out = a.ng;
  switch (a.ng) {
    case 1: return launch_ega_fast_t<1, JRB_MASK>(a, stream, sm_count);
    case 2: return launch_ega_fast_t<2, JRB_MASK>(a, stream, sm_count);
    case 3: return launch_ega_fast_t<3, JRB_MASK>(a, stream, sm_count);
    case 4: return launch_ega_fast_t<4, JRB_MASK>(a, stream, sm_count);
    case 5: return launch_ega_fast_t<5, JRB_MASK>(a, stream, sm_count);
    case 6: return launch_ega_fast_t<6, JRB_MASK>(a, stream, sm_count);
    case 7: return launch_ega_fast_t<7, JRB_MASK>(a, stream, sm_count);
    case 8: return launch_ega_fast_t<8, JRB_MASK>(a, stream, sm_count);
    default: return cudaErrorInvalidValue;
  }
}

} // namespace jrb
