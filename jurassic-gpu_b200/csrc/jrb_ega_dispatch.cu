// jrb_ega_dispatch.cu -- picks the specialised EGA kernel for (ng, continuum mask).
// JRB_MASK_LIST (bit m set = mask m was built) lets development builds compile a subset.
#include "jrb_ega_fast.cuh"
#include "jrb_ega_tiled.cuh"

#ifndef JRB_MASK_LIST
#define JRB_MASK_LIST 0xffff
#endif

namespace jrb {

#define JRB_DECL(m)                                                                           \
  template <> cudaError_t launch_ega_fast_mask<m>(const EgaArgs &, cudaStream_t, int, int *); \
  template <> cudaError_t launch_ega_tiled_mask<m>(const EgaArgs &, cudaStream_t, int, int *);
JRB_DECL(0) JRB_DECL(1) JRB_DECL(2) JRB_DECL(3) JRB_DECL(4) JRB_DECL(5) JRB_DECL(6) JRB_DECL(7)
JRB_DECL(8) JRB_DECL(9) JRB_DECL(10) JRB_DECL(11) JRB_DECL(12) JRB_DECL(13) JRB_DECL(14) JRB_DECL(15)
#undef JRB_DECL

bool ega_fast_available(int ng, int ctm_mask) {
  return ng >= 0 && ng <= 32 && ctm_mask >= 0 && ctm_mask < 16 && ((JRB_MASK_LIST >> ctm_mask) & 1);
}

// can the specialised kernel run at all for this shape?  (a one-warp CTA is its smallest configuration: rpw staged
// records per warp plus 16 B of state per gas and thread must fit the opt-in shared memory)
bool ega_fast_fits(int ng, int los_head, int cpw, size_t smem_max) {
  const int rpw = cpw < 32 ? 32 / (cpw < 1 ? 1 : cpw) : 1;
  return ega_fast_smem_bytes(ng, los_head, 32, rpw) <= smem_max;
}

template <int M>
static cudaError_t call_mask(const EgaArgs &a, cudaStream_t s, int sm, int *ngb) {
  if constexpr (((JRB_MASK_LIST) >> M) & 1) return launch_ega_fast_mask<M>(a, s, sm, ngb);
  else return cudaErrorInvalidValue;
}

bool ega_tiled_fits(int ng, int los_rec, size_t smem_max) { return ega_tiled_smem_bytes(ng, los_rec, 256) <= smem_max; }

template <int M>
static cudaError_t call_tiled(const EgaArgs &a, cudaStream_t s, int sm, int *nl) {
  if constexpr (((JRB_MASK_LIST) >> M) & 1) return launch_ega_tiled_mask<M>(a, s, sm, nl);
  else return cudaErrorInvalidValue;
}

// segment-tiled form of the specialised kernel (one ray x 32 channels per warp, shared (p,T) axes, fused gas loop)
cudaError_t launch_ega_tiled(const EgaArgs &a, cudaStream_t stream, int *n_launched) {
  if (!ega_fast_available(a.ng, a.ctm_mask) || a.per_channel_axes || a.cpw != 32) return cudaErrorInvalidValue;
  int dev = 0, sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  switch (a.ctm_mask) {
    case 0: return call_tiled<0>(a, stream, sm, n_launched);
    case 1: return call_tiled<1>(a, stream, sm, n_launched);
    case 2: return call_tiled<2>(a, stream, sm, n_launched);
    case 3: return call_tiled<3>(a, stream, sm, n_launched);
    case 4: return call_tiled<4>(a, stream, sm, n_launched);
    case 5: return call_tiled<5>(a, stream, sm, n_launched);
    case 6: return call_tiled<6>(a, stream, sm, n_launched);
    case 7: return call_tiled<7>(a, stream, sm, n_launched);
    case 8: return call_tiled<8>(a, stream, sm, n_launched);
    case 9: return call_tiled<9>(a, stream, sm, n_launched);
    case 10: return call_tiled<10>(a, stream, sm, n_launched);
    case 11: return call_tiled<11>(a, stream, sm, n_launched);
    case 12: return call_tiled<12>(a, stream, sm, n_launched);
    case 13: return call_tiled<13>(a, stream, sm, n_launched);
    case 14: return call_tiled<14>(a, stream, sm, n_launched);
    case 15: return call_tiled<15>(a, stream, sm, n_launched);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_ega_fast(const EgaArgs &a, cudaStream_t stream, int *ngb_out) {
  if (!ega_fast_available(a.ng, a.ctm_mask)) return cudaErrorInvalidValue;
  int dev = 0, sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  switch (a.ctm_mask) {
    case 0: return call_mask<0>(a, stream, sm, ngb_out);
    case 1: return call_mask<1>(a, stream, sm, ngb_out);
    case 2: return call_mask<2>(a, stream, sm, ngb_out);
    case 3: return call_mask<3>(a, stream, sm, ngb_out);
    case 4: return call_mask<4>(a, stream, sm, ngb_out);
    case 5: return call_mask<5>(a, stream, sm, ngb_out);
    case 6: return call_mask<6>(a, stream, sm, ngb_out);
    case 7: return call_mask<7>(a, stream, sm, ngb_out);
    case 8: return call_mask<8>(a, stream, sm, ngb_out);
    case 9: return call_mask<9>(a, stream, sm, ngb_out);
    case 10: return call_mask<10>(a, stream, sm, ngb_out);
    case 11: return call_mask<11>(a, stream, sm, ngb_out);
    case 12: return call_mask<12>(a, stream, sm, ngb_out);
    case 13: return call_mask<13>(a, stream, sm, ngb_out);
    case 14: return call_mask<14>(a, stream, sm, ngb_out);
    case 15: return call_mask<15>(a, stream, sm, ngb_out);
  }
  return cudaErrorInvalidValue;
}

} // namespace jrb
