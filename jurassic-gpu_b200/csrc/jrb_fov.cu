// jrb_fov.cu -- field-of-view convolution as a device epilogue.
//
// What the reference's formod_fov computes (src/jurassic.c:214-258) when it is called after formod(): every ray's
// radiance and transmittance become the weighted mean, over the n points (dz_i, w_i) of the FOV shape file, of the
// pencil-beam values interpolated linearly in view-point altitude between the rays of the same package and time that
// lie within +-NFOV ray indices.  The interpolation extrapolates beyond the outermost neighbour and the index search is
// the reference's direction-aware bisection (src/jr_common.h:87-104), so ascending and descending scans both work.
//
// Layout: one thread per (ray, channel), channel fastest, so the gathers rad[ir2][id] of a warp are contiguous; the
// neighbour list and the bisection are uniform over the channels of a ray.  The pencil-beam values are read from the
// EGA kernel's output and the result goes to a scratch buffer that the runtime copies back (a ray is its neighbours'
// input, so the update cannot be done in place).
#include "jrb_internal.h"

namespace jrb {

namespace {

constexpr int kNFOV = 5;             // src/jurassic.h:175
constexpr int kNB = 2 * kNFOV + 1;

__device__ __forceinline__ int locate_any(const double *xx, int n, double x) {
  int ilo = 0, ihi = n - 1, i = (n - 1) >> 1;
  if (xx[i] < xx[i + 1]) {
    while (ihi > ilo + 1) { i = (ihi + ilo) >> 1; if (xx[i] > x) ihi = i; else ilo = i; }
  } else {
    while (ihi > ilo + 1) { i = (ihi + ilo) >> 1; if (xx[i] <= x) ihi = i; else ilo = i; }
  }
  return ilo;
}

__global__ void __launch_bounds__(256) fov_kernel(FovArgs a) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= a.n_rays * a.nd) return;
  const long long ir = tid / a.nd;
  const int id = (int)(tid - ir * a.nd);
  const int pk = a.ray_pkg[ir];
  const double t = a.time[ir];
  double z[kNB], r[kNB], ta[kNB];
  int nz = 0;
  const long long lo = ir - kNFOV > 0 ? ir - kNFOV : 0, hi = ir + 1 + kNFOV < a.n_rays ? ir + 1 + kNFOV : a.n_rays;
  for (long long i2 = lo; i2 < hi; i2++)
    if (a.ray_pkg[i2] == pk && a.time[i2] == t) {
      z[nz] = a.vpz[i2];
      r[nz] = a.rad_in[i2 * a.nd + id];
      ta[nz] = a.tau_in[i2 * a.nd + id];
      nz++;
    }
  if (nz < 2) { // "Cannot apply FOV convolution!" (src/jurassic.c:236): fatal in the reference, reported by the runtime
    if (id == 0) atomicOr(a.error, 1);
    return;
  }
  const double z0 = a.vpz[ir];
  double srad = 0, stau = 0, wsum = 0;
  for (int i = 0; i < a.n_shape; i++) {
    const double zf = z0 + a.dz[i], w = a.w[i];
    const int k = locate_any(z, nz, zf);
    const double f = (zf - z[k]), d = (z[k + 1] - z[k]);
    srad += w * (r[k] + f * (r[k + 1] - r[k]) / d);
    stau += w * (ta[k] + f * (ta[k + 1] - ta[k]) / d);
    wsum += w;
  }
  a.rad_out[tid] = srad / wsum;
  a.tau_out[tid] = stau / wsum;
}

// apply_mask (src/jr_common.h:203-210) on the device: needed before the convolution because formod() masks first
__global__ void nan_mask_kernel(double *rad, const long long *flat, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rad[flat[i]] = __longlong_as_double(0x7ff8000000000000LL);
}

} // namespace

cudaError_t launch_fov(const FovArgs &a, cudaStream_t stream) {
  const long long n = a.n_rays * a.nd;
  if (n <= 0 || a.n_shape <= 0) return cudaSuccess;
  fov_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_nan_mask(double *rad, const long long *flat, long long n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  nan_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rad, flat, n);
  return cudaGetLastError();
}

} // namespace jrb
