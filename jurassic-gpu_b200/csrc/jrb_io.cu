// jrb_io.cu -- package I/O kernels: how inputs reach the device arrays and how results reach host memory.
//
// The reference moves whole structs per call: cudaMemcpyAsync of atm_t (2.8 MB) and obs_t (1.8 MB) in, obs_t out, from
// pageable memory (src/GPUdrivers.cu:222-223,244).  Here
//   * results are stored by the compute kernels themselves into host-mapped memory as each ray finishes (per-ray row
//     addresses prepared by stage_kernel), so there is no device-to-host copy phase at all;
//   * when the caller's atm_t / obs_t blocks are page-locked (jrb_host_register), stage_kernel gathers the populated
//     prefixes of their arrays directly over PCIe into the device SoA -- no host-side packing, no staging buffer --
//     and the result rows above are the caller's own obs_t rows.
#include "jrb_internal.h"

namespace jrb {

namespace {

// grid: x = package, y = field.  y < n_geo: obs_t input arrays (field 0 also fills ray_pkg); n_geo <= y < n_geo +
// n_atm_fields: atm_t arrays (the first also records np); y == n_geo + n_atm_fields: per-ray output addresses.
__global__ void __launch_bounds__(256) stage_kernel(const StageArgs a) {
  const int pk = blockIdx.x, f = blockIdx.y;
  const long long r0 = a.ray_off[pk], nr = a.ray_off[pk + 1] - r0;
  const long long a0 = a.atm_off[pk], np = a.atm_off[pk + 1] - a0;
  const int nf = a.n_geo + a.n_atm_fields;
  if (f < a.n_geo) {
    if (a.src) {
      const double *__restrict__ s = a.src[(size_t)pk * nf + f];
      double *__restrict__ d = a.geo + (size_t)f * a.R + r0;
      for (long long i = threadIdx.x; i < nr; i += blockDim.x) d[i] = s[i];
    }
    if (f == 0)
      for (long long i = threadIdx.x; i < nr; i += blockDim.x) a.ray_pkg[r0 + i] = pk;
  } else if (f < nf) {
    if (a.src) {
      const double *__restrict__ s = a.src[(size_t)pk * nf + f];
      double *__restrict__ d = a.atm + (size_t)(f - a.n_geo) * a.A + a0;
      for (long long i = threadIdx.x; i < np; i += blockDim.x) d[i] = s[i];
    }
    if (f == a.n_geo && threadIdx.x == 0) a.pkg_atm_np[pk] = (int)np;
  } else {
    const OutTab o = a.out[pk];
    for (long long i = threadIdx.x; i < nr; i += blockDim.x) {
      a.ray_out[0 * a.R + r0 + i] = o.rad + i * o.stride;
      a.ray_out[1 * a.R + r0 + i] = o.tau + i * o.stride;
      a.ray_out[2 * a.R + r0 + i] = o.tpz + i;
      a.ray_out[3 * a.R + r0 + i] = o.tplon + i;
      a.ray_out[4 * a.R + r0 + i] = o.tplat + i;
    }
  }
}

__global__ void __launch_bounds__(256) publish_kernel(const double *__restrict__ rad, const double *__restrict__ tau,
                                                      double *const *__restrict__ rad_host, double *const *__restrict__ tau_host,
                                                      const long long n, const int nd) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const long long ir = idx / nd;
  const int id = (int)(idx - ir * nd);
  rad_host[ir][id] = rad[idx];
  tau_host[ir][id] = tau[idx];
}

} // namespace

cudaError_t launch_stage(const StageArgs &a, cudaStream_t stream) {
  if (a.npk <= 0) return cudaSuccess;
  dim3 grid((unsigned)a.npk, (unsigned)(a.n_geo + a.n_atm_fields + 1));
  stage_kernel<<<grid, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_publish(const double *rad, const double *tau, double *const *rad_host, double *const *tau_host, long long n_rays,
                           int nd, cudaStream_t stream) {
  const long long n = n_rays * nd;
  if (n <= 0) return cudaSuccess;
  publish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rad, tau, rad_host, tau_host, n, nd);
  return cudaGetLastError();
}

} // namespace jrb
