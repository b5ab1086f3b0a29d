// jrb_host.h -- host-side helpers of the runtime (table packing, channel constants).
#pragma once
#include <jurassic_b200.h>
#include "jrb_internal.h"
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

namespace jrb {

// tbl_t view -> position-independent blob (host copy); see jrb_device.cuh for the layout
int pack_tables(const jrb_tbl_view &v, int ng, int nd, std::vector<unsigned char> &blob, std::string &err);
// header (host copy) + device base address -> resolved device pointers
int resolve_tables(const void *blob_host_header, const unsigned char *dev_base, TblHeader &h, TblDev &t, std::string &err);
// channel-only continuum coefficients, SoA [CH_NFIELDS][nd]; mask = CO2*8+H2O*4+N2*2+O2
void channel_constants(int nd, const double *nus, int mask, std::vector<double> &chan);

// host threads for packing tables and packing / scattering packages.  Launchers such as torchrun export
// OMP_NUM_THREADS=1, which would serialise the host side of the end-to-end path, so the count is taken from
// JRB_HOST_THREADS or the hardware.
inline int host_threads() {
  static int n = 0;
  if (n == 0) {
    n = 8;
    if (const char *s = getenv("JRB_HOST_THREADS")) { int v = atoi(s); if (v > 0) n = v; }
    else {
      int hw = (int)std::thread::hardware_concurrency();
      // one process per GPU: share the cores with the other ranks of this node (torchrun exports LOCAL_WORLD_SIZE)
      if (const char *w = getenv("LOCAL_WORLD_SIZE")) { int lw = atoi(w); if (lw > 1 && hw > 0) hw = hw / lw; }
      if (hw > 0 && hw < n) n = hw;
      if (n < 1) n = 1;
    }
  }
  return n;
}

} // namespace jrb
