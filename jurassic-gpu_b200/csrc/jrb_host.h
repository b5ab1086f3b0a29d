// jrb_host.h -- host-side helpers of the runtime (table packing, channel constants).
#pragma once
#include <jurassic_b200.h>
#include "jrb_internal.h"
#include <string>
#include <vector>

namespace jrb {

// tbl_t view -> position-independent blob (host copy); see jrb_device.cuh for the layout
int pack_tables(const jrb_tbl_view &v, int ng, int nd, std::vector<unsigned char> &blob, std::string &err);
// header (host copy) + device base address -> resolved device pointers
int resolve_tables(const void *blob_host_header, const unsigned char *dev_base, TblHeader &h, TblDev &t, std::string &err);
// channel-only continuum coefficients, SoA [CH_NFIELDS][nd]; mask = CO2*8+H2O*4+N2*2+O2
void channel_constants(int nd, const double *nus, int mask, std::vector<double> &chan);

} // namespace jrb
