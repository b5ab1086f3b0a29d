// jrb_tables.cpp -- host side: re-layout of the reference's tbl_t into the packed device blob (see jrb_device.cuh)
// and the channel-only continuum coefficients.
//
// tbl_t (src/jurassic.h:390-425) stores u/eps as [g][p][T][u][d] floats with the channel innermost; the hot loop of
// the reference walks one (g,p,T,d) column with stride ND*4 bytes (src/jr_common.h:116-125).  Here every column
// becomes a contiguous run of float4 brackets {u_k, eps_k, u_k+1, eps_k+1}.
#include "jrb_host.h"
#include "jrb_ctm_data.h" // generated at build time from the reference's src/ctm*.tbl (tools/gen_ctm_data.py)

#include <cmath>
#include <cstring>
#include <vector>

namespace jrb {

namespace {
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
} // namespace

int pack_tables(const jrb_tbl_view &v, int ng, int nd, std::vector<unsigned char> &blob, std::string &err) {
  if (ng > v.dim_g || nd > v.dim_d) { err = "pack_tables: ng/nd exceed table dimensions"; return JRB_ERR_ARG; }
  if (v.dim_s != kTBLNS) { err = "pack_tables: source table must have TBLNS=1201 temperatures"; return JRB_ERR_ARG; }
  const size_t D = v.dim_d, P = v.dim_p, T = v.dim_t, U = v.dim_u;
  auto NPv = [&](int g, int d) { return v.np[(size_t)g * D + d]; };
  auto NTv = [&](int g, int ip, int d) { return v.nt[((size_t)g * P + ip) * D + d]; };
  auto NUv = [&](int g, int ip, int it, int d) { return v.nu[(((size_t)g * P + ip) * T + it) * D + d]; };
  auto PXv = [&](int g, int ip, int d) { return v.p[((size_t)g * P + ip) * D + d]; };
  auto TXv = [&](int g, int ip, int it, int d) { return v.t[(((size_t)g * P + ip) * T + it) * D + d]; };

  int npmax = 2, ntmax = 2;
  for (int g = 0; g < ng; g++)
    for (int d = 0; d < nd; d++) {
      const int np = NPv(g, d);
      if (np > (int)P || np < 0) { err = "pack_tables: np out of range"; return JRB_ERR_ARG; }
      if (np > npmax) npmax = np;
      for (int ip = 0; ip < np; ip++) {
        const int nt = NTv(g, ip, d);
        if (nt > (int)T || nt < 0) { err = "pack_tables: nt out of range"; return JRB_ERR_ARG; }
        if (nt > ntmax) ntmax = nt;
      }
    }
  if (npmax > 255 || ntmax > 255) { err = "pack_tables: more than 255 p or T levels"; return JRB_ERR_LIMIT; }

  // --- channel-independent axes? (per gas, over the channels that have a table) ---
  int all_shared = 1;
  std::vector<int> gnp(ng, 0), gnt((size_t)ng * npmax, 0);
  std::vector<double> gp((size_t)ng * npmax, 0.0), gt((size_t)ng * npmax * ntmax, 0.0);
  for (int g = 0; g < ng; g++) {
    int d0 = -1;
    for (int d = 0; d < nd; d++) if (NPv(g, d) >= 2) { d0 = d; break; }
    if (d0 < 0) continue; // gas without any table: factor 1 everywhere
    gnp[g] = NPv(g, d0);
    for (int ip = 0; ip < gnp[g]; ip++) {
      gp[(size_t)g * npmax + ip] = PXv(g, ip, d0);
      const int nt = NTv(g, ip, d0);
      gnt[(size_t)g * npmax + ip] = nt;
      for (int it = 0; it < nt; it++) gt[((size_t)g * npmax + ip) * ntmax + it] = TXv(g, ip, it, d0);
    }
    for (int d = 0; d < nd && all_shared; d++) {
      if (NPv(g, d) < 2) continue;
      if (NPv(g, d) != gnp[g]) { all_shared = 0; break; }
      for (int ip = 0; ip < gnp[g] && all_shared; ip++) {
        if (PXv(g, ip, d) != gp[(size_t)g * npmax + ip] || NTv(g, ip, d) != gnt[(size_t)g * npmax + ip]) { all_shared = 0; break; }
        for (int it = 0; it < gnt[(size_t)g * npmax + ip]; it++)
          if (TXv(g, ip, it, d) != gt[((size_t)g * npmax + ip) * ntmax + it]) { all_shared = 0; break; }
      }
    }
  }

  // --- do all gases (that have tables) share one (p,T) grid? ---
  int gas_axes_same = all_shared;
  {
    int g0 = -1;
    for (int g = 0; g < ng && gas_axes_same; g++) {
      if (gnp[g] < 2) continue;
      if (g0 < 0) { g0 = g; continue; }
      if (gnp[g] != gnp[g0]) { gas_axes_same = 0; break; }
      for (int ip = 0; ip < gnp[g] && gas_axes_same; ip++) {
        if (gp[(size_t)g * npmax + ip] != gp[(size_t)g0 * npmax + ip] || gnt[(size_t)g * npmax + ip] != gnt[(size_t)g0 * npmax + ip]) { gas_axes_same = 0; break; }
        for (int it = 0; it < gnt[(size_t)g * npmax + ip]; it++)
          if (gt[((size_t)g * npmax + ip) * ntmax + it] != gt[((size_t)g0 * npmax + ip) * ntmax + it]) { gas_axes_same = 0; break; }
      }
    }
  }

  // --- column descriptors ---
  const size_t ncol = (size_t)ng * npmax * ntmax * nd;
  std::vector<uint32_t> col_first(ncol, 0), col_nu(ncol, 0);
  uint64_t n_entries = 0;
  int max_nu = 0;
  for (int g = 0; g < ng; g++)
    for (int d = 0; d < nd; d++) { // pair-major order: one (gas,channel) slab is contiguous
      const int np = NPv(g, d);
      for (int ip = 0; ip < np; ip++) {
        const int nt = NTv(g, ip, d);
        for (int it = 0; it < nt; it++) {
          const int nu = NUv(g, ip, it, d);
          if (nu < 0 || nu > (int)U) { err = "pack_tables: nu out of range"; return JRB_ERR_ARG; }
          const size_t c = (((size_t)g * npmax + ip) * ntmax + it) * nd + d;
          col_nu[c] = (uint32_t)nu;
          if (nu > max_nu) max_nu = nu;
          if (nu >= 2) { col_first[c] = (uint32_t)n_entries; n_entries += (uint64_t)nu; }
        }
      }
    }
  if (n_entries >= 0xffffffffull) { err = "pack_tables: more than 2^32 table entries"; return JRB_ERR_LIMIT; }

  // --- blob layout ---
  TblHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = kTblMagic;
  h.ng = ng; h.nd = nd; h.npmax = npmax; h.ntmax = ntmax; h.all_shared = all_shared; h.monotone = 1;
  h.n_entries = n_entries;
  h.max_nu = max_nu;
  h.gas_axes_same = gas_axes_same;
  size_t off = align_up(sizeof(TblHeader), 256);
  auto place = [&](uint64_t &dst, size_t bytes) { dst = off; off = align_up(off + bytes, 256); };
  place(h.off_np, sizeof(int32_t) * ng * nd);
  place(h.off_nt, sizeof(int32_t) * (size_t)ng * npmax * nd);
  place(h.off_pax, sizeof(double) * (size_t)ng * npmax * nd);
  place(h.off_tax, sizeof(double) * ncol);
  place(h.off_col, sizeof(uint32_t) * 2 * ncol);
  place(h.off_brk, sizeof(float) * 4 * (size_t)(n_entries ? n_entries : 1));
  place(h.off_sr, sizeof(double) * (size_t)kTBLNS * nd);
  place(h.off_gnp, sizeof(int32_t) * ng);
  place(h.off_gnt, sizeof(int32_t) * (size_t)ng * npmax);
  place(h.off_gp, sizeof(double) * (size_t)ng * npmax);
  place(h.off_gt, sizeof(double) * (size_t)ng * npmax * ntmax);
  h.nbytes = off;
  blob.assign(off, 0);
  unsigned char *B = blob.data();

  int32_t *o_np = (int32_t *)(B + h.off_np);
  int32_t *o_nt = (int32_t *)(B + h.off_nt);
  double *o_pax = (double *)(B + h.off_pax);
  double *o_tax = (double *)(B + h.off_tax);
  uint32_t *o_col = (uint32_t *)(B + h.off_col);
  float *o_brk = (float *)(B + h.off_brk);
  double *o_sr = (double *)(B + h.off_sr);

  for (int g = 0; g < ng; g++)
    for (int d = 0; d < nd; d++) {
      const int np = NPv(g, d);
      o_np[(size_t)g * nd + d] = np;
      for (int ip = 0; ip < np; ip++) {
        const int nt = NTv(g, ip, d);
        o_nt[((size_t)g * npmax + ip) * nd + d] = nt;
        o_pax[((size_t)g * npmax + ip) * nd + d] = PXv(g, ip, d);
        for (int it = 0; it < nt; it++) o_tax[(((size_t)g * npmax + ip) * ntmax + it) * nd + d] = TXv(g, ip, it, d);
      }
    }
  for (size_t c = 0; c < ncol; c++) { o_col[2 * c] = col_first[c]; o_col[2 * c + 1] = col_nu[c]; }

  // --- brackets: read tbl_t with the channel innermost (its contiguous direction), write per-column runs ---
  int monotone = 1;
#pragma omp parallel for collapse(2) schedule(dynamic, 4) reduction(&& : monotone) num_threads(host_threads())
  for (int g = 0; g < ng; g++)
    for (int ip = 0; ip < npmax; ip++)
      for (int it = 0; it < ntmax; it++) {
        const size_t cbase = (((size_t)g * npmax + ip) * ntmax + it) * nd;
        uint32_t numax = 0;
        for (int d = 0; d < nd; d++) if (col_nu[cbase + d] >= 2 && col_nu[cbase + d] > numax) numax = col_nu[cbase + d];
        for (uint32_t iu = 0; iu < numax; iu++) {
          const size_t src = ((((size_t)g * P + ip) * T + it) * U + iu) * D;
          const size_t srcn = src + D; // next u entry
          for (int d = 0; d < nd; d++) {
            const uint32_t nu = col_nu[cbase + d];
            if (nu < 2 || iu >= nu) continue;
            float *b = o_brk + 4 * ((size_t)col_first[cbase + d] + iu);
            const float u0 = v.u[src + d], e0 = v.eps[src + d];
            b[0] = u0; b[1] = e0;
            if (iu + 1 < nu) {
              const float u1 = v.u[srcn + d], e1 = v.eps[srcn + d];
              b[2] = u1; b[3] = e1;
              if (!(u1 >= u0) || !(e1 >= e0)) { // the column is searched by full bisection on the device (flag = top bit of nu)
                monotone = 0;
#pragma omp atomic
                o_col[2 * (cbase + d) + 1] |= kColNonMonotone;
              }
            } else { b[2] = u0; b[3] = e0; }
          }
        }
      }
  h.monotone = monotone;

  for (int it = 0; it < kTBLNS; it++)
    for (int d = 0; d < nd; d++) o_sr[(size_t)it * nd + d] = v.sr[(size_t)it * D + d];
  std::memcpy(B + h.off_gnp, gnp.data(), sizeof(int32_t) * ng);
  std::memcpy(B + h.off_gnt, gnt.data(), sizeof(int32_t) * gnt.size());
  std::memcpy(B + h.off_gp, gp.data(), sizeof(double) * gp.size());
  std::memcpy(B + h.off_gt, gt.data(), sizeof(double) * gt.size());
  std::memcpy(B, &h, sizeof(h));
  return JRB_OK;
}

int resolve_tables(const void *blob_host_header, const unsigned char *dev_base, TblHeader &h, TblDev &t, std::string &err) {
  std::memcpy(&h, blob_host_header, sizeof(h));
  if (h.magic != kTblMagic) { err = "table blob: bad magic"; return JRB_ERR_ARG; }
  t.ng = h.ng; t.nd = h.nd; t.npmax = h.npmax; t.ntmax = h.ntmax;
  t.np = (const int32_t *)(dev_base + h.off_np);
  t.nt = (const int32_t *)(dev_base + h.off_nt);
  t.pax = (const double *)(dev_base + h.off_pax);
  t.tax = (const double *)(dev_base + h.off_tax);
  t.col = (const uint2 *)(dev_base + h.off_col);
  t.brk = (const float4 *)(dev_base + h.off_brk);
  t.sr = (const double *)(dev_base + h.off_sr);
  t.gnp = (const int32_t *)(dev_base + h.off_gnp);
  t.gnt = (const int32_t *)(dev_base + h.off_gnt);
  t.gp = (const double *)(dev_base + h.off_gp);
  t.gt = (const double *)(dev_base + h.off_gt);
  return JRB_OK;
}

// Channel-only parts of the continua (src/jr_common.h:315-390), evaluated once per channel in the reference's own
// operation order so the products are bit-identical to what the reference recomputes per segment.
void channel_constants(int nd, const double *nus, int mask, std::vector<double> &chan) {
  chan.assign((size_t)CH_NFIELDS * nd, 0.0);
  for (int id = 0; id < nd; id++) {
    const double nu = nus[id];
    chan[(size_t)CH_NU * nd + id] = nu;
    if ((mask & 8) && !(nu < 0 || nu >= 4000)) { // ctmco2 (:318-325)
      const double xw = nu * 0.5 + 1;
      const int iw = (int)xw;
      const double dw = xw - iw, ew = 1 - dw;
      chan[(size_t)CH_CO2_296 * nd + id] = ew * jrb_co2296[iw - 1] + dw * jrb_co2296[iw];
      chan[(size_t)CH_CO2_260 * nd + id] = ew * jrb_co2260[iw - 1] + dw * jrb_co2260[iw];
      chan[(size_t)CH_CO2_230 * nd + id] = ew * jrb_co2230[iw - 1] + dw * jrb_co2230[iw];
    }
    if ((mask & 4) && !(nu < 0 || nu >= 20000)) { // ctmh2o (:336-357)
      const double xw = nu / 10 + 1;
      const int iw = (int)xw;
      const double dw = xw - iw, ew = 1 - dw;
      const double cw296 = ew * jrb_h2o296[iw - 1] + dw * jrb_h2o296[iw];
      const double cw260 = ew * jrb_h2o260[iw - 1] + dw * jrb_h2o260[iw];
      const double cwfrn = ew * jrb_h2ofrn[iw - 1] + dw * jrb_h2ofrn[iw];
      double sfac = 1.;
      if ((nu > 820.) && (nu < 960.)) { // CKD self-continuum correction 820-960 cm^-1, in single precision (:345-351)
        static const signed char xfc_milli[16] = {3, 9, 15, 23, 29, 33, 37, 39, 40, 46, 36, 27, 10, 2, 0, 0};
        const float xx = (float)(nu * 0.1 - 82);
        const int ix = (int)xx;
        const float dx = xx - (float)ix;
        const float corr = (1 - dx) * (float)xfc_milli[ix] + dx * (float)xfc_milli[ix + 1];
        sfac += .001 * corr;
      }
      const double vf1 = nu - 370.;
      const double vf2 = vf1 * vf1;
      const double vf6 = vf2 * vf2 * vf2;
      const double fscal = 36100. / (vf2 + vf6 * 1e-8 + 36100.) * -.25 + 1.;
      chan[(size_t)CH_H2O_S296 * nd + id] = sfac * cw296;
      chan[(size_t)CH_H2O_RATIO * nd + id] = cw260 / cw296;
      chan[(size_t)CH_H2O_LNRATIO * nd + id] = std::log(cw260 / cw296);
      chan[(size_t)CH_H2O_FRN * nd + id] = cwfrn * fscal;
    }
    if ((mask & 2) && !(nu < 2120 || nu > 2605)) { // ctmn2 (:367-372)
      const double xnu = nu * 0.2 - 424;
      const int idx = (int)xnu;
      const double a1 = xnu - idx, a0 = 1 - a1;
      const double b1 = idx + 1 < 98 ? jrb_n2_ba[idx + 1] : 0.0, be1 = idx + 1 < 98 ? jrb_n2_betaa[idx + 1] : 0.0; // (a1==0 there)
      chan[(size_t)CH_N2_B * nd + id] = a0 * jrb_n2_ba[idx] + a1 * b1;
      chan[(size_t)CH_N2_BETA * nd + id] = a0 * jrb_n2_betaa[idx] + a1 * be1;
    }
    if ((mask & 1) && !(nu < 1360 || nu > 1805)) { // ctmo2 (:381-386)
      const double xnu = nu * 0.2 - 272;
      const int idx = (int)xnu;
      const double a1 = xnu - idx, a0 = 1 - a1;
      const double b1 = idx + 1 < 90 ? jrb_o2_ba[idx + 1] : 0.0, be1 = idx + 1 < 90 ? jrb_o2_betaa[idx + 1] : 0.0;
      chan[(size_t)CH_O2_B * nd + id] = a0 * jrb_o2_ba[idx] + a1 * b1;
      chan[(size_t)CH_O2_BETA * nd + id] = a0 * jrb_o2_betaa[idx] + a1 * be1;
    }
  }
}

} // namespace jrb
