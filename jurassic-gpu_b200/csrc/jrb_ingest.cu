// jrb_ingest.cu -- native ingest of the reference's ASCII emissivity tables and filter files.
//
// Reads "<tblbase>_<nu %.4f>_<GAS>.tab" (rows: p[hPa] T[K] u[molec/cm^2] eps) and "<tblbase>_<nu %.4f>.filt" (rows: nu f)
// with the acceptance rules of the reference's init_tbl (src/jurassic.c:326-416, 612-667) and builds the table arrays
// directly in compact form -- extents = what the files contain -- instead of the 8.8 GB tbl_t.  The result is handed to
// jrb_set_tables through a jrb_tbl_view like any other table set.  Host code only (OpenMP over the table files).
#include "jrb_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace jrb;

namespace {

struct PairTable { // one (gas, channel) file
  int np = 0;
  std::vector<double> p;               // [np]
  std::vector<int> nt;                 // [np]
  std::vector<std::vector<double>> t;  // [np][nt]
  std::vector<std::vector<std::vector<float>>> u, e; // [np][nt][nu]
  int status = 0;                      // 0 ok / missing, < 0 error
  std::string err;
};

// parse one .tab file exactly like the reference: new p level when the pressure value changes, new T level when the
// temperature value changes, a row extends the u axis only if both eps and u grow (or it is the first of its column);
// a rejected row still overwrites the current entry; at most max_u entries per column, further rows are ignored
void parse_tab(const char *filename, int max_p, int max_t, int max_u, PairTable &out) {
  FILE *in = std::fopen(filename, "r");
  if (!in) return; // a missing table is tolerated: the gas contributes the factor 1 (:341-345)
  char line[5000];
  double eps_old = -999, press_old = -999, temp_old = -999, u_old = -999;
  int ip = -1, it = -1, iu = -1;
  while (std::fgets(line, sizeof(line), in)) {
    double press = 0, temp = 0, u = 0, eps = 0;
    if (std::sscanf(line, "%lg %lg %lg %lg", &press, &temp, &u, &eps) != 4) continue;
    if (press != press_old) {
      press_old = press;
      if (++ip >= max_p) { out.status = -1; out.err = std::string("Too many pressure levels in ") + filename; break; }
      out.p.push_back(press); out.nt.push_back(0); out.t.emplace_back(); out.u.emplace_back(); out.e.emplace_back();
      it = -1;
    }
    if (temp != temp_old) {
      temp_old = temp;
      if (++it >= max_t) { out.status = -1; out.err = std::string("Too many temperatures in ") + filename; break; }
      out.t[ip].push_back(temp); out.u[ip].emplace_back(); out.e[ip].emplace_back();
      iu = -1;
    }
    if (it < 0) { // first temperature of a pressure block equals the last of the previous one: the reference would index [-1]
      out.status = -1; out.err = std::string("temperature axis does not restart at a new pressure level in ") + filename; break;
    }
    if ((eps > eps_old && u > u_old) || iu < 0) {
      eps_old = eps; u_old = u;
      if (++iu >= max_u) { iu--; continue; } // column full: row ignored (:373-378)
      out.u[ip][it].push_back(0.f); out.e[ip][it].push_back(0.f);
    }
    out.p[ip] = press; out.t[ip][it] = temp;
    out.u[ip][it][iu] = (float)u; out.e[ip][it][iu] = (float)eps;
  }
  std::fclose(in);
  if (out.status < 0) return;
  out.np = ip + 1;
  for (int i = 0; i < out.np; i++) out.nt[i] = (int)out.t[i].size();
}

} // namespace

struct jrb_host_tables {
  int ng = 0, nd = 0, dim_p = 1, dim_t = 1, dim_u = 1;
  std::vector<int32_t> np, nt, nu;
  std::vector<double> p, t, sr, st;
  std::vector<float> u, eps;
  int n_missing = 0;
};

extern "C" {

static std::string g_ingest_error;
const char *jrb_ingest_last_error(void) { return g_ingest_error.c_str(); }

int jrb_tables_read_ascii(const char *tblbase, int ng, const char *const *emitters, int nd, const double *nu, int max_p,
                          int max_t, int max_u, jrb_host_tables **out) {
  if (!tblbase || !out || ng < 0 || nd < 1 || (ng > 0 && !emitters) || !nu) return JRB_ERR_ARG;
  if (max_p <= 0) max_p = 40;   // TBLNP, TBLNT, TBLNU of the reference (src/jurassic.h:178-184)
  if (max_t <= 0) max_t = 30;
  if (max_u <= 0) max_u = 304;
  *out = nullptr;
  std::vector<PairTable> pairs((size_t)ng * nd);
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < ng * nd; k++) {
    const int ig = k / nd, id = k % nd;
    char fn[6000];
    std::snprintf(fn, sizeof(fn), "%s_%.4f_%s.tab", tblbase, nu[id], emitters[ig]);
    parse_tab(fn, max_p, max_t, max_u, pairs[k]);
  }
  jrb_host_tables *T = new jrb_host_tables();
  T->ng = ng; T->nd = nd;
  for (auto &pt : pairs) {
    if (pt.status < 0) { g_ingest_error = pt.err; delete T; return JRB_ERR_ARG; }
    if (pt.np == 0) T->n_missing++;
    T->dim_p = std::max(T->dim_p, pt.np);
    for (int ip = 0; ip < pt.np; ip++) {
      T->dim_t = std::max(T->dim_t, pt.nt[ip]);
      for (int it = 0; it < pt.nt[ip]; it++) T->dim_u = std::max(T->dim_u, (int)pt.u[ip][it].size());
    }
  }
  const size_t G = ng ? ng : 1, P = T->dim_p, TT = T->dim_t, U = T->dim_u, D = nd;
  T->np.assign(G * D, 0); T->nt.assign(G * P * D, 0); T->nu.assign(G * P * TT * D, 0);
  T->p.assign(G * P * D, 0.0); T->t.assign(G * P * TT * D, 0.0);
  T->u.assign(G * P * TT * U * D, 0.f); T->eps.assign(G * P * TT * U * D, 0.f);
  for (int ig = 0; ig < ng; ig++)
    for (int id = 0; id < nd; id++) {
      const PairTable &pt = pairs[(size_t)ig * nd + id];
      T->np[(size_t)ig * D + id] = pt.np;
      for (int ip = 0; ip < pt.np; ip++) {
        T->nt[((size_t)ig * P + ip) * D + id] = pt.nt[ip];
        T->p[((size_t)ig * P + ip) * D + id] = pt.p[ip];
        for (int it = 0; it < pt.nt[ip]; it++) {
          const size_t c = (((size_t)ig * P + ip) * TT + it) * D + id;
          const int n = (int)pt.u[ip][it].size();
          T->nu[c] = n;
          T->t[c] = pt.t[ip][it];
          for (int iu = 0; iu < n; iu++) {
            const size_t e = ((((size_t)ig * P + ip) * TT + it) * U + iu) * D + id;
            T->u[e] = pt.u[ip][it][iu]; T->eps[e] = pt.e[ip][it][iu];
          }
        }
      }
    }
  // source function: st[it] = 100 + 0.25 it K; sr = filter-weighted mean of the Planck function (:612-615, 645-667)
  T->st.resize(kTBLNS); T->sr.assign((size_t)kTBLNS * D, 0.0);
  for (int it = 0; it < kTBLNS; it++) T->st[it] = 100.0 + ((double)it - 0.0) * (400.0 - 100.0) / ((kTBLNS - 1.0) - 0.0);
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int id = 0; id < nd; id++) {
    char fn[6000];
    std::snprintf(fn, sizeof(fn), "%s_%.4f.filt", tblbase, nu[id]);
    FILE *in = std::fopen(fn, "r");
    if (!in) {
#pragma omp critical
      { bad = 1; g_ingest_error = std::string("missing filter file ") + fn; }
      continue;
    }
    std::vector<double> x, f;
    char line[5000];
    double a, b;
    while (std::fgets(line, sizeof(line), in))
      if (std::sscanf(line, "%lg %lg", &a, &b) == 2) { x.push_back(a); f.push_back(b); }
    std::fclose(in);
    if (x.empty() || x.size() > 2048) { // NSHAPE (src/jurassic.h:172)
#pragma omp critical
      { bad = 1; g_ingest_error = std::string("filter file empty or longer than NSHAPE=2048: ") + fn; }
      continue;
    }
    for (int it = 0; it < kTBLNS; it++) {
      double fsum = 0, fpsum = 0;
      for (size_t i = 0; i < x.size(); i++) {
        fsum += f[i];
        fpsum += f[i] * (kC1 * (x[i] * x[i] * x[i]) / std::expm1(kC2 * x[i] / T->st[it])); // planck(), :860
      }
      T->sr[(size_t)it * D + id] = fpsum / fsum;
    }
  }
  if (bad) { delete T; return JRB_ERR_ARG; }
  *out = T;
  return JRB_OK;
}

int jrb_host_tables_view(const jrb_host_tables *T, jrb_tbl_view *v, int *n_missing) {
  if (!T || !v) return JRB_ERR_ARG;
  v->dim_g = T->ng ? T->ng : 1; v->dim_p = T->dim_p; v->dim_t = T->dim_t; v->dim_u = T->dim_u; v->dim_d = T->nd; v->dim_s = kTBLNS;
  v->np = T->np.data(); v->nt = T->nt.data(); v->nu = T->nu.data();
  v->p = T->p.data(); v->t = T->t.data(); v->u = T->u.data(); v->eps = T->eps.data();
  v->sr = T->sr.data(); v->st = T->st.data();
  if (n_missing) *n_missing = T->n_missing;
  return JRB_OK;
}

void jrb_host_tables_free(jrb_host_tables *T) { delete T; }

} // extern "C"
