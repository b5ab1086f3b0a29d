// jrb_ingest.cu -- native ingest of the reference's ASCII emissivity tables and filter files, and reader / writer of its
// binary table cache.
//
// Reads "<tblbase>_<nu %.4f>_<GAS>.tab" (rows: p[hPa] T[K] u[molec/cm^2] eps) and "<tblbase>_<nu %.4f>.filt" (rows: nu f)
// with the acceptance rules of the reference's init_tbl (src/jurassic.c:326-416, 612-667) and builds the table arrays
// directly in compact form -- extents = what the files contain -- instead of the 8.8 GB tbl_t.  The result is handed to
// jrb_set_tables through a jrb_tbl_view like any other table set.  Host code only (OpenMP over the table files).
#include "jrb_host.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

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
  // binary cache: the arrays stay in the (read-only, lazily paged) file mapping
  void *map = nullptr;
  size_t map_len = 0;
  jrb_tbl_view mapped{};
  ~jrb_host_tables() { if (map) munmap(map, map_len); }
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
#pragma omp parallel for schedule(dynamic, 1) num_threads(host_threads())
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
#pragma omp parallel for schedule(dynamic, 1) num_threads(host_threads())
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
  if (T->map) { *v = T->mapped; if (n_missing) *n_missing = 0; return JRB_OK; }
  v->dim_g = T->ng ? T->ng : 1; v->dim_p = T->dim_p; v->dim_t = T->dim_t; v->dim_u = T->dim_u; v->dim_d = T->nd; v->dim_s = kTBLNS;
  v->np = T->np.data(); v->nt = T->nt.data(); v->nu = T->nu.data();
  v->p = T->p.data(); v->t = T->t.data(); v->u = T->u.data(); v->eps = T->eps.data();
  v->sr = T->sr.data(); v->st = T->st.data();
  if (n_missing) *n_missing = T->n_missing;
  return JRB_OK;
}

void jrb_host_tables_free(jrb_host_tables *T) { delete T; }

// ---- the reference's binary table cache (src/jr_binary_tables_io.h) --------------------------------------------------
// File = 16 384-byte NUL-padded text header ("key value" lines) + the raw tbl_t of the build that wrote it.  The header
// names the compile-time extents, so the struct layout can be reconstructed for any (NG, ND, TBLN*) without recompiling.

namespace {

constexpr long long kBinVersion = 20200211; // BINARY_TABLES_VERSION (:10)
constexpr size_t kBinHeaderLen = 16384;     // BINARY_TABLES_HEADER_LEN (:8)

struct BinLayout { // byte offsets of the tbl_t members (src/jurassic.h:390-425) for given extents
  size_t np, nt, nu, p, t, u, eps, sr, st, total;
};
BinLayout bin_layout(size_t G, size_t P, size_t T, size_t U, size_t D, size_t S) {
  BinLayout L;
  size_t o = 0;
  auto al8 = [](size_t x) { return (x + 7) / 8 * 8; };
  L.np = o; o += 4 * G * D;
  L.nt = o; o += 4 * G * P * D;
  L.nu = o; o += 4 * G * P * T * D;
  o = al8(o);
  L.p = o; o += 8 * G * P * D;
  L.t = o; o += 8 * G * P * T * D;
  L.u = o; o += 4 * G * P * T * U * D;
  L.eps = o; o += 4 * G * P * T * U * D;
  o = al8(o);
  L.sr = o; o += 8 * S * D;
  L.st = o; o += 8 * S;
  L.total = al8(o);
  return L;
}

bool is_number(const char *s) {
  char *end = nullptr;
  std::strtod(s, &end);
  return end != s && *end == 0;
}

} // namespace

int jrb_binary_tables_filename(char *out, size_t cap, int NG, int TBLNP, int TBLNT, int TBLNU, int ND) {
  if (!out) return JRB_ERR_ARG; // naming convention of jr_binary_tables_filename (:12-16), float payload
  const int n = std::snprintf(out, cap, "bin.jurassic-fp32-tables-g%d-p%d-T%d-u%d-d%d", NG, TBLNP, TBLNT, TBLNU, ND);
  return (n > 0 && (size_t)n < cap) ? JRB_OK : JRB_ERR_ARG;
}

size_t jrb_binary_tables_size(int NG, int TBLNP, int TBLNT, int TBLNU, int ND) {
  return kBinHeaderLen + bin_layout(NG, TBLNP, TBLNT, TBLNU, ND, kTBLNS).total;
}

// Header acceptance follows jr_binary_tables_check_header (:65-211): version not newer than ours, float payload, at least
// ng gases / nd channels, the gas at every index < ng and the channel ("%.4f") at every index < nd equal to the caller's.
int jrb_tables_read_binary(const char *filename, int ng, const char *const *emitters, int nd, const double *nu,
                           jrb_host_tables **out) {
  if (!filename || !out || ng < 0 || nd < 1 || (ng > 0 && !emitters) || !nu) return JRB_ERR_ARG;
  *out = nullptr;
  const int fd = open(filename, O_RDONLY);
  if (fd < 0) { g_ingest_error = std::string("cannot open binary tables file ") + filename; return JRB_ERR_ARG; }
  struct stat sb;
  if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < kBinHeaderLen) {
    close(fd); g_ingest_error = std::string("binary tables file shorter than its header: ") + filename; return JRB_ERR_ARG;
  }
  std::vector<char> header(kBinHeaderLen + 1, 0);
  if (pread(fd, header.data(), kBinHeaderLen, 0) != (ssize_t)kBinHeaderLen) {
    close(fd); g_ingest_error = "cannot read the binary tables header"; return JRB_ERR_ARG;
  }
  long long G = -1, P = -1, T = -1, U = -1, D = -1, S = -1, header_size = -1, table_size = -1;
  std::vector<char> gas_found(ng ? ng : 1, 0), nu_found(nd, 0);
  std::string problems;
  auto bad = [&](const std::string &m) { problems += (problems.empty() ? "" : "; ") + m; };
  for (char *h = header.data(); *h;) {
    char key[64] = "";
    long long v = 0;
    const int got = std::sscanf(h, "%63s %lld", key, &v);
    if (got >= 2) {
      const std::string k(key);
      if (k == "JURASSIC" || k == "git_key" || k == "file_size" || k == "header_end" || k == "FAST_INVERSE_OF_U") {
      } else if (k == "version") { if (v > kBinVersion) bad("file version is newer than " + std::to_string(kBinVersion));
      } else if (k == "float") { if (v != 4) bad("float payload must be 4 bytes");
      } else if (k == "double") { bad("double-precision table payload is not supported");
      } else if (k == "NG") G = v; else if (k == "TBLNP") P = v; else if (k == "TBLNT") T = v; else if (k == "TBLNU") U = v;
      else if (k == "ND") D = v; else if (k == "TBLNS") S = v;
      else if (k == "header_size") header_size = v; else if (k == "table_size") table_size = v;
      else if (k == "ng") { if (v < ng) bad("file holds fewer gases than requested"); }
      else if (k == "nd") { if (v < nd) bad("file holds fewer channels than requested"); }
      else if (is_number(key)) { // channel list entry "<nu %.4f> <index>"
        if (v < 0 || (D >= 0 && v >= D)) bad(std::string("channel index out of range for ") + key);
        else if (v < nd) {
          char want[32];
          std::snprintf(want, sizeof(want), "%.4f", nu[v]);
          if (k == want) nu_found[v] = 1; else bad(std::string("channel ") + std::to_string(v) + " is " + key + ", requested " + want);
        }
      } else { // gas list entry "<NAME> <index>"
        if (v < 0 || (G >= 0 && v >= G)) bad(std::string("gas index out of range for ") + key);
        else if (v < ng) {
          if (k == emitters[v]) gas_found[v] = 1; else bad(std::string("gas ") + std::to_string(v) + " is " + key + ", requested " + emitters[v]);
        }
      }
    } else if (got == 1) {
      bad(std::string("header line without a value: ") + key);
    }
    while (*h && *h != '\n') ++h;
    while (*h == '\n') ++h;
  }
  for (int ig = 0; ig < ng; ig++) if (!gas_found[ig]) { bad(std::string("gas not in file: ") + emitters[ig]); break; }
  for (int id = 0; id < nd; id++) if (!nu_found[id]) { bad("channel not in file: " + std::to_string(nu[id])); break; }
  if (G < 1 || P < 1 || T < 1 || U < 1 || D < 1) bad("extents NG/TBLNP/TBLNT/TBLNU/ND missing");
  if (S != kTBLNS) bad("TBLNS must be 1201");
  if (header_size < 0) header_size = (long long)kBinHeaderLen; // key absent: the reference's fixed header length
  // the header is untrusted input: the reference writes exactly one header length, and every extent is a small
  // compile-time constant there (TBLNU = 304 is its largest) -- bounded here so that the size products cannot overflow
  // and the array views behind the header stay 8-byte aligned
  if (header_size != (long long)kBinHeaderLen) bad("header_size must be 16384");
  if (G > 4096 || P > 4096 || T > 4096 || U > 4096 || D > 4096) bad("extents NG/TBLNP/TBLNT/TBLNU/ND above 4096");
  else if (G >= 1 && P >= 1 && T >= 1 && U >= 1 && D >= 1 &&
           (long double)G * P * T * U * D * 8.0L + (long double)G * P * T * D * 12.0L > 4.0e12L) bad("table extents describe more than 4 TB");
  BinLayout L{};
  if (problems.empty()) {
    L = bin_layout(G, P, T, U, D, S);
    if (table_size != (long long)L.total) bad("table_size does not match the extents in the header");
    else if ((size_t)sb.st_size < (size_t)header_size + L.total) bad("file is truncated");
    if (ng > G || nd > D) bad("requested ng/nd exceed the extents of the file");
  }
  if (!problems.empty()) { close(fd); g_ingest_error = std::string(filename) + ": " + problems; return JRB_ERR_ARG; }
  const size_t len = (size_t)header_size + L.total;
  void *m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) { g_ingest_error = "mmap of the binary tables file failed"; return JRB_ERR_ARG; }
  jrb_host_tables *H = new jrb_host_tables();
  H->ng = ng; H->nd = nd; H->map = m; H->map_len = len;
  const unsigned char *b = (const unsigned char *)m + header_size;
  jrb_tbl_view &v = H->mapped;
  v.dim_g = (int)G; v.dim_p = (int)P; v.dim_t = (int)T; v.dim_u = (int)U; v.dim_d = (int)D; v.dim_s = (int)S;
  v.np = (const int32_t *)(b + L.np); v.nt = (const int32_t *)(b + L.nt); v.nu = (const int32_t *)(b + L.nu);
  v.p = (const double *)(b + L.p); v.t = (const double *)(b + L.t);
  v.u = (const float *)(b + L.u); v.eps = (const float *)(b + L.eps);
  v.sr = (const double *)(b + L.sr); v.st = (const double *)(b + L.st);
  // counts come from a file: make sure they cannot index outside the arrays
  for (long long ig = 0; ig < ng; ig++)
    for (long long id = 0; id < nd; id++) {
      const int n_p = v.np[ig * D + id];
      bool ok = n_p >= 0 && n_p <= P;
      for (long long ip = 0; ok && ip < n_p; ip++) {
        const int n_t = v.nt[(ig * P + ip) * D + id];
        ok = n_t >= 0 && n_t <= T;
        for (long long it = 0; ok && it < n_t; it++) {
          const int n_u = v.nu[((ig * P + ip) * T + it) * D + id];
          ok = n_u >= 0 && n_u <= U;
        }
      }
      if (!ok) { delete H; g_ingest_error = std::string(filename) + ": table counts exceed the extents"; return JRB_ERR_ARG; }
    }
  *out = H;
  return JRB_OK;
}

// Writes `tbl` (any allocated extents <= the target's) as the cache file of a reference build with the extents
// NG/TBLNP/TBLNT/TBLNU/ND; regions the view does not cover are holes of the (sparse) file and read back as zeros.
int jrb_tables_write_binary(const char *filename, const jrb_tbl_view *tbl, int ng, const char *const *emitters, int nd,
                            const double *nu, int NG, int TBLNP, int TBLNT, int TBLNU, int ND) {
  if (!filename || !tbl || ng < 0 || nd < 1 || (ng > 0 && !emitters) || !nu) return JRB_ERR_ARG;
  const jrb_tbl_view &v = *tbl;
  if (ng > NG || nd > ND || ng > v.dim_g || nd > v.dim_d || v.dim_s != kTBLNS) {
    g_ingest_error = "jrb_tables_write_binary: ng/nd exceed the extents"; return JRB_ERR_ARG;
  }
  // populated extents of the view must fit the target
  int mp = 0, mt = 0, mu = 0;
  for (int ig = 0; ig < ng; ig++)
    for (int id = 0; id < nd; id++) {
      const int n_p = v.np[(size_t)ig * v.dim_d + id];
      mp = std::max(mp, n_p);
      for (int ip = 0; ip < n_p; ip++) {
        const int n_t = v.nt[((size_t)ig * v.dim_p + ip) * v.dim_d + id];
        mt = std::max(mt, n_t);
        for (int it = 0; it < n_t; it++) mu = std::max(mu, v.nu[(((size_t)ig * v.dim_p + ip) * v.dim_t + it) * v.dim_d + id]);
      }
    }
  if (mp > TBLNP || mt > TBLNT || mu > TBLNU) { g_ingest_error = "jrb_tables_write_binary: tables exceed TBLNP/TBLNT/TBLNU"; return JRB_ERR_ARG; }
  const size_t G = NG, P = TBLNP, T = TBLNT, U = TBLNU, D = ND;
  const BinLayout L = bin_layout(G, P, T, U, D, kTBLNS);

  std::string h;
  char line[256];
  std::snprintf(line, sizeof(line), "JURASSIC 0  Binary Emissivity Tables\nversion %lld\nfloat   4\nNG     %d gases\nTBLNP  %d pressures\n"
                "TBLNT  %d temperatures\nTBLNU  %d column_densities\nND     %d detector_channels\nTBLNS  %d radiances",
                kBinVersion, NG, TBLNP, TBLNT, TBLNU, ND, kTBLNS);
  h += line;
  std::snprintf(line, sizeof(line), "\n\nfile_size   %lld\nheader_size %lld\ntable_size  %lld", (long long)(kBinHeaderLen + L.total),
                (long long)kBinHeaderLen, (long long)L.total);
  h += line;
  std::snprintf(line, sizeof(line), "\n\n\nng %d emitter gases:\n", ng);
  h += line;
  for (int ig = 0; ig < ng; ig++) { std::snprintf(line, sizeof(line), "%s %i\n", emitters[ig], ig); h += line; }
  std::snprintf(line, sizeof(line), "\n\n\nnd %d channels [cm^-1]:\n", nd);
  h += line;
  for (int id = 0; id < nd; id++) { std::snprintf(line, sizeof(line), "%.4f %i\n", nu[id], id); h += line; }
  h += "\n\nFAST_INVERSE_OF_U 0\n\n\nheader_end 0\n";
  if (h.size() > kBinHeaderLen - 16) { g_ingest_error = "jrb_tables_write_binary: header does not fit 16 KiB"; return JRB_ERR_LIMIT; }
  h.resize(kBinHeaderLen, '\0');

  const int fd = open(filename, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) { g_ingest_error = std::string("cannot create ") + filename; return JRB_ERR_ARG; }
  bool ok = pwrite(fd, h.data(), kBinHeaderLen, 0) == (ssize_t)kBinHeaderLen;
  ok = ok && ftruncate(fd, (off_t)(kBinHeaderLen + L.total)) == 0;
  // one row = the D innermost (channel) entries of one array element group; rows of the view are copied, the rest stays a hole
  std::vector<unsigned char> row;
  auto put_rows = [&](size_t base, size_t elem, const void *src, size_t n_rows_view, auto map_row) {
    // map_row(view_row) -> target row index, or (size_t)-1 to skip
    row.assign(D * elem, 0);
    for (size_t r = 0; ok && r < n_rows_view; r++) {
      const size_t tr = map_row(r);
      if (tr == (size_t)-1) continue;
      std::memset(row.data(), 0, row.size());
      std::memcpy(row.data(), (const unsigned char *)src + r * (size_t)v.dim_d * elem, (size_t)nd * elem);
      ok = pwrite(fd, row.data(), row.size(), (off_t)(kBinHeaderLen + base + tr * D * elem)) == (ssize_t)row.size();
    }
  };
  const size_t vg = ng, vp = std::min<size_t>(v.dim_p, P), vt = std::min<size_t>(v.dim_t, T), vu = std::min<size_t>(v.dim_u, U);
  put_rows(L.np, 4, v.np, (size_t)v.dim_g, [&](size_t r) { return r < vg ? r : (size_t)-1; });
  auto map_gp = [&](size_t r) { const size_t g = r / v.dim_p, p = r % v.dim_p; return (g < vg && p < vp) ? g * P + p : (size_t)-1; };
  auto map_gpt = [&](size_t r) {
    const size_t t = r % v.dim_t, gp = r / v.dim_t, g = gp / v.dim_p, p = gp % v.dim_p;
    return (g < vg && p < vp && t < vt) ? (g * P + p) * T + t : (size_t)-1;
  };
  auto map_gptu = [&](size_t r) {
    const size_t u = r % v.dim_u, gpt = r / v.dim_u, t = gpt % v.dim_t, gp = gpt / v.dim_t, g = gp / v.dim_p, p = gp % v.dim_p;
    if (!(g < vg && p < vp && t < vt && u < vu)) return (size_t)-1;
    // skip rows beyond every channel's nu: they hold nothing (keeps the file sparse)
    bool used = false;
    for (int id = 0; id < nd && !used; id++) used = (int)u < v.nu[((g * v.dim_p + p) * v.dim_t + t) * v.dim_d + id];
    return used ? ((g * P + p) * T + t) * U + u : (size_t)-1;
  };
  put_rows(L.nt, 4, v.nt, (size_t)v.dim_g * v.dim_p, map_gp);
  put_rows(L.nu, 4, v.nu, (size_t)v.dim_g * v.dim_p * v.dim_t, map_gpt);
  put_rows(L.p, 8, v.p, (size_t)v.dim_g * v.dim_p, map_gp);
  put_rows(L.t, 8, v.t, (size_t)v.dim_g * v.dim_p * v.dim_t, map_gpt);
  put_rows(L.u, 4, v.u, (size_t)v.dim_g * v.dim_p * v.dim_t * v.dim_u, map_gptu);
  put_rows(L.eps, 4, v.eps, (size_t)v.dim_g * v.dim_p * v.dim_t * v.dim_u, map_gptu);
  put_rows(L.sr, 8, v.sr, (size_t)kTBLNS, [&](size_t r) { return r; });
  ok = ok && pwrite(fd, v.st, 8 * (size_t)kTBLNS, (off_t)(kBinHeaderLen + L.st)) == (ssize_t)(8 * (size_t)kTBLNS);
  ok = (close(fd) == 0) && ok;
  if (!ok) { g_ingest_error = std::string("write error on ") + filename; return JRB_ERR_ARG; }
  return JRB_OK;
}

} // extern "C"
