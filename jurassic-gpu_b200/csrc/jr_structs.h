/* jr_structs.h -- layout-identical mirrors of the reference's interface structs (src/jurassic.h:215-425) for one
 * compile-time dimension set.  These four structs ARE the binary interface of formod(); field order, types and array
 * extents must match the caller's jurassic.h exactly (tests/test_abi.py compares every offset with the reference
 * build).  Nothing else of the reference header is reproduced here. */
#ifndef JR_STRUCTS_H
#define JR_STRUCTS_H
#include <stdint.h>

#ifndef ND
#define ND 100 /* channels  (src/jurassic.h:138) */
#endif
#ifndef NG
#define NG 30 /* emitters  (:143) */
#endif
#define NP 9600  /* atmospheric points (:148) */
#define NR 1088  /* rays per obs_t     (:151) */
#define NW 1     /* spectral windows   (:154) */
#define LEN 5000 /* string length      (:157) */
#define NLOS 400
#define NSHAPE 2048 /* points of a shape file (:172) */
#define TBLNP 40
#define TBLNT 30
#define TBLNU 304
#define TBLNS 1201

typedef struct {
  double time[NP], z[NP], lon[NP], lat[NP], p[NP], t[NP];
  double q[NG][NP];
  double k[NW][NP];
  int np, init;
} atm_t;

typedef struct {
  int ng;
  char emitter[NG][LEN];
  int nd, nw;
  double nu[ND];
  int window[ND];
  char tblbase[LEN];
  double hydz;
  int ctm_co2, ctm_h2o, ctm_n2, ctm_o2;
  int ip;
  double cz, cx;
  int refrac;
  double rayds, raydz;
  char fov[LEN];
  double retp_zmin, retp_zmax, rett_zmin, rett_zmax;
  double retq_zmin[NG], retq_zmax[NG];
  double retk_zmin[NW], retk_zmax[NW];
  int write_bbt, write_matrix, formod;
  char rfmbin[LEN], rfmhit[LEN];
  char rfmxsc[NG][LEN];
  int useGPU, checkmode;
  int MPIglobrank, MPIlocalrank;
  int read_binary, write_binary;
  int gpu_nbytes_shared_memory;
} ctl_t;

typedef struct {
  double time[NR], obsz[NR], obslon[NR], obslat[NR], vpz[NR], vplon[NR], vplat[NR];
  double tpz[NR], tplon[NR], tplat[NR];
  double tau[NR][ND];
  double rad[NR][ND];
  int nr;
} obs_t;

typedef struct {
  int32_t np[NG][ND];
  int32_t nt[NG][TBLNP][ND];
  int32_t nu[NG][TBLNP][TBLNT][ND];
  double p[NG][TBLNP][ND];
  double t[NG][TBLNP][TBLNT][ND];
  float u[NG][TBLNP][TBLNT][TBLNU][ND];
  float eps[NG][TBLNP][TBLNT][TBLNU][ND];
  double sr[TBLNS][ND];
  double st[TBLNS];
} tbl_t;

#endif
