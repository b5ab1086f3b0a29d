/* jurassic_b200.h -- C ABI of the B200-native EGA forward model (core library libjurassic_b200.so).
 *
 * Plain C: pointers, sizes and strides only; no CUDA or torch types.  The core library is independent of the
 * reference's compile-time dimensions (ND, NG, ...): callers describe their ctl_t / atm_t / obs_t / tbl_t through
 * the "views" below (base pointers + strides into the caller's structs, nothing is copied on the host side).
 * The reference-facing entry points (formod_GPU and the batched extension, taking the reference's own structs)
 * live in the thin per-(ND,NG) layer declared in jurassic_b200_dropin.h.
 *
 * Reference interface each piece replaces (paths below /root/reference):
 *   jrb_ctl_view        <- the ctl_t fields the path reads            src/jurassic.h:229-347
 *   jrb_atm_view        <- atm_t                                      src/jurassic.h:215-226
 *   jrb_obs_view        <- obs_t                                      src/jurassic.h:371-385
 *   jrb_tbl_view        <- tbl_t                                      src/jurassic.h:390-425
 *   jrb_set_tables      <- get_tbl_on_GPU                             src/GPUdrivers.cu:78-93
 *   jrb_formod_batch    <- formod_GPU / formod_one_package            src/GPUdrivers.cu:187-250, 262-360
 *                          (same work as formod_CPU, src/CPUdrivers.c:108-151, for npk packages at once)
 *   jrb_stage / jrb_run_staged / jrb_fetch_staged : the three phases of jrb_formod_batch, separately callable so
 *                          that the device path can be timed with inputs resident in HBM.
 *
 * All functions return 0 on success, a negative code on failure; jrb_last_error() gives the message.
 * There is NO CPU fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef JURASSIC_B200_H
#define JURASSIC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JRB_OK 0
#define JRB_ERR_ARG (-1)
#define JRB_ERR_CUDA (-2)
#define JRB_ERR_STATE (-3)
#define JRB_ERR_LIMIT (-4)

#define JRB_NLOS 400   /* max. LOS points per ray  (NLOS,  src/jurassic.h:169) */
#define JRB_TBLNS 1201 /* source-function temperatures (TBLNS, src/jurassic.h:187) */
#define JRB_MAX_NG 64  /* limits of the packed formats */
#define JRB_MAX_NW 8

/* Control parameters read by the path (values AFTER read_ctl's automatic CTM_* switch-off, src/jurassic.c:954-968). */
typedef struct {
  int ng, nd, nw;
  const double *nu;  /* [nd] channel centroid wavenumbers [cm^-1] */
  const int *window; /* [nd] extinction window per channel */
  int ctm_co2, ctm_h2o, ctm_n2, ctm_o2;
  int ig_co2, ig_h2o; /* find_emitter(ctl,"CO2"/"H2O"), -1 if absent (src/jurassic.c:198-207) */
  int refrac;
  double rayds, raydz;
  double hydz; /* < 0: no hydrostatic adjustment */
  int write_bbt;
  int formod; /* must be 2 (EGA) */
  int ip;     /* atmosphere interpolation (src/jurassic.c:685-691): 1 = 1-D profile, 2 = 2-D (profiles along a track), 3 = 3-D (weighted average) */
  double cz, cx; /* ip == 3: vertical / horizontal influence radius [km] (ctl->cz, ctl->cx) */
} jrb_ctl_view;

typedef struct {
  int np;
  double *time, *z, *lon, *lat, *p, *t; /* [np]; p is written when hydz >= 0 */
  double *q; long q_stride;             /* q[ig*q_stride + ip], ig < ng */
  double *k; long k_stride;             /* k[iw*k_stride + ip], iw < nw */
  double **q_rows;                      /* optional: if non-NULL, profile of gas ig is q_rows[ig][ip] (q/q_stride ignored); */
  double **k_rows;                      /* lets many atmospheres share all profiles but one (batched Jacobians)          */
} jrb_atm_view;

typedef struct {
  int nr;
  const double *time, *obsz, *obslon, *obslat, *vpz, *vplon, *vplat; /* [nr] inputs */
  double *tpz, *tplon, *tplat;                                       /* [nr] outputs */
  double *rad, *tau; /* outputs, element (ir,id) at [ir*row_stride + id]; rad is also read for the NaN mask */
  long row_stride;   /* = ND of the caller's obs_t */
  int nd_reset;      /* columns [nd, nd_reset) are reset to rad=0, tau=1 like the reference (= ND) */
} jrb_obs_view;

/* Row-major [g][p][t][u][d] arrays exactly as in tbl_t; dim_* are the ALLOCATED extents (NG, TBLNP, ...). */
typedef struct {
  int dim_g, dim_p, dim_t, dim_u, dim_d, dim_s;
  const int32_t *np; /* [g][d] */
  const int32_t *nt; /* [g][p][d] */
  const int32_t *nu; /* [g][p][t][d] */
  const double *p;   /* [g][p][d] */
  const double *t;   /* [g][p][t][d] */
  const float *u;    /* [g][p][t][u][d] */
  const float *eps;  /* [g][p][t][u][d] */
  const double *sr;  /* [s][d] */
  const double *st;  /* [s] */
} jrb_tbl_view;

typedef struct {
  long long n_packages, n_rays, n_ray_channels, n_los_points; /* of the last run */
  long long n_kernel_launches;                                /* kernels of this library launched by the last run */
  float ms_raytrace, ms_ega, ms_total_device;                 /* CUDA-event times of the last run (device path) */
  long long h2d_bytes, d2h_bytes;                             /* of the last stage / fetch */
  int ega_kernel_variant;                                     /* 0 generic, 1 specialised (template <continuum mask>) */
  int ega_ngb, ega_ctm_mask;                                  /* gases handled by the specialised kernel; continuum mask */
  long long table_blob_bytes;
  float host_ms_pack, host_ms_h2d, host_ms_d2h, host_ms_scatter; /* wall-clock phases of the last stage / fetch */
  int n_chunks, pipelined; /* LOS chunks of the last run; 1 if the tracer of chunk c+1 ran beside the EGA kernel of chunk c */
  int ega_phase_lock;      /* 1 if the specialised kernel ran its CTAs in lock step (rays of equal length) */
  int ega_channels_per_warp; /* channels of a ray handled by one warp of the specialised kernel (32, or fewer = several rays per warp) */
  int io_direct;           /* 1: inputs were gathered from / results stored into the caller's page-locked structs (jrb_host_register) */
  float host_ms_stage;     /* wall clock of the whole staging phase (tables of addresses, packing or input gather, allocation) */
  /* accumulated over all runs since the batch was staged (CUDA-event times; one EGA launch per run and LOS chunk) */
  long long cum_runs, cum_launches, cum_ega_launches;
  double cum_ms_ega, cum_ms_raytrace, cum_ms_device;
  int ega_per_channel_axes; /* 1: the (p,T) axes of the tables depend on the channel; the specialised kernel located the cells per lane */
  int ega_tiled;           /* 1: the segment-tiled form of the specialised kernel ran (large batches) */
  int ega_gas_blocks;      /* > 1: split mode -- gas-block passes + combine kernel (many gases, or a batch too small to fill the GPU) */
} jrb_stats;

typedef struct jrb_context jrb_context;

int jrb_device_count(void);
int jrb_create(jrb_context **out, int device);
void jrb_destroy(jrb_context *ctx);
const char *jrb_last_error(const jrb_context *ctx); /* ctx may be NULL: last error of jrb_create */

int jrb_context_device(const jrb_context *ctx);

int jrb_set_control(jrb_context *ctx, const jrb_ctl_view *ctl);
/* pack tbl_t into per-(gas,channel) slabs and upload (requires jrb_set_control first) */
int jrb_set_tables(jrb_context *ctx, const jrb_tbl_view *tbl);
/* host-only (no GPU needed): properties of the packed form of a table set.  all_shared: the (p,T) axes of every gas do
 * not depend on the channel (else the specialised kernel locates the table cell per channel); monotone: every column is
 * non-decreasing in u and eps (else the flagged columns are searched by plain bisection); gas_axes_same: all gases share one (p,T) grid (one table cell per LOS segment instead of ng) */
int jrb_tables_pack_info(const jrb_tbl_view *tbl, int ng, int nd, size_t *nbytes, int *all_shared, int *monotone,
                         unsigned long long *n_entries, int *gas_axes_same);
/* host-only: the packed blob itself (out == NULL: query the size); deterministic, so ranks can compare checksums */
int jrb_tables_pack_host(const jrb_tbl_view *tbl, int ng, int nd, void *out, size_t capacity, size_t *nbytes);
/* upload a host blob made by jrb_tables_pack_host (e.g. received from another rank) and use it */
int jrb_tables_upload_blob(jrb_context *ctx, const void *host_blob, size_t nbytes);
/* multi-GPU: the packed slabs are one position-independent device blob that can be broadcast (e.g. NCCL) */
int jrb_tables_blob(jrb_context *ctx, void **dev_ptr, size_t *nbytes);
int jrb_tables_alloc_blob(jrb_context *ctx, size_t nbytes, void **dev_ptr); /* receiver side */
int jrb_tables_adopt_blob(jrb_context *ctx);                                 /* after the blob has been filled */

/* lanes: ctx (same device as `from`, control already set) uses the packed tables of `from` -- no copy, shared ownership */
int jrb_tables_share(jrb_context *ctx, jrb_context *from);
/* upper limit of this context's line-of-sight scratch in GB (0: JRB_LOS_GB or the default of 72); larger batches are chunked */
int jrb_set_los_limit_gb(jrb_context *ctx, double gb);

/* Page-locking of caller memory (process-wide registry; replaces nothing in the reference, which copies whole structs
 * from pageable memory, src/GPUdrivers.cu:222-223,244).  A batch whose atm/obs arrays ALL lie in registered memory runs in
 * "direct" mode: a staging kernel gathers the inputs straight from the caller's structs over PCIe and the compute kernels
 * store every ray's rad/tau row and tangent point straight into the caller's obs rows while they run -- no host-side
 * packing, no copy phase, no scatter.  Anything else runs "staged" (pinned staging buffers, same results).  The caller
 * must keep registered memory mapped until jrb_host_unregister_all(); blocks may overlap or share pages. */
int jrb_host_register(void *ptr, size_t bytes);
int jrb_host_unregister(void *ptr, size_t bytes); /* the registered ranges lying completely inside the block */
int jrb_host_unregister_all(void);
int jrb_host_is_registered(const void *ptr, size_t bytes);

/* Native ingest of the reference's ASCII inputs "<tblbase>_<nu %.4f>_<GAS>.tab" / "<tblbase>_<nu %.4f>.filt" with the
 * acceptance rules of init_tbl (src/jurassic.c:311-416, 612-667), into compact host arrays (no 8.8 GB tbl_t).  Host only.
 * max_p/max_t/max_u <= 0 select the reference's TBLNP/TBLNT/TBLNU (40/30/304).  Missing .tab files are tolerated. */
typedef struct jrb_host_tables jrb_host_tables;
int jrb_tables_read_ascii(const char *tblbase, int ng, const char *const *emitters, int nd, const double *nu, int max_p,
                          int max_t, int max_u, jrb_host_tables **out);
int jrb_host_tables_view(const jrb_host_tables *t, jrb_tbl_view *view, int *n_missing);
void jrb_host_tables_free(jrb_host_tables *t);
const char *jrb_ingest_last_error(void);

/* The reference's binary table cache (src/jr_binary_tables_io.h:12-290; read/written by init_tbl when READ_BINARY /
 * WRITE_BINARY are set, src/jurassic.c:312-320, 669-671): a 16 KiB text header naming the writer's compile-time extents,
 * followed by its raw tbl_t.  Host only.  The reader maps the file instead of loading it (only the populated entries are
 * ever touched, by the packer) and accepts files of ANY extents that contain the requested gases and channels at the
 * same indices; the writer produces the file a reference build with the extents dim_g = NG, dim_p = TBLNP, dim_t = TBLNT,
 * dim_u = TBLNU, dim_d = ND would write (sparse where empty).
 * jrb_binary_tables_filename: the reference's naming convention "bin.jurassic-fp32-tables-g<NG>-p..-T..-u..-d<ND>". */
int jrb_binary_tables_filename(char *out, size_t cap, int dim_g, int dim_p, int dim_t, int dim_u, int dim_d);
size_t jrb_binary_tables_size(int dim_g, int dim_p, int dim_t, int dim_u, int dim_d); /* header + sizeof(tbl_t) */
int jrb_tables_read_binary(const char *filename, int ng, const char *const *emitters, int nd, const double *nu,
                           jrb_host_tables **out);
int jrb_tables_write_binary(const char *filename, const jrb_tbl_view *tbl, int ng, const char *const *emitters, int nd,
                            const double *nu, int dim_g, int dim_p, int dim_t, int dim_u, int dim_d);

/* select kernel: -1 auto, 0 force generic, 1 force fast (fails if tables do not allow it) */
int jrb_set_kernel_variant(jrb_context *ctx, int variant);

/* Optional epilogue: field-of-view convolution.  With n > 0 every later run returns what the reference's
 * formod(ctl,atm,obs); formod_fov(ctl,obs); (src/jurassic.c:214-258) returns: rad/tau of a ray become the w-weighted mean
 * of the pencil-beam values interpolated in view-point altitude at vpz + dz[i] over the rays of the same package and time
 * within +-NFOV(5) ray indices.  dz/w are the two columns of the ctl->fov shape file (n <= NSHAPE = 2048).  A ray with
 * fewer than two such neighbours makes the run fail ("Cannot apply FOV convolution!").  n = 0 switches it off. */
int jrb_set_fov(jrb_context *ctx, int n, const double *dz, const double *w);

/* complete forward model for npk packages, host buffers in and out (H2D + kernels + D2H, synchronous) */
int jrb_formod_batch(jrb_context *ctx, int npk, const jrb_atm_view *atm, const jrb_obs_view *obs);

/* the same in three phases */
int jrb_stage(jrb_context *ctx, int npk, const jrb_atm_view *atm, const jrb_obs_view *obs);
int jrb_run_staged(jrb_context *ctx);
int jrb_fetch_staged(jrb_context *ctx, int npk, const jrb_obs_view *obs);

/* device pointers of the staged results (rad/tau compact [n_rays][nd]); for device-side gathers */
int jrb_staged_results(jrb_context *ctx, double **rad_dev, double **tau_dev, long long *n_rays, int *nd);
/* debug/test: copy the LOS of staged ray r (after jrb_run_staged) into out[np][rec]; returns np via *np_out */
int jrb_debug_los(jrb_context *ctx, long long ray, double *out, int max_doubles, int *np_out, int *rec_doubles,
                  double *tsurf_out);

/* the compact device results of the last run as one block: rad[R][nd], tau[R][nd], tpz[R], tplon[R], tplat[R] */
int jrb_staged_results_blob(jrb_context *ctx, void **dev, size_t *bytes, long long *n_rays, int *nd);

int jrb_get_stats(const jrb_context *ctx, jrb_stats *out);
const char *jrb_version(void);

/* ---- devices and lanes behind one handle (jrb_group.cu) -----------------------------------------------------------------
 * Replaces the lane hand-out and the per-device loop of the reference (src/GPUdrivers.cu:275-358; device = ctl->MPIlocalrank
 * :288, `omp parallel for num_threads(numDevices)` :351-357).  A group owns ndev devices x nlanes contexts:
 *   - concurrent callers (host threads calling with one package each) get a lane each, their kernels overlap on the device;
 *   - a large batch is cut into contiguous package slices, one per device, processed by one host thread per device; every
 *     device stores its results straight into the caller's obs rows;
 *   - the packed tables are broadcast once with NCCL (ncclCommInitAll + ncclBroadcast; NCCL is loaded at run time and only
 *     needed when ndev > 1 or in rank style). */
typedef struct jrb_group jrb_group;
typedef struct {
  int ndev, nlanes, n_slices; /* n_slices: devices the last batch was cut over */
  int nccl_nranks;            /* ranks of the NCCL communicator the tables were broadcast over (0: one device, no NCCL) */
  int dist_rank, dist_nranks; /* rank style (jrb_group_dist_init), else -1 / 0 */
  long long table_bytes, gather_bytes;
  float ms_tables, ms_last_call, ms_gather, ms_gather_scatter;
} jrb_group_stats;

/* ndev <= 0: all visible devices; devices == NULL: ordinals 0..ndev-1; nlanes in 1..8 */
int jrb_group_create(jrb_group **out, int ndev, const int *devices, int nlanes);
void jrb_group_destroy(jrb_group *g);
const char *jrb_group_last_error(const jrb_group *g);
int jrb_group_size(const jrb_group *g, int *ndev, int *nlanes);
jrb_context *jrb_group_context(jrb_group *g, int dev, int lane); /* for tests / introspection */
int jrb_group_set_control(jrb_group *g, const jrb_ctl_view *ctl);
int jrb_group_set_fov(jrb_group *g, int n, const double *dz, const double *w); /* shape used by calls with use_fov != 0 */
/* pack once, upload to the first device, ncclBroadcast to the others, share among the lanes (ctl may be NULL: keep) */
int jrb_group_set_tables(jrb_group *g, const jrb_ctl_view *ctl, const jrb_tbl_view *tbl);
/* jrb_formod_batch over the group; ctl == NULL: the control set last; thread safe */
int jrb_group_formod_batch(jrb_group *g, const jrb_ctl_view *ctl, int npk, const jrb_atm_view *atm, const jrb_obs_view *obs,
                           int use_fov);
int jrb_group_get_stats(jrb_group *g, jrb_group_stats *out);

/* Rank style: one process per GPU (MPI / torchrun launchers).  The launcher distributes the 128-byte id made by
 * jrb_dist_unique_id on one rank; each process has a group of ONE device.
 *   jrb_group_dist_set_tables : the root passes the tables, everybody receives the packed blob by ncclBroadcast;
 *   jrb_group_dist_gather     : every rank sends the compact device results of its last batch to the root, which lands
 *                               them in the obs views of all packages (counts[r] = packages of rank r, rank order). */
int jrb_dist_unique_id(void *id, size_t capacity /* >= 128 */);
int jrb_group_dist_init(jrb_group *g, int rank, int nranks, const void *id, size_t id_bytes);
int jrb_group_dist_set_tables(jrb_group *g, const jrb_ctl_view *ctl, const jrb_tbl_view *tbl /* root only */, int root);
int jrb_group_dist_gather(jrb_group *g, int root, const int *counts, int npk_all, const jrb_obs_view *obs_all /* root only */);

/* Node-shared page-locked memory (POSIX shm + jrb_host_register): obs_t blocks placed here by all ranks of a node are
 * written by every rank's GPU over its own PCIe link while its kernels run, so the root sees all results without any
 * gather step.  create != 0 makes the segment, else it is attached. */
int jrb_shared_alloc(const char *name, size_t bytes, int create, void **out);
int jrb_shared_free(const char *name, void *ptr, size_t bytes, int unlink_it);

#ifdef __cplusplus
}
#endif
#endif
