/* jurassic_b200_dropin.h -- reference-facing entry points of libjurassic_b200_dropin_nd<ND>_ng<NG>.so.
 *
 * This thin C layer is compiled once per compile-time dimension set (-DND=.. -DNG=.., like the reference itself,
 * src/jurassic.h:137-145) because the layouts of ctl_t / atm_t / obs_t / tbl_t depend on ND and NG.  It only turns
 * the reference's structs into the pointer/stride views of jurassic_b200.h and forwards to the core library.
 *
 *   formod_GPU             replaces  formod_GPU()            src/GPUdrivers.cu:253-262 (declared src/CPUdrivers.c:153-155)
 *                          called by formod()                src/CPUdrivers.c:189-190 when ctl->useGPU != 0
 *   jr_b200_init           replaces  the first-call block    src/GPUdrivers.cu:275-329 (+ get_tbl_on_GPU :78-93)
 *   jr_b200_formod_batch   new: many (atm_t, obs_t) packages per call -- a single obs_t holds at most NR = 1088 rays
 *                          (src/jurassic.h:151), far too few to occupy a B200
 *   jr_b200_kernel         replaces  kernel()                src/jurassic.c:812-857 (finite-difference Jacobian; the caller of
 *                          formod in retrievals): all perturbed forward models of a Jacobian are one device batch
 *   jr_b200_formod_fov_batch  replaces  formod() + formod_fov()  src/jurassic.c:214-258 (FOV convolution as a device epilogue)
 *   jr_b200_finalize       new: releases what the reference never frees (src/GPUdrivers.cu:309)
 *
 * Error behaviour mirrors the reference: fatal conditions print a message and exit(EXIT_FAILURE) (ERRMSG,
 * src/jurassic.h:68-72); ctl->checkmode != 0 prints a note and returns without computing (src/GPUdrivers.cu:337).
 * The struct types are the reference's own (include its jurassic.h before this header) or the layout-identical
 * mirrors of jurassic-gpu_b200/csrc/jr_structs.h.
 */
#ifndef JURASSIC_B200_DROPIN_H
#define JURASSIC_B200_DROPIN_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Drop-in for the reference symbol.  Tables: if jr_b200_init() was not called, they are taken from the
 * reference's get_tbl(ctl) (src/jr_common.h:60-78) when that symbol is linked in, else the call is fatal.
 * ctl->ip = 2, 3 (2-D / 3-D atmosphere, ctl->cz, ctl->cx) selects intpol_atm_2d / _3d (src/jurassic.c:685-804) inside the
 * tracer, where the reference's own formod_GPU stops at an assert (src/jr_common.h:573,581). */
void formod_GPU(ctl_t const *ctl, atm_t *atm, obs_t *obs);

/* Explicit initialisation with a caller-owned table (kept only until this call returns).
 * device < 0 selects ctl->MPIlocalrank like the reference (src/GPUdrivers.cu:288). Returns 0 or exits. */
int jr_b200_init(ctl_t const *ctl, tbl_t const *tbl, int device);

/* The same with tables read natively from "<ctl->tblbase>_<nu>_<GAS>.tab" / ".filt" (parsing rules of init_tbl,
 * src/jurassic.c:311-416, 612-667) -- no tbl_t is ever allocated on the host. */
int jr_b200_init_from_files(ctl_t const *ctl, int device);

/* formod_GPU semantics for each of npackages (atm[i], obs[i]) pairs, one device batch. */
void jr_b200_formod_batch(ctl_t const *ctl, atm_t *const atm[], obs_t *const obs[], int npackages);

/* Finite-difference Jacobian, what the reference's kernel() computes (src/jurassic.c:812-857) with the 1+n forward models
 * run as device batches.  k is row-major m x n (gsl_matrix data with tda = n): m = finite radiances of obs (obs2y order),
 * n = retrieved state elements (atm2x order, ranges ctl->ret*_zmin/zmax).  obs holds the undisturbed result on return. */
void jr_b200_kernel(ctl_t const *ctl, atm_t *atm, obs_t *obs, double *k, size_t m, size_t n);
/* n (return value) and m for the above */
size_t jr_b200_kernel_dims(ctl_t const *ctl, atm_t *atm, obs_t const *obs, size_t *m_out);

/* Per package what the reference's  formod(ctl, atm, obs); formod_fov(ctl, obs);  (src/jurassic.c:214-258) return: the
 * forward model followed by the field-of-view convolution with the shape file ctl->fov ("-": no convolution), the
 * convolution running on the device on the resident results.  formod_GPU itself ignores ctl->fov, like formod(). */
void jr_b200_formod_fov_batch(ctl_t const *ctl, atm_t *const atm[], obs_t *const obs[], int npackages);

/* ---- more than one GPU ---------------------------------------------------------------------------------------------------
 * The reference only pretends (device = ctl->MPIlocalrank, never set, src/GPUdrivers.cu:288; per-device loop over pointers of
 * one device, :344-358).  Two real forms here (SURVEY.md 8b "needed extension", 8e):
 *
 * (1) one process, all GPUs: jr_b200_init_multi() makes one context set and one host thread per device, packs the tables
 *     once and broadcasts them with NCCL (ncclCommInitAll + ncclBroadcast); jr_b200_formod_batch() then cuts every batch
 *     into contiguous package slices, one per device, and each device stores its results straight into the caller's obs_t
 *     rows.  ndevices <= 0: all visible devices.  Returns the number of devices in use.
 *
 * (2) one process per GPU (MPI or torchrun launchers): the launcher distributes the 128-byte id from
 *     jr_b200_dist_unique_id() of one rank; jr_b200_dist_init() joins the NCCL communicator and receives the packed tables
 *     from `root` (the only rank that passes tbl) by ncclBroadcast; every rank then calls jr_b200_formod_batch() on its own
 *     contiguous slice of packages; jr_b200_dist_gather() finally lands all radiances, transmittances and tangent points in
 *     the root's obs_t array (obs_all: root only, all packages in rank order; counts[r] = packages of rank r). */
int jr_b200_init_multi(ctl_t const *ctl, tbl_t const *tbl, int ndevices);
int jr_b200_dist_unique_id(char id[128]);
int jr_b200_dist_init(ctl_t const *ctl, tbl_t const *tbl, int rank, int nranks, char const id[128], int device, int root);
void jr_b200_dist_gather(obs_t *const obs_all[], int const counts[], int nranks, int root);

/* ---- page-locked caller structs -------------------------------------------------------------------------------------------
 * The reference copies whole atm_t / obs_t structs from pageable memory on every call (src/GPUdrivers.cu:222-223,244).
 * Callers that keep their packages alive across calls (retrieval iterations, orbit loops) can page-lock them once: batches
 * whose packages are ALL pinned run in "direct" mode -- the device gathers the populated prefixes of the inputs straight
 * from the structs and stores every ray's results straight into obs->rad / obs->tau / obs->tp* while the kernels run, with
 * no host-side packing, copy phase or scatter.  The structs must stay allocated until jr_b200_unpin_all().
 * jr_b200_shared_alloc(): the same for memory shared by the processes of a node (POSIX shm): obs_t arrays placed there are
 * filled by every rank's GPU over its own PCIe link, so the root rank sees all results without a gather step. */
int jr_b200_pin_packages(atm_t *const atm[], obs_t *const obs[], int npackages);
void jr_b200_unpin_all(void);
void *jr_b200_shared_alloc(char const *name, size_t bytes, int create);
void jr_b200_shared_free(char const *name, void *ptr, size_t bytes, int unlink_it);

void jr_b200_finalize(void);

/* introspection for tests: dimension macros this layer was compiled with: {ND, NG, NP, NR, NW, NLOS, TBLNP, TBLNT,
 * TBLNU, TBLNS, LEN} and sizeof(ctl_t, atm_t, obs_t, tbl_t) */
void jr_b200_dims(int dims[11], long long sizes[4]);
/* the core context behind the drop-in (struct jrb_context*, see jurassic_b200.h), NULL before initialisation */
void *jr_b200_core_context(void);
/* the device/lane group behind the drop-in (struct jrb_group*), NULL before initialisation */
void *jr_b200_core_group(void);

#ifdef __cplusplus
}
#endif
#endif
