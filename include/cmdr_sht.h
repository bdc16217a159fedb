/* cmdr_sht.h -- C ABI of the B200-native SHT engine for Commander3.
 *
 * Part 1 is the drop-in boundary: exactly the symbols that Commander3's
 * ISO_C_BINDING shim `commander3/src/sharp.f90` binds.  Linking
 * Commander against libcmdr_sht.so instead of libsharp2 needs no Fortran change.
 * Each entry point cites the reference interface it replaces.
 *
 * Part 2 is additive (device-resident pointers, fused IQU calls, multi-GPU
 * bootstrap, counters).  Nothing in part 1 depends on part 2 being called.
 *
 * All data is FP64.  There is no CPU fallback: every execute call runs CUDA
 * kernels on the current device and aborts with a message if that is impossible
 * (libsharp2 has no error channel either; it aborts on internal assertions).
 *
 * Threading: calls on one device must be serialised -- one calling thread and one
 * in-flight stream per process (Commander calls sharp_execute from the single main
 * thread of each MPI rank).  The work buffers (phases, FFT work area,
 * staging, event pool) are per process and device, not per stream: two transforms
 * enqueued concurrently on different streams of one device would share them.  One
 * process drives one GPU.
 */
#ifndef CMDR_SHT_H
#define CMDR_SHT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque handles (commander3/src/sharp.f90:22-30 keeps them as type(c_ptr)). */
typedef struct sharp_alm_info sharp_alm_info;
typedef struct sharp_geom_info sharp_geom_info;

/* Job types, commander3/src/sharp.f90:8-14 */
enum { SHARP_YtW = 0, SHARP_Y = 1, SHARP_Yt = 2, SHARP_WY = 3, SHARP_ALM2MAP_DERIV1 = 4 };
/* alm_info flags, commander3/src/sharp.f90:5 and libsharp2 sharp_almhelpers */
enum { SHARP_PACKED = 1, SHARP_REAL_HARMONICS = 1 << 6 };
/* job flags, commander3/src/sharp.f90:17-20 */
enum { SHARP_DP = 1 << 4, SHARP_ADD = 1 << 5, SHARP_NO_FFT = 1 << 7 };

/* ------------------------------------------------------------------ part 1 */

/* commander3/src/sharp.f90:35-42 (declared by the reference, never called).
 * Complex alm, entry (l,m) at alm[2*(mvstart[im] + stride*l) + {0,1}].  Only
 * stride==1 is supported. */
void sharp_make_general_alm_info(int lmax, int nm, int stride, const int *mval,
                                 const ptrdiff_t *mvstart, int flags,
                                 sharp_alm_info **alm_info);

/* commander3/src/sharp.f90:44-50, called at :128,:131 from
 * commander3/src/comm_map_mod.f90:264.  m-major real-packed layout: for each
 * m in `ms` (NULL -> 0..nm-1) in order, m=0 stores lmax+1 reals, m>0 stores
 * (lmax+1-m) pairs (alm[+m], alm[-m]) with a_lm = (alm[+m] + i alm[-m])/sqrt2. */
void sharp_make_mmajor_real_packed_alm_info(int lmax, int stride, int nm, const int *ms,
                                            sharp_alm_info **alm_info);

/* commander3/src/sharp.f90:52-56 */
ptrdiff_t sharp_alm_count(const sharp_alm_info *self);

/* commander3/src/sharp.f90:58-61 */
void sharp_destroy_alm_info(sharp_alm_info *info);

/* commander3/src/sharp.f90:64-71, called at :158,:162 from
 * commander3/src/comm_map_mod.f90:266-282.  `rings` are 1-based HEALPix ring
 * numbers (NULL -> 1..nrings), stored in the map in the given order with
 * stride `stride` (only 1 supported); `weight` has 2*nside entries indexed by
 * northern-ring number - 1 (NULL -> 1.0); ring weight = 4 pi / npix * weight. */
void sharp_make_subset_healpix_geom_info(int nside, int stride, int nrings, const int *rings,
                                         const double *weight, sharp_geom_info **geom_info);

/* commander3/src/sharp.f90:73-76 */
void sharp_destroy_geom_info(sharp_geom_info *info);

/* commander3/src/sharp.f90:78-82 */
ptrdiff_t sharp_map_size(const sharp_geom_info *info);

/* commander3/src/sharp.f90:86-94 (called at :234).  `alm` and `map` are arrays of
 * pointers, one per component (spin 0: 1, spin>0: 2), each to a contiguous
 * double array.  The pointers may be host memory (pageable or pinned; the
 * Fortran case) or device memory (detected with cudaPointerGetAttributes).
 * spin 0, or any spin 1..32 (two components; conviqt calls spin j, comm_conviqt_mod.f90:254).  `time` (seconds) and `opcnt` (nominal flops) are
 * optional outputs. */
void sharp_execute(int type, int spin, void *alm, void *map, const sharp_geom_info *geom_info,
                   const sharp_alm_info *alm_info, int flags, double *time,
                   unsigned long long *opcnt);

/* commander3/src/sharp.f90:96-104 (called at :227).  `comm` is an MPI_Fint.  This
 * library does not link MPI: the communicator is looked up in the table filled by
 * cmdr_sht_comm_register() (part 2) -- EVERY communicator that reaches this call must
 * have been registered (comm_map passes info%comm, comm_conviqt its own self%comm,
 * commander3/src/comm_conviqt_mod.f90:234-239).  An unknown communicator is accepted
 * only when the handles cover the whole sphere (all 4 nside - 1 rings and all m: a
 * group of one rank); otherwise the call aborts with a message instead of silently
 * transforming local m's onto local rings.  `time` receives the wall seconds of the
 * call in both modes.  See INTEGRATION.md for the MPI bootstrap stub. */
void sharp_execute_mpi_fortran(int comm, int type, int spin, void *alm, void *map,
                               const sharp_geom_info *geom_info, const sharp_alm_info *alm_info,
                               int flags, double *time, unsigned long long *opcnt);

/* ------------------------------------------------------------------ part 2 */

int cmdr_sht_version(void);

/* Device-resident variant of sharp_execute: all pointers are device pointers on
 * the current device, work is enqueued on `stream` (a cudaStream_t; NULL = the
 * legacy default stream) and the call returns without synchronising. */
void cmdr_sht_execute_dev(int type, int spin, double *const *alm, double *const *map,
                          const sharp_geom_info *geom_info, const sharp_alm_info *alm_info,
                          int flags, void *stream);

/* Fused IQU transform = what comm_map%Y / Yt / YtW / WY do with pol=.true.,
 * nmaps=3 (commander3/src/comm_map_mod.f90:437-564): spin-0 on component 0 with
 * geom_T and spin-2 on components 1,2 with geom_P, sharing one phase exchange.
 * Pointers may be host or device as for sharp_execute; stream as above (host
 * pointers force a synchronisation before returning). */
void cmdr_sht_execute_iqu(int type, double *const *alm3, double *const *map3,
                          const sharp_geom_info *geom_T, const sharp_geom_info *geom_P,
                          const sharp_alm_info *alm_info, int flags, void *stream);

/* nbatch independent IQU transforms sharing the handles: the per-band loop around comm_map%Y /
 * %Yt / %YtW in cr_matmulA and the map updates (commander3/src/comm_cr_mod.f90:880-918; BASELINE
 * config 5: 30 bands per Gibbs step).  alm3 / map3 hold 3*nbatch column pointers, band-major.
 * With host buffers the bands are software-pipelined (upload of band b+1 and download of band b-1
 * beside the kernels of band b; use pinned buffers for the full overlap); device pointers simply
 * loop.  Single GPU. */
void cmdr_sht_execute_iqu_batch(int type, int nbatch, double *const *alm3, double *const *map3,
                                const sharp_geom_info *geom_T, const sharp_geom_info *geom_P,
                                const sharp_alm_info *alm_info, int flags, void *stream);

/* ---- multi-GPU (one process per GPU), libsharp-MPI layout: m's and ring pairs
 * round-robin per rank (commander3/src/comm_map_mod.f90:197-261), one
 * all-to-all of phases per transform. */

/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host program
 * (MPI_Bcast in Commander, torch.distributed in this repo's harness). */
void cmdr_sht_get_unique_id(void *id128);

/* Collective.  Binds this process (current CUDA device) as `rank` of `nranks`
 * and registers the group under `comm` (the MPI_Fint the Fortran side passes, or
 * any integer chosen by the harness).  Returns 0 on success. */
int cmdr_sht_comm_register(int comm, int rank, int nranks, const void *id128);
void cmdr_sht_comm_destroy(int comm);

/* The communicator-less handle of the part-2 entry points below that take a `comm`: one GPU,
 * no collective.  Any other value must be a registered communicator, or the handles must
 * cover the whole sphere (same rule as sharp_execute_mpi_fortran). */
#define CMDR_SHT_COMM_SELF (-1)

/* Collective (every rank, same value).  How the m <-> ring transpose of later transforms on `comm`
 * travels: 1 = fused into the Legendre kernels as peer stores / loads over NVLink (default where
 * CUDA IPC works and nranks <= 8), 0 = NCCL all-to-all between a send and a receive buffer,
 * -1 = back to the automatic choice ($CMDR_SHT_P2P).  For tests and benchmarks that compare the two. */
void cmdr_sht_comm_set_exchange(int comm, int mode);

/* Collective distributed transform on a registered comm.  geometry/alm handles
 * describe this rank's local rings and m's exactly as comm_mapinfo builds them.
 * Every rank must pass the same nside/lmax and the round-robin layout. */
void cmdr_sht_execute_dist(int comm, int type, int spin, void *alm, void *map,
                           const sharp_geom_info *geom_info, const sharp_alm_info *alm_info,
                           int flags, void *stream);
void cmdr_sht_execute_iqu_dist(int comm, int type, double *const *alm3, double *const *map3,
                               const sharp_geom_info *geom_T, const sharp_geom_info *geom_P,
                               const sharp_alm_info *alm_info, int flags, void *stream);

/* Pixel-space mixing, alm <- YtW( F .* Y(alm) ): the spatially varying branch of
 * evalDiffuseBand / projectDiffuseBand (commander3/src/comm_diffuse_comp_mod.f90:2078-2080
 * and :2148-2150 -- `call m%Y(); m%map = m%map * F%map; call m%YtW()`), which the reference
 * runs as four sharp_execute calls with a host-side multiply in between.  Here the map stays
 * on the device; only alm (and F, if it lives on the host) cross PCIe.
 * nmaps = 1 (T, geom_P ignored) or 3 (I,Q,U); alm and F are arrays of nmaps pointers to
 * n_alm / n_pix doubles, host or device.  comm: a registered communicator (collective call,
 * layout as for sharp_execute_mpi_fortran) or CMDR_SHT_COMM_SELF for one GPU. */
void cmdr_sht_mix(int comm, int nmaps, double *const *alm, const double *const *F,
                  const sharp_geom_info *geom_T, const sharp_geom_info *geom_P,
                  const sharp_alm_info *alm_info, void *stream);

/* Diagonal of the inverse noise covariance in harmonic space, the input of the diagonal CG
 * preconditioner: replaces compute_invN_lm (commander3/src/comm_N_mod.f90:127-197) after its
 * YtW_scalar and mpi_bcast.  a_l0[c] -> the lmax+1 m=0 coefficients of YtW(N^-1 map) of component c
 * (host memory, identical on every rank); out[c] <- N_lm in the local real-packed order of alm_info
 * (host or device), both entries of an m>0 pair equal (:176-181).  npix = 12 nside^2.  The 3j sum
 * of the reference is evaluated as an exact Gauss-Legendre quadrature of lambda_lm^2 times the
 * azimuthally averaged profile (see commander_b200/csrc/invn.cu); no collective. */
void cmdr_sht_invn_diag(int nmaps, const double *const *a_l0, double npix,
                        const sharp_alm_info *alm_info, double *const *out, void *stream);

/* The convolution cube of conviqt: replaces comm_conviqt%precompute_sky together with get_alms
 * (commander3/src/comm_conviqt_mod.f90:207-292 and :294-357), which the reference runs as bmax+1
 * sharp_execute calls through host arrays (:247-258) plus one FFTW c2r of length 2*bmax per pixel
 * (:267-282).  sky_alm: nmaps (1..3) pointers to the local real-packed a_lm of alm_info; beam: the
 * shared beam table alm_beam%a(nmaps, (lmax+1)(lmax+2)/2) exactly as the reference holds it (:94-115) --
 * single-precision complex, entry (l,m) at complex index (l(l+1)/2 + m)*nmaps + c; cube: 2*bmax rows
 * of n_pix local pixels (c%a(pix, psi) in Fortran order), float when cube_f64 == 0 (the reference's
 * real(sp) cube, :281) or double.  All three may be host or device memory.  1 <= bmax <= 32.
 * comm: a registered communicator (collective) or CMDR_SHT_COMM_SELF for one GPU. */
void cmdr_sht_conviqt_cube(int comm, int nmaps, int bmax, const double *const *sky_alm, const float *beam,
                           const sharp_geom_info *geom_T, const sharp_alm_info *alm_info, void *cube,
                           int cube_f64, void *stream);

/* NCCL sum-allreduce of n doubles (device pointer) on the comm: the collective
 * behind mpi_dot_product (commander3/src/comm_utils.f90:599-614).  No-op for CMDR_SHT_COMM_SELF
 * or a registered group of one; an unregistered communicator aborts (a silent no-op would leave
 * every rank with its partial sum). */
void cmdr_sht_allreduce_sum(int comm, double *dev_buf, int n, void *stream);

/* ---- constrained-realisation CG, device resident (commander_b200/csrc/cr.cu) ----------------------------
 * One diffuse signal component with a diagonal prior (or none) seen through nbands bands that share the
 * comm_mapinfo the handles come from: the CMB amplitude solve of comm_cr_mod.  The handle caches N^-1, the beams,
 * sqrt(S) and all work vectors on the device; between cmdr_cr_setup and cmdr_cr_destroy no vector crosses PCIe
 * except b on the way in and x on the way out of cmdr_cr_solve.
 *
 * Vectors are passed as sharp_execute passes a_lm: an array of nmaps column pointers, each to n_alm doubles in
 * the local real-packed order of alm_info, host or device memory.  nmaps = 1 (T) or 3 (T, E, B).
 * comm: the registered communicator of the comm_mapinfo (collective calls, m-distributed vectors, dot products
 * all-reduced) or CMDR_SHT_COMM_SELF. */
typedef struct cmdr_cr_system cmdr_cr_system;

/* invN : nbands * nmaps pointers (band-major) to N^-1 per local pixel = siN^2 x mask
 *        (commander3/src/comm_N_rms_mod.f90:264-273), host or device; copied.
 * b_l  : nbands * nmaps pointers to lmax+1 doubles (host): the beam of commander3/src/comm_B_bl_mod.f90:108-127
 *        times mb_eff and the mixing scalar F_mean of the band (comm_diffuse_comp_mod.f90:2077-2080).
 * sqrtS: nmaps pointers to lmax+1 doubles (host), sqrt(C_l) of the diagonal prior
 *        (commander3/src/comm_Cl_mod.f90:588-637), or NULL for a component without prior (P = 0, sqrt(S) = 1).
 * precond: 0 = none (identity), 1 = diagonal (comm_diffuse_comp_mod.f90:2186-2235 for npre = 1, with
 *        N^-1_{lm,lm} of every band from compute_invN_lm, comm_N_mod.f90:127-197, evaluated on the device). */
cmdr_cr_system *cmdr_cr_setup(int comm, int nbands, int nmaps, const sharp_geom_info *geom_T,
                              const sharp_geom_info *geom_P, const sharp_alm_info *alm_info,
                              const double *const *invN, const double *const *b_l,
                              const double *const *sqrtS, int precond);
void cmdr_cr_destroy(cmdr_cr_system *sys);

/* y = A x, cr_matmulA (commander3/src/comm_cr_mod.f90:771-1024):
 *   A = P + sqrt(S) sum_nu B^t Y^t N^-1 Y B sqrt(S),  P = 1 with a prior, 0 without. */
void cmdr_cr_matmulA(cmdr_cr_system *sys, const double *const *x, double *const *y, void *stream);

/* z = M^-1 r, cr_invM (commander3/src/comm_cr_mod.f90:1026-1077) with the diagonal preconditioner. */
void cmdr_cr_invM(cmdr_cr_system *sys, const double *const *r, double *const *z, void *stream);
void cmdr_cr_set_precond_diag(cmdr_cr_system *sys, const double *const *Minv);   /* NULL: identity */
void cmdr_cr_get_precond_diag(const cmdr_cr_system *sys, double *const *out);

/* b = sqrt(S) sum_nu B^t Y^t (N^-1 d_nu + N^-1/2 eta_nu) + eta_0, cr_computeRHS (commander3/src/comm_cr_mod.f90:542-769).
 * data, eta_pix: nbands * nmaps pointers to n_pix doubles (eta_pix may be NULL: mean-field term only);
 * eta_alm: nmaps pointers to n_alm doubles or NULL. */
void cmdr_cr_compute_rhs(cmdr_cr_system *sys, const double *const *data, const double *const *eta_pix,
                         const double *const *eta_alm, double *const *b, void *stream);

/* solve_cr_eqn_by_CG (commander3/src/comm_cr_mod.f90:201-348): same update order, same convergence test
 * (stop when r^t M^-1 r < cg_tol * b^t M^-1 b, checked every cg_check_conv_freq iterations, not before
 * cg_miniter; conv_crit 0 = 'residual', 1 = 'fixed_iter' = always maxiter iterations, the shipped setting).
 * x: initial guess when x0_given != 0 (else zero), solution on return.  hist: optional host array of
 * maxiter + 1 doubles, r^t M^-1 r before the loop and after every iteration.  Returns the iterations done.
 * With 'fixed_iter' the host does not wait for the device anywhere inside the loop. */
int cmdr_cr_solve(cmdr_cr_system *sys, const double *const *b, double *const *x, int x0_given, int maxiter,
                  double cg_tol, int conv_crit, int cg_miniter, int cg_check_conv_freq, double *hist,
                  void *stream);
unsigned long long cmdr_cr_matmul_count(const cmdr_cr_system *sys);

/* ---- introspection for benchmarks/tests */

/* Number of CUDA kernels this library has launched since load (cuFFT execs
 * count as one each). */
unsigned long long cmdr_sht_launch_count(void);

/* Event timing of the Legendre kernels of the most recent execute on this
 * thread: returns the number of entries written (<= max).  Each entry is
 * {spin, direction(0 synth,1 analysis), milliseconds}.  Only filled when
 * cmdr_sht_set_profiling(1) was called (adds event records to the stream). */
void cmdr_sht_set_profiling(int on);
int cmdr_sht_last_legendre_ms(double *entries3, int max);

/* Nominal flop count of one transform direction (SURVEY.md 8d convention:
 * 8 flops per (l,m,ring pair) for spin 0, 28 for spin 2, FMA = 2). */
unsigned long long cmdr_sht_nominal_flops(const sharp_geom_info *geom_info,
                                          const sharp_alm_info *alm_info, int spin);

/* FP64 FMA throughput of the current device in TFLOP/s (best of `reps` launches of a
 * register-only DFMA probe, `iters` x 64 FMAs per thread): the roofline denominator for the
 * Legendre kernels, which MEASURED_PEAKS.json does not carry. */
double cmdr_sht_measure_fp64_tflops(int iters, int reps);
/* The same probe with three distinct vector-register operands per DFMA: on B200 the register
 * file feeds one 64-bit operand per cycle per scheduler, so such a DFMA issues every 3 cycles
 * instead of 2 (2/3 of the peak above).  Diagnostic for the roofline discussion in DESIGN.md. */
double cmdr_sht_measure_fp64_tflops_3op(int iters, int reps);

/* Pageable caller arrays are staged through a pinned arena by a pool of copy threads (commander_b200/csrc/hostio.cu;
 * $CMDR_SHT_COPY_THREADS, default: the CPUs the process may run on minus one, at most 16).  Payload GB/s of that
 * pool for one copy of `bytes`: direction 0 = caller -> arena (write-combined when wc != 0), 1 = arena -> caller. */
double cmdr_sht_measure_host_copy(size_t bytes, int direction, int wc, int reps);
int cmdr_sht_host_copy_threads(void);

/* Frees cached device buffers, cuFFT plans and coefficient tables. */
void cmdr_sht_release_caches(void);

#ifdef __cplusplus
}
#endif
#endif /* CMDR_SHT_H */
