/* mithra_gpu.h -- C ABI of the B200-native MITHRA time-march (libmithra_gpu.so).
 *
 * This is the drop-in boundary for the hot path named by BASELINE.json: the Lorentz-boosted-frame
 * FDTD/PIC loop of MITHRA 2.0.  The reference has no FFI; its boundary is the C++ class surface
 * Solver / FdTd / FdTdSC (reference src/solver.h:25-178, src/fdtd.h:18-66, src/fdtdSC.h:18-66).  Every entry
 * point below replaces one of those methods (cited per function); the host-side classes in
 * mithra_b200/host/ keep the reference's names and call straight into this ABI.
 *
 * Conventions
 *   - plain C, opaque handle, no C++/torch types; all pointers are HOST pointers unless named d_*;
 *   - every function returns 0 on success, non-zero on failure; mithra_gpu_last_error() gives the text.
 *     (The reference's convention is "print and exit(1)"; the host classes do exactly that on non-zero.)
 *   - there is NO CPU fallback: without a CUDA device mithra_gpu_create fails.
 *   - host-side array layouts are the reference's own:
 *       vector field  : double[np*N0*N1][3], node m = N1*N0*k + N1*i + j      (solver.cpp:674, fieldvector.h:25)
 *       scalar field  : double[np*N0*N1]
 *       E / B         : float [np*N0*N1][3]                                    (solver.h:244-245)
 *       particles     : double[n][11] = { q, rnp[3], rnm[3], gb[3], e }        (stdinclude.h:130-144)
 *     The device layout (component-planar, padded rows) is private to the library, see DESIGN.md.
 */
#ifndef MITHRA_GPU_H_
#define MITHRA_GPU_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MITHRA_GPU_ABI_VERSION      4
#define MITHRA_MAX_UNDULATORS       16
#define MITHRA_MAX_EXTFIELDS        8
#define MITHRA_MAX_POWER_PLANES     256
#define MITHRA_MAX_POWER_LAMBDAS    64
#define MITHRA_MAX_SCREENS          64

/* enum values are the reference's (stdinclude.h:18-40) */
enum { MITHRA_SOLVER_FD = 0, MITHRA_SOLVER_NSFD = 1 };
enum { MITHRA_UNDULATOR_STATIC = 0, MITHRA_UNDULATOR_OPTICAL = 1 };
enum { MITHRA_SIGNAL_NEUMANN = 0, MITHRA_SIGNAL_GAUSSIAN = 1, MITHRA_SIGNAL_SECANT = 2, MITHRA_SIGNAL_FLATTOP = 3, MITHRA_SIGNAL_INVGAUSSIAN = 4 };
enum { MITHRA_BEAM_PLANEWAVE = 0, MITHRA_BEAM_PLANEWAVETRUNCATED = 1, MITHRA_BEAM_GAUSSIAN = 2, MITHRA_BEAM_SUPERGAUSSIAN = 3,
       MITHRA_BEAM_STANDINGPLANEWAVE = 4, MITHRA_BEAM_STANDINGPLANEWAVETRUNCATED = 5, MITHRA_BEAM_STANDINGGAUSSIAN = 6,
       MITHRA_BEAM_STANDINGSUPERGAUSSIAN = 7 };

/* Signal (classes.h Signal; classes.cpp:534-575), already converted to solver units. */
typedef struct MithraSignal
{
  int    type;
  double t0, s, f0, cep;
  int    nR;
  double sigma_inv_g[2];          /* inverse-gaussian only (classes.cpp:559-572)                     */
} MithraSignal;

/* One analytic beam: Seed, optical Undulator or ExtField share this shape (classes.h:208-262, 303-352, 383-427). */
typedef struct MithraBeam
{
  int          seed_type;
  double       position[3], direction[3], polarization[3];
  double       amplitude;
  double       radius[2];
  double       l;                 /* wavelength                                                      */
  double       zR[2];             /* Rayleigh lengths                                                */
  int          order[2];
  MithraSignal signal;
} MithraBeam;

/* Undulator module (classes.h:265-352), after Solver::setSimulationParameters (sorted, shifted, boosted). */
typedef struct MithraUndulator
{
  int        type;                /* MITHRA_UNDULATOR_*                                              */
  double     k, lu, rb, theta;
  double     length;              /* number of periods (unsigned in the reference)                   */
  double     dist;                /* bunch-head to entrance distance; [0].dist gates the e flag      */
  MithraBeam beam;                /* used when type == OPTICAL                                       */
} MithraUndulator;

/* Radiation-power sampling group (one FEL-OUTPUT/radiation-power block; radiation.cpp:18-121). */
typedef struct MithraPower
{
  int    enabled;
  int    N;                                     /* planes                                            */
  double z[MITHRA_MAX_POWER_PLANES];            /* boosted plane positions                           */
  int    Nl;                                    /* wavelengths                                       */
  double w[MITHRA_MAX_POWER_LAMBDAS];           /* angular frequencies 2 pi / dt_lambda              */
  int    Nf;                                    /* DFT window length                                 */
  double pc;                                    /* power prefactor                                   */
} MithraPower;

/* Screens (solver.cpp:2145-2257). */
/* power-visualization group (FreeElectronLaser::vtkPower_, Solver::initializePowerVisualize radiation.cpp:238-318):
 * one plane, one harmonic, the radiated power PER PIXEL of the plane                                  */
typedef struct MithraPowerMap
{
  int    enabled;
  int    Nf;                                    /* DFT window length                                 */
  double z;                                     /* boosted plane position                            */
  double w;                                     /* angular frequency 2 pi / dt_lambda                */
  double pc;                                    /* power prefactor                                   */
} MithraPowerMap;

typedef struct MithraScreens
{
  int    enabled;
  int    N;
  double pos[MITHRA_MAX_SCREENS];               /* lab-frame positions, sorted                       */
} MithraScreens;

/* Everything Solver::initialize() derives on the host (solver.cpp:547-842, 1050-1059) and the device needs. */
typedef struct MithraGpuParams
{
  int    abi_version;

  /* mesh and slab (solver.cpp:601-689) */
  int    N0, N1, N2;              /* global node counts                                              */
  int    np, k0;                  /* local planes and index of the first one                         */
  int    rank, size;              /* slab index / number of slabs                                    */
  double dx, dy, dz, dt;
  double xmin, xmax, ymin, ymax, zmin, zmax;
  double zp[2];                   /* ownership interval [zp0, zp1)                                   */
  double Lz;                      /* mesh_.meshLength_[2] (period of the z wrap, solver.cpp:1440)    */

  /* solver switches (classes.h Mesh) */
  int    solver;                  /* MITHRA_SOLVER_*                                                 */
  int    space_charge;
  int    truncation_order;

  /* update coefficients (solver.cpp:729-824) */
  double a[6], alpha, beta_nsfd;
  double bB[5], cB[5], dB[5], eE[5], fE[5], gE[5], hC[17];

  /* frame (solver.cpp:168-176, 322) and units (solver.cpp:59-61) */
  double c0, gamma, beta, dt_shift;

  /* bunch update (solver.cpp:268-275, 1053-1059) */
  double dt_bunch;                /* bunch_.timeStep_                                                */
  int    n_update_bunch;          /* nUpdateBunch_                                                   */
  double r1, r2, dtb;

  /* analytic fields seen by the particles */
  int             n_undulators;
  MithraUndulator undulator[MITHRA_MAX_UNDULATORS];
  int             n_ext_fields;
  MithraBeam      ext_field[MITHRA_MAX_EXTFIELDS];

  /* seed injected through the TF/SF shell (fdtd.cpp:307-373); amplitude 0 disables it */
  int        seed_enabled;
  MithraBeam seed;

  /* diagnostics */
  MithraPower   power;
  MithraScreens screens;

  /* capacity hints (0 = library default) */
  size_t max_particles;           /* capacity of the particle arrays                                 */
  size_t max_screen_records;      /* capacity of the per-screen record buffers                       */
  int    device;                  /* CUDA device ordinal, -1 = current                               */
  int    sort_interval;           /* field steps between two counting sorts of the bunch by cell: > 0 as given,
				     < 0 never, 0 = library default (16 for bunches of >= 4096 particles)     */
  MithraPowerMap power_map;       /* power-visualization (radiation.cpp:238-450)                      */
} MithraGpuParams;

typedef struct MithraGpu MithraGpu;

/* Counters of work done, used by bench.py. */
typedef struct MithraGpuCounters
{
  unsigned long long field_steps;
  unsigned long long cell_updates;        /* nodes advanced by fieldUpdate                            */
  unsigned long long particle_pushes;     /* particle sub-steps                                       */
  unsigned long long kernel_launches;     /* kernels of this library launched                         */
} MithraGpuCounters;

const char* mithra_gpu_last_error (void);
int         mithra_gpu_abi_version (void);
int         mithra_gpu_device_count (void);

/* Solver::Solver + Solver::initializeMesh allocation part (solver.cpp:14-62, 646-658). */
int  mithra_gpu_create  (const MithraGpuParams* params, MithraGpu** out);
void mithra_gpu_destroy (MithraGpu* h);

/* Solver state in: an_, anm1_, the current in anp1_ (may be NULL = zero) and, with space charge, fn_, fnm1_,
 * the charge in fnp1_ (solver.cpp:646-658, 828-839).                                                 */
int mithra_gpu_upload_fields (MithraGpu* h, const double* an, const double* anm1, const double* jn,
			      const double* fn, const double* fnm1, const double* rho);
/* Any pointer may be NULL. anp1/fnp1 name what the reference keeps in anp1_/fnp1_ at that moment: the new
 * potential after mithra_gpu_field_update, the deposited current/charge after mithra_gpu_current_update. */
int mithra_gpu_download_fields (MithraGpu* h, double* anp1, double* an, double* anm1,
				double* fnp1, double* fn, double* fnm1);
/* en_, bn_ as float[.][3]; nodes never evaluated in this step hold 0; mask (may be NULL) flags evaluated nodes. */
int mithra_gpu_download_eb (MithraGpu* h, float* en, float* bn, unsigned char* mask);

/* Solver::initializeField, solver.cpp:828-839: an_, anm1_ = Seed::fields(node, time_ / timem1_) inside the total-field
 * box; no-op without a seed. Call after mithra_gpu_set_time.                                          */
int mithra_gpu_seed_initial (MithraGpu* h);

/* ---- the bunch of Solver::initialize() on the device ----------------------------------------------
 * A MithraGpuBunch is a particle list double[n][11] (Charge: q, rnp[3], rnm[3], gb[3], e; stdinclude.h:130-144) in device
 * memory that exists before any slab does: generated, boosted and back-projected on the device, then handed to the
 * slabs without touching the host.  Formulas and operation order are the reference's; only the libm differs (CUDA's log /
 * cos / sin against glibc's, last ulp), so the list equals the reference's to 1e-15, not bit for bit.                 */
typedef struct MithraGpuBunch MithraGpuBunch;

/* The members of BunchInitialize that Bunch::initializeEllipsoid reads (classes.h, classes.cpp:104-298), for ONE bunch
 * position (position_[ia]) with the Halton generator and without shot noise.                                         */
typedef struct MithraBunchEllipsoid
{
  unsigned int number_of_particles;   /* numberOfParticles_, already rounded up to a multiple of four (classes.cpp:107-113) */
  unsigned int index_offset;          /* Np0: size of the charge vector before this bunch (Halton index offset)        */
  double       cloud_charge;          /* cloudCharge_                                                                  */
  double       initial_gamma;         /* initialGamma_                                                                 */
  double       beta_vector[3];        /* betaVector_                                                                   */
  double       position[3];           /* position_[ia]                                                                 */
  double       sigma_position[3];     /* sigmaPosition_                                                                */
  double       sigma_gamma_beta[3];   /* sigmaGammaBeta_                                                               */
  double       tran_trun, long_trun;  /* tranTrun_, longTrun_                                                          */
  double       lambda;                /* lambda_: bunching wavelength, 0 = no undulator (single particles, no copies)  */
  double       bunching_factor;       /* bF_                                                                           */
  double       bunching_phase;        /* bFP_ [degree]                                                                 */
  int          distribution;          /* 0 = uniform (with Gaussian tapers), 1 = gaussian                              */
  int          device;                /* CUDA device ordinal, -1 = current                                             */
} MithraBunchEllipsoid;

/* Bunch::initializeEllipsoid, classes.cpp:104-298 (rank 0 of 1): n = number of macro-particles generated.            */
int  mithra_gpu_bunch_generate    (const MithraBunchEllipsoid* init, MithraGpuBunch** out, size_t* n);
/* Solver::lorentzBoostBunch, solver.cpp:294-302: rnp[2] *= gamma, gb[2] = gamma g (bz - beta); zmax = max rnp[2] (zmaxG). */
int  mithra_gpu_bunch_boost       (MithraGpuBunch* b, double gamma, double beta, double* zmax);
/* solver.cpp:340-346: rnp += gb / g (rnp[2] - zu) beta.                                                               */
int  mithra_gpu_bunch_backproject (MithraGpuBunch* b, double zu, double beta);
/* the list in its order, for dumps and tests (n x 11 doubles).                                                        */
int  mithra_gpu_bunch_download    (MithraGpuBunch* b, double* aos11, size_t capacity, size_t* n);
void mithra_gpu_bunch_destroy     (MithraGpuBunch* b);
/* Solver::distributeParticles, solver.cpp:429-487, for the slab of `h`: the particles whose wrapped z lies in [zp0, zp1)
 * become the slab's bunch, in list order -- mithra_gpu_upload_particles without the host.                             */
int  mithra_gpu_upload_particles_device (MithraGpu* h, const MithraGpuBunch* b);

/* chargeVectorn_ (solver.h:271). */
int mithra_gpu_upload_particles   (MithraGpu* h, const double* aos11, size_t n);
int mithra_gpu_download_particles (MithraGpu* h, double* aos11, size_t capacity, size_t* n);
int mithra_gpu_num_particles      (MithraGpu* h, size_t* n);

/* Re-order the bunch by mesh cell now (mithra_gpu_bunch_update does it every sort_interval steps).  The reference
 * keeps a std::list in insertion order (solver.h:271); here the order in memory is free and every particle keeps
 * its upload index, in which downloads, screen records and mithra_gpu_particle_cells are returned.        */
int mithra_gpu_sort_particles     (MithraGpu* h);

/* Particle-to-cell assignment computed on the device with the arithmetic of the push (solver.cpp:1440-1469:
 * push_m[n] = gather cell (k-k0) N0 N1 + i N1 + j, or -1 when the particle gathers no mesh field) and of the
 * deposit (fdtd.cpp:70-77: ijk6[n] = ip, jp, kp, im, jm, km).  Either pointer may be NULL.  Diagnostic: this
 * is what the bit-exact index parity is checked on.                                                   */
int mithra_gpu_particle_cells (MithraGpu* h, long* push_m, int* ijk6, size_t capacity);

/* time_, timeBunch_, nTime_ (solver.h:265-275). */
int mithra_gpu_set_time (MithraGpu* h, double time, double time_bunch, unsigned int n_time);
int mithra_gpu_get_time (MithraGpu* h, double* time, double* time_bunch, unsigned int* n_time);

/* --- one entry point per reference method of the time march ------------------------------------------- */
int mithra_gpu_field_update        (MithraGpu* h);   /* FdTd::fieldUpdate  fdtd.cpp:231-800 / fdtdSC.cpp:260-1087 */
int mithra_gpu_bunch_update        (MithraGpu* h);   /* rnm = rnp + nUpdateBunch x Solver::bunchUpdate, solver.cpp:1311-1325 */
int mithra_gpu_screen_profile      (MithraGpu* h);   /* Solver::screenProfile solver.cpp:2205-2257        */
int mithra_gpu_power_sample        (MithraGpu* h);   /* Solver::powerSample  radiation.cpp:127-232        */
int mithra_gpu_power_visualize     (MithraGpu* h);   /* Solver::powerVisualize radiation.cpp:324-391 (the map; the .vts writer is the host's) */
int mithra_gpu_field_shift         (MithraGpu* h);   /* FdTd::fieldShift     fdtd.cpp:806-812             */
int mithra_gpu_current_reset       (MithraGpu* h);   /* FdTd::currentReset   fdtd.cpp:23-32               */
int mithra_gpu_current_update      (MithraGpu* h);   /* FdTd::currentUpdate  fdtd.cpp:38-185              */
int mithra_gpu_current_communicate (MithraGpu* h);   /* FdTd::currentCommunicate fdtd.cpp:191-225         */
int mithra_gpu_advance_time        (MithraGpu* h);   /* solver.cpp:1396-1399                              */

/* The body of the second while loop of Solver::solve (solver.cpp:1300-1399), nsteps times, without host
 * synchronisation between steps.  Same results as the calls above issued one by one; the library additionally
 * clears J and prepares the next seed tables on a side stream as soon as the field update has read them, and
 * tests the screens at the tail of the push instead of in a pass of its own.                            */
int mithra_gpu_step (MithraGpu* h, int nsteps);
/* Same, bracketed by CUDA events on the library's stream; *ms = elapsed device time.                  */
int mithra_gpu_step_timed (MithraGpu* h, int nsteps, float* ms);
int mithra_gpu_synchronize (MithraGpu* h);

/* Radiated power: one row of N*Nl doubles (pG[k*Nl+l], radiation.cpp:209-218) per sampled step since the
 * last fetch; *nrows rows are copied (at most capacity_rows).                                          */
int mithra_gpu_fetch_power (MithraGpu* h, double* rows, size_t capacity_rows, size_t* nrows);
/* Screen crossings: records of 6 doubles { x, y, t, gbx, gby, gbz_lab } (solver.cpp:2229-2252) since the
 * last fetch, in particle order within a step.                                                        */
int mithra_gpu_fetch_screen (MithraGpu* h, int screen, double* rec6, size_t capacity, size_t* n);

/* Power map of the last mithra_gpu_power_visualize call: pL[i*N1 + j] (radiation.cpp:388), N0*N1 doubles, zero
 * outside 1 <= i <= N0-2, 1 <= j <= N1-2.  *mine = 1 when the plane lies in this slab (else pL is not written). */
int mithra_gpu_fetch_power_map (MithraGpu* h, double* pL, size_t capacity, int* mine);

/* Solver::bunchSample (solver.cpp:1582-1608): the raw sums over the particles this slab owns, in the order
 * q, q r[3], q r^2[3], q gb[3], q gb^2[3]; the division by q, the standard deviations and the text line
 * (solver.cpp:1617-1640) stay with the host writer.                                                    */
int mithra_gpu_bunch_moments (MithraGpu* h, double sums[13]);

/* FdTd::fieldSample / FdTdSC::fieldSample (fdtd.cpp:851-950, fdtdSC.cpp:1147-1250), to be called where solve() calls
 * it (after the field update and the bunch update of the step, before the field shift: solver.cpp:1326-1328).
 * pos3 = n sampling points (x, y, z) in the moving frame, i.e. seed_.samplingPosition_ after initializeSeedSampling
 * (solver.cpp:862-866); out9[9 t ..] = et[3], bt[3], at[3]: E and B evaluated at the 8 nodes of the point's cell and
 * A^n, interpolated with the reference's weights and summation order (sf_.et, sf_.bt, sf_.at); mine[t] = 1 when the
 * point lies in this slab (solver.cpp:896), else the nine values are 0.  The lab-frame combinations, the unit factors
 * Ce, Cb, Ca and the text line (fdtd.cpp:914-942) stay with the host writer.                                   */
int mithra_gpu_field_sample (MithraGpu* h, const double* pos3, size_t n, double* out9, unsigned char* mine);

/* E, B (FdTd::fieldEvaluate, fdtd.cpp:818-845) and A^n at n mesh nodes, for the field visualisation writers
 * FdTd::fieldVisualizeInPlane{X,Y,Z}Normal (fdtd.cpp:1128-1540), called where solve() calls them (solver.cpp:1332-1340).
 * ijk3[3 t ..] = (i, j, k), k the plane index in this slab's reference numbering (global plane - k0 of the slab);
 * out9[9 t ..] = en_[m][0..2], bn_[m][0..2] (floats, widened) and (*an_)[m][0..2]; mine[t] = 1 when the node is one of
 * the slab's own planes (else zeros).  E/B of the mesh's two end planes are those of their interior neighbour (the
 * reference's fieldEvaluate reads beyond its arrays there).                                                     */
int mithra_gpu_field_nodes (MithraGpu* h, const int* ijk3, size_t n, double* out9, unsigned char* mine);

int mithra_gpu_counters (MithraGpu* h, MithraGpuCounters* out);

/* Per-kernel timing of the last mithra_gpu_step_profiled call (device ms by CUDA events around each
 * kernel group): stencil, boundary, clear, eval_eb, push, deposit, power, screens.                    */
#define MITHRA_GPU_NPHASES 8
int mithra_gpu_step_profiled (MithraGpu* h, int nsteps, float ms[MITHRA_GPU_NPHASES]);

/* Device self-test of the constant-divisor division the kernels use for cell indices and E/B (div_by,
 * mithra_b200/csrc/device_types.cuh): divides the n host doubles x[] by d on the current device both ways and
 * returns the number of results that differ bitwise from IEEE division (must be 0).  No reference counterpart. */
int mithra_gpu_selftest_divide (const double* x, size_t n, double d, unsigned long long* mismatches);

/* --- z-slab exchange between the GPUs of one box (one handle per slab / GPU) --------------------------- */
/* The slab partition is the reference's (Solver::initializeMesh, solver.cpp:619-641: np local planes from global
 * plane k0, two planes shared with each neighbour, ownership interval zp) with rank / size = slab index / count.
 * Export an opaque blob (addresses + CUDA IPC handles of the arrays the ring neighbours write into; call with
 * blob == NULL to get its size) ...                                                                    */
int mithra_gpu_ipc_export  (MithraGpu* h, void* blob, size_t capacity, size_t* nbytes);
/* ... and connect with the blobs of the ring neighbours (rank-1) mod size and (rank+1) mod size -- the ring closes
 * for the particles exactly like the reference's rankB_/rankF_ (solver.cpp:49-52), the potentials do not wrap.
 * Handles of one process (even on one device) and handles of different processes connect the same way.     */
int mithra_gpu_ipc_connect (MithraGpu* h, const void* blob_prev, const void* blob_next);
/* Particle hand-over between slabs once per field step, after the deposit (replaces the per-sub-step ring exchange
 * of solver.cpp:1544-1568, recycleParticles solver.cpp:493-503 and the purge of fdtd.cpp:214-224; a crossing
 * particle deposits both half-segments first, i.e. the k-slab run reproduces the single-slab result).
 * mithra_gpu_step calls both; a process driving several slabs issues every _begin before the first _end.  */
int mithra_gpu_migrate_begin (MithraGpu* h);
int mithra_gpu_migrate_end   (MithraGpu* h);

#ifdef __cplusplus
}
#endif

#endif
