/* engine.cu -- device state and the C ABI (include/mithra_gpu.h) of the B200-native MITHRA time-march.
 *
 * One MithraGpu handle owns one z-slab on one GPU: three rotating levels of the potentials, the current
 * buffer, the E/B node array, the particle struct-of-arrays, the power ring buffer and the screen records.
 * All work of a time step is enqueued on one stream without host synchronisation; the order is the
 * reference's (Solver::solve, solver.cpp:1300-1399).
 *
 * There is no CPU path in this file: every entry point either launches sm_100a kernels or fails.
 */
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>

#include <cuda_runtime.h>

#include "../../include/mithra_gpu.h"
#include "device_types.cuh"
#include "kernels_field.cuh"
#include "kernels_bunch.cuh"
#include "kernels_sort.cuh"
#include "kernels_seed.cuh"
#include "exchange.cuh"
#include "kernels_init.cuh"

using namespace mithra;

/* ---------------------------------------------------------------------------------------------------- */

static thread_local std::string g_error;

static int fail (const char* fmt, ...)
{
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  g_error = buf;
  return 1;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define TRY(call) do { int r_ = (call); if (r_) return r_; } while (0)

enum Phase { PH_STENCIL = 0, PH_BOUNDARY, PH_CLEAR, PH_EVAL, PH_PUSH, PH_DEPOSIT, PH_POWER, PH_SCREEN };

struct MithraGpu
{
  MithraGpuParams prm;
  FieldDev        fd;
  BunchDev        bd;
  int             device;
  int             num_sms;
  cudaStream_t    stream;
  /* housekeeping that only has to be finished by the NEXT consumer runs beside the particle kernels on a second,
   * low-priority stream when the loop is driven by mithra_gpu_step: the clear of the deposit box (needed by the
   * deposit) and the seed line table of the next time level (needed by the next field update)                */
  cudaStream_t    side;
  cudaEvent_t     ev_main, ev_clear, ev_seed;
  bool            overlap;                /* MITHRA_NO_OVERLAP unset                                       */
  bool            clear_ahead;            /* the box has been cleared (or is being cleared) on `side`      */
  bool            seed_ahead;             /* seed table + lines for seed_ahead_time are (being) computed   */
  double          seed_ahead_time;

  /* potentials */
  size_t          level_doubles;          /* ncomp * np * Pp                                              */
  double*         A[3];                   /* rotating: A[ip1], A[in], A[im1]                              */
  void*           Abase[4];               /* the allocations behind A[0..2] and J                          */
  int             ip1, in, im1;
  bool            anp1_is_current;        /* reference view: anp1_ currently holds J (after fieldShift)    */
  double*         J;
  double*         d_stage;                /* AoS staging of the field transfers (field_stage)              */
  size_t          stage_bytes;
  bool            stream_configured[2][2][3]; /* stencil_stream<NSFD, T, .., FACES>: dynamic shared memory limit raised on this device */
  int             face_nodes[3];          /* stencil_stream<.., FACES>: most nodes next to a y face in one tile, per variant (-1: not yet counted) */
  Box*            d_jbox;
  unsigned int*   d_done;

  /* E/B */
  float4*         eb;
  Box*            d_pbox;                 /* particle cell box (filled by push / particle_box)            */
  Box*            d_ebox;                 /* node box evaluated by eval_eb_box                            */
  Box             h_ebox_last;
  bool            fuse_screens;           /* set by mithra_gpu_step around its push                         */
  bool            ev_main_fresh;          /* ev_main was recorded inside the last field update (after J's last reader) */
  unsigned char*  d_emask_cells;          /* E/B pencil mask: cells that hold a particle (marked by push / particle_box) */
  unsigned char*  d_emask_nodes[2];       /* ... spread to the nodes those particles can gather from (spread_eb_mask); two
					     buffers: the one of the last field update also bounds the deposit that followed */
  int             emask_cur;              /* buffer of the last spread                                              */
  int             pushes_since_spread;    /* the mask bounds a deposit only after exactly one push                  */
  bool            emask_fresh;            /* no particle upload since the last spread                               */
  bool            jmask_valid;            /* J lies inside d_emask_nodes[jmask_buf] (plus the slab merge planes)       */
  bool            j_empty;                /* J has been cleared and nothing deposited or uploaded since               */
  bool            j_zeroed_by_update;     /* the last field update cleared J on every plane it read it (stencil_stream + rim_update) */
  int             jmask_buf;
  size_t          emask_bytes;

  /* particles */
  ParticlesDev    P, Palt;                /* the bunch and the copy the counting sort moves it into        */
  double*         pstore[2];              /* 11 * capacity doubles each                                    */
  unsigned int*   idstore[2];
  size_t          pcap, pn;
  unsigned int    next_id;                /* upload index given to the next arrival from a neighbouring slab */
  bool            ids_dense;              /* the ids are a permutation of 0 .. pn-1 (no migration yet)     */
  unsigned int*   d_noutside;

  /* counting sort by cell (kernels_sort.cuh) */
  long            sort_cap;               /* histogram bins                                                */
  unsigned int*   d_hist;
  unsigned int*   d_sums;
  unsigned int*   d_key;
  unsigned int*   d_rank;
  int             sort_interval;          /* MithraGpuParams.sort_interval                                 */
  int             steps_since_sort;

  /* power */
  PowerDev        pw;
  double*         d_fdt;
  double2*        d_ep;
  double*         d_partial;
  int             power_blocks;
  double*         d_rows;                 /* [rows_cap][N*Nl]                                             */
  size_t          rows_cap, rows_used;
  std::vector<double> h_rows;             /* rows already flushed to the host                             */
  int             power_k[MITHRA_MAX_POWER_PLANES];
  double          power_dzr[MITHRA_MAX_POWER_PLANES];
  bool            power_mine[MITHRA_MAX_POWER_PLANES];

  /* power map (power-visualization) */
  PowerDev        pm;
  double*         d_pm_fdt;               /* [Nf][4][P]                                                   */
  double2*        d_pm_ep;                /* [Nf]                                                         */
  double*         d_pm_pL;                /* [P]                                                          */
  int             pm_k;
  double          pm_dzr;
  bool            pm_mine;

  /* screens */
  double*         d_scr_pos;
  double*         d_scr_rec;
  unsigned int*   d_scr_cur;
  unsigned int    scr_cap;
  std::vector<unsigned int> scr_fetched;  /* records already handed out per screen                        */

  /* seed */
  SeedDev*        d_seed;
  double*         d_seed_tab;             /* per-plane table of a seed along +z (kernels_seed.cuh), else 0  */
  double*         d_seedu;                /* RimDev.seedu: seed scalar on the 8 shell source lines of every plane */
  int             seed_L;

  /* slab exchange */
  Exchange        xch;

  /* time */
  double          time, timem1, timep1, time_bunch;
  unsigned int    n_time, n_time_bunch;

  MithraGpuCounters cnt;

  /* profiling */
  bool            profiling;
  cudaEvent_t     pev[2];
  float           pms[MITHRA_GPU_NPHASES];
};

/* a particle list that exists before any slab does (kernels_init.cuh)                                  */
struct MithraGpuBunch
{
  int     device;
  double* aos;                            /* [n][11]                                                        */
  size_t  n;
};

/* ---------------------------------------------------------------------------------------------------- */

extern "C" const char* mithra_gpu_last_error (void) { return g_error.c_str(); }
extern "C" int mithra_gpu_abi_version (void) { return MITHRA_GPU_ABI_VERSION; }
extern "C" int mithra_gpu_device_count (void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static inline int grid_for (long n, int block, int cap)
{
  long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int) g;
}

struct PhaseTimer
{
  MithraGpu* h; int ph;
  PhaseTimer (MithraGpu* h_, int ph_) : h(h_), ph(ph_) { if (h->profiling) cudaEventRecord(h->pev[0], h->stream); }
  ~PhaseTimer ()
  {
    if (!h->profiling) return;
    cudaEventRecord(h->pev[1], h->stream); cudaEventSynchronize(h->pev[1]);
    float ms = 0.f; cudaEventElapsedTime(&ms, h->pev[0], h->pev[1]); h->pms[ph] += ms;
  }
};

/* ---------------------------------------------------------------------------------------------------- */
/* layout conversion kernels (host AoS <-> device component-planar)                                      */

/* `nodes` nodes in the reference's slab numbering; the device arrays hold np planes of which the reference's
 * plane 0 is plane kshift                                                                                 */
__global__ void aos_to_planar (const double* __restrict__ src, double* __restrict__ dst, int ncomp_src, int c0, int nc,
			       long nodes, int P, long Pp, int np, int kshift)
{
  for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < nodes * nc; t += (long) gridDim.x * blockDim.x)
    {
      const long m = t / nc; const int c = (int) (t - m * nc);
      const long k = m / P, r = m - k * P;
      dst[((long) (c0 + c) * np + k + kshift) * Pp + r] = src[m * ncomp_src + c];
    }
}

__global__ void planar_to_aos (const double* __restrict__ src, double* __restrict__ dst, int ncomp_dst, int c0, int nc,
			       long nodes, int P, long Pp, int np, int kshift)
{
  for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < nodes * nc; t += (long) gridDim.x * blockDim.x)
    {
      const long m = t / nc; const int c = (int) (t - m * nc);
      const long k = m / P, r = m - k * P;
      dst[m * ncomp_dst + c] = src[((long) (c0 + c) * np + k + kshift) * Pp + r];
    }
}

/* particles: host array-of-structures double[n][11] <-> device struct-of-arrays.  Row of particle t in the AoS:
 * t on upload (which also numbers the particles), order[t] (or P.id[t] when order == 0) on download, i.e. the
 * reference's list order whatever the counting sort did to the device order.                               */
__global__ void __launch_bounds__(256) aos_to_particles (const double* __restrict__ aos, ParticlesDev P, long n)
{
  const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double* o = aos + t * 11;
  P.q[t] = o[0];
  P.r[0][t] = o[1];  P.r[1][t] = o[2];  P.r[2][t] = o[3];
  P.rm[0][t] = o[4]; P.rm[1][t] = o[5]; P.rm[2][t] = o[6];
  P.gb[0][t] = o[7]; P.gb[1][t] = o[8]; P.gb[2][t] = o[9];
  P.e[t] = o[10];
  P.id[t] = (unsigned int) t;
}

__global__ void __launch_bounds__(256) particles_to_aos (ParticlesDev P, long n, const unsigned int* __restrict__ order, double* __restrict__ aos)
{
  const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double* o = aos + (long) (order ? order[t] : P.id[t]) * 11;
  o[0] = P.q[t];
  o[1] = P.r[0][t];  o[2] = P.r[1][t];  o[3] = P.r[2][t];
  o[4] = P.rm[0][t]; o[5] = P.rm[1][t]; o[6] = P.rm[2][t];
  o[7] = P.gb[0][t]; o[8] = P.gb[1][t]; o[9] = P.gb[2][t];
  o[10] = P.e[t];
}

__global__ void set_box (Box* b, int l0, int l1, int l2, int h0, int h1, int h2)
{ b->lo[0] = l0; b->lo[1] = l1; b->lo[2] = l2; b->hi[0] = h0; b->hi[1] = h1; b->hi[2] = h2; }

/* node box for the E/B evaluation from the particle cell box: pad by the cells a particle can cross in one
 * field step, add the +1 upper node, clamp to the interior transversally; then empty the particle box.   */
__global__ void make_eb_box (const Box* __restrict__ pbox_in, Box* __restrict__ ebox, Box* __restrict__ pbox_reset,
			     int padx, int pady, int padz, int N0, int N1, int np)
{
  Box p = *pbox_in;
  Box e;
  if (p.hi[0] < p.lo[0]) { e.lo[0] = e.lo[1] = e.lo[2] = 0; e.hi[0] = e.hi[1] = e.hi[2] = -1; }
  else
    {
      e.lo[0] = max(1, p.lo[0] - padx); e.hi[0] = min(N0 - 2, p.hi[0] + 1 + padx);
      e.lo[1] = max(1, p.lo[1] - pady); e.hi[1] = min(N1 - 2, p.hi[1] + 1 + pady);
      e.lo[2] = max(0, p.lo[2] - padz); e.hi[2] = min(np - 1, p.hi[2] + 1 + padz);
    }
  *ebox = e;
  if (pbox_reset)
    {
      pbox_reset->lo[0] = pbox_reset->lo[1] = pbox_reset->lo[2] = 0x7fffffff;
      pbox_reset->hi[0] = pbox_reset->hi[1] = pbox_reset->hi[2] = -1;
    }
}

/* ---------------------------------------------------------------------------------------------------- */

static void fill_field_dev (const MithraGpuParams& p, FieldDev& f)
{
  memset(&f, 0, sizeof(f));
  f.kshift = (p.size > 1 && p.rank > 0) ? 1 : 0;
  f.kb = f.kshift + 1;
  f.N0 = p.N0; f.N1 = p.N1; f.np = p.np + f.kshift; f.k0 = p.k0 - f.kshift;
  f.P  = p.N0 * p.N1;
  f.Pp = ((long) f.P + 15) / 16 * 16;
  if (const char* e = getenv("MITHRA_PLANE_PAD")) f.Pp += 16L * atoi(e);                 /* experiments */
  f.ncomp = p.space_charge ? 4 : 3;
  f.rank = p.rank; f.size = p.size;
  f.nsfd = (p.solver == MITHRA_SOLVER_NSFD) ? 1 : 0;
  f.order = p.truncation_order;
  memcpy(f.a, p.a, sizeof(f.a)); f.alpha = p.alpha; f.beta = p.beta_nsfd;
  memcpy(f.bB, p.bB, sizeof(f.bB)); memcpy(f.cB, p.cB, sizeof(f.cB)); memcpy(f.dB, p.dB, sizeof(f.dB));
  memcpy(f.eE, p.eE, sizeof(f.eE)); memcpy(f.fE, p.fE, sizeof(f.fE)); memcpy(f.gE, p.gE, sizeof(f.gE));
  memcpy(f.hC, p.hC, sizeof(f.hC));
  f.dt = p.dt; f.dx2 = 2.0 * p.dx; f.dy2 = 2.0 * p.dy; f.dz2 = 2.0 * p.dz;
  f.rmdt = reciprocal_of(- f.dt); f.rdx2 = reciprocal_of(f.dx2); f.rdy2 = reciprocal_of(f.dy2); f.rdz2 = reciprocal_of(f.dz2);
}

static void fill_bunch_dev (const MithraGpuParams& p, const FieldDev& f, BunchDev& b)
{
  const double PI = 3.1415926535, EC = 1.602e-19, EM = 9.109e-31;      /* stdinclude.h:43-52 */
  memset(&b, 0, sizeof(b));
  b.xmin = p.xmin; b.xmax = p.xmax; b.ymin = p.ymin; b.ymax = p.ymax; b.zmin = p.zmin; b.zmax = p.zmax;
  b.zp0 = p.zp[0]; b.zp1 = p.zp[1]; b.Lz = p.Lz;
  b.dx = p.dx; b.dy = p.dy; b.dz = p.dz;
  b.rdx = reciprocal_of(b.dx); b.rdy = reciprocal_of(b.dy); b.rdz = reciprocal_of(b.dz); b.rr2 = reciprocal_of(p.r2);
  b.c0 = p.c0; b.gamma = p.gamma; b.beta = p.beta; b.dt_shift = p.dt_shift;
  b.r1 = p.r1; b.r2 = p.r2; b.dtb = p.dtb; b.dt_bunch = p.dt_bunch; b.dt_field = p.dt;
  {
    const double cdt = p.c0 * p.dt * ( 1.0 + 1.0e-6 );
    b.reach[0] = cdt / p.dx + 1.0e-9; b.reach[1] = cdt / p.dy + 1.0e-9; b.reach[2] = cdt / p.dz + 1.0e-9;
  }
  b.N0 = p.N0; b.N1 = p.N1; b.np = f.np; b.k0 = f.k0; b.kshift = f.kshift; b.P = f.P; b.Pp = f.Pp; b.ncomp = f.ncomp;
  b.size = p.size;
  b.n_und = p.n_undulators;
  b.und0_dist = p.n_undulators > 0 ? p.undulator[0].dist : 0.0;
  for (int u = 0; u < p.n_undulators; u++)
    {
      const MithraUndulator& U = p.undulator[u];
      UndulatorDev& D = b.und[u];
      D.type = U.type;
      /* solver.cpp:1805-1812, same operation order */
      D.b0 = ( U.lu != 0.0 ) ? EM * p.c0 * 2 * PI / U.lu * U.k / EC : 0.0;
      D.ku = ( U.lu != 0.0 ) ? 2 * PI / U.lu : 0.0;
      D.ct = cos( U.theta ); D.st = sin( U.theta );
      D.rb = U.rb; D.len = U.length * U.lu;
      D.has_prev = (u > 0); D.has_next = (u + 1 < p.n_undulators);
      /* beam.cc:35 and :60 */
      D.r0_prev = D.has_prev ? p.undulator[u - 1].rb + p.undulator[u - 1].length * p.undulator[u - 1].lu - U.rb : 0.0;
      D.r0_next = D.has_next ? p.undulator[u + 1].rb - U.rb - U.length * U.lu : 0.0;
      D.beam = U.beam;
    }
  b.n_ext = p.n_ext_fields;
  for (int u = 0; u < p.n_ext_fields; u++) b.ext[u] = p.ext_field[u];
}

/* CUDA loads kernels lazily, and loading one while another kernel spins on a neighbour's flag can block the
 * launching host thread until the spinning kernel ends -- a deadlock when the awaited neighbour is driven by the same
 * thread.  Touching every kernel once at creation forces the loads up front.                               */
template <class K> static inline cudaError_t preload (K kernel) { cudaFuncAttributes a; return cudaFuncGetAttributes(&a, (const void*) kernel); }

static int preload_kernels ()
{
  cudaError_t e = cudaSuccess;
  #define PL(k) do { cudaError_t r_ = preload(k); if (r_ != cudaSuccess) e = r_; } while (0)
  PL(aos_to_planar); PL(planar_to_aos); PL(aos_to_particles); PL(particles_to_aos); PL(set_box); PL(make_eb_box);
  PL(sort_zero); PL(sort_count); PL(scan_chunk_sums); PL(scan_sums); PL(scan_chunks); PL(sort_permute);
  PL((stencil_interior<true, 128, 32>)); PL((stencil_interior<false, 128, 32>)); PL((stencil_stream<true, 512, 8, false>)); PL((stencil_stream<false, 512, 8, false>));
  PL((stencil_stream<true, 480, 8, false>)); PL((stencil_stream<false, 480, 8, false>)); PL((stencil_stream<true, 448, 8, true>)); PL((stencil_stream<false, 448, 8, true>)); PL((stencil_stream<true, 384, 8, true, true>)); PL((stencil_stream<false, 384, 8, true, true>));
  PL(boundary_faces); PL(boundary_edges); PL(boundary_corners); PL(clear_current_box);
  PL(eval_eb_box<true>); PL(eval_eb_box<false>); PL(eval_eb_march<true>); PL(eval_eb_march<false>); PL(spread_eb_mask);
  PL(particle_box); PL(particle_cells); PL(bunch_moments); PL(push_particles<true>); PL(push_particles<false>); PL(deposit_current<true>); PL(deposit_current<false>);
  PL(screen_cross); PL(power_dft<true>); PL(power_dft<false>); PL(power_finish); PL(power_map<true>); PL(power_map<false>); PL(field_sample<true>); PL(field_sample<false>); PL(field_nodes<true>); PL(field_nodes<false>);
  PL(seed_inject_scan); PL(seed_inject_shell); PL(seed_lines); PL(seed_xshell_rows); PL(seed_inject_zshell); PL(rim_update<true>); PL(rim_update<false>); PL(seed_initial_kernel); PL(seed_plane_table);
  PL(put_planes); PL(put_eb); PL(put_jmail); PL(add_jmail); PL(signal_flag); PL(wait_flag);
  PL(migrate_pack); PL(put_outbox); PL(fill_holes); PL(unpack_inbox);
  PL(ellipsoid_count); PL(scan_block_counts); PL(ellipsoid_write); PL(bunch_boost); PL(bunch_backproject); PL(owned_count); PL(owned_copy);
  #undef PL
  if (e != cudaSuccess) return fail("preloading the kernels failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int mithra_gpu_create (const MithraGpuParams* params, MithraGpu** out)
{
  if (!params || !out) return fail("mithra_gpu_create: null argument");
  if (params->abi_version != MITHRA_GPU_ABI_VERSION) return fail("mithra_gpu_create: ABI version %d, library has %d", params->abi_version, MITHRA_GPU_ABI_VERSION);
  if (params->N0 < 5 || params->N1 < 5 || params->np < 5) return fail("mithra_gpu_create: mesh too small (%d x %d x %d local planes)", params->N0, params->N1, params->np);
  if (params->n_undulators > MITHRA_MAX_UNDULATORS || params->n_ext_fields > MITHRA_MAX_EXTFIELDS) return fail("mithra_gpu_create: too many undulators / external fields");
  if (params->power.enabled && (params->power.N > MITHRA_MAX_POWER_PLANES || params->power.Nl > MITHRA_MAX_POWER_LAMBDAS || params->power.Nf < 1)) return fail("mithra_gpu_create: power sampling out of range");
  if (params->screens.enabled && params->screens.N > MITHRA_MAX_SCREENS) return fail("mithra_gpu_create: too many screens");

  if (params->size < 1 || params->rank < 0 || params->rank >= params->size) return fail("mithra_gpu_create: rank %d of %d slabs", params->rank, params->size);
  if (params->size > 1 && params->np < 6) return fail("mithra_gpu_create: a slab of %d planes is too thin to be split further", params->np);

  for (double d : { params->dt, params->dx, params->dy, params->dz, params->r2 })
    if (reciprocal_of(d) == 0.0) return fail("mithra_gpu_create: a cell size, the time step or r2 (%g) is outside (1e-100, 1e100) job units", d);

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    { cudaGetLastError(); return fail("mithra_gpu_create: no CUDA device available; this library has no CPU path"); }

  MithraGpu* h = new MithraGpu();
  h->prm = *params;
  h->device = params->device;
  if (h->device < 0) CU(cudaGetDevice(&h->device));
  CU(cudaSetDevice(h->device));
  cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major < 10) { delete h; return fail("mithra_gpu_create: device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor); }
  h->num_sms = prop.multiProcessorCount;
  if (preload_kernels()) { delete h; return 1; }
  {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi));
    CU(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, lo));
    CU(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_clear, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_seed, cudaEventDisableTiming));
    h->overlap = !getenv("MITHRA_NO_OVERLAP");
    h->clear_ahead = false; h->seed_ahead = false; h->seed_ahead_time = 0.0;
  }
  CU(cudaEventCreate(&h->pev[0])); CU(cudaEventCreate(&h->pev[1]));
  h->profiling = false; memset(h->pms, 0, sizeof(h->pms));

  fill_field_dev(*params, h->fd);
  fill_bunch_dev(*params, h->fd, h->bd);
  const FieldDev& f = h->fd;

  h->level_doubles = (size_t) f.ncomp * f.np * f.Pp;
  {
    /* experiments: MITHRA_LEVEL_SKEW = bytes by which consecutive levels are shifted against each other               */
    const size_t skew = getenv("MITHRA_LEVEL_SKEW") ? (size_t) atol(getenv("MITHRA_LEVEL_SKEW")) / 128 * 128 : 0;
    for (int l = 0; l < 4; l++)
      {
	CU(cudaMalloc(&h->Abase[l], h->level_doubles * sizeof(double) + 4 * skew));
	double* q = (double*) ((char*) h->Abase[l] + l * skew);
	CU(cudaMemsetAsync(q, 0, h->level_doubles * sizeof(double), h->stream));
	if (l < 3) h->A[l] = q; else h->J = q;
      }
  }
  h->ip1 = 0; h->in = 1; h->im1 = 2; h->anp1_is_current = true;
  memset(h->stream_configured, 0, sizeof(h->stream_configured)); h->face_nodes[0] = h->face_nodes[1] = h->face_nodes[2] = -1;
  h->d_stage = 0; h->stage_bytes = 0;
  CU(cudaMalloc(&h->d_jbox, sizeof(Box))); CU(cudaMalloc(&h->d_pbox, sizeof(Box))); CU(cudaMalloc(&h->d_ebox, sizeof(Box)));
  CU(cudaMalloc(&h->d_done, sizeof(unsigned int))); CU(cudaMemsetAsync(h->d_done, 0, sizeof(unsigned int), h->stream));
  set_box<<<1, 1, 0, h->stream>>>(h->d_jbox, 0x7fffffff, 0x7fffffff, 0x7fffffff, -1, -1, -1);
  set_box<<<1, 1, 0, h->stream>>>(h->d_pbox, 0x7fffffff, 0x7fffffff, 0x7fffffff, -1, -1, -1);
  set_box<<<1, 1, 0, h->stream>>>(h->d_ebox, 0, 0, 0, -1, -1, -1);
  h->fuse_screens = false; h->ev_main_fresh = false;
  h->h_ebox_last.lo[0] = h->h_ebox_last.lo[1] = h->h_ebox_last.lo[2] = 0; h->h_ebox_last.hi[0] = h->h_ebox_last.hi[1] = h->h_ebox_last.hi[2] = -1;

  h->d_emask_cells = 0; h->d_emask_nodes[0] = h->d_emask_nodes[1] = 0;
  h->emask_cur = 0; h->pushes_since_spread = 0; h->emask_fresh = false; h->jmask_valid = false; h->jmask_buf = 0; h->j_empty = true; h->j_zeroed_by_update = false;
  h->emask_bytes = ( (size_t) ((f.np + (1 << MITHRA_EB_CHUNK_LOG2) - 1) >> MITHRA_EB_CHUNK_LOG2) * f.P + 15 ) / 16 * 16;   /* whole 16-byte words: spread_eb_mask */
  if (!getenv("MITHRA_NO_EBMASK"))
    {
      CU(cudaMalloc(&h->d_emask_cells, h->emask_bytes)); CU(cudaMemsetAsync(h->d_emask_cells, 0, h->emask_bytes, h->stream));
      for (int w = 0; w < 2; w++) { CU(cudaMalloc(&h->d_emask_nodes[w], h->emask_bytes)); CU(cudaMemsetAsync(h->d_emask_nodes[w], 0, h->emask_bytes, h->stream)); }
    }
  const size_t nodes = (size_t) f.np * f.P;
  CU(cudaMalloc(&h->eb, nodes * 2 * sizeof(float4))); CU(cudaMemsetAsync(h->eb, 0, nodes * 2 * sizeof(float4), h->stream));

  h->pcap = params->max_particles ? params->max_particles : (size_t) 1 << 20;
  h->pn = 0; h->next_id = 0; h->ids_dense = true;
  for (int w = 0; w < 2; w++)
    {
      CU(cudaMalloc(&h->pstore[w], h->pcap * 11 * sizeof(double)));
      CU(cudaMalloc(&h->idstore[w], h->pcap * sizeof(unsigned int)));
      ParticlesDev& Q = w ? h->Palt : h->P;
      double* s = h->pstore[w]; const size_t c = h->pcap;
      Q.q = s; for (int a = 0; a < 3; a++) { Q.r[a] = s + (1 + a) * c; Q.rm[a] = s + (4 + a) * c; Q.gb[a] = s + (7 + a) * c; } Q.e = s + 10 * c;
      Q.id = h->idstore[w];
    }
  /* counting sort: one bin per cell of the particle box, as many as the slab has cells (at most 2^30)      */
  h->sort_cap = std::min<long>((long) f.np * f.P, 1L << 30);
  h->sort_interval = params->sort_interval;
  if (const char* e = getenv("MITHRA_SORT_INTERVAL")) h->sort_interval = atoi(e);      /* experiments: overrides the parameter block */
  h->steps_since_sort = 0;
  CU(cudaMalloc(&h->d_hist, (size_t) h->sort_cap * sizeof(unsigned int)));
  CU(cudaMalloc(&h->d_sums, (size_t) (h->sort_cap / MITHRA_SCAN_CHUNK + 2) * sizeof(unsigned int)));
  CU(cudaMalloc(&h->d_key,  h->pcap * sizeof(unsigned int)));
  CU(cudaMalloc(&h->d_rank, h->pcap * sizeof(unsigned int)));
  CU(cudaMalloc(&h->d_noutside, sizeof(unsigned int))); CU(cudaMemsetAsync(h->d_noutside, 0, sizeof(unsigned int), h->stream));

  /* power sampling, radiation.cpp:18-121 */
  memset(&h->pw, 0, sizeof(h->pw));
  h->d_fdt = 0; h->d_ep = 0; h->d_partial = 0; h->d_rows = 0; h->rows_cap = 0; h->rows_used = 0; h->power_blocks = 0;
  if (params->power.enabled)
    {
      const MithraPower& pp = params->power;
      PowerDev& pw = h->pw;
      pw.N = pp.N; pw.Nl = pp.Nl; pw.Nf = pp.Nf; pw.ni = f.N0 - 4; pw.nj = f.N1 - 4; pw.npx = pw.ni * pw.nj;
      pw.pc = pp.pc; pw.gamma = params->gamma; pw.beta = params->beta; pw.c0 = params->c0;
      int nmine = 0;
      for (int k = 0; k < pp.N; k++)
	{
	  h->power_mine[k] = ( pp.z[k] < params->zp[1] && pp.z[k] >= params->zp[0] );
	  double c; h->power_dzr[k] = modf( ( pp.z[k] - params->zmin ) / params->dz, &c ); h->power_k[k] = (int) c - f.k0;
	  if (h->power_mine[k]) nmine++;
	}
      const size_t ring = (size_t) pp.N * pp.Nf * 4 * pw.npx;
      CU(cudaMalloc(&h->d_fdt, ring * sizeof(double))); CU(cudaMemsetAsync(h->d_fdt, 0, ring * sizeof(double), h->stream));
      std::vector<double2> ep((size_t) pp.Nl * pp.Nf);
      for (int l = 0; l < pp.Nl; l++)
	for (int m = 0; m < pp.Nf; m++)
	  { ep[(size_t) l * pp.Nf + m].x = cos( pp.w[l] * m * params->dt ); ep[(size_t) l * pp.Nf + m].y = sin( pp.w[l] * m * params->dt ); }
      CU(cudaMalloc(&h->d_ep, ep.size() * sizeof(double2))); CU(cudaMemcpy(h->d_ep, ep.data(), ep.size() * sizeof(double2), cudaMemcpyHostToDevice));
      h->power_blocks = (pw.npx + MITHRA_POWER_PX - 1) / MITHRA_POWER_PX;
      CU(cudaMalloc(&h->d_partial, (size_t) pp.N * pp.Nl * h->power_blocks * sizeof(double)));
      CU(cudaMemsetAsync(h->d_partial, 0, (size_t) pp.N * pp.Nl * h->power_blocks * sizeof(double), h->stream));
      h->rows_cap = 4096;
      CU(cudaMalloc(&h->d_rows, h->rows_cap * pp.N * pp.Nl * sizeof(double)));
    }

  /* power map, radiation.cpp:238-318 */
  memset(&h->pm, 0, sizeof(h->pm));
  h->d_pm_fdt = 0; h->d_pm_ep = 0; h->d_pm_pL = 0; h->pm_mine = false; h->pm_k = 0; h->pm_dzr = 0.0;
  if (params->power_map.enabled)
    {
      const MithraPowerMap& pp = params->power_map;
      if (pp.Nf < 1) { mithra_gpu_destroy(h); return fail("mithra_gpu_create: power map window Nf = %d", pp.Nf); }
      PowerDev& pm = h->pm;
      pm.N = 1; pm.Nl = 1; pm.Nf = pp.Nf; pm.ni = f.N0 - 2; pm.nj = f.N1 - 2; pm.npx = f.P;
      pm.pc = pp.pc; pm.gamma = params->gamma; pm.beta = params->beta; pm.c0 = params->c0;
      h->pm_mine = ( pp.z < params->zp[1] && pp.z >= params->zp[0] );
      double c; h->pm_dzr = modf( ( pp.z - params->zmin ) / params->dz, &c ); h->pm_k = (int) c - f.k0;
      if (h->pm_mine)
	{
	  const size_t ring = (size_t) pp.Nf * 4 * f.P;
	  CU(cudaMalloc(&h->d_pm_fdt, ring * sizeof(double))); CU(cudaMemsetAsync(h->d_pm_fdt, 0, ring * sizeof(double), h->stream));
	  std::vector<double2> ep((size_t) pp.Nf);
	  for (int m = 0; m < pp.Nf; m++) { ep[m].x = cos( pp.w * m * params->dt ); ep[m].y = sin( pp.w * m * params->dt ); }
	  CU(cudaMalloc(&h->d_pm_ep, ep.size() * sizeof(double2))); CU(cudaMemcpy(h->d_pm_ep, ep.data(), ep.size() * sizeof(double2), cudaMemcpyHostToDevice));
	  CU(cudaMalloc(&h->d_pm_pL, (size_t) f.P * sizeof(double))); CU(cudaMemsetAsync(h->d_pm_pL, 0, (size_t) f.P * sizeof(double), h->stream));
	}
    }

  /* screens */
  h->d_scr_pos = 0; h->d_scr_rec = 0; h->d_scr_cur = 0; h->scr_cap = 0;
  if (params->screens.enabled && params->screens.N > 0)
    {
      const int ns = params->screens.N;
      h->scr_cap = (unsigned int) (params->max_screen_records ? params->max_screen_records : std::max<size_t>(h->pcap, 1024));
      CU(cudaMalloc(&h->d_scr_pos, ns * sizeof(double))); CU(cudaMemcpy(h->d_scr_pos, params->screens.pos, ns * sizeof(double), cudaMemcpyHostToDevice));
      CU(cudaMalloc(&h->d_scr_rec, (size_t) ns * h->scr_cap * 8 * sizeof(double)));
      CU(cudaMalloc(&h->d_scr_cur, ns * sizeof(unsigned int))); CU(cudaMemsetAsync(h->d_scr_cur, 0, ns * sizeof(unsigned int), h->stream));
      h->scr_fetched.assign(ns, 0u);
    }

  /* seed */
  h->d_seed = 0; h->d_seed_tab = 0; h->d_seedu = 0; h->seed_L = 0;
  if (params->seed_enabled)
    {
      SeedDev sd; memset(&sd, 0, sizeof(sd));
      sd.beam = params->seed; sd.c0 = params->c0; sd.gamma = params->gamma; sd.beta = params->beta; sd.dt_shift = params->dt_shift;
      sd.xmin = params->xmin; sd.ymin = params->ymin; sd.zmin = params->zmin; sd.dx = params->dx; sd.dy = params->dy; sd.dz = params->dz;
      sd.k0 = f.k0;
      const double* D = params->seed.direction; const double* Q = params->seed.polarization;
      sd.yv[0] = D[1] * Q[2] - D[2] * Q[1]; sd.yv[1] = D[2] * Q[0] - D[0] * Q[2]; sd.yv[2] = D[0] * Q[1] - D[1] * Q[0];   /* fieldvector.h cross() */
      const int st = params->seed.seed_type;
      sd.along_z = ( D[0] == 0.0 && D[1] == 0.0 && D[2] == 1.0 &&
		     ( st == MITHRA_BEAM_PLANEWAVE || st == MITHRA_BEAM_PLANEWAVETRUNCATED || st == MITHRA_BEAM_GAUSSIAN || st == MITHRA_BEAM_SUPERGAUSSIAN ) ) ? 1 : 0;
      if (getenv("MITHRA_SEED_GENERIC")) sd.along_z = 0;
      CU(cudaMalloc(&h->d_seed, sizeof(SeedDev))); CU(cudaMemcpy(h->d_seed, &sd, sizeof(SeedDev), cudaMemcpyHostToDevice));
      if (sd.along_z) CU(cudaMalloc(&h->d_seed_tab, (size_t) f.np * MITHRA_SEED_TAB * sizeof(double)));
      h->seed_L = (std::max(f.N0, f.N1) + 1) & ~1;
      CU(cudaMalloc(&h->d_seedu, (size_t) f.np * 8 * h->seed_L * sizeof(double)));
      CU(cudaMemsetAsync(h->d_seedu, 0, (size_t) f.np * 8 * h->seed_L * sizeof(double), h->stream));
    }

  if (exchange_init(h->xch, f, h->pcap, h->stream)) { std::string e = h->xch.error; mithra_gpu_destroy(h); return fail("mithra_gpu_create: %s", e.c_str()); }

  h->time = 0.0; h->timem1 = - params->dt; h->timep1 = params->dt; h->time_bunch = 0.0; h->n_time = 0; h->n_time_bunch = 0;
  memset(&h->cnt, 0, sizeof(h->cnt));
  CU(cudaStreamSynchronize(h->stream));
  *out = h;
  return 0;
}

extern "C" void mithra_gpu_destroy (MithraGpu* h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  exchange_destroy(h->xch);
  for (int l = 0; l < 4; l++) cudaFree(h->Abase[l]);
  cudaFree(h->d_stage);
  cudaFree(h->d_jbox); cudaFree(h->d_pbox); cudaFree(h->d_ebox); cudaFree(h->d_done);
  cudaFree(h->eb); cudaFree(h->d_noutside); cudaFree(h->d_emask_cells); cudaFree(h->d_emask_nodes[0]); cudaFree(h->d_emask_nodes[1]);
  for (int w = 0; w < 2; w++) { cudaFree(h->pstore[w]); cudaFree(h->idstore[w]); }
  cudaFree(h->d_hist); cudaFree(h->d_sums); cudaFree(h->d_key); cudaFree(h->d_rank);
  cudaFree(h->d_pm_fdt); cudaFree(h->d_pm_ep); cudaFree(h->d_pm_pL);
  cudaFree(h->d_fdt); cudaFree(h->d_ep); cudaFree(h->d_partial); cudaFree(h->d_rows);
  cudaFree(h->d_scr_pos); cudaFree(h->d_scr_rec); cudaFree(h->d_scr_cur); cudaFree(h->d_seed); cudaFree(h->d_seed_tab); cudaFree(h->d_seedu);
  cudaEventDestroy(h->pev[0]); cudaEventDestroy(h->pev[1]);
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  if (h->ev_clear) cudaEventDestroy(h->ev_clear);
  if (h->ev_seed) cudaEventDestroy(h->ev_seed);
  cudaStreamDestroy(h->stream);
  delete h;
}

#define USE(h) do { if (!(h)) return fail("null handle"); CU(cudaSetDevice((h)->device)); } while (0)

/* ---------------------------------------------------------------------------------------------------- */
/* state transfer                                                                                        */

/* Staging buffer of the AoS <-> planar field transfers (the ABI speaks the reference's [node][xyz] layout): one level
 * of potentials, allocated at the first transfer and kept -- cudaMalloc / cudaFree of 1.4 GB around every copy cost more
 * than the copy.                                                                                              */
static int field_stage (MithraGpu* h, size_t bytes, double** out)
{
  if (h->stage_bytes < bytes)
    {
      if (h->d_stage) { CU(cudaStreamSynchronize(h->stream)); CU(cudaFree(h->d_stage)); h->d_stage = 0; h->stage_bytes = 0; }
      CU(cudaMalloc(&h->d_stage, bytes));
      h->stage_bytes = bytes;
    }
  *out = h->d_stage;
  return 0;
}

static int upload_vec (MithraGpu* h, const double* src, double* dst, int ncomp_src, int c0, int nc)
{
  const FieldDev& f = h->fd;
  const long nodes = (long) h->prm.np * f.P;
  double* tmp = 0;
  TRY(field_stage(h, (size_t) nodes * ncomp_src * sizeof(double), &tmp));
  CU(cudaMemcpyAsync(tmp, src, (size_t) nodes * ncomp_src * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  aos_to_planar<<<grid_for(nodes * nc, 256, h->num_sms * 8), 256, 0, h->stream>>>(tmp, dst, ncomp_src, c0, nc, nodes, f.P, f.Pp, f.np, f.kshift);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

static int download_vec (MithraGpu* h, const double* src, double* dst, int ncomp_dst, int c0, int nc)
{
  const FieldDev& f = h->fd;
  const long nodes = (long) h->prm.np * f.P;
  double* tmp = 0;
  TRY(field_stage(h, (size_t) nodes * ncomp_dst * sizeof(double), &tmp));
  planar_to_aos<<<grid_for(nodes * nc, 256, h->num_sms * 8), 256, 0, h->stream>>>(src, tmp, ncomp_dst, c0, nc, nodes, f.P, f.Pp, f.np, f.kshift);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(dst, tmp, (size_t) nodes * ncomp_dst * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int mithra_gpu_upload_fields (MithraGpu* h, const double* an, const double* anm1, const double* jn,
					 const double* fn, const double* fnm1, const double* rho)
{
  USE(h);
  const FieldDev& f = h->fd;
  if (an)   TRY(upload_vec(h, an,   h->A[h->in],  3, 0, 3));
  if (anm1) TRY(upload_vec(h, anm1, h->A[h->im1], 3, 0, 3));
  if (jn)   TRY(upload_vec(h, jn,   h->J,         3, 0, 3));
  if (f.ncomp == 4)
    {
      if (fn)   TRY(upload_vec(h, fn,   h->A[h->in],  1, 3, 1));
      if (fnm1) TRY(upload_vec(h, fnm1, h->A[h->im1], 1, 3, 1));
      if (rho)  TRY(upload_vec(h, rho,  h->J,         1, 3, 1));
    }
  if (jn || rho)
    {
      set_box<<<1, 1, 0, h->stream>>>(h->d_jbox, 0, 0, 0, f.N0 - 1, f.N1 - 1, f.np - 1);
      h->jmask_valid = false; h->j_empty = false;          /* an arbitrary source: the box alone bounds it               */
    }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  h->anp1_is_current = true;
  return 0;
}

extern "C" int mithra_gpu_download_fields (MithraGpu* h, double* anp1, double* an, double* anm1,
					   double* fnp1, double* fn, double* fnm1)
{
  USE(h);
  const FieldDev& f = h->fd;
  const double* np1 = h->anp1_is_current ? h->J : h->A[h->ip1];
  if (anp1) TRY(download_vec(h, np1,          anp1, 3, 0, 3));
  if (an)   TRY(download_vec(h, h->A[h->in],  an,   3, 0, 3));
  if (anm1) TRY(download_vec(h, h->A[h->im1], anm1, 3, 0, 3));
  if (f.ncomp == 4)
    {
      if (fnp1) TRY(download_vec(h, np1,          fnp1, 1, 3, 1));
      if (fn)   TRY(download_vec(h, h->A[h->in],  fn,   1, 3, 1));
      if (fnm1) TRY(download_vec(h, h->A[h->im1], fnm1, 1, 3, 1));
    }
  return 0;
}

/* The reach mask (device_types.cuh) assumes that a particle crosses at most one cell face per axis transversally and
 * stays within the neighbouring plane chunk in one field step -- true for every stable time step (c dt < dx, dy; c dt ~ dz). */
static bool reach_mask_usable (const MithraGpu* h)
{
  if (!h->d_emask_nodes[0] || getenv("MITHRA_EB_BOX")) return false;
  const double cdt = h->prm.c0 * h->prm.dt * ( 1.0 + 1.0e-6 );
  return cdt + 1.0e-9 * h->prm.dx < h->prm.dx && cdt + 1.0e-9 * h->prm.dy < h->prm.dy &&
	 (int) ceil(cdt / h->prm.dz) + 2 <= (1 << MITHRA_EB_CHUNK_LOG2);
}

extern "C" int mithra_gpu_download_eb (MithraGpu* h, float* en, float* bn, unsigned char* mask)
{
  USE(h);
  const FieldDev& f = h->fd;
  const size_t nodes_int = (size_t) f.np * f.P;
  const size_t nodes = (size_t) h->prm.np * f.P;                  /* reference slab numbering                     */
  std::vector<float4> tmp(nodes_int * 2);
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(tmp.data(), h->eb, nodes_int * 2 * sizeof(float4), cudaMemcpyDeviceToHost));
  Box b; CU(cudaMemcpy(&b, h->d_ebox, sizeof(Box), cudaMemcpyDeviceToHost));
  std::vector<unsigned char> pencil;
  if (reach_mask_usable(h))
    { pencil.resize(h->emask_bytes); CU(cudaMemcpy(pencil.data(), h->d_emask_nodes[h->emask_cur], h->emask_bytes, cudaMemcpyDeviceToHost)); }
  for (size_t m = 0; m < nodes; m++)
    {
      const int kr = (int) (m / f.P), r = (int) (m % f.P), i = r / f.N1, j = r % f.N1;
      const int k = kr + f.kshift;
      const size_t mi = (size_t) k * f.P + r;
      bool in = ( i >= b.lo[0] && i <= b.hi[0] && j >= b.lo[1] && j <= b.hi[1] && k >= b.lo[2] && k <= b.hi[2] );
      /* inside the box only the marked pencils are evaluated (the two copied end planes: the whole box)            */
      if (in && !pencil.empty() && k >= f.kb && k <= f.np - 2) in = pencil[(size_t) (k >> MITHRA_EB_CHUNK_LOG2) * f.P + r] != 0;
      /* ghost planes hold what the neighbour evaluated on the whole plane                                     */
      if (f.size > 1 && ((k < f.kb && f.rank != 0) || (k == f.np - 1 && f.rank != f.size - 1)))
	in = h->xch.connected && i >= 1 && i <= f.N0 - 2 && j >= 1 && j <= f.N1 - 2;
      if (mask) mask[m] = in ? 1 : 0;
      if (en) { en[3 * m] = in ? tmp[2 * mi].x : 0.f; en[3 * m + 1] = in ? tmp[2 * mi].y : 0.f; en[3 * m + 2] = in ? tmp[2 * mi].z : 0.f; }
      if (bn) { bn[3 * m] = in ? tmp[2 * mi + 1].x : 0.f; bn[3 * m + 1] = in ? tmp[2 * mi + 1].y : 0.f; bn[3 * m + 2] = in ? tmp[2 * mi + 1].z : 0.f; }
    }
  return 0;
}

extern "C" int mithra_gpu_seed_initial (MithraGpu* h)
{
  USE(h);
  if (!h->d_seed) return 0;
  seed_initial_kernel<<<h->num_sms * 8, 128, 0, h->stream>>>(h->d_seed, h->fd, h->A[h->in], h->A[h->im1], h->time, h->timem1);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

static int refresh_particle_box (MithraGpu* h)
{
  h->emask_fresh = false;                                /* new positions: the node mask of the last field update is not theirs */
  set_box<<<1, 1, 0, h->stream>>>(h->d_pbox, 0x7fffffff, 0x7fffffff, 0x7fffffff, -1, -1, -1);
  if (h->d_emask_cells) CU(cudaMemsetAsync(h->d_emask_cells, 0, h->emask_bytes, h->stream));
  if (h->pn > 0)
    particle_box<<<grid_for((long) h->pn, 256, h->num_sms * 8), 256, 0, h->stream>>>(h->bd, h->P, 0L, (long) h->pn, h->d_pbox, h->d_emask_cells);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 2;
  return 0;
}

/* The copy the sort is not using at the moment doubles as the staging buffer of the AoS transfers.           */
/* the n particles staged as double[n][11] in the idle copy become the slab's bunch                              */
static int adopt_staged_particles (MithraGpu* h, size_t n)
{
  if (n)
    {
      aos_to_particles<<<(int) ((n + 255) / 256), 256, 0, h->stream>>>(h->Palt.q, h->P, (long) n);
      CU(cudaGetLastError());
      h->cnt.kernel_launches += 1;
    }
  h->pn = n; h->next_id = (unsigned int) n; h->ids_dense = true;
  h->steps_since_sort = 1 << 30;                        /* sort before the next push (if sorting is on)             */
  TRY(refresh_particle_box(h));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int mithra_gpu_upload_particles (MithraGpu* h, const double* aos11, size_t n)
{
  USE(h);
  if (n > h->pcap) return fail("mithra_gpu_upload_particles: %zu particles exceed the capacity %zu (MithraGpuParams.max_particles)", n, h->pcap);
  if (n) CU(cudaMemcpyAsync(h->Palt.q, aos11, n * 11 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return adopt_staged_particles(h, n);
}

/* Solver::distributeParticles (solver.cpp:429-487) without the host: the particles of the device-resident list that this
 * slab owns, in list order.  The list may live on another device of the box (peer copy).                           */
extern "C" int mithra_gpu_upload_particles_device (MithraGpu* h, const MithraGpuBunch* b)
{
  USE(h);
  if (!b) return fail("mithra_gpu_upload_particles_device: null bunch");
  const MithraGpuParams& p = h->prm;
  size_t n = 0;
  if (b->n > 0)
    {
      const double* src = b->aos;
      double* tmp = 0;
      if (b->device != h->device)
	{
	  CU(cudaMalloc(&tmp, b->n * 11 * sizeof(double)));
	  CU(cudaMemcpyPeer(tmp, h->device, b->aos, b->device, b->n * 11 * sizeof(double)));
	  src = tmp;
	}
      const unsigned int nblocks = (unsigned int) ((b->n + 255) / 256);
      unsigned int* d_cnt = 0; unsigned long long* d_tot = 0;
      CU(cudaMalloc(&d_cnt, (size_t) nblocks * sizeof(unsigned int))); CU(cudaMalloc(&d_tot, sizeof(unsigned long long)));
      owned_count<<<nblocks, 256, 0, h->stream>>>(src, b->n, p.zmin, p.Lz, p.zp[0], p.zp[1], d_cnt);
      scan_block_counts<<<1, 1024, 0, h->stream>>>(d_cnt, nblocks, d_tot);
      CU(cudaGetLastError());
      unsigned long long own = 0;
      CU(cudaMemcpyAsync(&own, d_tot, sizeof(own), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      n = (size_t) own;
      if (n > h->pcap)
	{ cudaFree(d_cnt); cudaFree(d_tot); cudaFree(tmp); return fail("mithra_gpu_upload_particles_device: %zu particles exceed the capacity %zu (MithraGpuParams.max_particles)", n, h->pcap); }
      if (n) owned_copy<<<nblocks, 256, 0, h->stream>>>(src, b->n, p.zmin, p.Lz, p.zp[0], p.zp[1], d_cnt, h->Palt.q);
      CU(cudaGetLastError());
      h->cnt.kernel_launches += 3;
      CU(cudaStreamSynchronize(h->stream));
      cudaFree(d_cnt); cudaFree(d_tot); cudaFree(tmp);
    }
  return adopt_staged_particles(h, n);
}

/* ---------------------------------------------------------------------------------------------------- */
/* the bunch of Solver::initialize() on the device (kernels_init.cuh)                                    */

extern "C" void mithra_gpu_bunch_destroy (MithraGpuBunch* b)
{
  if (!b) return;
  cudaSetDevice(b->device);
  cudaFree(b->aos);
  delete b;
}

extern "C" int mithra_gpu_bunch_generate (const MithraBunchEllipsoid* init, MithraGpuBunch** out, size_t* n)
{
  if (!init || !out) return fail("mithra_gpu_bunch_generate: null argument");
  if (init->number_of_particles % 4 != 0) return fail("mithra_gpu_bunch_generate: the number of particles must be a multiple of four (classes.cpp:107-113)");
  if (init->bunching_factor > 2.0 || init->bunching_factor < 0.0) return fail("mithra_gpu_bunch_generate: the bunching factor can not be larger than one or a negative value");
  if (init->distribution != 0 && init->distribution != 1) return fail("mithra_gpu_bunch_generate: unknown longitudinal distribution");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    { cudaGetLastError(); return fail("mithra_gpu_bunch_generate: no CUDA device available; this library has no CPU path"); }
  MithraGpuBunch* b = new MithraGpuBunch();
  b->device = init->device; b->aos = 0; b->n = 0;
  if (b->device < 0) CU(cudaGetDevice(&b->device));
  CU(cudaSetDevice(b->device));
  const double PI = 3.1415926535;                        /* stdinclude.h:43 */
  const unsigned int ng = ( init->lambda == 0.0 ) ? 1u : 4u;
  const unsigned int nbody = init->number_of_particles / ng;
  /* body + the tapers of a uniform profile, classes.cpp:203,268 */
  unsigned int ncand = nbody;
  if (init->distribution == 0)
    ncand = std::max(nbody, (unsigned int) ( nbody * ( 1.0 + 2.0 * init->lambda * sqrt( 2.0 * PI ) / ( 2.0 * init->sigma_position[2] ) ) ));
  const unsigned int nblocks = (ncand + 255u) / 256u;
  unsigned int* d_cnt = 0; unsigned long long* d_tot = 0;
  CU(cudaMalloc(&d_cnt, (size_t) std::max(1u, nblocks) * sizeof(unsigned int))); CU(cudaMalloc(&d_tot, sizeof(unsigned long long)));
  unsigned long long accepted = 0;
  if (ncand > 0)
    {
      ellipsoid_count<<<nblocks, 256>>>(*init, ncand, d_cnt);
      scan_block_counts<<<1, 1024>>>(d_cnt, nblocks, d_tot);
      CU(cudaGetLastError());
      CU(cudaMemcpy(&accepted, d_tot, sizeof(accepted), cudaMemcpyDeviceToHost));
    }
  b->n = (size_t) accepted * ng;
  if (b->n > 0)
    {
      CU(cudaMalloc(&b->aos, b->n * 11 * sizeof(double)));
      ellipsoid_write<<<nblocks, 256>>>(*init, ncand, d_cnt, b->aos);
      CU(cudaGetLastError());
      CU(cudaDeviceSynchronize());
    }
  cudaFree(d_cnt); cudaFree(d_tot);
  if (n) *n = b->n;
  *out = b;
  return 0;
}

extern "C" int mithra_gpu_bunch_boost (MithraGpuBunch* b, double gamma, double beta, double* zmax)
{
  if (!b) return fail("mithra_gpu_bunch_boost: null bunch");
  CU(cudaSetDevice(b->device));
  double zm = -1.0e100;
  if (b->n > 0)
    {
      const unsigned int nblocks = (unsigned int) ((b->n + 255) / 256);
      double* d_z = 0; CU(cudaMalloc(&d_z, (size_t) nblocks * sizeof(double)));
      bunch_boost<<<nblocks, 256>>>(b->aos, b->n, gamma, beta, d_z);
      CU(cudaGetLastError());
      std::vector<double> z(nblocks);
      CU(cudaMemcpy(z.data(), d_z, (size_t) nblocks * sizeof(double), cudaMemcpyDeviceToHost));
      cudaFree(d_z);
      for (double v : z) zm = std::max(zm, v);
    }
  if (zmax) *zmax = zm;
  return 0;
}

extern "C" int mithra_gpu_bunch_backproject (MithraGpuBunch* b, double zu, double beta)
{
  if (!b) return fail("mithra_gpu_bunch_backproject: null bunch");
  CU(cudaSetDevice(b->device));
  if (b->n > 0)
    {
      bunch_backproject<<<(unsigned int) ((b->n + 255) / 256), 256>>>(b->aos, b->n, zu, beta);
      CU(cudaGetLastError());
      CU(cudaDeviceSynchronize());
    }
  return 0;
}

extern "C" int mithra_gpu_bunch_download (MithraGpuBunch* b, double* aos11, size_t capacity, size_t* n)
{
  if (!b) return fail("mithra_gpu_bunch_download: null bunch");
  if (n) *n = b->n;
  if (!aos11) return 0;
  if (capacity < b->n) return fail("mithra_gpu_bunch_download: capacity %zu < %zu particles", capacity, b->n);
  CU(cudaSetDevice(b->device));
  if (b->n > 0) CU(cudaMemcpy(aos11, b->aos, b->n * 11 * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

/* rank of every particle in the order of the upload indices, for bunches that lost or received particles     */
static int id_ranks (MithraGpu* h, unsigned int** d_order)
{
  *d_order = 0;
  if (h->ids_dense || h->pn == 0) return 0;
  const size_t n = h->pn;
  std::vector<unsigned int> ids(n), order(n), rank(n);
  CU(cudaMemcpy(ids.data(), h->P.id, n * sizeof(unsigned int), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; i++) order[i] = (unsigned int) i;
  std::sort(order.begin(), order.end(), [&](unsigned int a, unsigned int b) { return ids[a] < ids[b]; });
  for (size_t i = 0; i < n; i++) rank[order[i]] = (unsigned int) i;
  CU(cudaMalloc(d_order, n * sizeof(unsigned int)));
  CU(cudaMemcpy(*d_order, rank.data(), n * sizeof(unsigned int), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int mithra_gpu_download_particles (MithraGpu* h, double* aos11, size_t capacity, size_t* n)
{
  USE(h);
  CU(cudaStreamSynchronize(h->stream));
  if (n) *n = h->pn;
  if (!aos11) return 0;
  if (capacity < h->pn) return fail("mithra_gpu_download_particles: capacity %zu < %zu particles", capacity, h->pn);
  const size_t np = h->pn;
  if (np == 0) return 0;
  unsigned int* d_order = 0;
  TRY(id_ranks(h, &d_order));
  double* stage = h->Palt.q;
  particles_to_aos<<<(int) ((np + 255) / 256), 256, 0, h->stream>>>(h->P, (long) np, d_order, stage);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  CU(cudaMemcpyAsync(aos11, stage, np * 11 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (d_order) cudaFree(d_order);
  return 0;
}

/* Solver::bunchSample's sums (solver.cpp:1582-1608) over the particles of this slab, reduced on the device           */
extern "C" int mithra_gpu_bunch_moments (MithraGpu* h, double sums[13])
{
  USE(h);
  if (!sums) return fail("mithra_gpu_bunch_moments: null argument");
  for (int q = 0; q < MITHRA_MOMENTS; q++) sums[q] = 0.0;
  if (h->pn == 0) return 0;
  const int blocks = grid_for((long) h->pn, 256, h->num_sms * 4);
  double* d_part = 0;
  CU(cudaMalloc(&d_part, (size_t) blocks * MITHRA_MOMENTS * sizeof(double)));
  bunch_moments<<<blocks, 256, 0, h->stream>>>(h->bd, h->P, (long) h->pn, d_part);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  std::vector<double> part((size_t) blocks * MITHRA_MOMENTS);
  CU(cudaMemcpyAsync(part.data(), d_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaFree(d_part));
  for (int bl = 0; bl < blocks; bl++)
    for (int q = 0; q < MITHRA_MOMENTS; q++) sums[q] += part[(size_t) bl * MITHRA_MOMENTS + q];
  return 0;
}

/* FdTd::fieldSample / FdTdSC::fieldSample (fdtd.cpp:851-950, fdtdSC.cpp:1147-1250): interpolated E, B, A^n at n points  */
extern "C" int mithra_gpu_field_sample (MithraGpu* h, const double* pos3, size_t n, double* out9, unsigned char* mine)
{
  USE(h);
  if (n == 0) return 0;
  if (!pos3 || !out9 || !mine) return fail("mithra_gpu_field_sample: null argument");
  double* d_pos = 0; double* d_out = 0; unsigned char* d_mine = 0;
  CU(cudaMalloc(&d_pos, n * 3 * sizeof(double))); CU(cudaMalloc(&d_out, n * 9 * sizeof(double))); CU(cudaMalloc(&d_mine, n));
  CU(cudaMemcpyAsync(d_pos, pos3, n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const int grid = (int) ((n + 63) / 64);
  if (h->fd.ncomp == 4) field_sample<true ><<<grid, 64, 0, h->stream>>>(h->fd, h->bd, h->A[h->ip1], h->A[h->in], h->eb, d_pos, (int) n, d_out, d_mine);
  else                  field_sample<false><<<grid, 64, 0, h->stream>>>(h->fd, h->bd, h->A[h->ip1], h->A[h->in], h->eb, d_pos, (int) n, d_out, d_mine);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  CU(cudaMemcpyAsync(out9, d_out, n * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(mine, d_mine, n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  cudaFree(d_pos); cudaFree(d_out); cudaFree(d_mine);
  return 0;
}

/* E, B, A^n at a list of nodes for the field visualisation writers (fdtd.cpp:1128-1540)                              */
extern "C" int mithra_gpu_field_nodes (MithraGpu* h, const int* ijk3, size_t n, double* out9, unsigned char* mine)
{
  USE(h);
  if (n == 0) return 0;
  if (!ijk3 || !out9 || !mine) return fail("mithra_gpu_field_nodes: null argument");
  int* d_ijk = 0; double* d_out = 0; unsigned char* d_mine = 0;
  CU(cudaMalloc(&d_ijk, n * 3 * sizeof(int))); CU(cudaMalloc(&d_out, n * 9 * sizeof(double))); CU(cudaMalloc(&d_mine, n));
  CU(cudaMemcpyAsync(d_ijk, ijk3, n * 3 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  const int grid = (int) ((n + 127) / 128);
  if (h->fd.ncomp == 4) field_nodes<true ><<<grid, 128, 0, h->stream>>>(h->fd, h->A[h->ip1], h->A[h->in], h->eb, d_ijk, (long) n, d_out, d_mine);
  else                  field_nodes<false><<<grid, 128, 0, h->stream>>>(h->fd, h->A[h->ip1], h->A[h->in], h->eb, d_ijk, (long) n, d_out, d_mine);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  CU(cudaMemcpyAsync(out9, d_out, n * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(mine, d_mine, n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  cudaFree(d_ijk); cudaFree(d_out); cudaFree(d_mine);
  return 0;
}

extern "C" int mithra_gpu_particle_cells (MithraGpu* h, long* push_m, int* ijk6, size_t capacity)
{
  USE(h);
  if (capacity < h->pn) return fail("mithra_gpu_particle_cells: capacity %zu < %zu particles", capacity, h->pn);
  if (h->pn == 0) return 0;
  const size_t n = h->pn;
  long* dm = 0; int* dd = 0;
  if (push_m) CU(cudaMalloc(&dm, n * sizeof(long)));
  if (ijk6)   CU(cudaMalloc(&dd, n * 6 * sizeof(int)));
  particle_cells<<<(int) ((n + 255) / 256), 256, 0, h->stream>>>(h->bd, h->P, (long) n, dm, dd);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  CU(cudaStreamSynchronize(h->stream));
  /* device order -> order of the upload indices                                                              */
  std::vector<unsigned int> ids(n), order(n);
  CU(cudaMemcpy(ids.data(), h->P.id, n * sizeof(unsigned int), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; i++) order[i] = (unsigned int) i;
  if (!h->ids_dense) std::sort(order.begin(), order.end(), [&](unsigned int a, unsigned int b) { return ids[a] < ids[b]; });
  else for (size_t i = 0; i < n; i++) order[ids[i]] = (unsigned int) i;
  if (push_m)
    {
      std::vector<long> t(n); CU(cudaMemcpy(t.data(), dm, n * sizeof(long), cudaMemcpyDeviceToHost)); cudaFree(dm);
      for (size_t i = 0; i < n; i++) push_m[i] = t[order[i]];
    }
  if (ijk6)
    {
      std::vector<int> t(n * 6); CU(cudaMemcpy(t.data(), dd, n * 6 * sizeof(int), cudaMemcpyDeviceToHost)); cudaFree(dd);
      for (size_t i = 0; i < n; i++) memcpy(ijk6 + 6 * i, t.data() + 6 * (size_t) order[i], 6 * sizeof(int));
    }
  return 0;
}

extern "C" int mithra_gpu_num_particles (MithraGpu* h, size_t* n) { USE(h); if (n) *n = h->pn; return 0; }

extern "C" int mithra_gpu_set_time (MithraGpu* h, double time, double time_bunch, unsigned int n_time)
{
  USE(h);
  h->time = time; h->timem1 = time - h->prm.dt; h->timep1 = time + h->prm.dt; h->time_bunch = time_bunch; h->n_time = n_time;
  return 0;
}

extern "C" int mithra_gpu_get_time (MithraGpu* h, double* time, double* time_bunch, unsigned int* n_time)
{
  USE(h);
  if (time) *time = h->time; if (time_bunch) *time_bunch = h->time_bunch; if (n_time) *n_time = h->n_time;
  return 0;
}

/* ---------------------------------------------------------------------------------------------------- */
/* the time march                                                                                        */

/* the pencil mask that bounds the J of the last deposit (kernels_field.cuh source_planes), or 0: the box alone            */
static const unsigned char* source_mask (const MithraGpu* h)
{
  static const bool off = getenv("MITHRA_NO_JMASK") != 0;
  return (h->jmask_valid && !off) ? h->d_emask_nodes[h->jmask_buf] : 0;
}

/* bulk-async plane pipeline (kernels_field.cuh stencil_stream); false when its stages do not fit in shared memory.
 * faces: the variant that also does the x and y absorbing faces (no rim_update afterwards)                         */
template <bool NSFD, bool FACES, int T, int NB, bool SEED>
static bool launch_stencil_stream_t (MithraGpu* h, bool skiprim, const RimDev& rz)
{
  const FieldDev& f = h->fd;
  constexpr int slot = (T == 512) ? 1 : SEED ? 2 : 0;
  static const int KC = getenv("MITHRA_STENCIL_KC") ? std::min(64, std::max(1, atoi(getenv("MITHRA_STENCIL_KC")))) : 64;   /* <= 64: source_planes */
  if (FACES)
    {
      /* a lane of the face warp for every node next to a y face                                                   */
      if (h->face_nodes[slot] < 0) h->face_nodes[slot] = stencil_stream_face_nodes(f.N0, f.N1, T, SEED);
      if (h->face_nodes[slot] > 32) return false;
    }
  const size_t smem = stencil_stream_smem(T, f.N1, NB, FACES);
  if (smem > (size_t) (FACES ? 112 : 200) * 1024) return false;      /* with the faces: two CTAs per SM or not at all     */
  /* the attribute belongs to the device: one process may drive several (one handle per slab)                */
  if (!h->stream_configured[NSFD][FACES][slot])
    {
      if (cudaFuncSetAttribute((const void*) stencil_stream<NSFD, T, NB, FACES, SEED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) { cudaGetLastError(); return false; }
      h->stream_configured[NSFD][FACES][slot] = true;
    }
  dim3 grid((unsigned) ((f.P + T - 1) / T), (unsigned) ((f.np - 1 - f.kb + KC - 1) / KC), (unsigned) f.ncomp);
  stencil_stream<NSFD, T, NB, FACES, SEED><<<grid, T + (FACES ? 64 : 32), smem, h->stream>>>(f, h->A[h->ip1], h->A[h->in], h->A[h->im1], h->J, h->d_jbox, KC, skiprim ? 1 : 0, source_mask(h), rz);
  return true;
}

/* Tile sizes: the consumer loop must not spill (a reload from local memory in it costs more than a plane of the march), and
 * two CTAs share an SM: 480 consumers + the producer warp leave 64 registers per thread, so do 448 + the producer and
 * the face warp; with a seed the face warp carries the seed scalars of the next plane as well: 384 consumers, 72 registers.
 * MITHRA_STREAM_T=512: the 512-node tiles of round 1 (56 registers, a few spilled words) for comparison. */
template <bool NSFD, bool FACES>
static bool launch_stencil_stream_as (MithraGpu* h, bool skiprim, const RimDev& rz)
{
  if (getenv("MITHRA_STENCIL_PLAIN")) return false;
  static const int tsel = getenv("MITHRA_STREAM_T") ? atoi(getenv("MITHRA_STREAM_T")) : 0;
  if constexpr (!FACES)
    {
      if (tsel == 512) return launch_stencil_stream_t<NSFD, false, 512, 8, false>(h, skiprim, rz);
      return launch_stencil_stream_t<NSFD, false, 480, 8, false>(h, skiprim, rz);
    }
  else
    {
      if (rz.seed) return launch_stencil_stream_t<NSFD, true, 384, 8, true>(h, skiprim, rz);
      return launch_stencil_stream_t<NSFD, true, 448, 8, false>(h, skiprim, rz);
    }
}

/* the interior sweep; *faces: in, the x / y faces may be fused into it (a rim path without seed); out, they were.
 * The plain-load interior kernel always covers every interior node (rim_update afterwards simply rewrites the rim) */
template <bool NSFD>
static void launch_stencil (MithraGpu* h, bool skiprim, bool* faces, const RimDev& rz)
{
  h->j_zeroed_by_update = true;
  if (*faces && launch_stencil_stream_as<NSFD, true>(h, false, rz)) return;
  *faces = false;
  if (launch_stencil_stream_as<NSFD, false>(h, skiprim, rz)) return;
  h->j_zeroed_by_update = false;
  const FieldDev& f = h->fd;
  constexpr int BX = 128, KC = 32;
  dim3 grid((f.P + BX - 1) / BX, (f.np - 1 - f.kb + KC - 1) / KC, f.ncomp);
  stencil_interior<NSFD, BX, KC><<<grid, BX, 0, h->stream>>>(f, h->A[h->ip1], h->A[h->in], h->A[h->im1], h->J, h->d_jbox, source_mask(h));
}

/* per-plane seed table and the line table rim_update reads, for the time level `time`, on stream `st`           */
static int launch_seed_lines (MithraGpu* h, double time, cudaStream_t st)
{
  const FieldDev& f = h->fd;
  const bool zlo = (f.rank == 0), zhi = (f.rank == f.size - 1);
  const int KI = zlo ? 2 : f.kb, KF = zhi ? f.np - 2 : f.np - 1;
  if (h->d_seed_tab) { seed_plane_table<<<(f.np + 127) / 128, 128, 0, st>>>(h->d_seed, f.np, time, h->d_seed_tab); h->cnt.kernel_launches += 1; }
  const long tot = (long) (4 * (f.N1 - 4) + 4 * (f.N0 - 4)) * (KF - KI);
  seed_lines<<<grid_for(tot, 128, h->num_sms * 16), 128, 0, st>>>(h->d_seed, h->d_seed_tab, f, KI, KF, h->seed_L, h->d_seedu, time);
  h->cnt.kernel_launches += 1;
  CU(cudaGetLastError());
  return 0;
}

/* FdTd::fieldUpdate in two halves: the potentials (stencil, rim, boundaries, ghost exchange: no particle involved) and the
 * E/B evaluation for the bunch (needs the particle box and masks).  mithra_gpu_step enqueues the first half of the NEXT
 * step before it blocks in the particle hand-over between slabs, so the device is not idle while the host waits.        */
static int field_update_potentials (MithraGpu* h)
{
  USE(h);
  const FieldDev& f = h->fd;
  double* ap = h->A[h->ip1]; const double* a = h->A[h->in]; const double* am = h->A[h->im1];
  const bool zlo = (f.rank == 0), zhi = (f.rank == f.size - 1);

  /* Rim path (N0, N1, np >= 8): rim_update owns the two outer node layers in x and y -- interior value, x / y shell
   * seed terms (from the line table) and the x / y faces in one pass; what remains afterwards is the z shell, the z
   * faces, the edges and the corners.  MITHRA_NO_FUSE keeps the reference's three passes apart (the parity tests
   * compare the two bit for bit).                                                                              */
  const bool rim = f.N0 >= 8 && f.N1 >= 8 && f.np >= 8 && !getenv("MITHRA_NO_FUSE");
  /* Without a seed stencil_stream does the y faces itself and one pass over whole rows the x faces (MITHRA_NO_FACEWARP:
   * stencil_stream on the inner nodes + rim_update, as for seeded jobs).  With a seed the same split works -- the face warp
   * owns the four y-shell nodes of every row and applies their seed terms, seed_xshell_rows the x-shell terms of the rows
   * in between -- and is bit-identical, but it LOSES on FEL-SEEDED (sweep 1.58 ms against 0.76 + 0.23 of rim_update saved:
   * the face warp needs 72 registers, so 384-node tiles, a third fewer bytes in flight per SM, and it is the slowest
   * consumer of every stage): opt-in with MITHRA_SEEDWARP=1, kept for the parity tests                               */
  /* The face warp pays where the rim is a noticeable part of the plane: rim_update costs about ten times its share of the
   * nodes (FEL-LCLS, 102 x 102: 3.9 % of the nodes on the perimeter, 1.64 of 6.4 ms), the face warp a fixed 10-15 % of the
   * sweep (448- instead of 480-node tiles, one more warp per CTA).  FEL-ICS (402 x 402, 1 %): 7.2 ms per step with
   * rim_update, 7.6 with the face warp.  MITHRA_FACEWARP=1 forces it.                                                  */
  const bool narrow = 4.0 * (f.N0 + f.N1) >= 0.02 * (double) f.N0 * f.N1 || getenv("MITHRA_FACEWARP");
  bool faces = rim && narrow && !getenv("MITHRA_NO_FACEWARP") && (!h->d_seed || getenv("MITHRA_SEEDWARP"));
  RimDev rz; memset(&rz, 0, sizeof(rz));
  if (rim && h->d_seed)
    {
      PhaseTimer t(h, PH_BOUNDARY);
      rz.seed = 1; rz.seedu = h->d_seedu; rz.L = h->seed_L;
      rz.KI = zlo ? 2 : f.kb; rz.KF = zhi ? f.np - 2 : f.np - 1;
      const MithraBeam& B = h->prm.seed;
      rz.supergaussian = (B.seed_type == MITHRA_BEAM_SUPERGAUSSIAN) ? 1 : 0;
      rz.ni = rz.supergaussian ? ( 2 * B.order[0] + 1 ) * ( 2 * B.order[1] + 1 ) : 1;
      rz.pol[0] = B.polarization[0]; rz.pol[1] = B.polarization[1]; rz.pol[2] = B.polarization[2]; rz.gamma = h->prm.gamma;
      if (h->seed_ahead && h->seed_ahead_time == h->time)
	CU(cudaStreamWaitEvent(h->stream, h->ev_seed, 0));        /* computed beside the last particle phase      */
      else
	{
	  if (h->seed_ahead) CU(cudaStreamWaitEvent(h->stream, h->ev_seed, 0));      /* stale: let it finish first */
	  TRY(launch_seed_lines(h, h->time, h->stream));
	}
      h->seed_ahead = false;
    }
  {
    PhaseTimer t(h, PH_STENCIL);
    if (f.nsfd) launch_stencil<true>(h, rim, &faces, rz); else launch_stencil<false>(h, rim, &faces, rz);
    CU(cudaGetLastError());
    h->cnt.kernel_launches += 1;
  }
  {
    PhaseTimer t(h, PH_BOUNDARY);
    if (rim)
      {
	if (faces && rz.seed)
	  {
	    const long tot = 4L * (f.N1 - 6) * (rz.KF - rz.KI) * 3;
	    seed_xshell_rows<<<grid_for(tot, 128, h->num_sms * 8), 128, 0, h->stream>>>(f, rz, ap);
	    h->cnt.kernel_launches += 1;
	  }
	if (faces)
	  {
	    /* stencil_stream has done the y faces; the x faces are whole rows: one coalesced pass                       */
	    const long nface = 2L * (f.N1 - 2) * (f.np - 1 - f.kb) * f.ncomp;
	    boundary_faces<<<grid_for(nface, 256, h->num_sms * 8), 256, 0, h->stream>>>(f, ap, a, am, 2);
	    h->cnt.kernel_launches += 1;
	  }
	if (!faces)
	  {
	const int per = 4 * (f.N1 - 2) + 4 * (f.N0 - 6);
	static const int KC = getenv("MITHRA_RIM_KC") ? std::min(64, std::max(1, atoi(getenv("MITHRA_RIM_KC")))) : 64;   /* planes per CTA (fitting the grid to whole waves changes nothing: measured) */
	dim3 grid((unsigned) ((per + 127) / 128), (unsigned) ((f.np - 1 - f.kb + KC - 1) / KC), (unsigned) f.ncomp);
	if (f.nsfd) rim_update<true ><<<grid, 128, 0, h->stream>>>(f, rz, ap, a, am, h->J, h->d_jbox, KC, source_mask(h), h->j_zeroed_by_update ? 1 : 0);
	else        rim_update<false><<<grid, 128, 0, h->stream>>>(f, rz, ap, a, am, h->J, h->d_jbox, KC, source_mask(h), h->j_zeroed_by_update ? 1 : 0);
	h->cnt.kernel_launches += 1;
	  }
	if (h->d_seed && (zlo || zhi))
	  {
	    const long tot = 4L * (f.N0 - 4) * (f.N1 - 4);
	    seed_inject_zshell<<<grid_for(tot, 128, h->num_sms * 4), 128, 0, h->stream>>>(h->d_seed, h->d_seed_tab, f, ap, h->time);
	    h->cnt.kernel_launches += 1;
	  }
	if (zlo || zhi)
	  {
	    const long nface = 2L * (f.N0 - 2) * (f.N1 - 2) * f.ncomp;
	    boundary_faces<<<grid_for(nface, 256, h->num_sms * 8), 256, 0, h->stream>>>(f, ap, a, am, 1);
	    h->cnt.kernel_launches += 1;
	  }
      }
    else
      {
	if (h->d_seed)
	  {
	    TRY(seed_inject(h->d_seed, h->d_seed_tab, f, ap, h->time, h->stream, h->num_sms));
	    h->cnt.kernel_launches += h->d_seed_tab ? 2 : 1;
	  }
	const long nface = (2L * (f.N1 - 2) * (f.np - 2) + 2L * (f.N0 - 2) * (f.np - 2) + 2L * (f.N0 - 2) * (f.N1 - 2)) * f.ncomp;
	boundary_faces<<<grid_for(nface, 256, h->num_sms * 8), 256, 0, h->stream>>>(f, ap, a, am, 0);
	h->cnt.kernel_launches += 1;
      }
    /* J, its box and the seed tables have had their last reader: housekeeping_ahead may clear / refill them from here
     * on, beside the edges, the ghost exchange and the E/B evaluation                                              */
    if (h->overlap && !h->profiling) { CU(cudaEventRecord(h->ev_main, h->stream)); h->ev_main_fresh = true; }
    if (f.order == 2)
      {
	const long nedge = (4L * (f.np - 2) + 4L * (f.N0 - 2) + 4L * (f.N1 - 2)) * f.ncomp;
	boundary_edges<<<grid_for(nedge, 256, h->num_sms * 4), 256, 0, h->stream>>>(f, ap, a, am);
	boundary_corners<<<1, 32, 0, h->stream>>>(f, ap, a, am);
	h->cnt.kernel_launches += 2;
      }
    CU(cudaGetLastError());
    if (f.size > 1)
      {
	if (!h->xch.connected) return fail("mithra_gpu_field_update: slab %d of %d is not connected to its neighbours (mithra_gpu_ipc_connect)", f.rank, f.size);
	if (exchange_potentials(h->xch, f, ap, h->ip1, h->stream, h->num_sms, &h->cnt.kernel_launches)) return fail("A ghost exchange: %s", h->xch.error.c_str());
      }
  }
  h->anp1_is_current = false;
  h->cnt.cell_updates += (unsigned long long) (f.np - 1 - f.kb + (f.rank == 0 ? 1 : 0) + (f.rank == f.size - 1 ? 1 : 0)) * f.P;
  return 0;
}

static int field_update_eb (MithraGpu* h)
{
  USE(h);
  const FieldDev& f = h->fd;
  double* ap = h->A[h->ip1]; const double* a = h->A[h->in];
  const bool zlo = (f.rank == 0), zhi = (f.rank == f.size - 1);
  {
    PhaseTimer t(h, PH_EVAL);
    const double cdt = h->prm.c0 * h->prm.dt;
    const int padx = (int) ceil(cdt / h->prm.dx), pady = (int) ceil(cdt / h->prm.dy), padz = (int) ceil(cdt / h->prm.dz);
    make_eb_box<<<1, 1, 0, h->stream>>>(h->d_pbox, h->d_ebox, (Box*) 0, padx, pady, padz, f.N0, f.N1, f.np);
    if (getenv("MITHRA_EB_BOX"))
      {
	/* the node-at-a-time kernel over the whole box (the parity tests compare the two bit for bit)                  */
	if (f.ncomp == 4) eval_eb_box<true ><<<h->num_sms * 4, 256, 0, h->stream>>>(f, ap, a, h->eb, h->d_ebox, 0);
	else              eval_eb_box<false><<<h->num_sms * 4, 256, 0, h->stream>>>(f, ap, a, h->eb, h->d_ebox, 0);
      }
    else
      {
	/* only the node pencils within some particle's reach (device_types.cuh "Reach mask")                          */
	const unsigned char* mask = 0;
	const bool masked = reach_mask_usable(h);
	if (masked)
	  {
	    /* into the buffer the source mask of this step's stencil and clear is NOT (they may still be reading it)  */
	    h->emask_cur ^= 1;
	    unsigned char* nodes = h->d_emask_nodes[h->emask_cur];
	    CU(cudaMemsetAsync(nodes, 0, h->emask_bytes, h->stream));
	    spread_eb_mask<<<h->num_sms * 8, 256, 0, h->stream>>>(f, h->d_emask_cells, nodes, (long) h->emask_bytes);
	    h->cnt.kernel_launches += 1;
	    mask = nodes;
	    h->emask_fresh = true; h->pushes_since_spread = 0;
	  }
	else h->emask_fresh = false;
	if (f.ncomp == 4) eval_eb_march<true ><<<h->num_sms * 8, 256, 0, h->stream>>>(f, ap, a, h->eb, h->d_ebox, mask);
	else              eval_eb_march<false><<<h->num_sms * 8, 256, 0, h->stream>>>(f, ap, a, h->eb, h->d_ebox, mask);
	if (zlo || zhi)
	  {
	    /* planes 0 / np-1 of the global ends copy planes 1 / np-2 (fdtd.cpp:754-773)                               */
	    if (f.ncomp == 4) eval_eb_box<true ><<<16, 256, 0, h->stream>>>(f, ap, a, h->eb, h->d_ebox, 1);
	    else              eval_eb_box<false><<<16, 256, 0, h->stream>>>(f, ap, a, h->eb, h->d_ebox, 1);
	    h->cnt.kernel_launches += 1;
	  }
      }
    CU(cudaGetLastError());
    h->cnt.kernel_launches += 2;
    if (f.size > 1)
      {
	/* whole planes the neighbours gather from and sample power on: kb -> prev; np-3, np-2 -> next           */
	if (h->xch.prev.chain)
	  {
	    if (f.ncomp == 4) eval_eb_box<true ><<<h->num_sms, 256, 0, h->stream>>>(f, ap, a, h->eb, h->xch.d_planes_lo, 0);
	    else              eval_eb_box<false><<<h->num_sms, 256, 0, h->stream>>>(f, ap, a, h->eb, h->xch.d_planes_lo, 0);
	    h->cnt.kernel_launches += 1;
	  }
	if (h->xch.next.chain)
	  {
	    if (f.ncomp == 4) eval_eb_box<true ><<<h->num_sms, 256, 0, h->stream>>>(f, ap, a, h->eb, h->xch.d_planes_hi, 0);
	    else              eval_eb_box<false><<<h->num_sms, 256, 0, h->stream>>>(f, ap, a, h->eb, h->xch.d_planes_hi, 0);
	    h->cnt.kernel_launches += 1;
	  }
	CU(cudaGetLastError());
	if (exchange_eb(h->xch, f, h->eb, h->stream, h->num_sms, &h->cnt.kernel_launches)) return fail("E/B ghost exchange: %s", h->xch.error.c_str());
      }
  }
  return 0;
}

extern "C" int mithra_gpu_field_update (MithraGpu* h)
{
  TRY(field_update_potentials(h));
  return field_update_eb(h);
}

/* Counting sort of the bunch by cell (kernels_sort.cuh); the particle box of the last push / upload bounds the keys. */
extern "C" int mithra_gpu_sort_particles (MithraGpu* h)
{
  USE(h);
  h->steps_since_sort = 0;
  if (h->pn < 2) return 0;
  const long n = (long) h->pn, cap = h->sort_cap;
  const int pgrid = (int) ((n + 255) / 256);
  const int cgrid = (int) ((cap + MITHRA_SCAN_CHUNK - 1) / MITHRA_SCAN_CHUNK);
  sort_zero<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_pbox, cap, h->d_hist);
  sort_count<<<pgrid, 256, 0, h->stream>>>(h->bd, h->P, n, h->d_pbox, cap, h->d_hist, h->d_key, h->d_rank);
  scan_chunk_sums<<<cgrid, 256, 0, h->stream>>>(h->d_pbox, cap, h->d_hist, h->d_sums);
  scan_sums<<<1, 1024, 0, h->stream>>>(h->d_pbox, cap, h->d_sums);
  scan_chunks<<<cgrid, 256, 0, h->stream>>>(h->d_pbox, cap, h->d_hist, h->d_sums);
  sort_permute<<<pgrid, 256, 0, h->stream>>>(h->P, h->Palt, n, h->d_pbox, cap, h->d_hist, h->d_key, h->d_rank);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 6;
  std::swap(h->P, h->Palt);
  return 0;
}

static ScreensDev screens_dev (MithraGpu* h, double time_bunch)
{
  ScreensDev S;
  S.n = h->prm.screens.N; S.pos = h->d_scr_pos; S.rec = h->d_scr_rec; S.cursor = h->d_scr_cur; S.capacity = h->scr_cap;
  S.step_id = (double) h->n_time; S.time_bunch = time_bunch;
  return S;
}

extern "C" int mithra_gpu_bunch_update (MithraGpu* h)
{
  USE(h);
  PhaseTimer t(h, PH_PUSH);
  const int nsub = h->prm.n_update_bunch;
  {
    /* MithraGpuParams.sort_interval: > 0 field steps between two sorts, < 0 never, 0 = every 16 steps for bunches
     * large enough for the order to matter                                                                   */
    const int every = h->sort_interval > 0 ? h->sort_interval : (h->sort_interval == 0 && h->pn >= 4096 ? 16 : 0);
    if (every > 0 && h->steps_since_sort >= every) TRY(mithra_gpu_sort_particles(h));
    if (h->steps_since_sort < (1 << 30)) ++h->steps_since_sort;
  }
  if (h->pn > 0)
    {
      /* the particle box is rebuilt by the push (it is read by the next field update)                      */
      set_box<<<1, 1, 0, h->stream>>>(h->d_pbox, 0x7fffffff, 0x7fffffff, 0x7fffffff, -1, -1, -1);
      if (h->d_emask_cells) CU(cudaMemsetAsync(h->d_emask_cells, 0, h->emask_bytes, h->stream));
      const int grid = (int) ((h->pn + 127) / 128);
      bool beams = h->bd.n_ext > 0;
      for (int u = 0; u < h->bd.n_und; u++) if (h->bd.und[u].type != MITHRA_UNDULATOR_STATIC) beams = true;
      /* mithra_gpu_step lets the push test the screens on its way out (h->fuse_screens); the time is the one
       * mithra_gpu_screen_profile would see: the sub-steps added one by one                                       */
      ScreensDev scr; memset(&scr, 0, sizeof(scr));
      if (h->fuse_screens && h->d_scr_pos)
	{
	  double ta = h->time_bunch; for (int s = 0; s < nsub; s++) ta += h->prm.dt_bunch;
	  scr = screens_dev(h, ta);
	}
      if (beams) push_particles<true ><<<grid, 128, 0, h->stream>>>(h->bd, h->P, (long) h->pn, h->eb, h->time_bunch, nsub, 1, h->d_pbox, h->d_noutside, h->d_emask_cells, scr);
      else       push_particles<false><<<grid, 128, 0, h->stream>>>(h->bd, h->P, (long) h->pn, h->eb, h->time_bunch, nsub, 1, h->d_pbox, h->d_noutside, h->d_emask_cells, scr);
      h->cnt.kernel_launches += 2;
      CU(cudaGetLastError());
      ++h->pushes_since_spread;
    }
  for (int s = 0; s < nsub; s++) { h->time_bunch += h->prm.dt_bunch; ++h->n_time_bunch; }
  h->cnt.particle_pushes += (unsigned long long) h->pn * nsub;
  return 0;
}

extern "C" int mithra_gpu_screen_profile (MithraGpu* h)
{
  USE(h);
  if (!h->d_scr_pos || h->pn == 0) return 0;
  PhaseTimer t(h, PH_SCREEN);
  screen_cross<<<(int) ((h->pn + 255) / 256), 256, 0, h->stream>>>(h->bd, h->P, (long) h->pn, screens_dev(h, h->time_bunch));
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  return 0;
}

static int flush_power_rows (MithraGpu* h)
{
  if (h->rows_used == 0) return 0;
  const size_t w = (size_t) h->pw.N * h->pw.Nl;
  const size_t old = h->h_rows.size();
  h->h_rows.resize(old + h->rows_used * w);
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(h->h_rows.data() + old, h->d_rows, h->rows_used * w * sizeof(double), cudaMemcpyDeviceToHost));
  h->rows_used = 0;
  return 0;
}

extern "C" int mithra_gpu_power_sample (MithraGpu* h)
{
  USE(h);
  if (!h->d_fdt) return 0;
  PhaseTimer t(h, PH_POWER);
  const FieldDev& f = h->fd;
  if (h->rows_used == h->rows_cap) TRY(flush_power_rows(h));
  const int slot = (int) (h->n_time % (unsigned int) h->pw.Nf);
  const double* np1 = h->A[h->ip1]; const double* an = h->A[h->in];
  for (int k = 0; k < h->pw.N; k++)
    {
      if (!h->power_mine[k]) continue;
      if (f.ncomp == 4)
	power_dft<true ><<<h->power_blocks, MITHRA_POWER_PX * MITHRA_POWER_MS, 0, h->stream>>>(f, h->pw, np1, an, h->eb, h->d_fdt, h->d_ep, k, h->power_k[k], h->power_dzr[k], slot, h->d_partial);
      else
	power_dft<false><<<h->power_blocks, MITHRA_POWER_PX * MITHRA_POWER_MS, 0, h->stream>>>(f, h->pw, np1, an, h->eb, h->d_fdt, h->d_ep, k, h->power_k[k], h->power_dzr[k], slot, h->d_partial);
      h->cnt.kernel_launches += 1;
    }
  const int nout = h->pw.N * h->pw.Nl;
  power_finish<<<(nout + 63) / 64, 64, 0, h->stream>>>(h->pw, h->d_partial, h->power_blocks, h->d_rows + h->rows_used * nout);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  h->rows_used++;
  return 0;
}

extern "C" int mithra_gpu_power_visualize (MithraGpu* h)
{
  USE(h);
  if (!h->prm.power_map.enabled || !h->pm_mine) return 0;
  PhaseTimer t(h, PH_POWER);
  const FieldDev& f = h->fd;
  const int slot = (int) (h->n_time % (unsigned int) h->pm.Nf);
  const int npx = (f.N0 - 2) * (f.N1 - 2);
  const int k = h->pm_k;                                 /* internal plane index (FieldDev.k0 is internal)          */
  if (f.ncomp == 4) power_map<true ><<<(npx + 127) / 128, 128, 0, h->stream>>>(f, h->pm, h->A[h->ip1], h->A[h->in], h->eb, h->d_pm_fdt, h->d_pm_ep, k, h->pm_dzr, slot, h->d_pm_pL);
  else              power_map<false><<<(npx + 127) / 128, 128, 0, h->stream>>>(f, h->pm, h->A[h->ip1], h->A[h->in], h->eb, h->d_pm_fdt, h->d_pm_ep, k, h->pm_dzr, slot, h->d_pm_pL);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  return 0;
}

extern "C" int mithra_gpu_fetch_power_map (MithraGpu* h, double* pL, size_t capacity, int* mine)
{
  USE(h);
  if (mine) *mine = (h->prm.power_map.enabled && h->pm_mine) ? 1 : 0;
  if (!h->prm.power_map.enabled || !h->pm_mine || !pL) return 0;
  if (capacity < (size_t) h->fd.P) return fail("mithra_gpu_fetch_power_map: capacity %zu < %d pixels", capacity, h->fd.P);
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(pL, h->d_pm_pL, (size_t) h->fd.P * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int mithra_gpu_field_shift (MithraGpu* h)
{
  USE(h);
  const int t = h->im1; h->im1 = h->in; h->in = h->ip1; h->ip1 = t;
  h->anp1_is_current = true;
  return 0;
}

extern "C" int mithra_gpu_current_reset (MithraGpu* h)
{
  USE(h);
  PhaseTimer t(h, PH_CLEAR);
  if (h->clear_ahead)
    {
      /* mithra_gpu_step cleared the box beside the particle kernels: the deposit only has to wait for it          */
      CU(cudaStreamWaitEvent(h->stream, h->ev_clear, 0));
      h->clear_ahead = false;
      return 0;
    }
  clear_current_box<<<h->j_zeroed_by_update ? h->num_sms : h->num_sms * 8, 256, 0, h->stream>>>(h->fd, h->J, h->d_jbox, h->d_done, source_mask(h), h->j_zeroed_by_update ? 1 : 0);
  h->jmask_valid = false; h->j_empty = true;             /* J is zero: whatever is put there next decides            */
  h->j_zeroed_by_update = false;
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  return 0;
}

/* After the field update of a step driven by mithra_gpu_step: J has been consumed and nothing reads it or the seed
 * tables before the next deposit / field update, so clear the box and prepare the next seed lines on the side stream. */
static int housekeeping_ahead (MithraGpu* h)
{
  if (!h->overlap || h->profiling) return 0;
  const FieldDev& f = h->fd;
  if (!h->ev_main_fresh) CU(cudaEventRecord(h->ev_main, h->stream));
  h->ev_main_fresh = false;
  CU(cudaStreamWaitEvent(h->side, h->ev_main, 0));
  clear_current_box<<<h->j_zeroed_by_update ? h->num_sms : h->num_sms * 8, 256, 0, h->side>>>(f, h->J, h->d_jbox, h->d_done, source_mask(h), h->j_zeroed_by_update ? 1 : 0);
  h->jmask_valid = false; h->j_empty = true; h->j_zeroed_by_update = false;
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  CU(cudaEventRecord(h->ev_clear, h->side));
  h->clear_ahead = true;
  if (h->d_seed && h->d_seedu && f.N0 >= 8 && f.N1 >= 8 && f.np >= 8 && !getenv("MITHRA_NO_FUSE"))
    {
      const double tnext = h->time + h->prm.dt;              /* exactly what mithra_gpu_advance_time will make of it */
      TRY(launch_seed_lines(h, tnext, h->side));
      CU(cudaEventRecord(h->ev_seed, h->side));
      h->seed_ahead = true; h->seed_ahead_time = tnext;
    }
  return 0;
}

extern "C" int mithra_gpu_current_update (MithraGpu* h)
{
  USE(h);
  PhaseTimer t(h, PH_DEPOSIT);
  if (h->pn == 0) return 0;
  /* Does the node mask of this step's E/B evaluation bound what is deposited now?  It does when the particles are the
   * ones it was spread from, moved by exactly one push (rm and r both within the padding of the cell at the start of
   * the step), and J was empty or bounded by the same mask before.                                                  */
  {
    const bool bounded = h->d_emask_nodes[0] && h->emask_fresh && h->pushes_since_spread == 1;
    if (h->j_empty) { h->jmask_valid = bounded; h->jmask_buf = h->emask_cur; }
    else if (!(h->jmask_valid && bounded && h->jmask_buf == h->emask_cur)) h->jmask_valid = false;
    h->j_empty = false;
  }
  static const int run = getenv("MITHRA_DEP_RUN") ? std::max(1, atoi(getenv("MITHRA_DEP_RUN"))) : MITHRA_DEP_RUN;
  const int grid = (int) ((h->pn + (size_t) 128 * run - 1) / ((size_t) 128 * run));
  if (h->fd.ncomp == 4) deposit_current<true ><<<grid, 128, 0, h->stream>>>(h->bd, h->P, (long) h->pn, h->J, h->d_jbox, run);
  else                  deposit_current<false><<<grid, 128, 0, h->stream>>>(h->bd, h->P, (long) h->pn, h->J, h->d_jbox, run);
  CU(cudaGetLastError());
  h->cnt.kernel_launches += 1;
  return 0;
}

extern "C" int mithra_gpu_current_communicate (MithraGpu* h)
{
  USE(h);
  if (h->fd.size > 1)
    {
      if (!h->xch.connected) return fail("mithra_gpu_current_communicate: slab is not connected to its neighbours");
      if (exchange_current(h->xch, h->fd, h->J, h->d_jbox, h->stream, h->num_sms, &h->cnt.kernel_launches)) return fail("J merge: %s", h->xch.error.c_str());
    }
  return 0;
}

/* Particle hand-over between slabs, once per field step after the deposit (solver.cpp:1544-1568, 493-503 and the
 * purge of fdtd.cpp:214-224).  _begin enqueues the sends, _end receives (one host synchronisation) -- split so that
 * one process driving several slabs can issue every _begin before the first _end.                          */
extern "C" int mithra_gpu_migrate_begin (MithraGpu* h)
{
  USE(h);
  if (h->fd.size <= 1) return 0;
  if (!h->xch.connected) return fail("mithra_gpu_migrate_begin: slab is not connected to its neighbours");
  if (migrate_begin(h->xch, h->bd, h->P, h->pn, h->stream, h->num_sms, &h->cnt.kernel_launches)) return fail("particle migration: %s", h->xch.error.c_str());
  return 0;
}

extern "C" int mithra_gpu_migrate_end (MithraGpu* h)
{
  USE(h);
  if (h->fd.size <= 1) return 0;
  const size_t before = h->pn;
  if (migrate_end(h->xch, h->P, &h->pn, h->pcap, &h->next_id, h->stream, &h->cnt.kernel_launches)) return fail("particle migration: %s", h->xch.error.c_str());
  /* arrivals sit near the slab faces: grow the box the next E/B evaluation covers                           */
  const size_t kept = before - h->xch.h_counts[2];
  if (h->xch.h_counts[2] || h->pn != kept) h->ids_dense = false;
  if (h->pn > kept)
    {
      particle_box<<<grid_for((long) (h->pn - kept), 256, h->num_sms), 256, 0, h->stream>>>(h->bd, h->P, (long) kept, (long) h->pn, h->d_pbox, h->d_emask_cells);
      CU(cudaGetLastError());
      h->cnt.kernel_launches += 1;
    }
  return 0;
}

extern "C" int mithra_gpu_advance_time (MithraGpu* h)
{
  USE(h);
  h->timem1 += h->prm.dt; h->time += h->prm.dt; h->timep1 += h->prm.dt; ++h->n_time;
  h->cnt.field_steps += 1;
  return 0;
}

extern "C" int mithra_gpu_step (MithraGpu* h, int nsteps)
{
  USE(h);
  /* With neighbours the particle hand-over ends in a host synchronisation (the number of arrivals).  The first half of
   * the NEXT field update needs no particle, so it is enqueued before the host blocks: the device works on the
   * potentials while the host waits and then queues the particle phase (MITHRA_NO_LOOKAHEAD: the plain order).        */
  static const bool lookahead = getenv("MITHRA_NO_LOOKAHEAD") == 0;
  bool potentials_done = false;
  for (int s = 0; s < nsteps; s++)
    {
      if (!potentials_done) TRY(field_update_potentials(h));
      potentials_done = false;
      TRY(field_update_eb(h));
      TRY(housekeeping_ahead(h));
      /* the screens ride on the push (one pass over the bunch less); phase profiling keeps the two kernels apart     */
      static const bool nofuse = getenv("MITHRA_NO_FUSE") != 0;
      h->fuse_screens = !h->profiling && !nofuse;
      const int rb = mithra_gpu_bunch_update(h);
      const bool fused = h->fuse_screens;
      h->fuse_screens = false;
      if (rb) return rb;
      if (!fused) TRY(mithra_gpu_screen_profile(h));
      TRY(mithra_gpu_power_sample(h));
      TRY(mithra_gpu_power_visualize(h));
      TRY(mithra_gpu_field_shift(h));
      TRY(mithra_gpu_current_reset(h));
      TRY(mithra_gpu_current_update(h));
      TRY(mithra_gpu_current_communicate(h));
      TRY(mithra_gpu_migrate_begin(h));
      TRY(mithra_gpu_advance_time(h));                       /* host-side counters only: nothing below reads them     */
      if (lookahead && h->fd.size > 1 && s + 1 < nsteps && !h->profiling)
	{
	  TRY(field_update_potentials(h));
	  potentials_done = true;
	}
      TRY(mithra_gpu_migrate_end(h));
    }
  /* whatever was started ahead on the side stream belongs to this call: later work on the main stream (and the
   * caller's timing events) come after it                                                                        */
  if (h->seed_ahead) CU(cudaStreamWaitEvent(h->stream, h->ev_seed, 0));
  return 0;
}

extern "C" int mithra_gpu_synchronize (MithraGpu* h)
{
  USE(h);
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaStreamSynchronize(h->side));
  CU(cudaGetLastError());
  if (h->xch.d_err)
    {
      int e = 0; CU(cudaMemcpy(&e, h->xch.d_err, sizeof(int), cudaMemcpyDeviceToHost));
      if (e) return fail("mithra_gpu_synchronize: timed out waiting for a neighbouring slab (flag %d, sequence %d)", (e & 0xff) - 1, e >> 8);
    }
  return 0;
}

extern "C" int mithra_gpu_step_timed (MithraGpu* h, int nsteps, float* ms)
{
  USE(h);
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, h->stream));
  int r = mithra_gpu_step(h, nsteps);
  CU(cudaEventRecord(e1, h->stream));
  CU(cudaEventSynchronize(e1));
  float t = 0.f; CU(cudaEventElapsedTime(&t, e0, e1));
  if (ms) *ms = t;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return r;
}

extern "C" int mithra_gpu_step_profiled (MithraGpu* h, int nsteps, float ms[MITHRA_GPU_NPHASES])
{
  USE(h);
  CU(cudaStreamSynchronize(h->stream));
  memset(h->pms, 0, sizeof(h->pms));
  h->profiling = true;
  int r = mithra_gpu_step(h, nsteps);
  h->profiling = false;
  CU(cudaStreamSynchronize(h->stream));
  if (ms) memcpy(ms, h->pms, sizeof(h->pms));
  return r;
}

extern "C" int mithra_gpu_fetch_power (MithraGpu* h, double* rows, size_t capacity_rows, size_t* nrows)
{
  USE(h);
  if (!h->d_fdt) { if (nrows) *nrows = 0; return 0; }
  TRY(flush_power_rows(h));
  const size_t w = (size_t) h->pw.N * h->pw.Nl;
  const size_t have = h->h_rows.size() / w;
  const size_t take = std::min(have, capacity_rows);
  if (rows && take) memcpy(rows, h->h_rows.data(), take * w * sizeof(double));
  if (rows) h->h_rows.erase(h->h_rows.begin(), h->h_rows.begin() + take * w);
  if (nrows) *nrows = rows ? take : have;
  return 0;
}

extern "C" int mithra_gpu_fetch_screen (MithraGpu* h, int screen, double* rec6, size_t capacity, size_t* n)
{
  USE(h);
  if (n) *n = 0;
  if (!h->d_scr_pos) return 0;
  if (screen < 0 || screen >= h->prm.screens.N) return fail("mithra_gpu_fetch_screen: screen %d out of range", screen);
  CU(cudaStreamSynchronize(h->stream));
  unsigned int cur = 0;
  CU(cudaMemcpy(&cur, h->d_scr_cur + screen, sizeof(unsigned int), cudaMemcpyDeviceToHost));
  if (cur > h->scr_cap) return fail("mithra_gpu_fetch_screen: screen %d overflowed its %u-record buffer (MithraGpuParams.max_screen_records)", screen, h->scr_cap);
  const unsigned int done = h->scr_fetched[screen];
  const size_t fresh = cur - done;
  if (!rec6) { if (n) *n = fresh; return 0; }
  if (capacity < fresh) return fail("mithra_gpu_fetch_screen: capacity %zu < %zu records", capacity, fresh);
  std::vector<double> raw(fresh * 8);
  if (fresh) CU(cudaMemcpy(raw.data(), h->d_scr_rec + ((size_t) screen * h->scr_cap + done) * 8, fresh * 8 * sizeof(double), cudaMemcpyDeviceToHost));
  /* restore the reference's order: by step, then by particle index */
  std::vector<size_t> order(fresh);
  for (size_t i = 0; i < fresh; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    if (raw[a * 8 + 7] != raw[b * 8 + 7]) return raw[a * 8 + 7] < raw[b * 8 + 7];
    return raw[a * 8 + 6] < raw[b * 8 + 6]; });
  for (size_t i = 0; i < fresh; i++) memcpy(rec6 + i * 6, raw.data() + order[i] * 8, 6 * sizeof(double));
  h->scr_fetched[screen] = cur;
  if (n) *n = fresh;
  return 0;
}

__global__ void selftest_divide_kernel (const double* __restrict__ x, long n, double d, double rd, unsigned long long* __restrict__ bad)
{
  for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long) gridDim.x * blockDim.x)
    {
      const double a = div_by(x[t], d, rd), b = x[t] / d;
      if (__double_as_longlong(a) != __double_as_longlong(b) && !(a != a && b != b)) atomicAdd(bad, 1ULL);
    }
}

extern "C" int mithra_gpu_selftest_divide (const double* x, size_t n, double d, unsigned long long* mismatches)
{
  if (!x || !mismatches) return fail("mithra_gpu_selftest_divide: null argument");
  double* dx = 0; unsigned long long* dbad = 0;
  CU(cudaMalloc(&dx, n * sizeof(double))); CU(cudaMalloc(&dbad, sizeof(unsigned long long)));
  CU(cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice)); CU(cudaMemset(dbad, 0, sizeof(unsigned long long)));
  selftest_divide_kernel<<<1024, 256>>>(dx, (long) n, d, reciprocal_of(d), dbad);
  CU(cudaGetLastError());
  CU(cudaMemcpy(mismatches, dbad, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dbad);
  return 0;
}

extern "C" int mithra_gpu_counters (MithraGpu* h, MithraGpuCounters* out)
{
  USE(h);
  if (out) *out = h->cnt;
  return 0;
}

extern "C" int mithra_gpu_ipc_export (MithraGpu* h, void* blob, size_t capacity, size_t* nbytes)
{
  USE(h);
  return exchange_export(h->xch, h->fd, h->device, h->A, h->eb, blob, capacity, nbytes) ? fail("mithra_gpu_ipc_export: %s", h->xch.error.c_str()) : 0;
}

extern "C" int mithra_gpu_ipc_connect (MithraGpu* h, const void* blob_prev, const void* blob_next)
{
  USE(h);
  return exchange_connect(h->xch, h->fd, h->device, blob_prev, blob_next) ? fail("mithra_gpu_ipc_connect: %s", h->xch.error.c_str()) : 0;
}
