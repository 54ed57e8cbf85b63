/* device_types.cuh -- device-side parameter blocks and index helpers of the B200 MITHRA time-march.
 *
 * Device layout (private to the library, see DESIGN.md "Data layout in HBM"):
 *   potentials : component-planar doubles, comp c in {Ax, Ay, Az [, phi]}:
 *                  idx(c,k,i,j) = (c*np + k) * Pp + i*N1 + j,   Pp = roundup(N0*N1, 16)   (128-byte planes)
 *                three time levels (n+1, n, n-1) are three such allocations that rotate (FdTd::fieldShift).
 *   current    : J = {Jx, Jy, Jz [, rho]} in the same layout, in its own allocation (the reference aliases it
 *                onto anp1_, fdtd.cpp:27,244); only the device-tracked bounding box of the deposits is read
 *                by the stencil and cleared afterwards.
 *   E, B       : one float4 pair per node (Ex,Ey,Ez,0 | Bx,By,Bz,0), node m = N0*N1*k + N1*i + j -- the
 *                reference's FieldVector<float> en_, bn_ (solver.h:244-245) interleaved for the 8-node gather.
 *   particles  : struct-of-arrays of doubles q, r[3], rm[3], gb[3], e (reference Charge, stdinclude.h:130-144) plus
 *                the upload index id; two copies, the counting sort by cell (kernels_sort.cuh) moves the bunch
 *                from one to the other.
 *
 * z-slabs (one handle per GPU): the ABI speaks the reference's slab numbering (np local planes from global plane
 * k0, two planes shared with each neighbour, solver.cpp:619-641).  Internally every slab except the first keeps
 * ONE MORE plane below (kshift = 1): a particle is advanced and deposited for a whole field step by the slab that
 * owns it at the start of the step, and it can drift at most one cell below the slab in that time.  Internal plane
 * l = reference plane + kshift; planes [0, kb) and, except on the last slab, plane np-1 are ghosts.
 */
#ifndef MITHRA_DEVICE_TYPES_CUH_
#define MITHRA_DEVICE_TYPES_CUH_

#include <cuda_runtime.h>
#include "../../include/mithra_gpu.h"

namespace mithra
{
  /* Bounding box of node indices, inclusive; empty when lo > hi. Lives in device memory.              */
  struct Box { int lo[3]; int hi[3]; };   /* axis 0 = i (x), 1 = j (y), 2 = k (z, local plane index)   */

  /* Field-side constants, passed to kernels by value.                                                 */
  struct FieldDev
  {
    int    N0, N1, np, k0;        /* np, k0: INTERNAL plane count / global index of internal plane 0       */
    int    kshift;                /* internal plane = reference local plane + kshift (0 on the first slab) */
    int    kb;                    /* first plane this slab updates (1 on the first slab, else kshift + 1); the
				     last one is np-2; plane np-1 is the z face (last slab) or a ghost      */
    int    P;                     /* N0*N1                                                              */
    long   Pp;                    /* padded plane stride (doubles)                                      */
    int    ncomp;                 /* 3, or 4 with space charge                                          */
    int    rank, size;
    int    nsfd;                  /* 1 = NSFD, 0 = FD                                                   */
    int    order;                 /* truncation order                                                   */
    double a[6], alpha, beta;
    double bB[5], cB[5], dB[5], eE[5], fE[5], gE[5], hC[17];
    double dt, dx2, dy2, dz2;     /* uf_.dt, 2dx, 2dy, 2dz (solver.cpp:714-722)                         */
    double rmdt, rdx2, rdy2, rdz2; /* reciprocal_of(-dt), (2dx), (2dy), (2dz) for div_by                  */
  };

  /* The rim of every plane -- rows i = 1, 2, N0-3, N0-2 and columns j = 1, 2, N1-3, N1-2 -- is where the TF/SF
   * seed corrections of the x / y shells land and from where the x / y absorbing faces are updated; rim_update
   * (kernels_field.cuh) owns those nodes and stencil_stream skips them.  seedu: the line table seed_lines writes
   * (kernels_seed.cuh), each entry the scalar u of Seed::fields at a source node of the shells.                 */
  struct RimDev
  {
    int    seed;                  /* apply the x / y shell corrections                                              */
    const double* seedu;          /* [np][8][L]                                                                     */
    int    L;                     /* entries per line (even, >= max(N0, N1))                                        */
    int    KI, KF;                /* planes [KI, KF) carry x / y shell corrections (fdtd.cpp:310-311)               */
    int    ni, supergaussian;     /* number of equal terms the seed vector is summed from (quirk Q8)                */
    double pol[3], gamma;
  };

  /* One static-undulator module with the per-module constants of Solver::undulatorField precomputed on
   * the host in the reference's operation order (solver.cpp:1805-1812).                               */
  struct UndulatorDev
  {
    int    type;
    double b0, ku, ct, st, rb, len;     /* len = length_ * lu_                                          */
    double r0_prev, r0_next;            /* gaps to the neighbouring modules (beam.cc:35,60)             */
    int    has_prev, has_next;
    MithraBeam beam;
  };

  /* Bunch-side constants.  Passed to the particle kernels BY VALUE as a __grid_constant__ parameter (about 6 KB;
   * CUDA 12 allows 32 KB of kernel parameters): every field is then a constant-bank operand or a uniform register
   * instead of a global load in front of the first use.                                                  */
  struct BunchDev
  {
    double xmin, xmax, ymin, ymax, zmin, zmax;
    double zp0, zp1, Lz;
    int    size;                        /* number of slabs; with more than one the particle list of a slab IS its
					   ownership (migration once per field step) and the z tests are global */
    double dx, dy, dz;
    double rdx, rdy, rdz, rr2;          /* reciprocal_of(dx), (dy), (dz), (r2) for div_by                */
    double c0, gamma, beta, dt_shift;
    double r1, r2, dtb, dt_bunch, dt_field;
    double reach[3];                    /* distance a particle can travel in one field step, in cells per axis, with a
					   safety margin (particle_reach, kernels_bunch.cuh)                       */
    int    N0, N1, np, k0, P;          /* np, k0 internal (see FieldDev)                                */
    int    kshift;
    long   Pp;
    int    ncomp;
    int    n_und;
    double und0_dist;                   /* undulator_[0].dist_ (entrance flag, solver.cpp:1510-1511)    */
    UndulatorDev und[MITHRA_MAX_UNDULATORS];
    int    n_ext;
    MithraBeam ext[MITHRA_MAX_EXTFIELDS];
  };

  /* Particle struct-of-arrays.                                                                        */
  struct ParticlesDev
  {
    double* q;
    double* r[3];
    double* rm[3];
    double* gb[3];
    double* e;
    unsigned int* id;                   /* index the particle had when it was uploaded (arrivals from a neighbouring
					   slab get fresh ones): the reference's list order, kept across the sorts */
  };

  /* Correctly rounded x / d for a divisor that is a run-time CONSTANT of the job (cell sizes, time step): with
   * rd = RN(1/d) from the host, q = RN(x rd) is within one ulp of x/d, the residual r = x - q d is exact in an FMA,
   * and RN(q + r rd) is the correctly rounded quotient (Markstein's theorem; checked against x/d over 4.5e8 random
   * operands on the CPU and by mithra_gpu_selftest_divide on the device).  Three FP64 issue slots instead of the
   * ~25 of the generic division routine; bit-identical to IEEE division, which the cell indices and the E/B floats
   * need.  Operands whose residual could leave the normal range (and NaN / Inf / zero) take the true division;
   * rd == 0 marks a divisor the host would not vouch for.                                                      */
  inline double reciprocal_of (double d)
  { const double a = d < 0 ? -d : d; return (a > 1.0e-100 && a < 1.0e100) ? 1.0 / d : 0.0; }

  __device__ __noinline__ double div_true (double x, double d) { return x / d; }    /* out of line: see div_by */

  __device__ __forceinline__ double div_by (double x, double d, double rd)
  {
    const double q = x * rd;
    const double r = __fma_rn(-q, d, x);
    double res = __fma_rn(r, rd, q);
    /* |x| in [2^-498, 2^498) -- inside (1e-150, 1e150) -- read off the exponent field with integer instructions: the
     * test costs the FP64 pipe nothing (three DSETP per division before)                                          */
    const unsigned ex = ( (unsigned) __double2hiint(x) >> 20 ) & 0x7ffu;
    if (ex - 525u >= 996u || rd == 0.0)                      /* rare: keep the straight-line path free of calls */
      res = (x == 0.0 && rd != 0.0) ? q : div_true(x, d);    /* (+-0) rd has the sign of (+-0) / d; a real call, or
								the compiler runs the division's inline part speculatively */
    return res;
  }

  /* The same three operations without the guard, for quotients that end up in a FLOAT (the E/B evaluation): the guard
   * only matters for |x| below 2^-969, where the residual would be subnormal -- such a quotient is far below the
   * smallest float (and any sum it enters is unchanged by its last bit), so the float is the one IEEE division gives.
   * (A zero keeps its value, possibly not its sign.)  mithra_gpu_create refuses divisors reciprocal_of() will not take. */
  __device__ __forceinline__ double div_fast (double x, double d, double rd)
  {
    const double q = x * rd;
    const double r = __fma_rn(-q, d, x);
    return __fma_rn(r, rd, q);
  }

  /* ------------------------------------------------------------------------------------------------
   * Reach mask.  The E/B evaluation, the source read of the stencil and the clear of J only have to touch the nodes a
   * particle can gather from / deposit on during ONE field step: it travels less than c dt (|v| < c, the sub-steps add up
   * to dt), so from its position at the start of the step it stays within the cells [lo, hi] per axis with
   * lo = cell(r - c dt), hi = cell(r + c dt), and touches the nodes lo .. hi + 1.  The mask has one byte per CELL PENCIL
   * (cell column (i, j) x 2^MITHRA_EB_CHUNK_LOG2 planes); a particle ORs into the byte of its own cell:
   *   PRESENT, and which neighbouring cells / plane chunks its reach extends into (XLO .. ZHI).
   * spread_eb_mask (kernels_field.cuh) turns the cell bytes into a node-pencil mask.
   * ------------------------------------------------------------------------------------------------ */
  #define MITHRA_EB_CHUNK_LOG2 3                /* planes per pencil of the masks                                       */
  enum { REACH_PRESENT = 1, REACH_XLO = 2, REACH_XHI = 4, REACH_YLO = 8, REACH_YHI = 16, REACH_ZLO = 32, REACH_ZHI = 64 };

  __host__ __device__ inline long fidx (const long Pp, const int np, const int N1, int c, int k, int i, int j)
  { return ((long) c * np + k) * Pp + (long) i * N1 + j; }
}

#endif
