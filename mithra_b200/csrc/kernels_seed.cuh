/* kernels_seed.cuh -- TF/SF seed injection (FdTd::fieldUpdate, fdtd.cpp:307-373).
 *
 * After the interior sweep the analytic seed potential S = Seed::fields(rc(node), time_) (classes.cpp:740-855)
 * is subtracted / added on the two node layers around the total-field box:
 *   A+(i=1)    -= a1 S(i=2)      A+(i=2)    += a1 S(i=1)      A+(i=N0-2) -= a1 S(N0-3)    A+(i=N0-3) += a1 S(N0-2)
 * for j in [2,N1-3], k in [KI,KF); the same in y with a2; in z with a3 on the first / last slab only.
 * One thread owns one shell node and applies the x, then y, then z corrections in the reference's order, so
 * nodes that sit on two shells get bit-identical sums.
 */
#ifndef MITHRA_KERNELS_SEED_CUH_
#define MITHRA_KERNELS_SEED_CUH_

#include "device_types.cuh"
#include "beams.cuh"

namespace mithra
{
  struct SeedDev
  {
    MithraBeam beam;
    double c0, gamma, beta, dt_shift;
    double xmin, ymin, zmin, dx, dy, dz;
    int    along_z;                 /* direction == (0, 0, 1): everything that depends on z and t only is tabulated
				       per plane (seed_plane_table) instead of being recomputed at every node   */
    int    k0;                      /* global index of the slab's internal plane 0                              */
    double yv[3];                   /* direction x polarization                                                */
  };

  /* Per-plane part of Seed::fields (classes.cpp:740-855) for a beam along +z.  With direction = (0,0,1) the beam
   * coordinate z = rv . direction is rv_z bit for bit, so the retarded time tl, the carrier phase without the
   * transverse term, the signal envelope, the Gouy phase, the curvature factor and the beam widths depend on the
   * plane only.  They are evaluated with the reference's operation order once per plane and step; a node then costs
   * one cos, one exp and four divisions instead of 2 cos, 3 exp, 2 atan, 3 sqrt and fifteen divisions.          */
  #define MITHRA_SEED_TAB 12
  enum { ST_ACTIVE = 0, ST_PHASE, ST_ENV, ST_P0, ST_CURV, ST_DXP, ST_DYS, ST_RXW, ST_RYW, ST_AMP, ST_RVZ, ST_TS0 };

  /* Signal::self split into carrier phase base 2 pi f0 (t - t0) + cep and the factor that multiplies the cosine  */
  __device__ inline void signal_split (const MithraSignal& g, double t, double& base, double& env)
  {
    const double PI = MITHRA_PI;
    const double d = t - g.t0;
    base = 2 * PI * g.f0 * d + g.cep;
    env = 0.0;
    if (fabs(d) > 10.0 * g.s) return;
    switch (g.type)
      {
      case MITHRA_SIGNAL_NEUMANN:  env = - 2.7724 * d / ( g.s * g.s ) * exp( -1.3863 * d * d / ( g.s * g.s ) ); break;
      case MITHRA_SIGNAL_GAUSSIAN: { const double u = d / g.s; env = exp( -1.3863 * ( u * u ) ); } break;
      case MITHRA_SIGNAL_SECANT:   env = 1.0 / cosh( d / g.s ); break;
      case MITHRA_SIGNAL_FLATTOP:
      case MITHRA_SIGNAL_INVGAUSSIAN:
	{
	  double e = 1.0;
	  if (g.type == MITHRA_SIGNAL_INVGAUSSIAN)
	    { const double u0 = d / g.sigma_inv_g[0], u1 = d / g.sigma_inv_g[1]; e = pow( ( 1.0 + u0 * u0 ) * ( 1.0 + u1 * u1 ), 0.25 ); }
	  if (d <= - g.s / 2.0)     { const double u = ( d + g.s / 2.0 ) * g.f0 / g.nR; e *= exp( - ( u * u ) ); }
	  else if (d > g.s / 2.0)   { const double u = ( d - g.s / 2.0 ) * g.f0 / g.nR; e *= exp( - ( u * u ) ); }
	  env = e;
	}
	break;
      }
  }

  __global__ void __launch_bounds__(128)
  seed_plane_table (const SeedDev* __restrict__ sp, int np, double time, double* __restrict__ tab)
  {
    const SeedDev& s = *sp;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= np) return;
    const double PI = MITHRA_PI;
    const MithraBeam& B = s.beam;
    double* T = tab + (long) k * MITHRA_SEED_TAB;
    const double pz = s.zmin + ( k + s.k0 ) * s.dz;
    const double rlz = s.gamma * ( pz + s.beta * s.c0 * ( time + s.dt_shift ) );
    double tl = s.gamma * ( time + s.dt_shift + s.beta / s.c0 * pz );
    const double rvz = rlz - B.position[2];
    const double z = rvz;
    tl -= z / s.c0;
    const double ts0 = signal_self(B.signal, tl, 0.0);
    double base, env; signal_split(B.signal, tl, base, env);
    T[ST_ACTIVE] = ( fabs(ts0) < 1.0e-6 ) ? 0.0 : 1.0;
    T[ST_TS0] = ts0; T[ST_PHASE] = base; T[ST_ENV] = env; T[ST_RVZ] = rvz;
    const double l = s.c0 / B.signal.f0;
    const double zRp = PI * B.radius[0] * B.radius[0] / l;
    const double wrp = sqrt( 1.0 + z * z / ( zRp * zRp ) );
    const double zRs = PI * B.radius[1] * B.radius[1] / l;
    const double wrs = sqrt( 1.0 + z * z / ( zRs * zRs ) );
    T[ST_P0]   = 0.5 * ( atan( z / zRp ) + atan( z / zRs ) - PI );
    T[ST_CURV] = PI * z / l;
    T[ST_DXP]  = zRp * wrp; T[ST_DYS] = zRs * wrs;
    T[ST_RXW]  = B.radius[0] * wrp; T[ST_RYW] = B.radius[1] * wrs;
    T[ST_AMP]  = 1.0 / sqrt( wrs * wrp ) * B.amplitude;
  }

  /* The scalar u of Seed::fields (beams.cuh seed_assemble) at node (i, j) of the tabulated plane                 */
  __device__ __forceinline__ double seed_scalar_from_table (const SeedDev& s, const double* __restrict__ T, int i, int j)
  {
    const MithraBeam& B = s.beam;
    if (T[ST_ACTIVE] == 0.0) return 0.0;
    if (B.seed_type == MITHRA_BEAM_PLANEWAVE) return B.amplitude * T[ST_TS0];
    const V3 pol = v3a(B.polarization);
    const V3 rv = v3(s.xmin + i * s.dx - B.position[0], s.ymin + j * s.dy - B.position[1], T[ST_RVZ]);
    const double x = dot3(rv, pol), y = dot3(rv, v3a(s.yv));
    if (B.seed_type == MITHRA_BEAM_PLANEWAVETRUNCATED)
      return (!(fabs(x) > B.radius[0] || fabs(y) > B.radius[1])) ? B.amplitude * T[ST_TS0] : 0.0;
    const double p  = T[ST_P0] - T[ST_CURV] * ( sq( x / T[ST_DXP] ) + sq( y / T[ST_DYS] ) );
    const double ts = cos_wide( T[ST_PHASE] + p ) * T[ST_ENV];
    const double t  = exp( - sq( x / T[ST_RXW] ) - sq( y / T[ST_RYW] ) ) * T[ST_AMP];
    return t * ts;
  }

  /* tab == 0: evaluate Seed::fields in full at the node (any beam direction); node = Solver::rc, solver.cpp:2302-2315 */
  __device__ __forceinline__ double seed_scalar_at (const SeedDev& s, const double* __restrict__ tab, int i, int j, int kglob, double time)
  {
    if (tab) return seed_scalar_from_table(s, tab + (long) ( kglob - s.k0 ) * MITHRA_SEED_TAB, i, j);
    return seed_scalar(s.beam, s.c0, s.gamma, s.beta, s.dt_shift, s.xmin + i * s.dx, s.ymin + j * s.dy, s.zmin + kglob * s.dz, time);
  }

  __device__ __forceinline__ V3 seed_at (const SeedDev& s, const double* __restrict__ tab, int i, int j, int kglob, double time)
  { return seed_assemble(s.beam, s.gamma, seed_scalar_at(s, tab, i, j, kglob, time)); }

  __device__ __forceinline__ void seed_apply (double* __restrict__ ap, long cs, long m, double coef, const V3& S, bool minus)
  {
    if (minus) { ap[m] -= coef * S.x; ap[cs + m] -= coef * S.y; ap[2 * cs + m] -= coef * S.z; }
    else       { ap[m] += coef * S.x; ap[cs + m] += coef * S.y; ap[2 * cs + m] += coef * S.z; }
  }

  __device__ __forceinline__ bool on_shell (int v, int N) { return v == 1 || v == 2 || v == N - 2 || v == N - 3; }

  /* All TF/SF corrections of one node, x then y then z like the reference's loop nest (fdtd.cpp:307-373). */
  __device__ __forceinline__ void seed_node (const SeedDev& s, const double* __restrict__ tab, const FieldDev& f, double* __restrict__ anp1, int i, int j, int k, double time)
  {
    const bool zlo = (f.rank == 0), zhi = (f.rank == f.size - 1);
    const int KI = zlo ? 2 : f.kb, KF = zhi ? f.np - 2 : f.np - 1;
    const long cs = (long) f.np * f.Pp;
    const bool sx = on_shell(i, f.N0), sy = on_shell(j, f.N1);
    const bool sz = (zlo && (k == 1 || k == 2)) || (zhi && (k == f.np - 2 || k == f.np - 3));
    if (!sx && !sy && !sz) return;
    const long m = (long) k * f.Pp + (long) i * f.N1 + j;
    const int kg = k + f.k0;

    if (sx && j >= 2 && j <= f.N1 - 3 && k >= KI && k < KF)
      {
	if (i == 1)        seed_apply(anp1, cs, m, f.a[1], seed_at(s, tab, i + 1, j, kg, time), true);
	if (i == 2)        seed_apply(anp1, cs, m, f.a[1], seed_at(s, tab, i - 1, j, kg, time), false);
	if (i == f.N0 - 2) seed_apply(anp1, cs, m, f.a[1], seed_at(s, tab, i - 1, j, kg, time), true);
	if (i == f.N0 - 3) seed_apply(anp1, cs, m, f.a[1], seed_at(s, tab, i + 1, j, kg, time), false);
      }
    if (sy && i >= 2 && i <= f.N0 - 3 && k >= KI && k < KF)
      {
	if (j == 1)        seed_apply(anp1, cs, m, f.a[2], seed_at(s, tab, i, j + 1, kg, time), true);
	if (j == 2)        seed_apply(anp1, cs, m, f.a[2], seed_at(s, tab, i, j - 1, kg, time), false);
	if (j == f.N1 - 2) seed_apply(anp1, cs, m, f.a[2], seed_at(s, tab, i, j - 1, kg, time), true);
	if (j == f.N1 - 3) seed_apply(anp1, cs, m, f.a[2], seed_at(s, tab, i, j + 1, kg, time), false);
      }
    if (sz && i >= 2 && i <= f.N0 - 3 && j >= 2 && j <= f.N1 - 3)
      {
	if (zlo && k == 1)        seed_apply(anp1, cs, m, f.a[3], seed_at(s, tab, i, j, kg + 1, time), true);
	if (zlo && k == 2)        seed_apply(anp1, cs, m, f.a[3], seed_at(s, tab, i, j, kg - 1, time), false);
	if (zhi && k == f.np - 2) seed_apply(anp1, cs, m, f.a[3], seed_at(s, tab, i, j, kg - 1, time), true);
	if (zhi && k == f.np - 3) seed_apply(anp1, cs, m, f.a[3], seed_at(s, tab, i, j, kg + 1, time), false);
      }
  }

  /* Generic (tiny meshes): walk every interior node, test for the shell. */
  __global__ void __launch_bounds__(128)
  seed_inject_scan (const SeedDev* __restrict__ sp, const double* __restrict__ tab, const FieldDev f, double* __restrict__ anp1, double time)
  {
    const long nin = (long) (f.N0 - 2) * (f.N1 - 2) * (f.np - 1 - f.kb);
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < nin; t += (long) gridDim.x * blockDim.x)
      {
	long r = t;
	const int k = f.kb + (int) (r / ((long) (f.N0 - 2) * (f.N1 - 2))); r -= (long) (k - f.kb) * (f.N0 - 2) * (f.N1 - 2);
	const int i = 1 + (int) (r / (f.N1 - 2)), j = 1 + (int) (r % (f.N1 - 2));
	seed_node(*sp, tab, f, anp1, i, j, k, time);
      }
  }

  /* Compact enumeration of the shell (N0, N1, np >= 8): every shell node exactly once, so that all lanes of a
   * warp do the transcendental work of Seed::fields.
   *   X: i in {1, 2, N0-3, N0-2}, j in [1, N1-2], k in [1, np-2]                 (j fastest: coalesced)
   *   Y: j in {1, 2, N1-3, N1-2}, i in [3, N0-4], k in [1, np-2]
   *   Z: k in {1, 2} on the first slab and {np-3, np-2} on the last, i in [3, N0-4], j in [3, N1-4]          */
  __global__ void __launch_bounds__(128)
  seed_inject_shell (const SeedDev* __restrict__ sp, const double* __restrict__ tab, const FieldDev f, double* __restrict__ anp1, double time)
  {
    const bool zlo = (f.rank == 0), zhi = (f.rank == f.size - 1);
    const int  nk = f.np - 1 - f.kb;
    const long nX = 4L * (f.N1 - 2) * nk;
    const long nY = 4L * (f.N0 - 6) * nk;
    const int  nzs = (zlo ? 2 : 0) + (zhi ? 2 : 0);
    const long nZ = (long) nzs * (f.N0 - 6) * (f.N1 - 6);
    const long tot = nX + nY + nZ;
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	long r = t; int i, j, k;
	if (r < nX)
	  {
	    const int q = (int) (r / ((long) (f.N1 - 2) * nk)); r -= (long) q * (f.N1 - 2) * nk;
	    k = f.kb + (int) (r / (f.N1 - 2)); j = 1 + (int) (r % (f.N1 - 2));
	    i = (q == 0) ? 1 : (q == 1) ? 2 : (q == 2) ? f.N0 - 3 : f.N0 - 2;
	  }
	else if (r < nX + nY)
	  {
	    r -= nX;
	    const int q = (int) (r / ((long) (f.N0 - 6) * nk)); r -= (long) q * (f.N0 - 6) * nk;
	    k = f.kb + (int) (r / (f.N0 - 6)); i = 3 + (int) (r % (f.N0 - 6));
	    j = (q == 0) ? 1 : (q == 1) ? 2 : (q == 2) ? f.N1 - 3 : f.N1 - 2;
	  }
	else
	  {
	    r -= nX + nY;
	    const int q = (int) (r / ((long) (f.N0 - 6) * (f.N1 - 6))); r -= (long) q * (f.N0 - 6) * (f.N1 - 6);
	    i = 3 + (int) (r / (f.N1 - 6)); j = 3 + (int) (r % (f.N1 - 6));
	    if (zlo && q < 2) k = 1 + q; else k = f.np - 3 + (q - (zlo ? 2 : 0));
	  }
	seed_node(*sp, tab, f, anp1, i, j, k, time);
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * Rim path (rim_update, kernels_field.cuh, applies the x / y shell corrections itself): the seed scalar on the
   * eight source lines of every plane, RimDev.seedu[k][line][idx]: lines 0..3 = rows i = 1, 2, N0-3, N0-2 indexed
   * by j, lines 4..7 = columns j = 1, 2, N1-3, N1-2 indexed by i.  Only the entries the shells read are evaluated
   * (j in [2, N1-3] on the rows, i in [2, N0-3] on the columns, k in [KI, KF)); the rest is never read.
   * ------------------------------------------------------------------------------------------------ */
  __global__ void __launch_bounds__(128)
  seed_lines (const SeedDev* __restrict__ sp, const double* __restrict__ tab, const FieldDev f, int KI, int KF, int L,
	      double* __restrict__ seedu, double time)
  {
    const int  nr = f.N1 - 4, nc = f.N0 - 4;              /* entries per row line / per column line           */
    const int  per = 4 * nr + 4 * nc;
    const long tot = (long) per * (KF - KI);
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	const int k = KI + (int) (t / per); int r = (int) (t - (long) (k - KI) * per);
	int line, idx, i, j;
	if (r < 4 * nr) { line = r / nr; idx = 2 + r % nr; j = idx; i = (line == 0) ? 1 : (line == 1) ? 2 : (line == 2) ? f.N0 - 3 : f.N0 - 2; }
	else { r -= 4 * nr; line = 4 + r / nc; idx = 2 + r % nc; i = idx; j = (line == 4) ? 1 : (line == 5) ? 2 : (line == 6) ? f.N1 - 3 : f.N1 - 2; }
	seedu[((long) k * 8 + line) * L + idx] = seed_scalar_at(*sp, tab, i, j, k + f.k0, time);
      }
  }

  /* z shell only (first / last slab): A+(k=1) -= a3 S(k=2), A+(k=2) += a3 S(k=1), likewise at np-2 / np-3, for
   * i in [2, N0-3], j in [2, N1-3] (fdtd.cpp:353-373); runs after rim_update, i.e. after the x and y terms.      */
  __global__ void __launch_bounds__(128)
  seed_inject_zshell (const SeedDev* __restrict__ sp, const double* __restrict__ tab, const FieldDev f, double* __restrict__ anp1, double time)
  {
    const bool zlo = (f.rank == 0), zhi = (f.rank == f.size - 1);
    const int  nzs = (zlo ? 2 : 0) + (zhi ? 2 : 0);
    const long per = (long) (f.N0 - 4) * (f.N1 - 4), tot = per * nzs;
    const long cs = (long) f.np * f.Pp;
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	const int q = (int) (t / per); const long r = t - (long) q * per;
	const int i = 2 + (int) (r / (f.N1 - 4)), j = 2 + (int) (r % (f.N1 - 4));
	int k; if (zlo && q < 2) k = 1 + q; else k = f.np - 3 + (q - (zlo ? 2 : 0));
	const long m = (long) k * f.Pp + (long) i * f.N1 + j;
	const int kg = k + f.k0;
	if (zlo && k == 1)        seed_apply(anp1, cs, m, f.a[3], seed_at(*sp, tab, i, j, kg + 1, time), true);
	if (zlo && k == 2)        seed_apply(anp1, cs, m, f.a[3], seed_at(*sp, tab, i, j, kg - 1, time), false);
	if (zhi && k == f.np - 2) seed_apply(anp1, cs, m, f.a[3], seed_at(*sp, tab, i, j, kg - 1, time), true);
	if (zhi && k == f.np - 3) seed_apply(anp1, cs, m, f.a[3], seed_at(*sp, tab, i, j, kg + 1, time), false);
      }
  }

  /* Initial condition: A^n and A^{n-1} = S inside the total-field box (solver.cpp:828-839). */
  __global__ void __launch_bounds__(128)
  seed_initial_kernel (const SeedDev* __restrict__ sp, const FieldDev f, double* __restrict__ an, double* __restrict__ anm1,
		       double time, double timem1)
  {
    const SeedDev& s = *sp;
    const int kb = (f.rank == 0) ? 2 : 0, ke = (f.rank == f.size - 1) ? f.np - 2 : f.np;
    const long cs = (long) f.np * f.Pp;
    const long tot = (long) (f.N0 - 4) * (f.N1 - 4) * (ke - kb);
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	long r = t;
	const int k = kb + (int) (r / ((long) (f.N0 - 4) * (f.N1 - 4))); r -= (long) (k - kb) * (f.N0 - 4) * (f.N1 - 4);
	const int i = 2 + (int) (r / (f.N1 - 4)), j = 2 + (int) (r % (f.N1 - 4));
	const long m = (long) k * f.Pp + (long) i * f.N1 + j;
	const V3 a = seed_at(s, 0, i, j, k + f.k0, time), b = seed_at(s, 0, i, j, k + f.k0, timem1);
	an[m] = a.x; an[cs + m] = a.y; an[2 * cs + m] = a.z;
	anm1[m] = b.x; anm1[cs + m] = b.y; anm1[2 * cs + m] = b.z;
      }
  }

  /* tab: device array of np * MITHRA_SEED_TAB doubles when the beam runs along +z (SeedDev.along_z), else 0      */
  static inline int seed_inject (const SeedDev* d_seed, double* tab, const FieldDev& f, double* anp1, double time, cudaStream_t stream, int num_sms)
  {
    if (tab) seed_plane_table<<<(f.np + 127) / 128, 128, 0, stream>>>(d_seed, f.np, time, tab);
    if (f.N0 < 8 || f.N1 < 8 || f.np < 8)
      seed_inject_scan<<<num_sms * 8, 128, 0, stream>>>(d_seed, tab, f, anp1, time);
    else
      {
	const long tot = 4L * (f.N1 - 2) * (f.np - 2) + 4L * (f.N0 - 6) * (f.np - 2) + 4L * (f.N0 - 6) * (f.N1 - 6);
	long g = (tot + 127) / 128; if (g > num_sms * 16L) g = num_sms * 16L; if (g < 1) g = 1;
	seed_inject_shell<<<(int) g, 128, 0, stream>>>(d_seed, tab, f, anp1, time);
      }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
}

#endif
