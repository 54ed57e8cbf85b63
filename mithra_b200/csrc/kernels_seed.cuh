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
  };

  __device__ __forceinline__ V3 seed_at (const SeedDev& s, int i, int j, int kglob, double time)
  {
    /* Solver::rc, solver.cpp:2302-2315 */
    return seed_fields(s.beam, s.c0, s.gamma, s.beta, s.dt_shift, s.xmin + i * s.dx, s.ymin + j * s.dy, s.zmin + kglob * s.dz, time);
  }

  __device__ __forceinline__ void seed_apply (double* __restrict__ ap, long cs, long m, double coef, const V3& S, bool minus)
  {
    if (minus) { ap[m] -= coef * S.x; ap[cs + m] -= coef * S.y; ap[2 * cs + m] -= coef * S.z; }
    else       { ap[m] += coef * S.x; ap[cs + m] += coef * S.y; ap[2 * cs + m] += coef * S.z; }
  }

  __device__ __forceinline__ bool on_shell (int v, int N) { return v == 1 || v == 2 || v == N - 2 || v == N - 3; }

  /* All TF/SF corrections of one node, x then y then z like the reference's loop nest (fdtd.cpp:307-373). */
  __device__ __forceinline__ void seed_node (const SeedDev& s, const FieldDev& f, double* __restrict__ anp1, int i, int j, int k, double time)
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
	if (i == 1)        seed_apply(anp1, cs, m, f.a[1], seed_at(s, i + 1, j, kg, time), true);
	if (i == 2)        seed_apply(anp1, cs, m, f.a[1], seed_at(s, i - 1, j, kg, time), false);
	if (i == f.N0 - 2) seed_apply(anp1, cs, m, f.a[1], seed_at(s, i - 1, j, kg, time), true);
	if (i == f.N0 - 3) seed_apply(anp1, cs, m, f.a[1], seed_at(s, i + 1, j, kg, time), false);
      }
    if (sy && i >= 2 && i <= f.N0 - 3 && k >= KI && k < KF)
      {
	if (j == 1)        seed_apply(anp1, cs, m, f.a[2], seed_at(s, i, j + 1, kg, time), true);
	if (j == 2)        seed_apply(anp1, cs, m, f.a[2], seed_at(s, i, j - 1, kg, time), false);
	if (j == f.N1 - 2) seed_apply(anp1, cs, m, f.a[2], seed_at(s, i, j - 1, kg, time), true);
	if (j == f.N1 - 3) seed_apply(anp1, cs, m, f.a[2], seed_at(s, i, j + 1, kg, time), false);
      }
    if (sz && i >= 2 && i <= f.N0 - 3 && j >= 2 && j <= f.N1 - 3)
      {
	if (zlo && k == 1)        seed_apply(anp1, cs, m, f.a[3], seed_at(s, i, j, kg + 1, time), true);
	if (zlo && k == 2)        seed_apply(anp1, cs, m, f.a[3], seed_at(s, i, j, kg - 1, time), false);
	if (zhi && k == f.np - 2) seed_apply(anp1, cs, m, f.a[3], seed_at(s, i, j, kg - 1, time), true);
	if (zhi && k == f.np - 3) seed_apply(anp1, cs, m, f.a[3], seed_at(s, i, j, kg + 1, time), false);
      }
  }

  /* Generic (tiny meshes): walk every interior node, test for the shell. */
  __global__ void __launch_bounds__(128)
  seed_inject_scan (const SeedDev* __restrict__ sp, const FieldDev f, double* __restrict__ anp1, double time)
  {
    const long nin = (long) (f.N0 - 2) * (f.N1 - 2) * (f.np - 1 - f.kb);
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < nin; t += (long) gridDim.x * blockDim.x)
      {
	long r = t;
	const int k = f.kb + (int) (r / ((long) (f.N0 - 2) * (f.N1 - 2))); r -= (long) (k - f.kb) * (f.N0 - 2) * (f.N1 - 2);
	const int i = 1 + (int) (r / (f.N1 - 2)), j = 1 + (int) (r % (f.N1 - 2));
	seed_node(*sp, f, anp1, i, j, k, time);
      }
  }

  /* Compact enumeration of the shell (N0, N1, np >= 8): every shell node exactly once, so that all lanes of a
   * warp do the transcendental work of Seed::fields.
   *   X: i in {1, 2, N0-3, N0-2}, j in [1, N1-2], k in [1, np-2]                 (j fastest: coalesced)
   *   Y: j in {1, 2, N1-3, N1-2}, i in [3, N0-4], k in [1, np-2]
   *   Z: k in {1, 2} on the first slab and {np-3, np-2} on the last, i in [3, N0-4], j in [3, N1-4]          */
  __global__ void __launch_bounds__(128)
  seed_inject_shell (const SeedDev* __restrict__ sp, const FieldDev f, double* __restrict__ anp1, double time)
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
	seed_node(*sp, f, anp1, i, j, k, time);
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
	const V3 a = seed_at(s, i, j, k + f.k0, time), b = seed_at(s, i, j, k + f.k0, timem1);
	an[m] = a.x; an[cs + m] = a.y; an[2 * cs + m] = a.z;
	anm1[m] = b.x; anm1[cs + m] = b.y; anm1[2 * cs + m] = b.z;
      }
  }

  static inline int seed_inject (const SeedDev* d_seed, const FieldDev& f, double* anp1, double time, cudaStream_t stream, int num_sms)
  {
    if (f.N0 < 8 || f.N1 < 8 || f.np < 8)
      seed_inject_scan<<<num_sms * 8, 128, 0, stream>>>(d_seed, f, anp1, time);
    else
      {
	const long tot = 4L * (f.N1 - 2) * (f.np - 2) + 4L * (f.N0 - 6) * (f.np - 2) + 4L * (f.N0 - 6) * (f.N1 - 6);
	long g = (tot + 127) / 128; if (g > num_sms * 16L) g = num_sms * 16L; if (g < 1) g = 1;
	seed_inject_shell<<<(int) g, 128, 0, stream>>>(d_seed, f, anp1, time);
      }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
}

#endif
