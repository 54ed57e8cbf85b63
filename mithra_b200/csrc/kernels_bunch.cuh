/* kernels_bunch.cuh -- particle-side sm_100a kernels: Boris push with analytic undulator / external fields and
 * trilinear E,B gather, charge-conserving ZigZag deposition, bounding boxes, screens and radiated power.
 *
 * Reference: Solver::bunchUpdate solver.cpp:1424-1576, undulatorField / externalField solver.cpp:1798-1947,
 * FdTd::currentUpdate fdtd.cpp:38-185 (+rho fdtdSC.cpp:141-160), Solver::screenProfile solver.cpp:2205-2257,
 * Solver::powerSample radiation.cpp:127-232.
 *
 * Index arithmetic (cell index, weights, ownership) uses true IEEE division, modf / floor and no FMA
 * contraction (-fmad=false), exactly as the reference does, so particle-to-cell assignment is bit-exact for
 * identical positions.
 */
#ifndef MITHRA_KERNELS_BUNCH_CUH_
#define MITHRA_KERNELS_BUNCH_CUH_

#include "device_types.cuh"
#include "beams.cuh"

namespace mithra
{
  /* pmod, stdinclude.cpp:88-93 */
  __device__ __forceinline__ double pmod (double a, double b)
  {
    if (a >= 0.0 && a < b) return a;                   /* fmod(a, b) == a exactly: skip its division loop        */
    double x = fmod(a, b);
    x += ( x < 0.0 ) ? b : 0.0;
    return x;
  }

  __device__ __forceinline__ void prefetch_l1 (const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

  __device__ __forceinline__ void warp_box_merge (Box* box, bool valid, int i0, int i1, int j0, int j1, int k0, int k1)
  {
    const unsigned full = 0xffffffffu;
    int lo0 = valid ? i0 : 0x7fffffff, lo1 = valid ? j0 : 0x7fffffff, lo2 = valid ? k0 : 0x7fffffff;
    int hi0 = valid ? i1 : -1,         hi1 = valid ? j1 : -1,         hi2 = valid ? k1 : -1;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      {
	lo0 = min(lo0, __shfl_xor_sync(full, lo0, o)); lo1 = min(lo1, __shfl_xor_sync(full, lo1, o)); lo2 = min(lo2, __shfl_xor_sync(full, lo2, o));
	hi0 = max(hi0, __shfl_xor_sync(full, hi0, o)); hi1 = max(hi1, __shfl_xor_sync(full, hi1, o)); hi2 = max(hi2, __shfl_xor_sync(full, hi2, o));
      }
    if ((threadIdx.x & 31) == 0 && hi0 >= lo0)
      {
	/* the box only grows during a launch: a (possibly stale) look at it tells most warps that they have nothing to add,
	 * and the six same-address reductions of every warp of a large bunch do not queue up in one L2 slice            */
	const volatile Box* g = box;
	if (lo0 < g->lo[0]) atomicMin(&box->lo[0], lo0);
	if (lo1 < g->lo[1]) atomicMin(&box->lo[1], lo1);
	if (lo2 < g->lo[2]) atomicMin(&box->lo[2], lo2);
	if (hi0 > g->hi[0]) atomicMax(&box->hi[0], hi0);
	if (hi1 > g->hi[1]) atomicMax(&box->hi[1], hi1);
	if (hi2 > g->hi[2]) atomicMax(&box->hi[2], hi2);
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * Bounding box (cell indices) of the particles that will gather mesh fields, used to size the E/B
   * evaluation of the next step.  The host pads it by the distance a particle can travel in one field step.
   * ------------------------------------------------------------------------------------------------ */
  /* Reach of a particle at (x, y, z) during the next field step (device_types.cuh "Reach mask"): its cell (i, j, k) -- k a
   * local plane index --, clamped into the mesh, and the byte it ORs into the cell-pencil mask.  false: nothing of the
   * mesh's gather / deposit region is within reach.  The margin is c dt plus a relative 1e-6 and a 1e-9 of the cell, far
   * above the rounding of the index arithmetic.                                                                      */
  __device__ __forceinline__ bool particle_reach (const BunchDev& b, double x, double y, double z, int& i, int& j, int& k, unsigned int& bits)
  {
    const double cdt = b.c0 * b.dt_field * ( 1.0 + 1.0e-6 );
    const double mx = cdt + 1.0e-9 * b.dx, my = cdt + 1.0e-9 * b.dy, mz = cdt + 1.0e-9 * b.dz;
    if (!( x + mx > b.xmin + b.dx && x - mx < b.xmax - b.dx && y + my > b.ymin + b.dy && y - my < b.ymax - b.dy &&
	   z + mz >= b.zmin && z - mz < b.zmax )) return false;
    /* position in cells (the quotient the cell index is the floor of), the margin in cells from the host               */
    const double qx = div_by( x - b.xmin, b.dx, b.rdx ), qy = div_by( y - b.ymin, b.dy, b.rdy ), qz = div_by( z - b.zmin, b.dz, b.rdz );
    const int ic = (int) floor( qx ), jc = (int) floor( qy ), kc = (int) floor( qz ) - b.k0;
    const int ilo = (int) floor( qx - b.reach[0] ), ihi = (int) floor( qx + b.reach[0] );
    const int jlo = (int) floor( qy - b.reach[1] ), jhi = (int) floor( qy + b.reach[1] );
    const int klo = (int) floor( qz - b.reach[2] ) - b.k0, khi = (int) floor( qz + b.reach[2] ) - b.k0;
    i = min(max(ic, 0), b.N0 - 2); j = min(max(jc, 0), b.N1 - 2); k = min(max(kc, 0), b.np - 1);
    constexpr int L = MITHRA_EB_CHUNK_LOG2;
    bits = REACH_PRESENT;
    if (ilo < i) bits |= REACH_XLO;
    if (ihi > i) bits |= REACH_XHI;
    if (jlo < j) bits |= REACH_YLO;
    if (jhi > j) bits |= REACH_YHI;
    if ((max(klo, 0) >> L) < (k >> L)) bits |= REACH_ZLO;
    if ((min(khi + 1, b.np - 1) >> L) > (k >> L)) bits |= REACH_ZHI;
    return true;
  }

  /* OR the reach byte into the cell-pencil mask (one 32-bit reduction; the byte array is a multiple of 4 bytes long)  */
  __device__ __forceinline__ void mark_eb_pencil (const BunchDev& b, unsigned char* __restrict__ emask, int i, int j, int k, unsigned int bits)
  {
    if (!emask) return;
    const long idx = ((long) (k >> MITHRA_EB_CHUNK_LOG2) * b.N0 + i) * b.N1 + j;
    atomicOr(reinterpret_cast<unsigned int*>(emask + (idx & ~3L)), bits << (8 * (int) (idx & 3L)));
  }

  __global__ void __launch_bounds__(256)
  particle_box (const __grid_constant__ BunchDev b, ParticlesDev P, long start, long n, Box* __restrict__ box,
		unsigned char* __restrict__ emask)
  {
    for (long base = start + (long) blockIdx.x * blockDim.x; base < n; base += (long) gridDim.x * blockDim.x)
      {
	const long t = base + threadIdx.x;
	bool valid = false; int i = 0, j = 0, k = 0;
	if (t < n)
	  {
	    unsigned int bits;
	    valid = particle_reach(b, P.r[0][t], P.r[1][t], P.r[2][t], i, j, k, bits);
	    if (valid) mark_eb_pencil(b, emask, i, j, k, bits);
	  }
	warp_box_merge(box, valid, i, i, j, j, k, k);
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * Particle-to-cell assignment exactly as the push (solver.cpp:1440-1469) and the deposit (fdtd.cpp:70-77)
   * compute it, written out for the bit-exactness check of the parity tests.
   * push_m[t]  = gather cell m = (k-k0) P + i N1 + j (reference node numbering), -1 when the particle gathers
   *              no mesh field in this sub-step;   dep[t][6] = ip, jp, kp, im, jm, km.
   * ------------------------------------------------------------------------------------------------ */
  __global__ void __launch_bounds__(256)
  particle_cells (const __grid_constant__ BunchDev b, ParticlesDev P, long n, long* __restrict__ push_m, int* __restrict__ dep)
  {
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double x = P.r[0][t], y = P.r[1][t], z = P.r[2][t];
    if (push_m)
      {
	long m = -1;
	const double zr = pmod( z - b.zmin, b.Lz ) + b.zmin;
	const double zlo = ( b.size == 1 ) ? b.zp0 : b.zmin, zhi = ( b.size == 1 ) ? b.zp1 : b.zmax;
	if ( ( b.size > 1 || ( ( zr >= b.zp0 ) && ( zr < b.zp1 ) ) ) && P.e[t] == 1.0 &&
	     x < b.xmax - b.dx && x > b.xmin + b.dx && y < b.ymax - b.dy && y > b.ymin + b.dy && z < zhi && z >= zlo )
	  {
	    double d1;
	    modf( div_by( x - b.xmin, b.dx, b.rdx ), &d1 ); const int i = (int) d1;
	    modf( div_by( y - b.ymin, b.dy, b.rdy ), &d1 ); const int j = (int) d1;
	    modf( div_by( z - b.zmin, b.dz, b.rdz ), &d1 ); const int k = (int) d1;
	    m = (long) ( k - b.k0 - b.kshift ) * b.P + (long) i * b.N1 + j;      /* reference slab numbering     */
	  }
	push_m[t] = m;
      }
    if (dep)
      {
	dep[6 * t + 0] = (int) floor( div_by( x - b.xmin, b.dx, b.rdx ) );
	dep[6 * t + 1] = (int) floor( div_by( y - b.ymin, b.dy, b.rdy ) );
	dep[6 * t + 2] = (int) floor( div_by( z - b.zmin, b.dz, b.rdz ) );
	dep[6 * t + 3] = (int) floor( div_by( P.rm[0][t] - b.xmin, b.dx, b.rdx ) );
	dep[6 * t + 4] = (int) floor( div_by( P.rm[1][t] - b.ymin, b.dy, b.rdy ) );
	dep[6 * t + 5] = (int) floor( div_by( P.rm[2][t] - b.zmin, b.dz, b.rdz ) );
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * Screens (solver.cpp:2205-2257).  One thread per particle, loop over screens; a crossing appends a record
   * { x, y, t, gbx, gby, gbz_lab, upload index of the particle, step } to the screen's buffer through an atomic cursor.
   * screen_records is the test for one particle with its positions at the start (m) and at the end (p) of the field
   * step, called by the stand-alone kernel screen_cross and from the tail of push_particles (mithra_gpu_step).
   * ------------------------------------------------------------------------------------------------ */
  struct ScreensDev
  {
    int                    n;                  /* number of screens, 0 = none                                    */
    const double*          pos;                /* lab-frame positions                                            */
    double*                rec;                /* [screen][capacity][8]                                          */
    unsigned int*          cursor;
    unsigned int           capacity;
    double                 step_id;
    double                 time_bunch;         /* bunch time AFTER the field step's sub-steps                    */
  };

  __device__ __forceinline__ void screen_records (const BunchDev& b, const ScreensDev& S, double xm, double ym, double zm,
						  double xp, double yp, double zp, double gx, double gy, double gz, unsigned int id)
  {
    if (b.size == 1)
      {
	const double zr = pmod( zp - b.zmin, b.Lz ) + b.zmin;
	if ( ! ( ( zr >= b.zp0 ) && ( zr < b.zp1 ) ) ) return;
      }
    const double time_bunch = S.time_bunch;
    const double lzm = b.gamma * ( zm + b.beta * b.c0 * ( time_bunch - b.dt_field + b.dt_shift ) );
    const double lzp = b.gamma * ( zp + b.beta * b.c0 * ( time_bunch + b.dt_shift ) );
    for (int s = 0; s < S.n; s++)
      {
	const double lzs = S.pos[s];
	if (lzm >= lzs) continue;
	if (lzp <  lzs) continue;
	const unsigned int slot = atomicAdd(&S.cursor[s], 1u);
	if (slot >= S.capacity) continue;
	double* r = S.rec + ( (size_t) s * S.capacity + slot ) * 8;
	const double fr = ( lzs - lzm ) / ( lzp - lzm );
	r[0] = xm + fr * ( xp - xm );
	r[1] = ym + fr * ( yp - ym );
	const double tm = b.gamma * ( time_bunch + b.dt_shift - b.dt_field + b.beta / b.c0 * zm );
	const double tp = b.gamma * ( time_bunch + b.dt_shift              + b.beta / b.c0 * zp );
	r[2] = tm + fr * ( tp - tm );
	r[3] = gx; r[4] = gy;
	r[5] = b.gamma * ( gz + b.beta * sqrt( 1.0 + ( gx * gx + gy * gy + gz * gz ) ) );
	r[6] = (double) id; r[7] = S.step_id;
      }
  }

  /* out of line for the push: a crossing is rare, and inlined the test costs the whole kernel 32 registers           */
  __device__ __noinline__ void screen_records_call (const BunchDev& b, const ScreensDev& S, double xm, double ym, double zm,
						    double xp, double yp, double zp, double gx, double gy, double gz, unsigned int id)
  { screen_records(b, S, xm, ym, zm, xp, yp, zp, gx, gy, gz, id); }

  __global__ void __launch_bounds__(256)
  screen_cross (const __grid_constant__ BunchDev b, ParticlesDev P, long n, const ScreensDev S)
  {
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    screen_records(b, S, P.rm[0][t], P.rm[1][t], P.rm[2][t], P.r[0][t], P.r[1][t], P.r[2][t],
		   P.gb[0][t], P.gb[1][t], P.gb[2][t], P.id[t]);
  }

  /* ------------------------------------------------------------------------------------------------
   * ZigZag deposition (fdtd.cpp:47-184): relay point, two segments, each scattering the three current components
   * to the 8 nodes of its cell (+ rho at the end point with space charge, fdtdSC.cpp:141-160).
   *
   * A thread walks MITHRA_DEP_RUN particles that are consecutive in memory -- after the counting sort by cell these
   * are particles of the same or of neighbouring cells -- and keeps the contributions to the 8 nodes of the cell
   * it is currently in in registers: 4 distinct values per component (the reference gives the same J_x to both x
   * nodes of a cell, fdtd.cpp:111-118, likewise y and z) and 8 charge weights.  Both segments of a particle that
   * stays in its cell and all particles of a run that share the cell are summed there; only when the cell changes
   * (and at the end of the run) the 24 (+8) sums go to L2 as FP64 reductions (RED.ADD.F64).  Measured on B200
   * this is what bounds the kernel: the atomics' issue rate, not HBM.  The bounding box of the touched nodes is
   * merged per warp for the stencil (which reads J only there) and the next clear.
   * ------------------------------------------------------------------------------------------------ */
  #define MITHRA_DEP_RUN 8

  template <bool SC>
  struct DepositAcc
  {
    long   m;                         /* first node of the cell the sums belong to, -1 = none                 */
    double jx[4], jy[4], jz[4];
    double rho[SC ? 8 : 1];
  };

  template <bool SC>
  __device__ __forceinline__ void deposit_flush (const BunchDev& b, double* __restrict__ jn, DepositAcc<SC>& a)
  {
    if (a.m < 0) return;
    const long N1 = b.N1, Pp = b.Pp, cs = (long) b.np * b.Pp;
    const long m = a.m;
    const long o[8] = { m, m + N1, m + 1, m + N1 + 1, m + Pp, m + Pp + N1, m + Pp + 1, m + Pp + N1 + 1 };
    double* J0 = jn; double* J1 = jn + cs; double* J2 = jn + 2 * cs;
    atomicAdd(J0 + o[0], a.jx[0]); atomicAdd(J0 + o[1], a.jx[0]); atomicAdd(J0 + o[2], a.jx[1]); atomicAdd(J0 + o[3], a.jx[1]);
    atomicAdd(J0 + o[4], a.jx[2]); atomicAdd(J0 + o[5], a.jx[2]); atomicAdd(J0 + o[6], a.jx[3]); atomicAdd(J0 + o[7], a.jx[3]);
    atomicAdd(J1 + o[0], a.jy[0]); atomicAdd(J1 + o[2], a.jy[0]); atomicAdd(J1 + o[1], a.jy[1]); atomicAdd(J1 + o[3], a.jy[1]);
    atomicAdd(J1 + o[4], a.jy[2]); atomicAdd(J1 + o[6], a.jy[2]); atomicAdd(J1 + o[5], a.jy[3]); atomicAdd(J1 + o[7], a.jy[3]);
    atomicAdd(J2 + o[0], a.jz[0]); atomicAdd(J2 + o[4], a.jz[0]); atomicAdd(J2 + o[1], a.jz[1]); atomicAdd(J2 + o[5], a.jz[1]);
    atomicAdd(J2 + o[2], a.jz[2]); atomicAdd(J2 + o[6], a.jz[2]); atomicAdd(J2 + o[3], a.jz[3]); atomicAdd(J2 + o[7], a.jz[3]);
    if constexpr (SC)
      {
	double* R = jn + 3 * cs;
	#pragma unroll
	for (int q = 0; q < 8; q++) atomicAdd(R + o[q], a.rho[q]);
      }
    a.m = -1;
  }

  template <bool SC>
  __device__ __forceinline__ void deposit_open (const BunchDev& b, double* __restrict__ jn, DepositAcc<SC>& a, long m)
  {
    if (a.m == m) return;
    deposit_flush<SC>(b, jn, a);
    a.m = m;
    #pragma unroll
    for (int q = 0; q < 4; q++) { a.jx[q] = 0.0; a.jy[q] = 0.0; a.jz[q] = 0.0; }
    if constexpr (SC) { _Pragma("unroll") for (int q = 0; q < 8; q++) a.rho[q] = 0.0; }
  }

  /* one segment with midpoint (mx,my,mz) and flux q (jx,jy,jz), fdtd.cpp:92-131                               */
  template <bool SC>
  __device__ __forceinline__ void deposit_segment (const BunchDev& b, DepositAcc<SC>& a, double q,
						   double mx, double my, double mz, double jx, double jy, double jz)
  {
    double c;
    const double dxp = modf( div_by( mx - b.xmin, b.dx, b.rdx ), &c );
    const double dyp = modf( div_by( my - b.ymin, b.dy, b.rdy ), &c );
    const double dzp = modf( div_by( mz - b.zmin, b.dz, b.rdz ), &c );
    const double x1 = 1.0 - dxp, x2 = dxp, y1 = 1.0 - dyp, y2 = dyp, z1 = 1.0 - dzp, z2 = dzp;
    const double h = q * 0.5;
    a.jx[0] += h * y1 * z1 * jx; a.jx[1] += h * y2 * z1 * jx; a.jx[2] += h * y1 * z2 * jx; a.jx[3] += h * y2 * z2 * jx;
    a.jy[0] += h * x1 * z1 * jy; a.jy[1] += h * x2 * z1 * jy; a.jy[2] += h * x1 * z2 * jy; a.jy[3] += h * x2 * z2 * jy;
    a.jz[0] += h * x1 * y1 * jz; a.jz[1] += h * x2 * y1 * jz; a.jz[2] += h * x1 * y2 * jz; a.jz[3] += h * x2 * y2 * jz;
  }

  /* one particle of FdTd::currentUpdate (fdtd.cpp:47-184): end point p, start point m, charge q; the sums go to the
   * accumulator of the cell they belong to (flushed when the cell changes), `valid` and the node box are updated       */
  template <bool SC>
  __device__ __forceinline__ void deposit_particle (const BunchDev& b, double* __restrict__ jn, DepositAcc<SC>& acc,
						    double rpx, double rpy, double rpz, double rmx, double rmy, double rmz, double q,
						    double zlo, double zhi, bool& valid, int& i0, int& i1, int& j0, int& j1, int& k0, int& k1)
  {
    const bool bpf = ( rpx < b.xmax - b.dx && rpx > b.xmin + b.dx && rpy < b.ymax - b.dy && rpy > b.ymin + b.dy &&
    		   rpz < zhi && rpz >= zlo );
    const bool bmf = ( rmx < b.xmax - b.dx && rmx > b.xmin + b.dx && rmy < b.ymax - b.dy && rmy > b.ymin + b.dy &&
    		   rmz < zhi && rmz >= zlo );
    if (!bpf && !bmf) return;

    const int ip = (int) floor( div_by( rpx - b.xmin, b.dx, b.rdx ) ), jp = (int) floor( div_by( rpy - b.ymin, b.dy, b.rdy ) ), kp = (int) floor( div_by( rpz - b.zmin, b.dz, b.rdz ) );
    const int im = (int) floor( div_by( rmx - b.xmin, b.dx, b.rdx ) ), jm = (int) floor( div_by( rmy - b.ymin, b.dy, b.rdy ) ), km = (int) floor( div_by( rmz - b.zmin, b.dz, b.rdz ) );

    /* relay point, fdtd.cpp:80-85 */
    const double rx = fmin( min(im, ip) * b.dx + b.dx + b.xmin, fmax( max(im, ip) * b.dx + b.xmin, 0.5 * ( rmx + rpx ) ) );
    const double ry = fmin( min(jm, jp) * b.dy + b.dy + b.ymin, fmax( max(jm, jp) * b.dy + b.ymin, 0.5 * ( rmy + rpy ) ) );
    const double rz = fmin( min(km, kp) * b.dz + b.dz + b.zmin, fmax( max(km, kp) * b.dz + b.zmin, 0.5 * ( rmz + rpz ) ) );

    valid = true;
    if (bpf)
      {
        deposit_open<SC>(b, jn, acc, (long) ( kp - b.k0 ) * b.Pp + (long) ip * b.N1 + jp);
        deposit_segment<SC>(b, acc, q, 0.5 * ( rpx + rx ), 0.5 * ( rpy + ry ), 0.5 * ( rpz + rz ), rpx - rx, rpy - ry, rpz - rz);
        if constexpr (SC)
          {
    	double c;
    	const double dxp = modf( div_by( rpx - b.xmin, b.dx, b.rdx ), &c ), dyp = modf( div_by( rpy - b.ymin, b.dy, b.rdy ), &c ), dzp = modf( div_by( rpz - b.zmin, b.dz, b.rdz ), &c );
    	const double x1 = 1.0 - dxp, x2 = dxp, y1 = 1.0 - dyp, y2 = dyp, z1 = 1.0 - dzp, z2 = dzp;
    	acc.rho[0] += q * x1 * y1 * z1; acc.rho[1] += q * x2 * y1 * z1; acc.rho[2] += q * x1 * y2 * z1; acc.rho[3] += q * x2 * y2 * z1;
    	acc.rho[4] += q * x1 * y1 * z2; acc.rho[5] += q * x2 * y1 * z2; acc.rho[6] += q * x1 * y2 * z2; acc.rho[7] += q * x2 * y2 * z2;
          }
        i0 = min(i0, ip); i1 = max(i1, ip + 1); j0 = min(j0, jp); j1 = max(j1, jp + 1); k0 = min(k0, kp - b.k0); k1 = max(k1, kp - b.k0 + 1);
      }
    if (bmf)
      {
        deposit_open<SC>(b, jn, acc, (long) ( km - b.k0 ) * b.Pp + (long) im * b.N1 + jm);
        deposit_segment<SC>(b, acc, q, 0.5 * ( rmx + rx ), 0.5 * ( rmy + ry ), 0.5 * ( rmz + rz ), rx - rmx, ry - rmy, rz - rmz);
        i0 = min(i0, im); i1 = max(i1, im + 1); j0 = min(j0, jm); j1 = max(j1, jm + 1); k0 = min(k0, km - b.k0); k1 = max(k1, km - b.k0 + 1);
      }
  }

  #ifndef MITHRA_DEP_MINBLOCKS
  #define MITHRA_DEP_MINBLOCKS 4
  #endif
  template <bool SC>
  __global__ void __launch_bounds__(128, SC ? MITHRA_DEP_MINBLOCKS - 1 : MITHRA_DEP_MINBLOCKS)
  deposit_current (const __grid_constant__ BunchDev b, ParticlesDev P, long n, double* __restrict__ jn, Box* __restrict__ jbox, int run)
  {
    const long t0 = ( (long) blockIdx.x * blockDim.x + threadIdx.x ) * run;
    bool valid = false; int i0 = 0x7fffffff, i1 = -1, j0 = 0x7fffffff, j1 = -1, k0 = 0x7fffffff, k1 = -1;
    DepositAcc<SC> acc; acc.m = -1;
    const double zlo = ( b.size == 1 ) ? b.zp0 : b.zmin, zhi = ( b.size == 1 ) ? b.zp1 : b.zmax;

    /* the particle of the NEXT iteration is loaded before the current one is worked on: a thread's particles are 64 bytes
     * apart in every array, the loads miss L1 and nothing else hides their latency (the flush is fire-and-forget)       */
    double nx = 0.0, ny = 0.0, nz = 0.0, nmx = 0.0, nmy = 0.0, nmz = 0.0, nq = 0.0;
    if (t0 < n)
      { nx = __ldg(P.r[0] + t0); ny = __ldg(P.r[1] + t0); nz = __ldg(P.r[2] + t0); nmx = __ldg(P.rm[0] + t0); nmy = __ldg(P.rm[1] + t0); nmz = __ldg(P.rm[2] + t0); nq = __ldg(P.q + t0); }

    for (int r = 0; r < run; r++)
      {
	const long t = t0 + r;
	if (t >= n) break;
	const double rpx = nx,  rpy = ny,  rpz = nz;
	const double rmx = nmx, rmy = nmy, rmz = nmz;
	const double q = nq;
	if (r + 1 < run && t + 1 < n)
	  { nx = __ldg(P.r[0] + t + 1); ny = __ldg(P.r[1] + t + 1); nz = __ldg(P.r[2] + t + 1); nmx = __ldg(P.rm[0] + t + 1); nmy = __ldg(P.rm[1] + t + 1); nmz = __ldg(P.rm[2] + t + 1); nq = __ldg(P.q + t + 1); }

	deposit_particle<SC>(b, jn, acc, rpx, rpy, rpz, rmx, rmy, rmz, q, zlo, zhi, valid, i0, i1, j0, j1, k0, k1);
      }
    deposit_flush<SC>(b, jn, acc);
    warp_box_merge(jbox, valid, i0, i1, j0, j1, k0, k1);
  }

  /* ------------------------------------------------------------------------------------------------
   * Push: `nsub` consecutive sub-steps of Solver::bunchUpdate for every particle (solver.cpp:1437-1549).
   * With first_of_step the start-of-step position is saved to rm first (solver.cpp:1311-1312); with scr.n > 0 the
   * lab-frame screens are tested at the end (solver.cpp:2205-2257), which saves screen_cross's pass over the bunch.
   * E,B of the mesh are gathered from the interleaved float4 pairs written by eval_eb_box.
   * ------------------------------------------------------------------------------------------------ */
  #ifndef MITHRA_PUSH_MINBLOCKS
  #define MITHRA_PUSH_MINBLOCKS 5                      /* 5 CTAs per SM = at most 102 registers: the kernel lives on occupancy */
  #endif
  template <bool BEAMS>                                /* false: static undulators only, no optical beam code in the kernel */
  __global__ void __launch_bounds__(128, MITHRA_PUSH_MINBLOCKS)
  push_particles (const __grid_constant__ BunchDev b, ParticlesDev P, long n, const float4* __restrict__ eb,
		  double time_bunch, int nsub, int first_of_step, Box* __restrict__ pbox, unsigned int* __restrict__ n_outside,
		  unsigned char* __restrict__ emask, const ScreensDev scr)
  {
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    bool boxvalid = false; int bi = 0, bj = 0, bk = 0;

    if (t < n)
      {
	double x = P.r[0][t], y = P.r[1][t], z = P.r[2][t];
	double gx = P.gb[0][t], gy = P.gb[1][t], gz = P.gb[2][t];
	double e = P.e[t];
	if (first_of_step) { P.rm[0][t] = x; P.rm[1][t] = y; P.rm[2][t] = z; }

	double tb = time_bunch;
	for (int s = 0; s < nsub; s++, tb += b.dt_bunch)
	  {
	    /* ownership with the periodic z wrap (solver.cpp:1440-1441)                                    */
	    if (b.size == 1)
	      {
		const double zr = pmod( z - b.zmin, b.Lz ) + b.zmin;
		if ( ! ( ( zr >= b.zp0 ) && ( zr < b.zp1 ) ) ) continue;
	      }
	    /* several slabs: the list of a slab is what it owns for this field step (migrate once per step), and
	     * "inside the mesh" is the single-slab test of the whole mesh, [zmin, zmax)                          */
	    const bool b1x = ( x < b.xmax - b.dx && x > b.xmin + b.dx );
	    const bool b1y = ( y < b.ymax - b.dy && y > b.ymin + b.dy );
	    const bool b1z = ( b.size == 1 ) ? ( z < b.zp1 && z >= b.zp0 ) : ( z < b.zmax && z >= b.zmin );

	    /* The eight E/B nodes of the cell (one 32-byte sector each) are asked for NOW, while the analytic undulator /
	     * beam fields below are computed (some hundred FP64 instructions): the gather further down then finds them in
	     * L1 instead of waiting for DRAM twice.  A prefetch takes no register and returns nothing.                  */
	    if (e == 1.0 && b1x && b1y && b1z)
	      {
		const int pi = (int) div_by( x - b.xmin, b.dx, b.rdx ), pj = (int) div_by( y - b.ymin, b.dy, b.rdy );
		const int pk = (int) div_by( z - b.zmin, b.dz, b.rdz ) - b.k0;
		const float4* q = eb + 2 * ( (long) pk * b.P + (long) pi * b.N1 + pj );
		const long sN = 2L * b.N1, sP = 2L * b.P;
		prefetch_l1(q);          prefetch_l1(q + 2);          prefetch_l1(q + sN);      prefetch_l1(q + sN + 2);
		prefetch_l1(q + sP);     prefetch_l1(q + sP + 2);     prefetch_l1(q + sP + sN); prefetch_l1(q + sP + sN + 2);
	      }

	    V3 et = v3(0.0, 0.0, 0.0), bt = v3(0.0, 0.0, 0.0);

	    /* undulatorField, solver.cpp:1798-1880                                                         */
	    for (int u = 0; u < b.n_und; u++)
	      {
		const UndulatorDev& U = b.und[u];
		if (!BEAMS || U.type == MITHRA_UNDULATOR_STATIC)
		  {
		    const double lz = b.gamma * ( z + b.beta * b.c0 * ( tb + b.dt_shift ) ) - U.rb;
		    const double ly = x * U.ct + y * U.st;
		    static_undulator(U, b.gamma, b.c0 * b.beta, lz, ly, et, bt);
		  }
		else
		  {
		    const V3 rl = v3(x, y, b.gamma * ( z + b.beta * b.c0 * ( tb + b.dt_shift ) ));
		    const double t0 = b.gamma * ( tb + b.dt_shift + b.beta / b.c0 * z );
		    const V3 rv = v3(rl.x - U.beam.position[0], rl.y - U.beam.position[1], rl.z - U.beam.position[2]);
		    const double zz = dot3(rv, v3a(U.beam.direction));
		    V3 eT, bT;
		    beam_fields(U.beam, b.c0, rv, zz, t0 - zz / b.c0, t0 + zz / b.c0, eT, bT);
		    bt.x += b.gamma * ( bT.x + b.beta / b.c0 * eT.y );
		    bt.y += b.gamma * ( bT.y - b.beta / b.c0 * eT.x );
		    bt.z += bT.z;
		    et.x += b.gamma * ( eT.x - b.beta * b.c0 * bT.y );
		    et.y += b.gamma * ( eT.y + b.beta * b.c0 * bT.x );
		    et.z += eT.z;
		  }
	      }

	    /* externalField, solver.cpp:1886-1947                                                          */
	    if (BEAMS && b.n_ext > 0)
	      {
		const V3 rl = v3(x, y, b.gamma * ( z + b.beta * b.c0 * ( tb + b.dt_shift ) ));
		const double t0 = b.gamma * ( tb + b.dt_shift + b.beta / b.c0 * z );
		for (int u = 0; u < b.n_ext; u++)
		  {
		    const MithraBeam& S = b.ext[u];
		    const V3 rv = v3(rl.x - S.position[0], rl.y - S.position[1], rl.z - S.position[2]);
		    const double zz = dot3(rv, v3a(S.direction));
		    V3 eT, bT;
		    beam_fields(S, b.c0, rv, zz, t0 - zz / b.c0, t0 + zz / b.c0, eT, bT);
		    bt.x += b.gamma * ( bT.x + b.beta / b.c0 * eT.y );
		    bt.y += b.gamma * ( bT.y - b.beta / b.c0 * eT.x );
		    bt.z += bT.z;
		    et.x += b.gamma * ( eT.x - b.beta * b.c0 * bT.y );
		    et.y += b.gamma * ( eT.y + b.beta * b.c0 * bT.x );
		    et.z += eT.z;
		  }
	      }

	    if (e == 1.0)
	      {
		if (b1x && b1y && b1z)
		  {
		    double d1;
		    const double dxr = modf( div_by( x - b.xmin, b.dx, b.rdx ), &d1 ); const int i = (int) d1;
		    const double dyr = modf( div_by( y - b.ymin, b.dy, b.rdy ), &d1 ); const int j = (int) d1;
		    const double dzr = modf( div_by( z - b.zmin, b.dz, b.rdz ), &d1 ); const int k = (int) d1;
		    const long m = (long) ( k - b.k0 ) * b.P + (long) i * b.N1 + j;
		    const long N1 = b.N1, Pn = b.P;
		    const long off[8] = { 0, N1, 1, N1 + 1, Pn, Pn + N1, Pn + 1, Pn + N1 + 1 };
		    const double w[8] = {
		      ( 1.0 - dxr ) * ( 1.0 - dyr ) * ( 1.0 - dzr ),         dxr   * ( 1.0 - dyr ) * ( 1.0 - dzr ),
		      ( 1.0 - dxr ) *         dyr   * ( 1.0 - dzr ),         dxr   *         dyr   * ( 1.0 - dzr ),
		      ( 1.0 - dxr ) * ( 1.0 - dyr ) *         dzr,           dxr   * ( 1.0 - dyr ) *         dzr,
		      ( 1.0 - dxr ) *         dyr   *         dzr,           dxr   *         dyr   *         dzr };
		    /* two batches of four nodes (plane k, then k+1): half the registers in flight, same sum order   */
		    double ex = et.x, ey = et.y, ez = et.z, bx = bt.x, by = bt.y, bz = bt.z;
		    #pragma unroll
		    for (int hb = 0; hb < 8; hb += 4)
		      {
			float4 fe[4], fb[4];
			#pragma unroll
			for (int q = 0; q < 4; q++) { fe[q] = __ldg(&eb[2 * (m + off[hb + q])]); fb[q] = __ldg(&eb[2 * (m + off[hb + q]) + 1]); }
			#pragma unroll
			for (int q = 0; q < 4; q++)
			  {
			    ex += w[hb + q] * fe[q].x; ey += w[hb + q] * fe[q].y; ez += w[hb + q] * fe[q].z;
			    bx += w[hb + q] * fb[q].x; by += w[hb + q] * fb[q].y; bz += w[hb + q] * fb[q].z;
			  }
		      }
		    et.x = ex; et.y = ey; et.z = ez;
		    bt.x = bx; bt.y = by; bt.z = bz;
		  }
		else if ( !b1x && !b1y && b1z )
		  atomicAdd(n_outside, 1u);
	      }
	    else if (b.n_und > 0)
	      {
		const double lz = b.gamma * ( z + b.beta * b.c0 * ( tb + b.dt_shift ) );
		e = ( lz > - b.und0_dist ) ? 1.0 : 0.0;
	      }
	    else
	      e = 1.0;

	    /* Boris rotation, solver.cpp:1519-1541                                                         */
	    const double mx = gx + b.r1 * et.x, my = gy + b.r1 * et.y, mz = gz + b.r1 * et.z;          /* gb-   */
	    double cx = my * bt.z - mz * bt.y, cy = mz * bt.x - mx * bt.z, cz = mx * bt.y - my * bt.x;
	    const double d1 = sqrt( 1.0 + ( mx * mx + my * my + mz * mz ) );
	    const double f1 = b.r2 / d1;
	    const double px = f1 * cx + mx, py = f1 * cy + my, pz = f1 * cz + mz;                      /* gb'   */
	    cx = py * bt.z - pz * bt.y; cy = pz * bt.x - px * bt.z; cz = px * bt.y - py * bt.x;
	    const double f2 = 2.0 / ( div_by( d1, b.r2, b.rr2 ) + b.r2 / d1 * ( bt.x * bt.x + bt.y * bt.y + bt.z * bt.z ) );
	    const double lx = f2 * cx + mx, ly = f2 * cy + my, lzz = f2 * cz + mz;                     /* gb+   */
	    gx = lx + b.r1 * et.x; gy = ly + b.r1 * et.y; gz = lzz + b.r1 * et.z;

	    const double f3 = b.dtb / sqrt( 1.0 + ( gx * gx + gy * gy + gz * gz ) );
	    const double drx = f3 * gx, dry = f3 * gy, drz = f3 * gz;
	    x += drx; y += dry; z += drz;
	  }

	P.r[0][t] = x; P.r[1][t] = y; P.r[2][t] = z;
	P.gb[0][t] = gx; P.gb[1][t] = gy; P.gb[2][t] = gz;
	P.e[t] = e;

	/* Solver::screenProfile on the way out (mithra_gpu_step): the start-of-step position is this thread's own store  */
	if (scr.n > 0)
	  {
	    /* nothing to do unless the step straddles a screen: the cheap part of the test stays inline              */
	    const double zm0 = P.rm[2][t];
	    const double lzm = b.gamma * ( zm0 + b.beta * b.c0 * ( scr.time_bunch - b.dt_field + b.dt_shift ) );
	    const double lzp = b.gamma * ( z   + b.beta * b.c0 * ( scr.time_bunch + b.dt_shift ) );
	    bool any = false;
	    for (int s = 0; s < scr.n; s++) any = any || ( lzm < scr.pos[s] && lzp >= scr.pos[s] );
	    if (any) screen_records_call(b, scr, P.rm[0][t], P.rm[1][t], zm0, x, y, z, gx, gy, gz, P.id[t]);
	  }

	/* where the particle can gather and deposit during the NEXT field step                                          */
	unsigned int bits;
	boxvalid = particle_reach(b, x, y, z, bi, bj, bk, bits);
	if (boxvalid) mark_eb_pencil(b, emask, bi, bj, bk, bits);
      }
    warp_box_merge(pbox, boxvalid, bi, bi, bj, bj, bk, bk);
  }

  /* ------------------------------------------------------------------------------------------------
   * Moments of the bunch (Solver::bunchSample, solver.cpp:1582-1608): the 13 raw sums q, q r, q r^2, q gb, q gb^2
   * over the particles this slab owns.  Block reduction in a fixed order (thread t sums particles t, t + stride, ...;
   * shared-memory tree), one partial row per block; the host adds the rows in order -- deterministic.
   * ------------------------------------------------------------------------------------------------ */
  #define MITHRA_MOMENTS 13
  __global__ void __launch_bounds__(256)
  bunch_moments (const __grid_constant__ BunchDev b, ParticlesDev P, long n, double* __restrict__ partial)
  {
    __shared__ double red[MITHRA_MOMENTS][256];
    double s[MITHRA_MOMENTS];
    #pragma unroll
    for (int q = 0; q < MITHRA_MOMENTS; q++) s[q] = 0.0;
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long) gridDim.x * blockDim.x)
      {
	const double z = P.r[2][t];
	if (b.size == 1)
	  {
	    const double zr = pmod( z - b.zmin, b.Lz ) + b.zmin;            /* particleInProcessor, solver.cpp:2292-2298 */
	    if ( ! ( ( zr >= b.zp0 ) && ( zr < b.zp1 ) ) ) continue;
	  }
	const double q = P.q[t];
	const double r[3] = { P.r[0][t], P.r[1][t], z }, g[3] = { P.gb[0][t], P.gb[1][t], P.gb[2][t] };
	s[0] += q;
	#pragma unroll
	for (int l = 0; l < 3; l++)
	  {
	    s[1 + l]  += q * r[l];
	    s[4 + l]  += r[l] * r[l] * q;
	    s[7 + l]  += q * g[l];
	    s[10 + l] += g[l] * g[l] * q;
	  }
      }
    #pragma unroll
    for (int q = 0; q < MITHRA_MOMENTS; q++) red[q][threadIdx.x] = s[q];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1)
      {
	if (threadIdx.x < w)
	  {
	    #pragma unroll
	    for (int q = 0; q < MITHRA_MOMENTS; q++) red[q][threadIdx.x] += red[q][threadIdx.x + w];
	  }
	__syncthreads();
      }
    if (threadIdx.x < MITHRA_MOMENTS) partial[(size_t) blockIdx.x * MITHRA_MOMENTS + threadIdx.x] = red[threadIdx.x][0];
  }

  /* ------------------------------------------------------------------------------------------------
   * Radiated power (radiation.cpp:154-215).  One CTA = PX pixels x MS window slices.  The thread whose slice
   * owns the current ring slot interpolates E,B between the two z planes, boosts to the lab frame and stores
   * the 4 values in the ring; every thread then accumulates its share of the length-Nf DFT sums.  Slices are
   * combined in shared memory in a fixed order and each CTA writes one partial per wavelength; power_finish
   * adds the partials in order (deterministic) and appends pc * sum to the output row.
   *
   * ring layout: fdt[plane][slot][4][npx]   (pixel fastest => coalesced), twiddles ep[l][m] = exp(+i w_l m dt).
   * ------------------------------------------------------------------------------------------------ */
  struct PowerDev
  {
    int    N, Nl, Nf, npx, ni, nj;          /* planes, wavelengths, window, pixels = ni*nj                  */
    double pc;
    double gamma, beta, c0;
  };

  #define MITHRA_POWER_PX 32
  #define MITHRA_POWER_MS 8

  template <bool SC>
  __global__ void __launch_bounds__(MITHRA_POWER_PX * MITHRA_POWER_MS)
  power_dft (const FieldDev f, const PowerDev pw, const double* __restrict__ anp1, const double* __restrict__ an,
	     const float4* __restrict__ ebn, double* __restrict__ fdt, const double2* __restrict__ ep, int plane, int kplane, double dzr, int slot,
	     double* __restrict__ partial)
  {
    __shared__ double red[4][2][MITHRA_POWER_MS][MITHRA_POWER_PX];
    __shared__ double pix[MITHRA_POWER_PX];
    const int lp = threadIdx.x % MITHRA_POWER_PX, ms = threadIdx.x / MITHRA_POWER_PX;
    const int px = blockIdx.x * MITHRA_POWER_PX + lp;
    const bool live = px < pw.npx;
    double* ring = fdt + (size_t) plane * pw.Nf * 4 * pw.npx;

    double cur[4] = { 0.0, 0.0, 0.0, 0.0 };
    if (live && (slot % MITHRA_POWER_MS) == ms)
      {
	const int i = 2 + px / pw.nj, j = 2 + px % pw.nj;
	/* fieldEvaluate on mi and mi + N1N0 with the z-end copy rule (planes 0 / np-1)                     */
	int ka = kplane, kb = kplane + 1;
	if (ka == 0 && f.rank == 0) ka = 1;
	if (kb == f.np - 1 && f.rank == f.size - 1) kb = f.np - 2;
	/* a plane this slab does not update is a ghost whose E/B came from the neighbour (full plane)     */
	EB A, B;
	if (ka < f.kb && f.rank != 0)
	  {
	    const float4 e = ebn[2 * ((long) ka * f.P + (long) i * f.N1 + j)], bb = ebn[2 * ((long) ka * f.P + (long) i * f.N1 + j) + 1];
	    A.e[0] = e.x; A.e[1] = e.y; A.e[2] = e.z; A.b[0] = bb.x; A.b[1] = bb.y; A.b[2] = bb.z;
	  }
	else A = eval_eb_node<SC>(f, anp1, an, i, j, ka);
	if (kb == f.np - 1 && f.rank != f.size - 1)
	  {
	    const float4 e = ebn[2 * ((long) kb * f.P + (long) i * f.N1 + j)], bb = ebn[2 * ((long) kb * f.P + (long) i * f.N1 + j) + 1];
	    B.e[0] = e.x; B.e[1] = e.y; B.e[2] = e.z; B.b[0] = bb.x; B.b[1] = bb.y; B.b[2] = bb.z;
	  }
	else B = eval_eb_node<SC>(f, anp1, an, i, j, kb);
	const double et0 = ( 1.0 - dzr ) * A.e[0] + dzr * B.e[0];
	const double et1 = ( 1.0 - dzr ) * A.e[1] + dzr * B.e[1];
	const double bt0 = ( 1.0 - dzr ) * A.b[0] + dzr * B.b[0];
	const double bt1 = ( 1.0 - dzr ) * A.b[1] + dzr * B.b[1];
	cur[0] = pw.gamma * ( et0 + pw.c0 * pw.beta * bt1 );
	cur[1] = pw.gamma * ( et1 - pw.c0 * pw.beta * bt0 );
	cur[2] = pw.gamma * ( bt0 - pw.beta / pw.c0 * et1 );
	cur[3] = pw.gamma * ( bt1 + pw.beta / pw.c0 * et0 );
	#pragma unroll
	for (int q = 0; q < 4; q++) ring[( (size_t) slot * 4 + q ) * pw.npx + px] = cur[q];
      }

    for (int l = 0; l < pw.Nl; l++)
      {
	double s[4][2] = { { 0.0, 0.0 }, { 0.0, 0.0 }, { 0.0, 0.0 }, { 0.0, 0.0 } };
	if (live)
	  for (int m = ms; m < pw.Nf; m += MITHRA_POWER_MS)
	    {
	      const double2 w = ep[(size_t) l * pw.Nf + m];
	      double v[4];
	      if (m == slot) { v[0] = cur[0]; v[1] = cur[1]; v[2] = cur[2]; v[3] = cur[3]; }
	      else
		{
		  #pragma unroll
		  for (int q = 0; q < 4; q++) v[q] = ring[( (size_t) m * 4 + q ) * pw.npx + px];
		}
	      /* ew1 += f0 ep; bw1 += f3 em; ew2 += f1 ep; bw2 += f2 em   (em = conj ep)                      */
	      s[0][0] += v[0] * w.x; s[0][1] += v[0] * w.y;
	      s[1][0] += v[3] * w.x; s[1][1] -= v[3] * w.y;
	      s[2][0] += v[1] * w.x; s[2][1] += v[1] * w.y;
	      s[3][0] += v[2] * w.x; s[3][1] -= v[2] * w.y;
	    }
	#pragma unroll
	for (int q = 0; q < 4; q++) { red[q][0][ms][lp] = s[q][0]; red[q][1][ms][lp] = s[q][1]; }
	__syncthreads();
	if (ms == 0)
	  {
	    double t[4][2];
	    #pragma unroll
	    for (int q = 0; q < 4; q++)
	      {
		double re = 0.0, im = 0.0;
		for (int a = 0; a < MITHRA_POWER_MS; a++) { re += red[q][0][a][lp]; im += red[q][1][a][lp]; }
		t[q][0] = re; t[q][1] = im;
	      }
	    /* Re(ew1 bw1) - Re(ew2 bw2) */
	    pix[lp] = live ? ( ( t[0][0] * t[1][0] - t[0][1] * t[1][1] ) - ( t[2][0] * t[3][0] - t[2][1] * t[3][1] ) ) : 0.0;
	  }
	__syncthreads();
	if (threadIdx.x == 0)
	  {
	    double acc = 0.0;
	    for (int a = 0; a < MITHRA_POWER_PX; a++) acc += pix[a];
	    partial[( (size_t) plane * pw.Nl + l ) * gridDim.x + blockIdx.x] = acc;
	  }
	__syncthreads();
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * Power map (Solver::powerVisualize, radiation.cpp:324-391): the same lab-frame fields and length-Nf DFT as the
   * power sampling, but kept per pixel for 1 <= i <= N0-2, 1 <= j <= N1-2 and one harmonic.  One thread per pixel
   * sums the window in the reference's order m = 0 .. Nf-1, so the map is bit-identical to the reference's for
   * identical E/B.  ring layout: fdt[slot][4][P] (pixel fastest), twiddles ep[m] = exp(+i w m dt).
   * ------------------------------------------------------------------------------------------------ */
  template <bool SC>
  __global__ void __launch_bounds__(128)
  power_map (const FieldDev f, const PowerDev pw, const double* __restrict__ anp1, const double* __restrict__ an,
	     const float4* __restrict__ ebn, double* __restrict__ fdt, const double2* __restrict__ ep, int kplane, double dzr, int slot,
	     double* __restrict__ pL)
  {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ni = f.N0 - 2, nj = f.N1 - 2;
    if (t >= ni * nj) return;
    const int i = 1 + t / nj, j = 1 + t % nj;
    const int px = i * f.N1 + j;
    int ka = kplane, kb = kplane + 1;
    if (ka == 0 && f.rank == 0) ka = 1;
    if (kb == f.np - 1 && f.rank == f.size - 1) kb = f.np - 2;
    EB A, B;
    if (ka < f.kb && f.rank != 0)
      {
	const float4 e = ebn[2 * ((long) ka * f.P + px)], bb = ebn[2 * ((long) ka * f.P + px) + 1];
	A.e[0] = e.x; A.e[1] = e.y; A.e[2] = e.z; A.b[0] = bb.x; A.b[1] = bb.y; A.b[2] = bb.z;
      }
    else A = eval_eb_node<SC>(f, anp1, an, i, j, ka);
    if (kb == f.np - 1 && f.rank != f.size - 1)
      {
	const float4 e = ebn[2 * ((long) kb * f.P + px)], bb = ebn[2 * ((long) kb * f.P + px) + 1];
	B.e[0] = e.x; B.e[1] = e.y; B.e[2] = e.z; B.b[0] = bb.x; B.b[1] = bb.y; B.b[2] = bb.z;
      }
    else B = eval_eb_node<SC>(f, anp1, an, i, j, kb);
    const double et0 = ( 1.0 - dzr ) * A.e[0] + dzr * B.e[0];
    const double et1 = ( 1.0 - dzr ) * A.e[1] + dzr * B.e[1];
    const double bt0 = ( 1.0 - dzr ) * A.b[0] + dzr * B.b[0];
    const double bt1 = ( 1.0 - dzr ) * A.b[1] + dzr * B.b[1];
    double cur[4];
    cur[0] = pw.gamma * ( et0 + pw.c0 * pw.beta * bt1 );
    cur[1] = pw.gamma * ( et1 - pw.c0 * pw.beta * bt0 );
    cur[2] = pw.gamma * ( bt0 - pw.beta / pw.c0 * et1 );
    cur[3] = pw.gamma * ( bt1 + pw.beta / pw.c0 * et0 );
    const size_t P = (size_t) f.P;
    #pragma unroll
    for (int q = 0; q < 4; q++) fdt[( (size_t) slot * 4 + q ) * P + px] = cur[q];

    double e1r = 0.0, e1i = 0.0, b1r = 0.0, b1i = 0.0, e2r = 0.0, e2i = 0.0, b2r = 0.0, b2i = 0.0;
    for (int m = 0; m < pw.Nf; m++)
      {
	const double2 w = ep[m];
	double v[4];
	if (m == slot) { v[0] = cur[0]; v[1] = cur[1]; v[2] = cur[2]; v[3] = cur[3]; }
	else
	  {
	    #pragma unroll
	    for (int q = 0; q < 4; q++) v[q] = fdt[( (size_t) m * 4 + q ) * P + px];
	  }
	e1r += v[0] * w.x; e1i += v[0] * w.y;
	b1r += v[3] * w.x; b1i += v[3] * ( - w.y );
	e2r += v[1] * w.x; e2i += v[1] * w.y;
	b2r += v[2] * w.x; b2i += v[2] * ( - w.y );
      }
    pL[px] = pw.pc * ( ( e1r * b1r - e1i * b1i ) - ( e2r * b2r - e2i * b2i ) );
  }

  /* ------------------------------------------------------------------------------------------------
   * FdTd::fieldSample / FdTdSC::fieldSample (fdtd.cpp:851-950): E, B (floats evaluated at the 8 nodes of the cell like
   * fieldEvaluate) and A^n interpolated to sampling points, one thread per point, weights and summation in the
   * reference's order (m, m+N1, m+1, m+N1+1, then plane k+1).  out[9 t ..] = et[3], bt[3], at[3]; mine[t] = 1 when the
   * point lies in this slab's [zp0, zp1) (solver.cpp:896).  Nodes on a ghost plane take the E/B the neighbour
   * evaluated for the whole plane.  The lab-frame combinations and unit factors stay with the host writer.
   * ------------------------------------------------------------------------------------------------ */
  template <bool SC>
  __global__ void __launch_bounds__(64)
  field_sample (const FieldDev f, const __grid_constant__ BunchDev b, const double* __restrict__ anp1, const double* __restrict__ an,
		const float4* __restrict__ ebn, const double* __restrict__ pos, int n, double* __restrict__ out,
		unsigned char* __restrict__ mine)
  {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double x = pos[3 * t], y = pos[3 * t + 1], z = pos[3 * t + 2];
    const bool own = ( z < b.zp1 && z >= b.zp0 );
    mine[t] = own ? 1 : 0;
    double* o = out + 9 * t;
    #pragma unroll
    for (int q = 0; q < 9; q++) o[q] = 0.0;
    if (!own) return;
    double c1;
    const double dxr = modf( ( x - b.xmin ) / b.dx, &c1 ); const int i = (int) c1;
    const double dyr = modf( ( y - b.ymin ) / b.dy, &c1 ); const int j = (int) c1;
    const double dzr = modf( ( z - b.zmin ) / b.dz, &c1 ); const int k = (int) c1 - b.k0;
    if (i < 1 || i > f.N0 - 3 || j < 1 || j > f.N1 - 3 || k < 0 || k > f.np - 2) { mine[t] = 0; return; }
    const double w[8] = {
      ( 1.0 - dxr ) * ( 1.0 - dyr ) * ( 1.0 - dzr ),         dxr   * ( 1.0 - dyr ) * ( 1.0 - dzr ),
      ( 1.0 - dxr ) *         dyr   * ( 1.0 - dzr ),         dxr   *         dyr   * ( 1.0 - dzr ),
      ( 1.0 - dxr ) * ( 1.0 - dyr ) *         dzr,           dxr   * ( 1.0 - dyr ) *         dzr,
      ( 1.0 - dxr ) *         dyr   *         dzr,           dxr   *         dyr   *         dzr };
    const int di[8] = { 0, 1, 0, 1, 0, 1, 0, 1 }, dj[8] = { 0, 0, 1, 1, 0, 0, 1, 1 }, dk[8] = { 0, 0, 0, 0, 1, 1, 1, 1 };
    const long cs = (long) f.np * f.Pp;
    double et[3] = { 0.0, 0.0, 0.0 }, bt[3] = { 0.0, 0.0, 0.0 }, at[3] = { 0.0, 0.0, 0.0 };
    for (int q = 0; q < 8; q++)
      {
	const int ii = i + di[q], jj = j + dj[q]; int kk = k + dk[q];
	EB v;
	const bool ghost = ( kk < f.kb && f.rank != 0 ) || ( kk == f.np - 1 && f.rank != f.size - 1 );
	if (ghost)
	  {
	    const long m = (long) kk * f.P + (long) ii * f.N1 + jj;
	    const float4 e = ebn[2 * m], bb = ebn[2 * m + 1];
	    v.e[0] = e.x; v.e[1] = e.y; v.e[2] = e.z; v.b[0] = bb.x; v.b[1] = bb.y; v.b[2] = bb.z;
	  }
	else
	  {
	    int ke = kk;
	    if (ke == 0 && f.rank == 0) ke = 1;                               /* the copied end planes, fdtd.cpp:754-773 */
	    if (ke == f.np - 1 && f.rank == f.size - 1) ke = f.np - 2;
	    v = eval_eb_node<SC>(f, anp1, an, ii, jj, ke);
	  }
	const long ma = (long) kk * f.Pp + (long) ii * f.N1 + jj;
	#pragma unroll
	for (int c = 0; c < 3; c++)
	  {
	    /* FieldVector::mv for the first node, pmv for the others (fieldvector.h): product, then sum             */
	    const double pe = w[q] * (double) v.e[c], pb = w[q] * (double) v.b[c], pa = w[q] * an[c * cs + ma];
	    et[c] = q ? et[c] + pe : pe;  bt[c] = q ? bt[c] + pb : pb;  at[c] = q ? at[c] + pa : pa;
	  }
      }
    #pragma unroll
    for (int c = 0; c < 3; c++) { o[c] = et[c]; o[3 + c] = bt[c]; o[6 + c] = at[c]; }
  }

  /* ------------------------------------------------------------------------------------------------
   * E, B (FdTd::fieldEvaluate) and A^n at a list of mesh nodes, for the host's field visualisation writers
   * (fdtd.cpp:1128-1540).  ijk[3 t ..] = (i, j, k) with k the plane index of the reference's slab numbering of THIS slab;
   * out[9 t ..] = en[3], bn[3] (the floats, widened) and an[3]; mine[t] = 1 when the slab holds the node as one of its
   * own planes.  Transverse boundary nodes (i, j = 0, N-1) have no E/B in the reference either (zero there).  On the
   * two end planes of the mesh the reference's fieldEvaluate reads beyond its arrays; here those planes take the values
   * of their interior neighbour, like every other consumer of E/B (fdtd.cpp:754-773).
   * ------------------------------------------------------------------------------------------------ */
  template <bool SC>
  __global__ void __launch_bounds__(128)
  field_nodes (const FieldDev f, const double* __restrict__ anp1, const double* __restrict__ an, const float4* __restrict__ ebn,
	       const int* __restrict__ ijk, long n, double* __restrict__ out, unsigned char* __restrict__ mine)
  {
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = ijk[3 * t], j = ijk[3 * t + 1], k = ijk[3 * t + 2] + f.kshift;
    double* o = out + 9 * t;
    #pragma unroll
    for (int q = 0; q < 9; q++) o[q] = 0.0;
    const bool inside = ( i >= 0 && i < f.N0 && j >= 0 && j < f.N1 && k >= 0 && k < f.np );
    /* own planes: kb .. np-2, plus the mesh's end planes on the first / last slab                                */
    const bool own = inside && ( ( k >= f.kb && k <= f.np - 2 ) || ( k < f.kb && f.rank == 0 ) || ( k == f.np - 1 && f.rank == f.size - 1 ) );
    mine[t] = own ? 1 : 0;
    if (!own) return;
    const long cs = (long) f.np * f.Pp, ma = (long) k * f.Pp + (long) i * f.N1 + j;
    o[6] = an[ma]; o[7] = an[cs + ma]; o[8] = an[2 * cs + ma];
    if (i < 1 || i > f.N0 - 2 || j < 1 || j > f.N1 - 2) return;
    int ke = k;
    if (ke < f.kb) ke = f.kb;                         /* rank 0 only (own): plane 0 <- plane 1                      */
    if (ke == f.np - 1) ke = f.np - 2;
    const EB v = eval_eb_node<SC>(f, anp1, an, i, j, ke);
    o[0] = v.e[0]; o[1] = v.e[1]; o[2] = v.e[2]; o[3] = v.b[0]; o[4] = v.b[1]; o[5] = v.b[2];
  }

  __global__ void power_finish (const PowerDev pw, const double* __restrict__ partial, int nblocks, double* __restrict__ row)
  {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= pw.N * pw.Nl) return;
    double acc = 0.0;
    for (int a = 0; a < nblocks; a++) acc += partial[(size_t) t * nblocks + a];
    row[t] = pw.pc * acc;
  }
}

#endif
