/* kernels_init.cuh -- the bunch of Solver::initialize() on the device: generation of the Halton ellipsoid
 * (Bunch::initializeEllipsoid, classes.cpp:104-298), Lorentz boost into the bunch frame (Solver::lorentzBoostBunch,
 * solver.cpp:286-302), ballistic back-projection to the start point (:340-346) and the hand-over to the slabs
 * (Solver::distributeParticles, :429-487).  The particles never exist on the host: 8.4 M of them are 3 s of one host core
 * and 740 MB over PCIe otherwise.
 *
 * Every formula is the reference's, operation by operation; the only difference to the host path is the libm: CUDA's log,
 * cos and sin differ from glibc's in the last ulp, so a device-generated bunch equals the reference's to 1e-15, not bit for
 * bit (tests/test_gpu_init.py).  Not generated here (the host path of mithra_b200/host/classes.cpp stays): shot noise
 * (a sequential minimum over the bunch numbers the buckets), the `random` generator (rand()), the other bunch types.
 */
#ifndef MITHRA_KERNELS_INIT_CUH_
#define MITHRA_KERNELS_INIT_CUH_

#include "device_types.cuh"

namespace mithra
{
  /* Halton radical inverse as the reference computes it, stdinclude.cpp:45-73: 1 - sum of digit / base^position       */
  __device__ inline double halton_dev (unsigned int dim, unsigned int j)
  {
    const int prime[20] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71 };
    const int p = prime[dim];
    int p0 = p, k = (int) j + 1;
    double x = 0.0;
    while (k > 0)
      {
	const int a = k % p;
	x += a / (double) p0;
	k  = k / p;
	p0 *= p;
      }
    return 1.0 - x;
  }

  /* one candidate of the ellipsoid: transverse and longitudinal offset r, momentum offset t; true when inside the
   * truncation (classes.cpp:219-246 for the body, :268-287 for the tapers of a uniform profile)                      */
  __device__ inline bool ellipsoid_candidate (const MithraBunchEllipsoid& b, unsigned int i, bool momentum, double r[3], double t[3])
  {
    const double PI = MITHRA_PI;
    const unsigned int m = i + b.index_offset;
    const unsigned int ng = ( b.lambda == 0.0 ) ? 1u : 4u;
    const unsigned int nbody = b.number_of_particles / ng;
    r[0] = b.sigma_position[0] * sqrt( - 2.0 * log( halton_dev(0, m) ) ) * cos( 2.0 * PI * halton_dev(1, m) );
    r[1] = b.sigma_position[1] * sqrt( - 2.0 * log( halton_dev(0, m) ) ) * sin( 2.0 * PI * halton_dev(1, m) );
    if (i < nbody)
      {
	if (b.distribution == 0) r[2] = ( 2.0 * halton_dev(2, m) - 1.0 ) * b.sigma_position[2];
	else                     r[2] = b.sigma_position[2] * sqrt( - 2.0 * log( halton_dev(2, m) ) ) * sin( 2.0 * PI * halton_dev(3, m) );
      }
    else
      {
	r[2]  = 2.0 * b.lambda * sqrt( - 2.0 * log( halton_dev(2, m) ) ) * sin( 2.0 * PI * halton_dev(3, m) );
	r[2] += ( r[2] < 0.0 ) ? ( - b.sigma_position[2] ) : ( b.sigma_position[2] );
      }
    if (momentum)
      {
	t[0] = b.sigma_gamma_beta[0] * sqrt( - 2.0 * log( halton_dev(4, m) ) ) * cos( 2.0 * PI * halton_dev(5, m) );
	t[1] = b.sigma_gamma_beta[1] * sqrt( - 2.0 * log( halton_dev(4, m) ) ) * sin( 2.0 * PI * halton_dev(5, m) );
	t[2] = b.sigma_gamma_beta[2] * sqrt( - 2.0 * log( halton_dev(6, m) ) ) * cos( 2.0 * PI * halton_dev(7, m) );
      }
    return fabs(r[0]) < b.tran_trun && fabs(r[1]) < b.tran_trun && fabs(r[2]) < b.long_trun;
  }

  /* pass 1: how many candidates of each block of 256 are accepted                                                    */
  __global__ void __launch_bounds__(256)
  ellipsoid_count (const __grid_constant__ MithraBunchEllipsoid b, unsigned int ncand, unsigned int* __restrict__ block_count)
  {
    const unsigned int i = blockIdx.x * 256u + threadIdx.x;
    double r[3], t[3];
    const bool ok = i < ncand && ellipsoid_candidate(b, i, false, r, t);
    const int n = __syncthreads_count(ok ? 1 : 0);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (unsigned int) n;
  }

  /* exclusive scan of up to a few 10^5 block counts by one block; total[0] = the sum                                */
  __global__ void __launch_bounds__(1024)
  scan_block_counts (unsigned int* __restrict__ v, unsigned int n, unsigned long long* __restrict__ total)
  {
    __shared__ unsigned int w[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (unsigned int base = 0; base < n; base += 1024u)
      {
	const unsigned int t = base + threadIdx.x;
	const unsigned int x = t < n ? v[t] : 0u;
	unsigned int incl = x;
	for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += y; }
	if ((threadIdx.x & 31) == 31) w[threadIdx.x >> 5] = incl;
	__syncthreads();
	unsigned int before = carry;
	for (int q = 0; q < (int) (threadIdx.x >> 5); q++) before += w[q];
	if (t < n) v[t] = before + incl - x;
	__syncthreads();
	if (threadIdx.x == 1023) carry = before + incl;
	__syncthreads();
      }
    if (threadIdx.x == 0) *total = carry;
  }

  /* pass 2: the accepted candidates write their group of ng particles, in the reference's order (candidate index, then
   * the quarter-wavelength copies ii = 0 .. 3 with the bunching modulation, classes.cpp:171-196)                       */
  __global__ void __launch_bounds__(256)
  ellipsoid_write (const __grid_constant__ MithraBunchEllipsoid b, unsigned int ncand, const unsigned int* __restrict__ block_offset,
		   double* __restrict__ aos)
  {
    __shared__ unsigned int wsum[8];
    const double PI = MITHRA_PI;
    const unsigned int i = blockIdx.x * 256u + threadIdx.x;
    double r[3], t[3];
    const bool ok = i < ncand && ellipsoid_candidate(b, i, true, r, t);
    const unsigned int bal = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    if (!ok) return;
    unsigned int rank = block_offset[blockIdx.x] + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u));
    for (int q = 0; q < (int) (threadIdx.x >> 5); q++) rank += wsum[q];
    const unsigned int ng = ( b.lambda == 0.0 ) ? 1u : 4u;
    const double q  = b.cloud_charge / b.number_of_particles;
    const double x  = b.position[0] + r[0], y = b.position[1] + r[1], z = b.position[2] + r[2];
    const double gx = b.initial_gamma * b.beta_vector[0] + t[0], gy = b.initial_gamma * b.beta_vector[1] + t[1], gz = b.initial_gamma * b.beta_vector[2] + t[2];
    for (unsigned int ii = 0; ii < ng; ii++)
      {
	double zz = z;
	if (b.lambda != 0.0)
	  {
	    zz  = z - b.lambda / 4 * ii;
	    zz -= b.lambda / PI * b.bunching_factor * sin( 2.0 * PI / b.lambda * zz + b.bunching_phase * PI / 180.0 );
	  }
	double* o = aos + ( (size_t) rank * ng + ii ) * 11;
	o[0] = q; o[1] = x; o[2] = y; o[3] = zz; o[4] = 0.0; o[5] = 0.0; o[6] = 0.0; o[7] = gx; o[8] = gy; o[9] = gz; o[10] = 0.0;
      }
  }

  /* Solver::lorentzBoostBunch, solver.cpp:294-302: z and gb_z into the frame moving with gamma; the largest z per block  */
  __global__ void __launch_bounds__(256)
  bunch_boost (double* __restrict__ aos, size_t n, double gamma, double beta, double* __restrict__ block_zmax)
  {
    __shared__ double red[256];
    const size_t t = (size_t) blockIdx.x * 256 + threadIdx.x;
    double zmax = -1.0e100;
    if (t < n)
      {
	double* o = aos + t * 11;
	const double g  = sqrt( 1.0 + ( o[7] * o[7] + o[8] * o[8] + o[9] * o[9] ) );
	const double bz = o[9] / g;
	o[3] *= gamma;
	o[9]  = gamma * g * ( bz - beta );
	zmax  = o[3];
      }
    red[threadIdx.x] = zmax;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) { if (threadIdx.x < w) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + w]); __syncthreads(); }
    if (threadIdx.x == 0) block_zmax[blockIdx.x] = red[0];
  }

  /* solver.cpp:340-346: the bunch properties hold at the start point; move every particle back along a straight line     */
  __global__ void __launch_bounds__(256)
  bunch_backproject (double* __restrict__ aos, size_t n, double zu, double beta)
  {
    const size_t t = (size_t) blockIdx.x * 256 + threadIdx.x;
    if (t >= n) return;
    double* o = aos + t * 11;
    const double g = sqrt( 1.0 + ( o[7] * o[7] + o[8] * o[8] + o[9] * o[9] ) );
    const double z = o[3];
    o[1] += o[7] / g * ( z - zu ) * beta;
    o[2] += o[8] / g * ( z - zu ) * beta;
    o[3] += o[9] / g * ( z - zu ) * beta;
  }

  /* Solver::distributeParticles for one slab: which particles it owns (wrapped z in [zp0, zp1), solver.cpp:1440-1441 /
   * 2292-2298), counted per block of 256 ...                                                                          */
  __device__ __forceinline__ bool slab_owns (double z, double zmin, double Lz, double zp0, double zp1)
  {
    double zr = fmod( z - zmin, Lz ); zr += ( zr < 0.0 ) ? Lz : 0.0; zr += zmin;
    return zr >= zp0 && zr < zp1;
  }

  __global__ void __launch_bounds__(256)
  owned_count (const double* __restrict__ aos, size_t n, double zmin, double Lz, double zp0, double zp1, unsigned int* __restrict__ block_count)
  {
    const size_t t = (size_t) blockIdx.x * 256 + threadIdx.x;
    const bool ok = t < n && slab_owns(aos[t * 11 + 3], zmin, Lz, zp0, zp1);
    const int c = __syncthreads_count(ok ? 1 : 0);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (unsigned int) c;
  }

  /* ... and copied, in list order, into the staging array the slab converts to its struct-of-arrays                     */
  __global__ void __launch_bounds__(256)
  owned_copy (const double* __restrict__ aos, size_t n, double zmin, double Lz, double zp0, double zp1,
	      const unsigned int* __restrict__ block_offset, double* __restrict__ out)
  {
    __shared__ unsigned int wsum[8];
    const size_t t = (size_t) blockIdx.x * 256 + threadIdx.x;
    const bool ok = t < n && slab_owns(aos[t * 11 + 3], zmin, Lz, zp0, zp1);
    const unsigned int bal = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    if (!ok) return;
    unsigned int rank = block_offset[blockIdx.x] + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u));
    for (int q = 0; q < (int) (threadIdx.x >> 5); q++) rank += wsum[q];
    const double* s = aos + t * 11; double* d = out + (size_t) rank * 11;
    #pragma unroll
    for (int c = 0; c < 11; c++) d[c] = s[c];
  }
}

#endif
