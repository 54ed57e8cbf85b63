/* kernels_sort.cuh -- counting sort of the macro-particles by mesh cell.
 *
 * The reference keeps the bunch in a std::list in insertion order (solver.h:271); the order has no meaning for the
 * physics, only for the order of the records in the output files.  On the GPU the order decides everything about
 * the memory system: the 8-node E/B gather of the push and the 24 (+8) atomic adds of the deposit hit the same
 * few cache lines for neighbouring lanes only when neighbouring lanes hold particles of neighbouring cells.
 * The bunch is therefore re-ordered every few field steps by the cell (k, i, j) that holds the particle
 * (k slowest: the order of the potentials in memory), and every particle carries the index `id` it had when it
 * was uploaded, so that downloads, screen records and the cell-assignment diagnostics are returned in the
 * reference's order.  In the boosted frame the bunch is almost at rest on the mesh (|gb| ~ 1e-2 cells per step),
 * so the order decays slowly and one sort serves many steps.
 *
 *   sort_zero   : clear the histogram over the cells of the particle box
 *   sort_count  : key[t] = cell relative to the particle box, rank[t] = atomicAdd(hist[key], 1)
 *   scan_*      : exclusive prefix sum of the histogram (three passes, 2048 bins per CTA)
 *   sort_permute: particle t moves to hist[key[t]] + rank[t] in the second copy of the struct-of-arrays
 *
 * The order inside a cell is the arrival order of the atomics, i.e. not reproducible from run to run; nothing
 * observable depends on it except the (already unordered) summation order of the deposit.
 */
#ifndef MITHRA_KERNELS_SORT_CUH_
#define MITHRA_KERNELS_SORT_CUH_

#include "device_types.cuh"

namespace mithra
{
  #define MITHRA_SCAN_CHUNK 2048              /* bins per CTA of the scan kernels (256 threads x 8)            */

  /* extent of the sort keys: the particle box (cells), clamped to `cap` bins                              */
  struct SortBox { int lo[3]; int n[3]; long vol; };

  __device__ __forceinline__ SortBox sort_box (const Box* __restrict__ pbox, long cap)
  {
    const Box b = *pbox;
    SortBox s;
    #pragma unroll
    for (int a = 0; a < 3; a++) { s.lo[a] = b.lo[a]; s.n[a] = b.hi[a] - b.lo[a] + 1; }
    if (s.n[0] <= 0 || s.n[1] <= 0 || s.n[2] <= 0) { s.n[0] = s.n[1] = s.n[2] = 0; s.vol = 0; return s; }
    s.vol = (long) s.n[0] * s.n[1] * s.n[2];
    if (s.vol > cap) s.vol = cap;               /* keys beyond the capacity share the last bin                   */
    return s;
  }

  __global__ void __launch_bounds__(256)
  sort_zero (const Box* __restrict__ pbox, long cap, unsigned int* __restrict__ hist)
  {
    const SortBox s = sort_box(pbox, cap);
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < s.vol; t += (long) gridDim.x * blockDim.x) hist[t] = 0u;
  }

  __global__ void __launch_bounds__(256)
  sort_count (const __grid_constant__ BunchDev b, ParticlesDev P, long n, const Box* __restrict__ pbox, long cap,
	      unsigned int* __restrict__ hist, unsigned int* __restrict__ key, unsigned int* __restrict__ rank)
  {
    const SortBox s = sort_box(pbox, cap);
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (s.vol == 0) { key[t] = 0u; rank[t] = (unsigned int) t; return; }
    const double x = P.r[0][t], y = P.r[1][t], z = P.r[2][t];
    /* the cell of the particle, clamped into the box (particles outside of the mesh sort to its rim)         */
    int i = (int) floor( ( x - b.xmin ) / b.dx ) - s.lo[0];
    int j = (int) floor( ( y - b.ymin ) / b.dy ) - s.lo[1];
    int k = (int) floor( ( z - b.zmin ) / b.dz ) - b.k0 - s.lo[2];
    i = min(max(i, 0), s.n[0] - 1); j = min(max(j, 0), s.n[1] - 1); k = min(max(k, 0), s.n[2] - 1);
    long c = ( (long) k * s.n[0] + i ) * s.n[1] + j;
    if (c >= s.vol) c = s.vol - 1;
    key[t]  = (unsigned int) c;
    rank[t] = atomicAdd(&hist[c], 1u);
  }

  /* pass 1: sum of every chunk                                                                             */
  __global__ void __launch_bounds__(256)
  scan_chunk_sums (const Box* __restrict__ pbox, long cap, const unsigned int* __restrict__ hist, unsigned int* __restrict__ sums)
  {
    const SortBox s = sort_box(pbox, cap);
    const long base = (long) blockIdx.x * MITHRA_SCAN_CHUNK;
    if (base >= s.vol) return;
    unsigned int v = 0u;
    #pragma unroll
    for (int q = 0; q < MITHRA_SCAN_CHUNK / 256; q++)
      {
	const long m = base + q * 256 + threadIdx.x;
	if (m < s.vol) v += hist[m];
      }
    __shared__ unsigned int w[8];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned int a = 0u; for (int q = 0; q < 8; q++) a += w[q]; sums[blockIdx.x] = a; }
  }

  /* pass 2: exclusive scan of the chunk sums by one CTA                                                    */
  __global__ void __launch_bounds__(1024)
  scan_sums (const Box* __restrict__ pbox, long cap, unsigned int* __restrict__ sums)
  {
    const SortBox s = sort_box(pbox, cap);
    const long nchunks = (s.vol + MITHRA_SCAN_CHUNK - 1) / MITHRA_SCAN_CHUNK;
    __shared__ unsigned int w[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (long base = 0; base < nchunks; base += 1024)
      {
	const long m = base + threadIdx.x;
	const unsigned int v = (m < nchunks) ? sums[m] : 0u;
	unsigned int incl = v;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += u; }
	if ((threadIdx.x & 31) == 31) w[threadIdx.x >> 5] = incl;
	__syncthreads();
	if (threadIdx.x < 32)
	  {
	    unsigned int a = w[threadIdx.x];
	    #pragma unroll
	    for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, a, o); if (threadIdx.x >= o) a += u; }
	    w[threadIdx.x] = a;
	  }
	__syncthreads();
	const unsigned int before = carry + ( (threadIdx.x >> 5) ? w[(threadIdx.x >> 5) - 1] : 0u );
	if (m < nchunks) sums[m] = before + incl - v;
	__syncthreads();
	if (threadIdx.x == 1023) carry = before + incl;
	__syncthreads();
      }
  }

  /* pass 3: exclusive scan inside every chunk + its offset, in place                                        */
  __global__ void __launch_bounds__(256)
  scan_chunks (const Box* __restrict__ pbox, long cap, unsigned int* __restrict__ hist, const unsigned int* __restrict__ sums)
  {
    const SortBox s = sort_box(pbox, cap);
    const long base = (long) blockIdx.x * MITHRA_SCAN_CHUNK;
    if (base >= s.vol) return;
    constexpr int PER = MITHRA_SCAN_CHUNK / 256;
    /* thread t owns the PER consecutive bins base + t*PER ...                                               */
    unsigned int v[PER]; unsigned int tot = 0u;
    #pragma unroll
    for (int q = 0; q < PER; q++)
      {
	const long m = base + (long) threadIdx.x * PER + q;
	v[q] = (m < s.vol) ? hist[m] : 0u;
	tot += v[q];
      }
    unsigned int incl = tot;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += u; }
    __shared__ unsigned int w[8];
    if ((threadIdx.x & 31) == 31) w[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned int before = sums[blockIdx.x];
    for (int q = 0; q < (int) (threadIdx.x >> 5); q++) before += w[q];
    unsigned int run = before + incl - tot;
    #pragma unroll
    for (int q = 0; q < PER; q++)
      {
	const long m = base + (long) threadIdx.x * PER + q;
	if (m < s.vol) hist[m] = run;
	run += v[q];
      }
  }

  __global__ void __launch_bounds__(256)
  sort_permute (ParticlesDev P, ParticlesDev Q, long n, const Box* __restrict__ pbox, long cap, const unsigned int* __restrict__ hist,
		const unsigned int* __restrict__ key, const unsigned int* __restrict__ rank)
  {
    const SortBox s = sort_box(pbox, cap);
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long d = (s.vol == 0) ? t : (long) hist[key[t]] + rank[t];      /* empty box: identity                   */
    Q.q[d] = P.q[t]; Q.e[d] = P.e[t]; Q.id[d] = P.id[t];
    #pragma unroll
    for (int a = 0; a < 3; a++) { Q.r[a][d] = P.r[a][t]; Q.rm[a][d] = P.rm[a][t]; Q.gb[a][d] = P.gb[a][t]; }
  }
}

#endif
