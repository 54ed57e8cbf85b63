/* kernels_field.cuh -- field-side sm_100a kernels: the 3-level leap-frog update of A (and phi), the
 * absorbing boundaries, the current-box clear and the E/B evaluation.
 *
 * Arithmetic follows the reference's association order operation by operation and the translation unit is
 * compiled with -fmad=false, so for identical inputs the potentials are BIT-IDENTICAL to the reference CPU
 * build (x86-64 g++ -O3 has no FMA contraction either).  Reference formulas:
 *   interior  : AdvanceField::advance{Magnetic,Scalar}Potential{NSFD,FD}   database.cpp:44-134
 *   faces     : AdvanceField::advanceBoundary{F,S}                          database.cpp:137-176
 *   edges     : AdvanceField::advanceEdge{F,S}                              database.cpp:179-229
 *   corners   : AdvanceField::advanceCorner{F,S}                            database.cpp:232-288
 *   call sites: FdTd::fieldUpdate                                           fdtd.cpp:231-725
 *   E/B       : FdTd::fieldEvaluate / FdTdSC::fieldEvaluate                 fdtd.cpp:818-845, fdtdSC.cpp:1110-1141
 */
#ifndef MITHRA_KERNELS_FIELD_CUH_
#define MITHRA_KERNELS_FIELD_CUH_

#include "device_types.cuh"
#include "beams.cuh"

namespace mithra
{
  /* ------------------------------------------------------------------------------------------------
   * Pencil mask of the source term.  J is non-zero only where the deposit of the last step put it: inside the box
   * `jbox`, and there only on the node pencils (node column x 8 planes, see spread_eb_mask) the particles could reach
   * -- the mask the E/B evaluation of the same step was given, because both ends of a particle's path of this step lie
   * within the padding around its cell at the start of the step.  On a slab with neighbours the planes kb, np-3 and
   * np-2 also receive the neighbours' deposits (exchange.cuh exchange_current) and are always taken.  jmask = 0: box only.
   * ------------------------------------------------------------------------------------------------ */
  /* bit (k - ks) of the result: does the thread of node column (i, j), p = i N1 + j, read the source on plane k of its
   * march ks .. ke-1 (at most 64 planes)?  Box, pencil mask and merge planes folded into one word before the march.    */
  __device__ __forceinline__ unsigned long long source_planes (const FieldDev& f, const Box& bx, const unsigned char* __restrict__ mask,
							       int i, int j, long p, int ks, int ke)
  {
    if (!(i >= bx.lo[0] && i <= bx.hi[0] && j >= bx.lo[1] && j <= bx.hi[1])) return 0ull;
    unsigned long long en = 0ull;
    int chunk = -1; bool on = true;
    for (int k = max(ks, bx.lo[2]); k < ke && k <= bx.hi[2]; k++)
      {
	if (mask && (k >> MITHRA_EB_CHUNK_LOG2) != chunk) { chunk = k >> MITHRA_EB_CHUNK_LOG2; on = mask[(long) chunk * f.P + p] != 0; }
	if (on || ( f.size > 1 && ( k == f.kb || k >= f.np - 3 ) )) en |= 1ull << (k - ks);
      }
    return en;
  }

  /* ------------------------------------------------------------------------------------------------
   * Interior stencil, z-marching register pipeline.
   *
   * grid  = ( ceil(P / BX), ceil((np-2) / KC), ncomp ),  block = BX threads.
   * A thread owns one in-plane position p = i*N1 + j of one component and marches KC planes in +z, keeping
   * the 5-point cross of the planes k-1, k, k+1 of A^n in registers: each A^n value is loaded once per
   * thread, in-plane neighbours come out of L1 (they are the centre loads of the neighbouring lanes /
   * warps of the same CTA).  A^{n-1} and J are streamed (read once), A^{n+1} is streamed out.
   * J is only read inside the deposit bounding box `jbox`; outside of it the source term is exactly zero.
   * ------------------------------------------------------------------------------------------------ */
  template <bool NSFD, int BX, int KC>
  __global__ void __launch_bounds__(BX)
  stencil_interior (const FieldDev f, double* __restrict__ anp1, const double* __restrict__ an,
		    const double* __restrict__ anm1, const double* __restrict__ jn, const Box* __restrict__ jbox,
		    const unsigned char* __restrict__ jmask)
  {
    const int p = blockIdx.x * BX + threadIdx.x;
    if (p >= f.P) return;
    const int i = p / f.N1, j = p - i * f.N1;
    if (i < 1 || i > f.N0 - 2 || j < 1 || j > f.N1 - 2) return;

    const int c  = blockIdx.z;
    const int ks = f.kb + blockIdx.y * KC;
    const int ke = min(ks + KC, f.np - 1);               /* exclusive                                    */
    if (ks >= ke) return;

    const long   cb  = (long) c * f.np * f.Pp;
    const double* A  = an   + cb + p;
    const double* Am = anm1 + cb + p;
    const double* Jn = jn   + cb + p;
    double*       Ap = anp1 + cb + p;
    const int    N1  = f.N1;
    const long   Pp  = f.Pp;

    const double a0 = f.a[0], a1 = f.a[1], a2 = f.a[2], a3 = f.a[3];
    const double as = (c < 3) ? f.a[4] : f.a[5];
    const double alpha = f.alpha, beta = f.beta;

    /* is this column inside the deposit box at all?                                                     */
    const Box bx = *jbox;
    const unsigned long long srcon = source_planes(f, bx, jmask, i, j, p, ks, ke);

    /* planes k-1 (suffix m), k (suffix 0), k+1 (suffix p) of the 5-point cross                          */
    double cm, c0, cp, xpm, xp0, xpp, xmm, xm0, xmp, ypm, yp0, ypp, ymm, ym0, ymp;
    {
      const double* q = A + (long) (ks - 1) * Pp;
      cm = q[0];
      if (NSFD) { xpm = q[N1]; xmm = q[-N1]; ypm = q[1]; ymm = q[-1]; }
      q += Pp;
      c0 = q[0]; xp0 = q[N1]; xm0 = q[-N1]; yp0 = q[1]; ym0 = q[-1];
    }

    for (int k = ks; k < ke; k++)
      {
	const double* q = A + (long) (k + 1) * Pp;
	cp = q[0];
	if (NSFD) { xpp = q[N1]; xmp = q[-N1]; ypp = q[1]; ymp = q[-1]; }
	const double vm1 = Am[(long) k * Pp];
	double src = 0.0;
	if ((srcon >> (k - ks)) & 1ull) src = Jn[(long) k * Pp];

	double r;
	if (NSFD)
	  r = a0 * c0 - vm1 + alpha * (
	      a1 * ( xp0 + xm0 + beta * ( xpp + xpm + xmp + xmm ) ) +
	      a2 * ( yp0 + ym0 + beta * ( ypp + ypm + ymp + ymm ) ) ) +
	      a3 * ( cp + cm ) +
	      as * src;
	else
	  r = a0 * c0 - vm1 +
	      a1 * ( xp0 + xm0 ) +
	      a2 * ( yp0 + ym0 ) +
	      a3 * ( cp + cm ) +
	      as * src;
	Ap[(long) k * Pp] = r;

	cm = c0; c0 = cp;
	if (NSFD) { xpm = xp0; xmm = xm0; ypm = yp0; ymm = ym0; xp0 = xpp; xm0 = xmp; yp0 = ypp; ym0 = ymp; }
	else      { const double* q0 = A + (long) (k + 1) * Pp; xp0 = q0[N1]; xm0 = q0[-N1]; yp0 = q0[1]; ym0 = q0[-1]; }
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * Interior stencil, bulk-async plane pipeline (the production path; stencil_interior above stays as the
   * plain-load fallback for meshes whose rows do not fit the staging buffers).
   *
   * A CTA owns T consecutive in-plane positions p = i*N1 + j of one component -- a CONTIGUOUS piece of every
   * z plane because rows are contiguous -- and marches KC planes in +z.  Plane k+1 of A^n (the T positions plus
   * one row of halo on either side) and plane k of A^{n-1} are fetched by ONE elected thread with
   * cp.async.bulk (1-D bulk copy through the TMA unit, completion on an mbarrier) into a ring of NB stages, NB-1
   * planes ahead of the plane being computed, so HBM latency is covered by the ring and not by occupancy.  A thread
   * keeps the 5-point cross of planes k-1, k in registers, takes the five values of plane k+1 from shared memory
   * and streams A^{n+1} out with a coalesced store; a stage goes back to the producer one step after it was taken, when
   * the A^{n-1} value that came with it has been read.  Arithmetic and association order are those of
   * stencil_interior (bit-identical results).
   * Variants (engine.cu launch_stencil_stream_as): <T = 480, FACES = false> inner nodes only, rim_update does the rim
   * (seeded jobs, wide meshes); <448, FACES> every interior node and the y faces (narrow meshes without a seed);
   * <384, FACES, SEED> the same with the y-shell seed terms (opt-in).  T is chosen so that two CTAs fit an SM without a
   * spill in the consumer loop: 64 / 64 / 72 registers.
   *
   * Alignment: bulk copies need 16-byte aligned addresses and sizes; Pp is a multiple of 16 doubles, T is even and
   * the halo is rounded up to an even number of doubles (N1e), the extra element is never read.
   * ------------------------------------------------------------------------------------------------ */
  __device__ __forceinline__ unsigned smem_u32 (const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
  __device__ __forceinline__ void mbar_init (unsigned long long* b, int count)
  { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count)); }
  __device__ __forceinline__ void mbar_expect_tx (unsigned long long* b, unsigned bytes)
  { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
  __device__ __forceinline__ void mbar_wait (unsigned long long* b, unsigned parity)
  {
    unsigned done;
    while (true)
      {
	asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
		     : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
	if (done) break;
	__nanosleep(32);                                   /* a waiting warp must not eat the issue slots of the working ones */
      }
  }
  __device__ __forceinline__ void bulk_g2s (void* dst, const void* src, unsigned bytes, unsigned long long* b)
  {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
  }

  __device__ __forceinline__ void mbar_arrive (unsigned long long* b)
  { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }

  /* shared memory of one CTA: a header (mbarriers, FACES: the counters of the face list), then NB stages, each
   * (T + 2 H) doubles of A^n and T + 2 HM doubles of A^{n-1}; FACES: then the face list.
   * H: halo of the A^n stage, N1e (FACES: N1e + 2); HM: 0 (FACES: 2)                                                       */
  static inline size_t stencil_stream_smem (int T, int N1, int NB, bool faces = false)
  {
    const int N1e = (N1 + 1) & ~1;
    if (!faces) return 256 + (size_t) NB * ( (size_t) T + 2 * N1e + T ) * sizeof(double);
    return 1024 + (size_t) NB * ( 2 * (size_t) T + 2 * (N1e + 2) + 4 ) * sizeof(double) + 32 * sizeof(int);
  }

  /* the nodes of a tile that sit next to a y face (the inward neighbours n of its nodes): their largest number over the
   * tiles of T consecutive in-plane positions; stencil_stream<.., FACES> gives each of them a lane of its face warp, so
   * it takes meshes with at most 32.  seeded: and the nodes one further in, the other half of the y shell of a TF/SF seed  */
  static inline int stencil_stream_face_nodes (int N0, int N1, int T, bool seeded = false)
  {
    const long P = (long) N0 * N1;
    int worst = 0;
    for (long p0 = 0; p0 < P; p0 += T)
      {
	int n = 0;
	for (long p = p0; p < p0 + T && p < P; p++)
	  {
	    const int i = (int) (p / N1), j = (int) (p - (long) i * N1);
	    if (i >= 1 && i <= N0 - 2 && j >= 1 && j <= N1 - 2 && (j == 1 || j == N1 - 2 || (seeded && (j == 2 || j == N1 - 3)))) n++;
	  }
	if (n > worst) worst = n;
      }
    return worst;
  }

  /* the 5-point cross of one plane around the thread's node                                                    */
  struct Cross { double c, xp, xm, yp, ym; };

  template <bool NSFD>
  __device__ __forceinline__ double stencil_value (const Cross& m, const Cross& z, const Cross& p, double vm1, double src,
						   double a0, double a1, double a2, double a3, double as, double alpha, double beta)
  {
    if (NSFD)
      return a0 * z.c - vm1 + alpha * (
	     a1 * ( z.xp + z.xm + beta * ( p.xp + m.xp + p.xm + m.xm ) ) +
	     a2 * ( z.yp + z.ym + beta * ( p.yp + m.yp + p.ym + m.ym ) ) ) +
	     a3 * ( p.c + m.c ) +
	     as * src;
    return a0 * z.c - vm1 +
	   a1 * ( z.xp + z.xm ) +
	   a2 * ( z.yp + z.ym ) +
	   a3 * ( p.c + m.c ) +
	   as * src;
  }

  /* AdvanceField::advanceBoundary{F,S} (database.cpp:137-176) for the face node s with its inward neighbour n:
   *   apn = A+_n, ams = A-_s, amn = A-_n, as_ = A_s, an_ = A_n, then the four tangential neighbours along t1 in the
   *   order (n+, n-, s+, s-) and likewise along z -- the association order of face_update below                     */
  __device__ __forceinline__ double face_value (const double* B, double ams, double apn, double amn, double as_, double an_,
						double n1p, double n1m, double s1p, double s1m,
						double n2p, double n2m, double s2p, double s2m)
  {
    return B[0] * ( ams + apn ) +
	   B[1] * amn +
	   B[2] * ( as_ + an_ ) +
	   B[3] * ( n1p + n1m + s1p + s1m ) +
	   B[4] * ( n2p + n2m + s2p + s2m );
  }

  /* T consumer threads (one in-plane position each) + one producer warp (+ FACES: one face warp)
   *
   * FACES (N0, N1, np >= 8; at most 32 nodes next to a y face per tile; by default meshes without a TF/SF seed whose
   * perimeter is at least 2 % of the plane -- engine.cu field_update_potentials): the y absorbing
   * faces (fdtd.cpp:449-520) are done here as well, the x faces -- whole rows -- by one coalesced pass of boundary_faces
   * afterwards, and rim_update is not launched at all.  A y face node s needs A+ of its inward neighbour n and, of A^n
   * and A^{n-1}, only values the ring holds anyway: s, n and their in-plane neighbours in plane k, s and n in the planes
   * k-1 and k+1 -- nearly all of them in the cross of n.  Doing the face in the consumer thread of n puts one or two
   * face lanes into most warps of a row-major tile, and every such warp then runs the face code (round 1: +1 ms); so the
   * nodes next to a y face get a WARP OF THEIR OWN: lane l of the face warp owns the l-th such node of the tile as a
   * consumer would (its consumer thread idles) -- interior value with the source term, store -- and then the face node
   * behind it from the same crosses plus three more values of the stage of plane k.  The face warp is one more consumer
   * to the ring (`empty` counts it in); nothing is handed over between warps.  rim_update moved 6.6 GB per FEL-LCLS step
   * for 2.8 GB of rim nodes; here the faces add nothing to the traffic of the sweep.  Stages carry two more doubles of
   * halo (the in-plane neighbours of a face node one row off the tile; the face node across the end of the tile in
   * A^{n-1}).  Same operations in the same order as face_update: bit-identical.
   * Shared memory is addressed as 32-bit shared-space addresses (one base per stage, constant offsets): the consumer loop
   * has to fit its registers without a spill -- a reload from local memory in it costs more than a plane.             */
  __device__ __forceinline__ double lds_f64 (unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
  __device__ __forceinline__ void mbar_wait_a (unsigned b, unsigned parity)
  {
    unsigned done;
    while (true)
      {
	asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
		     : "=r"(done) : "r"(b), "r"(parity) : "memory");
	if (done) break;
	__nanosleep(32);
      }
  }
  __device__ __forceinline__ void mbar_arrive_a (unsigned b)
  { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(b) : "memory"); }

  template <bool NSFD, int T, int NB, bool FACES, bool SEED = false>
  __global__ void __launch_bounds__(T + (FACES ? 64 : 32), 2)
  stencil_stream (const FieldDev f, double* __restrict__ anp1, const double* __restrict__ an,
		  const double* __restrict__ anm1, double* __restrict__ jn, const Box* __restrict__ jbox, int KC, int skiprim,
		  const unsigned char* __restrict__ jmask, const RimDev rz)
  {
    static_assert((NB & (NB - 1)) == 0 && NB >= 4 && NB <= 16, "stages: a power of two");
    static_assert(T % 32 == 0 && T <= 1024, "whole consumer warps, one prefix entry per lane of the face warp");
    constexpr int NW = T / 32;                            /* consumer warps                                            */
    constexpr int HDR = FACES ? 1024 : 256;
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned long long* full  = reinterpret_cast<unsigned long long*>(smraw);
    unsigned long long* empty = full + NB;
    int* wc = reinterpret_cast<int*>(smraw + 512);        /* FACES: [NW] counts -> prefixes, then the number of faces   */
    const int  N0 = f.N0, N1 = f.N1, N1e = (N1 + 1) & ~1;
    const int  H  = FACES ? N1e + 2 : N1e;                /* doubles of halo either side of the A^n part of a stage    */
    const int  W  = T + 2 * H;                            /* doubles of the A^n part                        */
    const int  HM = FACES ? 2 : 0, WM = T + 2 * HM;       /* the same for A^{n-1}                                      */
    const int  S  = W + WM;                               /* doubles of a stage                                        */
    double* st = reinterpret_cast<double*>(smraw + HDR);
    int*    faces = reinterpret_cast<int*>(st + (size_t) NB * S);         /* FACES: node | high side << 12              */
    const unsigned bars = smem_u32(smraw), st0 = smem_u32(st);
    const unsigned SB = (unsigned) S * 8u;                /* bytes of a stage                                          */

    const int  tid = threadIdx.x, c = blockIdx.z;
    const int  p0  = blockIdx.x * T;
    const int  ks  = f.kb + blockIdx.y * KC, ke = min(ks + KC, f.np - 1);      /* planes ks .. ke-1            */
    if (ks >= ke) return;
    const long Pp = f.Pp, cb = (long) c * f.np * Pp;
    const int  nq = ke - ks + 2;                          /* planes ks-1 .. ke travel through the ring       */

    if (tid == 0)
      {
	for (int s = 0; s < NB; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW + (FACES ? 1 : 0)); }
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }

    /* the node of a thread: a consumer's is its position in the tile                                            */
    int  node = tid;
    int  p = p0 + node, i = p / N1, j = p - i * N1;
    /* skiprim: the two outermost interior node layers in x and y belong to rim_update                          */
    const int  rim = (skiprim && !FACES) ? 2 : 0;
    bool interior = (tid < T && p < f.P && i >= 1 + rim && i <= N0 - 2 - rim && j >= 1 + rim && j <= N1 - 2 - rim);
    int  fs = 0;                                          /* FACES: offset in bytes of the y face node behind the node */

    if (FACES)
      {
	/* the nodes next to a y face, in node order, one per lane of the face warp: ballots, per-warp counts, a scan    */
	const unsigned below = (1u << (tid & 31)) - 1u;
	const int warp = tid >> 5;
	const bool on = interior && (j == 1 || j == N1 - 2 || (SEED && (j == 2 || j == N1 - 3)));
	unsigned bal = 0u;
	if (tid < T) { bal = __ballot_sync(0xffffffffu, on); if ((tid & 31) == 0) wc[warp] = __popc(bal); }
	__syncthreads();
	if (tid >= T + 32)
	  {
	    const int lane = tid & 31;
	    int v = lane < NW ? wc[lane] : 0, x = v;
	    #pragma unroll
	    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
	    if (lane < NW) wc[lane] = x - v;
	    if (lane == 31) wc[NW] = x;
	  }
	__syncthreads();
	if (on) { faces[wc[warp] + __popc(bal & below)] = tid | ((j == 1 ? 0 : j == N1 - 2 ? 1 : 2) << 12); interior = false; }    /* the face warp's */
	__syncthreads();
	if (tid >= T + 32)
	  {
	    const int lane = tid & 31;
	    if (lane < wc[NW])
	      {
		const int w = faces[lane];
		node = w & 0xfff; fs = (w >> 12) == 0 ? -8 : (w >> 12) == 1 ? 8 : 0;
		p = p0 + node; i = p / N1; j = p - i * N1; interior = true;
	      }
	    else node = 0;
	  }
      }
    else __syncthreads();

    /* FACES with a TF/SF seed: the x / y shell corrections (fdtd.cpp:312-350) of the nodes the face warp owns -- the four
     * y-shell nodes of every row; the x-shell nodes between them are whole rows and go to seed_xshell_rows afterwards.
     * Offsets into RimDev.seedu of plane 0 (line * L + index), -1: none; sign of the correction as in rim_update          */
    int sxo = -1, syo = -1;
    bool sxminus = false, syminus = false;
    if (FACES && SEED && c < 3 && tid >= T + 32 && interior)
      {
	if (j >= 2 && j <= N1 - 3) { const int l = (i == 1) ? 1 : (i == 2) ? 0 : (i == N0 - 2) ? 2 : (i == N0 - 3) ? 3 : -1; if (l >= 0) sxo = l * rz.L + j; }
	if (i >= 2 && i <= N0 - 3) { const int l = (j == 1) ? 5 : (j == 2) ? 4 : (j == N1 - 2) ? 6 : (j == N1 - 3) ? 7 : -1; if (l >= 0) syo = l * rz.L + i; }
	sxminus = (i == 1 || i == N0 - 2); syminus = (j == 1 || j == N1 - 2);
      }
    const bool seeded = SEED && (sxo >= 0 || syo >= 0);

    if (tid >= T && tid < T + 32)
      {
	/* ---- producer warp: one lane feeds the ring -------------------------------------------------------- */
	if (tid != T) return;
	/* source range of the A^n part of a stage, clipped to the plane; both ends are even                          */
	const long lo = max(0L, (long) p0 - H), hi = min(Pp, (long) p0 + T + H);
	const int  dstoff = (int) (lo - (p0 - H));
	const unsigned bytesA = (unsigned) ((hi - lo) * sizeof(double));
	const long mlo = max(0L, (long) p0 - HM), mhi = min(Pp, (long) p0 + T + HM);
	const int  dstoffM = W + (int) (mlo - (p0 - HM));
	const unsigned bytesM = (unsigned) ((mhi - mlo) * sizeof(double));
	const double* srcA = an   + cb + (long) (ks - 1) * Pp + lo;
	const double* srcM = anm1 + cb + (long) (ks - 1) * Pp + mlo;
	for (int q = 0; q < nq; q++, srcA += Pp, srcM += Pp)
	  {
	    const int s = q & (NB - 1);
	    if (q >= NB) mbar_wait(&empty[s], (unsigned) (((q / NB) - 1) & 1));
	    const bool needM = (q >= 1 && q < nq - 1);     /* A^{n-1} rides along for the planes that are updated */
	    mbar_expect_tx(&full[s], bytesA + (needM ? bytesM : 0u));
	    bulk_g2s(st + (size_t) s * S + dstoff, srcA, bytesA, &full[s]);
	    if (needM) bulk_g2s(st + (size_t) s * S + dstoffM, srcM, bytesM, &full[s]);
	  }
	return;
      }

    /* ---- consumers (and the face warp, whose lanes are consumers of the nodes next to a y face) -------------- */
    const bool lane0 = (tid & 31) == 0;
    const double as = (c < 3) ? f.a[4] : f.a[5];
    unsigned long long srcon;                             /* bit 0 = the next plane                          */
    { const Box bx = *jbox; srcon = interior ? source_planes(f, bx, jmask, i, j, p, ks, ke) : 0ull; }
    /* the node in plane k, in A^{n+1} and J: a base the whole CTA shares and a 32-bit offset (at most 64 planes)      */
    double* const apc = anp1 + cb + (long) ks * Pp + p0;
    double* const jnc = jn   + cb + (long) ks * Pp + p0;
    const unsigned PpU = (unsigned) Pp;
    unsigned off = (unsigned) node;
    const unsigned mine = st0 + (unsigned) (H + node) * 8u;                   /* this node in the A^n part of stage 0  */
    const unsigned n1b  = (unsigned) N1 * 8u;
    const unsigned dM   = (unsigned) (W - H + HM) * 8u;                       /* from there to its A^{n-1} value       */

    int q = 0;                                            /* ring position of the next plane to take          */
    /* take plane q out of the ring: its cross; returns the node's address in that stage.  The stage is handed back one
     * step later, after the A^{n-1} value that came with it has been read where it is needed (one live double less)      */
    auto take = [&] (Cross& x) {
      const unsigned s = (unsigned) q & (NB - 1);
      mbar_wait_a(bars + s * 8u, (unsigned) ((q / NB) & 1));
      const unsigned a = mine + s * SB;
      x.c = lds_f64(a); x.xp = lds_f64(a + n1b); x.xm = lds_f64(a - n1b); x.yp = lds_f64(a + 8u); x.ym = lds_f64(a - 8u);
      ++q; };
    /* hand the stage of ring position qq back (every lane of the warp has read what it needs of it)                     */
    auto release = [&] (unsigned qq) { __syncwarp(); if (lane0) mbar_arrive_a(bars + NB * 8u + (qq & (NB - 1)) * 8u); };

    Cross P0, P1, P2;
    take(P0);                                             /* plane ks-1                                       */
    release(0u);
    take(P1);                                             /* plane ks                                         */

    /* The source term.  J is loaded one plane ahead (its DRAM latency hides behind a whole plane of work) through the
     * non-coherent path: an ordinary load still in flight would hold up the release of the ring stage.  J's only
     * reader also clears it (FdTd::currentReset): the thread that loaded a value stores the zero afterwards, data-dependent
     * on the load, and nothing else touches that address during the launch.                                          */
    double srcn = 0.0;
    if (srcon & 1ull) srcn = __ldg(jnc + off);
    double uxn = 0.0, uyn = 0.0;                          /* SEED, face warp: the seed scalars of the next plane         */
    if (SEED && seeded && ks >= rz.KI && ks < rz.KF)
      {
	const double* su = rz.seedu + (long) ks * 8 * rz.L;
	if (sxo >= 0) uxn = __ldg(su + sxo);
	if (syo >= 0) uyn = __ldg(su + syo);
      }
    /* one plane: Z is plane k (ring position q - 1 on entry), M plane k-1, the new plane k+1 lands in Pn           */
    #define MITHRA_STREAM_STEP(M, Z, Pn)                                                                        \
      {                                                                                                         \
	const double src = srcn;                                                                                \
	srcon >>= 1;                                                                                            \
	srcn = 0.0;                                                                                             \
	if (srcon & 1ull)                                                                                       \
	  {                                                                                                     \
	    const double* jq = jnc + (off + PpU);                                                               \
	    srcn = __ldg(jq);                                                                                   \
	    asm volatile("prefetch.global.L2 [%0];" :: "l"(jq + 2u * PpU));   /* pencils are 8 planes tall: two planes on, the load then hits L2 */ \
	  }                                                                                                     \
	double ux = 0.0, uy = 0.0;                               /* face warp, seeded job: the seed scalars of plane k */ \
	if (SEED && seeded)                                      /* loaded a plane ahead like J, L2 four planes on     */ \
	  {                                                                                                     \
	    ux = uxn; uy = uyn; uxn = 0.0; uyn = 0.0;                                                           \
	    const int kn = ks + q - 1;                                                                          \
	    if (kn >= rz.KI && kn < rz.KF)                                                                      \
	      {                                                                                                 \
		const double* su = rz.seedu + (long) kn * 8 * rz.L;                                             \
		const long ahead = (kn + 4 < rz.KF) ? 32L * rz.L : 0L;                                         \
		if (sxo >= 0) { uxn = __ldg(su + sxo); asm volatile("prefetch.global.L2 [%0];" :: "l"(su + sxo + ahead)); } \
		if (syo >= 0) { uyn = __ldg(su + syo); asm volatile("prefetch.global.L2 [%0];" :: "l"(su + syo + ahead)); } \
	      }                                                                                                 \
	  }                                                                                                     \
	take(Pn);                                                                                               \
	const unsigned az = mine + (((unsigned) q - 2u) & (NB - 1)) * SB;        /* the node in the stage of plane k */ \
	const double vm1 = lds_f64(az + dM);                                             /* A^{n-1} of plane k */    \
	double r = stencil_value<NSFD>(M, Z, Pn, vm1, src, f.a[0], f.a[1], f.a[2], f.a[3], as, f.alpha, f.beta); \
	if (SEED && seeded)                                      /* x shell term, then y shell term (fdtd.cpp:312-350) */ \
	  {                                                                                                     \
	    const int k = ks + q - 3;                                                                           \
	    if (k >= rz.KI && k < rz.KF)                                                                        \
	      {                                                                                                 \
		const double polc = (c == 0 ? rz.pol[0] : c == 1 ? rz.pol[1] : rz.pol[2]);                      \
		if (sxo >= 0)                                                                                   \
		  { const double S = seed_assemble_comp(ux, polc, rz.ni, rz.supergaussian, c == 2, rz.gamma); r = sxminus ? r - f.a[1] * S : r + f.a[1] * S; } \
		if (syo >= 0)                                                                                   \
		  { const double S = seed_assemble_comp(uy, polc, rz.ni, rz.supergaussian, c == 2, rz.gamma); r = syminus ? r - f.a[2] * S : r + f.a[2] * S; } \
	      }                                                                                                 \
	  }                                                                                                     \
	if (interior) apc[off] = r;                                                                             \
	if (FACES && fs != 0)                                            /* face warp: the y face node behind */  \
	  {                                                                                                     \
	    const bool lo = fs < 0;                                                                             \
	    const double ams = lds_f64(az + fs + dM), s1p = lds_f64(az + fs + n1b), s1m = lds_f64(az + fs - n1b); \
	    apc[(int) off + (fs >> 3)] = face_value(f.cB, ams, r, vm1, lo ? Z.ym : Z.yp, Z.c, Z.xp, Z.xm, s1p, s1m, \
						     Pn.c, M.c, lo ? Pn.ym : Pn.yp, lo ? M.ym : M.yp);           \
	  }                                                                                                     \
	release((unsigned) q - 2u);                                                                             \
	if (src != 0.0) jnc[off] = src * 0.0;                    /* a (signed) zero, ordered after the load */ \
	off += PpU;                                                                                             \
      }
    while (true)
      {
	MITHRA_STREAM_STEP(P0, P1, P2); if (q >= nq) break;
	MITHRA_STREAM_STEP(P1, P2, P0); if (q >= nq) break;
	MITHRA_STREAM_STEP(P2, P0, P1); if (q >= nq) break;
      }
    #undef MITHRA_STREAM_STEP
  }

  /* The x-shell corrections of a TF/SF seed (fdtd.cpp:312-330) on the nodes stencil_stream<.., FACES> leaves to this pass:
   * rows i = 1, 2, N0-3, N0-2, j in [3, N1-4] (the face warp has done j = 2 and N1-3, which also take a y term), planes
   * [KI, KF): A+(1) -= a1 S(2), A+(2) += a1 S(1), A+(N0-2) -= a1 S(N0-3), A+(N0-3) += a1 S(N0-2).  Whole rows: coalesced.     */
  __global__ void __launch_bounds__(128)
  seed_xshell_rows (const FieldDev f, const RimDev rz, double* __restrict__ anp1)
  {
    const int  nj = f.N1 - 6;
    const long per = 4L * nj, tot = per * (rz.KF - rz.KI) * 3;
    const long cs = (long) f.np * f.Pp;
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	const int c = (int) (t / (per * (rz.KF - rz.KI)));
	long r = t - (long) c * per * (rz.KF - rz.KI);
	const int k = rz.KI + (int) (r / per); r -= (long) (k - rz.KI) * per;
	const int q = (int) (r / nj), j = 3 + (int) (r - (long) q * nj);
	const int i = (q == 0) ? 1 : (q == 1) ? 2 : (q == 2) ? f.N0 - 3 : f.N0 - 2;
	const int line = (q == 0) ? 1 : (q == 1) ? 0 : (q == 2) ? 3 : 2;
	const double S = seed_assemble_comp(rz.seedu[((long) k * 8 + line) * rz.L + j], rz.pol[c], rz.ni, rz.supergaussian, c == 2, rz.gamma);
	double* a = anp1 + (long) c * cs + (long) k * f.Pp + (long) i * f.N1 + j;
	const double v = *a;
	*a = (q == 0 || q == 3) ? v - f.a[1] * S : v + f.a[1] * S;
      }
  }

  #ifndef MITHRA_RIM_MINBLOCKS
  #define MITHRA_RIM_MINBLOCKS 6                        /* 80 registers: the kernel is pure load latency, 6 beats 5 and 8    */
  #endif

  /* ------------------------------------------------------------------------------------------------
   * Rim of every plane: rows i = 1, 2, N0-3, N0-2 and columns j = 1, 2, N1-3, N1-2 (N0, N1 >= 8).
   *
   * On these nodes fdtd.cpp does three things one after the other: the interior sweep (:270-303), the TF/SF seed
   * terms of the x shell and then of the y shell (:307-351), and -- reading the finished A+ of the nodes i = 1, N0-2
   * / j = 1, N1-2 -- the x / y absorbing faces (:377-520).  As three grid passes the last two are strided
   * read-modify-writes of single nodes per row; here ONE thread owns a rim node, marches KC planes in +z with the
   * 5-point crosses of the planes k-1, k, k+1 in registers (like stencil_interior), and does all of it in the
   * reference's order per node: interior value, x-shell term, y-shell term, store, then the face node(s) behind it
   * from values it already holds plus three loads.  stencil_stream skips the rim (skiprim).
   * Thread enumeration: 4 rows x (N1-2) nodes, j fastest, then 4 columns x rows [3, N0-4], column fastest.
   * Results are bit-identical to stencil + seed_inject + boundary_faces run one after the other.
   * ------------------------------------------------------------------------------------------------ */
  template <bool NSFD>
  __global__ void __launch_bounds__(128, MITHRA_RIM_MINBLOCKS)
  rim_update (const FieldDev f, const RimDev rz, double* __restrict__ anp1, const double* __restrict__ an,
	      const double* __restrict__ anm1, double* __restrict__ jn, const Box* __restrict__ jbox, int KC,
	      const unsigned char* __restrict__ jmask, int zero_j)
  {
    const int N0 = f.N0, N1 = f.N1;
    const int nr = N1 - 2, nrows = 4 * nr, per = nrows + 4 * (N0 - 6);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per) return;
    int i, j;
    if (t < nrows) { const int q = t / nr; j = 1 + t - q * nr; i = (q == 0) ? 1 : (q == 1) ? 2 : (q == 2) ? N0 - 3 : N0 - 2; }
    else           { const int u = t - nrows, q = u & 3; i = 3 + (u >> 2); j = (q == 0) ? 1 : (q == 1) ? 2 : (q == 2) ? N1 - 3 : N1 - 2; }

    const int c  = blockIdx.z;
    const int ks = f.kb + blockIdx.y * KC, ke = min(ks + KC, f.np - 1);    /* planes ks .. ke-1                */
    if (ks >= ke) return;
    const long Pp = f.Pp;
    const long nb = (long) c * f.np * Pp + (long) i * N1 + j;            /* the node in plane 0               */
    const double* A  = an   + nb;
    const double* Am = anm1 + nb;
    double*       Ap = anp1 + nb;
    double*       Jn = jn   + nb;

    const double a0 = f.a[0], a1 = f.a[1], a2 = f.a[2], a3 = f.a[3];
    const double as = (c < 3) ? f.a[4] : f.a[5];
    const double alpha = f.alpha, beta = f.beta;
    const Box bx = *jbox;
    unsigned long long srcon = source_planes(f, bx, jmask, i, j, (long) i * N1 + j, ks, ke);

    /* faces behind this node: offset of the face node s from n, 0 = none                                      */
    const int dsx = (i == 1) ? -N1 : (i == N0 - 2) ? N1 : 0;
    const int dsy = (j == 1) ? -1  : (j == N1 - 2) ? 1  : 0;
    /* seed terms (fdtd.cpp:312-350): A+(1) -= a S(2), A+(2) += a S(1), A+(N-2) -= a S(N-3), A+(N-3) += a S(N-2)   */
    int sxl = -1, syl = -1;                               /* source line in RimDev.seedu                      */
    if (rz.seed && c < 3)
      {
	if (j >= 2 && j <= N1 - 3) sxl = (i == 1) ? 1 : (i == 2) ? 0 : (i == N0 - 2) ? 2 : (i == N0 - 3) ? 3 : -1;
	if (i >= 2 && i <= N0 - 3) syl = (j == 1) ? 5 : (j == 2) ? 4 : (j == N1 - 2) ? 6 : (j == N1 - 3) ? 7 : -1;
      }
    const bool sxminus = (i == 1 || i == N0 - 2), syminus = (j == 1 || j == N1 - 2);
    const double polc = (c < 3) ? rz.pol[c] : 0.0;

    auto cross_at = [&] (int k) { Cross x; const double* q = A + (long) k * Pp; x.c = q[0]; x.xp = q[N1]; x.xm = q[-N1]; x.yp = q[1]; x.ym = q[-1]; return x; };

    Cross M = cross_at(ks - 1), Z = cross_at(ks);
    for (int k = ks; k < ke; k++)
      {
	const long ko = (long) k * Pp;
	const Cross Pn = cross_at(k + 1);
	const double vm1 = Am[ko];
	double src = 0.0;
	const bool hadj = srcon & 1ull;
	if (hadj) src = Jn[ko];
	srcon >>= 1;
	/* loads of the fused work, issued with the rest                                                          */
	const bool seedk = (k >= rz.KI && k < rz.KF);
	double ux = 0.0, uy = 0.0, amsx = 0.0, amsy = 0.0, dxp = 0.0, dxm = 0.0, dyp = 0.0, dym = 0.0;
	if (sxl >= 0 && seedk) ux = rz.seedu[((long) k * 8 + sxl) * rz.L + j];
	if (syl >= 0 && seedk) uy = rz.seedu[((long) k * 8 + syl) * rz.L + i];
	if (dsx != 0) { amsx = Am[ko + dsx]; dxp = A[ko + dsx + 1];  dxm = A[ko + dsx - 1];  }
	if (dsy != 0) { amsy = Am[ko + dsy]; dyp = A[ko + dsy + N1]; dym = A[ko + dsy - N1]; }

	double r = stencil_value<NSFD>(M, Z, Pn, vm1, src, a0, a1, a2, a3, as, alpha, beta);
	if (sxl >= 0 && seedk)
	  { const double S = seed_assemble_comp(ux, polc, rz.ni, rz.supergaussian, c == 2, rz.gamma); r = sxminus ? r - a1 * S : r + a1 * S; }
	if (syl >= 0 && seedk)
	  { const double S = seed_assemble_comp(uy, polc, rz.ni, rz.supergaussian, c == 2, rz.gamma); r = syminus ? r - a2 * S : r + a2 * S; }
	Ap[ko] = r;
	if (hadj && zero_j) Jn[ko] = 0.0;                        /* J's only reader clears it (with the other stores: a store
								    next to the load would split the batch of loads above) */
	if (dsx != 0)
	  {
	    const bool lo = dsx < 0;
	    Ap[ko + dsx] = face_value(f.bB, amsx, r, vm1, lo ? Z.xm : Z.xp, Z.c, Z.yp, Z.ym, dxp, dxm,
				      Pn.c, M.c, lo ? Pn.xm : Pn.xp, lo ? M.xm : M.xp);
	  }
	if (dsy != 0)
	  {
	    const bool lo = dsy < 0;
	    Ap[ko + dsy] = face_value(f.cB, amsy, r, vm1, lo ? Z.ym : Z.yp, Z.c, Z.xp, Z.xm, dyp, dym,
				      Pn.c, M.c, lo ? Pn.ym : Pn.yp, lo ? M.ym : M.yp);
	  }
	M = Z; Z = Pn;
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * Faces.  A+_s = B0 (A-_s + A+_n) + B1 A-_n + B2 (A_s + A_n)
   *              + B3 (A_{n+t1} + A_{n-t1} + A_{s+t1} + A_{s-t1}) + B4 (same along t2)
   * s = face node, n = s + dn its inward neighbour; x faces: (t1,t2) = (y,z), y faces: (x,z), z: (x,y).
   * ------------------------------------------------------------------------------------------------ */
  __device__ __forceinline__ void face_update (double* __restrict__ ap, const double* __restrict__ a,
					       const double* __restrict__ am, long s, long dn, long d1, long d2,
					       const double* B)
  {
    const long n = s + dn;
    ap[s] = B[0] * ( am[s] + ap[n] ) +
	    B[1] * am[n] +
	    B[2] * ( a[s] + a[n] ) +
	    B[3] * ( a[n + d1] + a[n - d1] + a[s + d1] + a[s - d1] ) +
	    B[4] * ( a[n + d2] + a[n - d2] + a[s + d2] + a[s - d2] );
  }

  /* grid-stride over all face nodes of all components; faces are enumerated x-, x+, y-, y+, [z-], [z+].  */
  __global__ void __launch_bounds__(256)
  boundary_faces (const FieldDev f, double* __restrict__ anp1, const double* __restrict__ an,
		  const double* __restrict__ anm1, int zonly)
  {
    const int  nk = f.np - 1 - f.kb;                      /* planes kb .. np-2                           */
    const long nx = (long) (f.N1 - 2) * nk;               /* per x face                                  */
    const long ny = (long) (f.N0 - 2) * nk;
    const long nz = (long) (f.N0 - 2) * (f.N1 - 2);
    const bool zlo = (f.rank == 0), zhi = (f.rank == f.size - 1);
    /* zonly = 1: rim_update (or stencil_stream) has done the x and y faces; 2: the x faces alone (stencil_stream does the
     * y faces and leaves these: whole rows, as cheap here as anywhere)                                                  */
    const long first = zonly == 1 ? 2 * nx + 2 * ny : 0;
    const long per = zonly == 2 ? 2 * nx : 2 * nx + 2 * ny + (zlo ? nz : 0) + (zhi ? nz : 0) - first;
    const long tot = per * f.ncomp;
    const long N1 = f.N1, Pp = f.Pp;

    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	const int  c = (int) (t / per);
	long       r = t - (long) c * per + first;
	const long cb = (long) c * f.np * Pp;
	double*       ap = anp1 + cb;
	const double* a  = an   + cb;
	const double* am = anm1 + cb;

	if (r < 2 * nx)
	  {
	    const bool hi = r >= nx; if (hi) r -= nx;
	    /* j fastest so that a warp walks along a row                                              */
	    const int k = f.kb + (int) (r / (f.N1 - 2)), j = 1 + (int) (r % (f.N1 - 2));
	    const int i = hi ? f.N0 - 1 : 0;
	    face_update(ap, a, am, (long) k * Pp + i * N1 + j, hi ? -N1 : N1, 1, Pp, f.bB);
	    continue;
	  }
	r -= 2 * nx;
	if (r < 2 * ny)
	  {
	    const bool hi = r >= ny; if (hi) r -= ny;
	    const int k = f.kb + (int) (r / (f.N0 - 2)), i = 1 + (int) (r % (f.N0 - 2));
	    const int j = hi ? f.N1 - 1 : 0;
	    face_update(ap, a, am, (long) k * Pp + i * N1 + j, hi ? -1 : 1, N1, Pp, f.cB);
	    continue;
	  }
	r -= 2 * ny;
	{
	  bool hi;
	  if (zlo && r < nz) hi = false; else { hi = true; if (zlo) r -= nz; }
	  const int i = 1 + (int) (r / (f.N1 - 2)), j = 1 + (int) (r % (f.N1 - 2));
	  const int k = hi ? f.np - 1 : 0;
	  face_update(ap, a, am, (long) k * Pp + i * N1 + j, hi ? -Pp : Pp, N1, 1, f.dB);
	}
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * Edges (second-order truncation only).  For the edge node e with inward neighbours u, v and diagonal d,
   * running along w:
   *   A+_e = E0 (A+_u + A-_v) + E1 (A-_u + A+_v) + E2 (A-_e + A+_d) + E3 (A_e + A_u + A_v + A_d)
   *        + E4 ( sum over {e,u,v,d} at w-1, then at w+1 ) - A-_d
   * z edges (eE): u = i-neighbour, v = j-neighbour; x edges (fE): u = j, v = k; y edges (gE): u = k, v = i.
   * ------------------------------------------------------------------------------------------------ */
  __device__ __forceinline__ void edge_update (double* __restrict__ ap, const double* __restrict__ a,
					       const double* __restrict__ am, long e, long du, long dv, long dw,
					       const double* E)
  {
    const long u = e + du, v = e + dv, d = e + du + dv;
    ap[e] = E[0] * ( ap[u] + am[v] ) +
	    E[1] * ( am[u] + ap[v] ) +
	    E[2] * ( am[e] + ap[d] ) +
	    E[3] * ( a[e] + a[u] + a[v] + a[d] ) +
	    E[4] * ( a[e - dw] + a[u - dw] + a[v - dw] + a[d - dw] + a[e + dw] + a[u + dw] + a[v + dw] + a[d + dw] ) -
	    am[d];
  }

  __global__ void __launch_bounds__(256)
  boundary_edges (const FieldDev f, double* __restrict__ anp1, const double* __restrict__ an,
		  const double* __restrict__ anm1)
  {
    const bool zlo = (f.rank == 0), zhi = (f.rank == f.size - 1);
    const long nze = 4L * (f.np - 1 - f.kb);
    const int  nends = (zlo ? 1 : 0) + (zhi ? 1 : 0);
    const long nxe = 2L * nends * (f.N0 - 2);
    const long nye = 2L * nends * (f.N1 - 2);
    const long per = nze + nxe + nye;
    const long tot = per * f.ncomp;
    const long N1 = f.N1, Pp = f.Pp;

    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	const int  c = (int) (t / per);
	long       r = t - (long) c * per;
	const long cb = (long) c * f.np * Pp;
	double*       ap = anp1 + cb;
	const double* a  = an   + cb;
	const double* am = anm1 + cb;

	if (r < nze)
	  {
	    const int q = (int) (r & 3); const int k = f.kb + (int) (r >> 2);
	    const bool ihi = q & 1, jhi = q & 2;
	    const long e = (long) k * Pp + (ihi ? f.N0 - 1 : 0) * N1 + (jhi ? f.N1 - 1 : 0);
	    edge_update(ap, a, am, e, ihi ? -N1 : N1, jhi ? -1 : 1, Pp, f.eE);
	    continue;
	  }
	r -= nze;
	if (r < nxe)
	  {
	    /* x edges: enumerate (end, jhi, i)                                                          */
	    const long per_end = 2L * (f.N0 - 2);
	    int end = (int) (r / per_end); r -= end * per_end;
	    const bool khi = zlo ? (end == 1) : true;
	    const bool jhi = r >= (f.N0 - 2); if (jhi) r -= (f.N0 - 2);
	    const int i = 1 + (int) r;
	    const long e = (long) (khi ? f.np - 1 : 0) * Pp + i * N1 + (jhi ? f.N1 - 1 : 0);
	    edge_update(ap, a, am, e, jhi ? -1 : 1, khi ? -Pp : Pp, N1, f.fE);
	    continue;
	  }
	r -= nxe;
	{
	  const long per_end = 2L * (f.N1 - 2);
	  int end = (int) (r / per_end); r -= end * per_end;
	  const bool khi = zlo ? (end == 1) : true;
	  const bool ihi = r >= (f.N1 - 2); if (ihi) r -= (f.N1 - 2);
	  const int j = 1 + (int) r;
	  const long e = (long) (khi ? f.np - 1 : 0) * Pp + (ihi ? f.N0 - 1 : 0) * N1 + j;
	  edge_update(ap, a, am, e, khi ? -Pp : Pp, ihi ? -N1 : N1, 1, f.gE);
	}
      }
  }

  /* Corners: A+_c = -[ h16 A_c + h8 A-_c + sum_{n=1..7} ( h_n A+_n + h16 A_n + h_{8+n} A-_n ) ] / h0,
   * n over the inward neighbours in the order (i), (j), (k), (ij), (ik), (jk), (ijk).                    */
  __global__ void boundary_corners (const FieldDev f, double* __restrict__ anp1, const double* __restrict__ an,
				    const double* __restrict__ anm1)
  {
    const int t = threadIdx.x;                         /* 8 corners x ncomp threads                        */
    if (t >= 8 * f.ncomp) return;
    const int c = t >> 3, q = t & 7;
    const bool ihi = q & 1, jhi = q & 2, khi = q & 4;
    if (!khi && f.rank != 0) return;
    if ( khi && f.rank != f.size - 1) return;
    const long N1 = f.N1, Pp = f.Pp;
    const long cb = (long) c * f.np * Pp;
    double*       ap = anp1 + cb;
    const double* a  = an   + cb;
    const double* am = anm1 + cb;
    const long m  = (long) (khi ? f.np - 1 : 0) * Pp + (ihi ? f.N0 - 1 : 0) * N1 + (jhi ? f.N1 - 1 : 0);
    const long si = ihi ? -N1 : N1, sj = jhi ? -1 : 1, sk = khi ? -Pp : Pp;
    const long nb[7] = { m + si, m + sj, m + sk, m + si + sj, m + si + sk, m + sj + sk, m + si + sj + sk };
    const double* h = f.hC;
    double s = a[m] * h[16] + am[m] * h[8];
    #pragma unroll
    for (int n = 0; n < 7; n++)
      {
	s = s + ap[nb[n]] * h[1 + n];
	s = s + a [nb[n]] * h[16];
	s = s + am[nb[n]] * h[9 + n];
      }
    ap[m] = - s / h[0];
  }

  /* ------------------------------------------------------------------------------------------------
   * Clear the deposit box of J (FdTd::currentReset, fdtd.cpp:23-32, restricted to the nodes that can be
   * non-zero) and leave the box empty for the next deposit.
   * ------------------------------------------------------------------------------------------------ */
  __global__ void __launch_bounds__(256)
  clear_current_box (const FieldDev f, double* __restrict__ jn, Box* __restrict__ jbox, unsigned int* __restrict__ done,
		     const unsigned char* __restrict__ jmask, int ends_only)
  {
    const Box b = *jbox;
    const int ni = b.hi[0] - b.lo[0] + 1, nj = b.hi[1] - b.lo[1] + 1;
    /* ends_only: stencil_stream and rim_update have cleared what they read -- every plane this slab updates; what is
     * left are the planes below kb (ghosts whose deposits went to the neighbour; the z face on the first slab) and np-1:
     * the planes lo[2] .. min(hi[2], kb-1), then np-1 if the box reaches it                                         */
    const int nlow = ends_only ? max(0, min(b.hi[2], f.kb - 1) - b.lo[2] + 1) : 0;
    const int nk = ends_only ? nlow + ( ( b.hi[2] >= f.np - 1 && b.lo[2] <= f.np - 1 ) ? 1 : 0 ) : b.hi[2] - b.lo[2] + 1;
    if (ni > 0 && nj > 0 && nk > 0)
      {
	/* one warp per row of the box (nj contiguous nodes): full sectors whatever the width of the box          */
	/* 32-bit row arithmetic (a slab has fewer than 2^31 rows): the 64-bit divisions were a good part of the kernel      */
	const long rows64 = (long) ni * nk * f.ncomp;
	const int  rows = rows64 < 0x7fffffffL ? (int) rows64 : 0x7fffffff;
	const int  lane = threadIdx.x & 31, per = ni * nk;
	const int  wstride = (int) (((long) gridDim.x * blockDim.x) >> 5);
	for (int w = (int) (((long) blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < rows; w += wstride)
	  {
	    const int c = w / per, r = w - c * per;
	    const int kr = r / ni, i = b.lo[0] + r % ni;
	    const int k = ends_only ? ( kr < nlow ? b.lo[2] + kr : f.np - 1 ) : b.lo[2] + kr;
	    double* row = jn + fidx(f.Pp, f.np, f.N1, c, k, i, b.lo[1]);
	    /* only the pencils that can hold a deposit (source_planes); every plane that takes the neighbours' deposits   */
	    const bool all = !jmask || ( f.size > 1 && ( k == f.kb || k >= f.np - 3 ) );
	    const unsigned char* mrow = jmask ? jmask + (long) (k >> MITHRA_EB_CHUNK_LOG2) * f.P + (long) i * f.N1 + b.lo[1] : 0;
	    for (int j = lane; j < nj; j += 32) if (all || mrow[j]) row[j] = 0.0;
	  }
      }
    /* the last block to finish empties the box                                                          */
    __syncthreads();
    if (threadIdx.x == 0)
      {
	__threadfence();
	const unsigned int prev = atomicAdd(done, 1u);
	if (prev == gridDim.x - 1)
	  {
	    *done = 0u;
	    jbox->lo[0] = jbox->lo[1] = jbox->lo[2] = 0x7fffffff;
	    jbox->hi[0] = jbox->hi[1] = jbox->hi[2] = -1;
	  }
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * E / B evaluation at one node (fdtd.cpp:818-845; fdtdSC.cpp:1110-1141), including the float roundings of
   * FieldVector<float>::dv / mdv (fieldvector.h:86-108): E is rounded to float twice (three times with phi).
   * ------------------------------------------------------------------------------------------------ */
  struct EB { float e[3]; float b[3]; };

  /* the arithmetic of one node from the values already loaded: the node's own A+ (p*) and A (q*), the centred
   * differences of phi (g*) and of the components of A (a**) and A+ (p**); every operation in the reference's order  */
  template <bool SC>
  __device__ __forceinline__ EB eb_assemble (const FieldDev& f, double p0, double p1, double p2, double q0, double q1, double q2,
					     double g0, double g1, double g2,
					     double azy, double ayz, double pzy, double pyz,
					     double axz, double azx, double pxz, double pzx,
					     double ayx, double axy, double pyx, double pxy)
  {
    const double mdt = - f.dt;
    EB o;
    {
      float e;
      e = (float) div_fast( p0, mdt, f.rmdt ); e = (float) ( (double) e - div_fast( q0, mdt, f.rmdt ) ); o.e[0] = e;
      e = (float) div_fast( p1, mdt, f.rmdt ); e = (float) ( (double) e - div_fast( q1, mdt, f.rmdt ) ); o.e[1] = e;
      e = (float) div_fast( p2, mdt, f.rmdt ); e = (float) ( (double) e - div_fast( q2, mdt, f.rmdt ) ); o.e[2] = e;
    }
    if (SC)
      {
	o.e[0] = (float) ( (double) o.e[0] - div_fast( g0, f.dx2, f.rdx2 ) );
	o.e[1] = (float) ( (double) o.e[1] - div_fast( g1, f.dy2, f.rdy2 ) );
	o.e[2] = (float) ( (double) o.e[2] - div_fast( g2, f.dz2, f.rdz2 ) );
      }
    o.b[0] = (float) ( 0.5 * ( div_fast( azy, f.dy2, f.rdy2 ) - div_fast( ayz, f.dz2, f.rdz2 ) + div_fast( pzy, f.dy2, f.rdy2 ) - div_fast( pyz, f.dz2, f.rdz2 ) ) );
    o.b[1] = (float) ( 0.5 * ( div_fast( axz, f.dz2, f.rdz2 ) - div_fast( azx, f.dx2, f.rdx2 ) + div_fast( pxz, f.dz2, f.rdz2 ) - div_fast( pzx, f.dx2, f.rdx2 ) ) );
    o.b[2] = (float) ( 0.5 * ( div_fast( ayx, f.dx2, f.rdx2 ) - div_fast( axy, f.dy2, f.rdy2 ) + div_fast( pyx, f.dx2, f.rdx2 ) - div_fast( pxy, f.dy2, f.rdy2 ) ) );
    return o;
  }

  template <bool SC>
  __device__ __forceinline__ EB eval_eb_node (const FieldDev& f, const double* __restrict__ anp1,
					      const double* __restrict__ an, int i, int j, int k)
  {
    const long N1 = f.N1, Pp = f.Pp;
    const long cs = (long) f.np * Pp;                    /* component stride                              */
    const long m  = (long) k * Pp + (long) i * N1 + j;
    const double* ax = an,   * ay = an   + cs, * az = an   + 2 * cs;
    const double* px = anp1, * py = anp1 + cs, * pz = anp1 + 2 * cs;
    /* all the loads first, in straight-line code: the divisions below contain (rare) branches the compiler will
     * not move loads across, and the kernel lives on having every load of a node in flight at once              */
    const double p0 = px[m], p1 = py[m], p2 = pz[m], q0 = ax[m], q1 = ay[m], q2 = az[m];
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    if (SC)
      {
	const double* fn = an + 3 * cs;
	g0 = fn[m + N1] - fn[m - N1]; g1 = fn[m + 1] - fn[m - 1]; g2 = fn[m + Pp] - fn[m - Pp];
      }
    const double azy = az[m + 1 ] - az[m - 1 ], ayz = ay[m + Pp] - ay[m - Pp], pzy = pz[m + 1 ] - pz[m - 1 ], pyz = py[m + Pp] - py[m - Pp];
    const double axz = ax[m + Pp] - ax[m - Pp], azx = az[m + N1] - az[m - N1], pxz = px[m + Pp] - px[m - Pp], pzx = pz[m + N1] - pz[m - N1];
    const double ayx = ay[m + N1] - ay[m - N1], axy = ax[m + 1 ] - ax[m - 1 ], pyx = py[m + N1] - py[m - N1], pxy = px[m + 1 ] - px[m - 1 ];
    return eb_assemble<SC>(f, p0, p1, p2, q0, q1, q2, g0, g1, g2, azy, ayz, pzy, pyz, axz, azx, pxz, pzx, ayx, axy, pyx, pxy);
  }

  /* Evaluate E/B on every node of `box` (node indices, already clamped to 1..N-2 transversally).  On the
   * global z ends plane 0 / np-1 take the values of plane 1 / np-2 (fdtd.cpp:754-773).                 */
  template <bool SC>
  __global__ void __launch_bounds__(256)
  eval_eb_box (const FieldDev f, const double* __restrict__ anp1, const double* __restrict__ an,
	       float4* __restrict__ eb, const Box* __restrict__ boxp, int ends_only)
  {
    const Box b = *boxp;
    const int ni = b.hi[0] - b.lo[0] + 1, nj = b.hi[1] - b.lo[1] + 1, nk = b.hi[2] - b.lo[2] + 1;
    if (ni <= 0 || nj <= 0 || nk <= 0) return;
    /* one warp per row of the box (lanes along j: neighbouring lanes share their y neighbours in L1 and the two
     * float4 stores of a warp are contiguous); 32-bit index arithmetic, once per row                            */
    /* ends_only: eval_eb_march has done the planes in between, what is left are the planes 0 and np-1 of the box  */
    const int rows = ends_only ? ni * 2 : ni * nk;
    const int lane = threadIdx.x & 31;
    const int wstride = (int) (((long) gridDim.x * blockDim.x) >> 5);
    for (int w = (int) (((long) blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < rows; w += wstride)
      {
	const int k = ends_only ? ( (w / ni) ? f.np - 1 : 0 ) : b.lo[2] + w / ni, i = b.lo[0] + w % ni;
	if (ends_only && (k < b.lo[2] || k > b.hi[2] || (k >= f.kb && k != f.np - 1))) continue;
	int ke = k;
	if (k < f.kb)      { if (f.rank != 0)          continue; ke = 1; }          /* ghosts come from the neighbour */
	if (k == f.np - 1) { if (f.rank != f.size - 1) continue; ke = f.np - 2; }
	for (int j = b.lo[1] + lane; j <= b.hi[1]; j += 32)
	  {
	    const EB o = eval_eb_node<SC>(f, anp1, an, i, j, ke);
	    const long m = (long) k * f.P + (long) i * f.N1 + j;
	    eb[2 * m]     = make_float4(o.e[0], o.e[1], o.e[2], 0.f);
	    eb[2 * m + 1] = make_float4(o.b[0], o.b[1], o.b[2], 0.f);
	  }
      }
  }
  /* ------------------------------------------------------------------------------------------------
   * E/B over the box, z-marching (the production path for the particle box; eval_eb_box above stays for the thin
   * plane boxes of the slab exchange and for the two copied end planes).
   *
   * eval_eb_box fetches the 24 (30) potentials of a node afresh for every node, eight (ten) of them from the planes
   * k-1 and k+1 that other CTAs own: fine for a bunch that fills a few dozen columns, but at FEL-LCLS scale the padded
   * particle box is 2/3 of a 347 M-node mesh and the evaluation costs more than the stencil.  Here a thread owns one
   * column (i, j) of the box and marches a chunk of 32 planes in +z: A_x, A_y (and phi) of the planes k-1, k, k+1 rotate through
   * registers, so every potential is loaded from memory once per thread and its in-plane neighbours are the centre
   * loads of the neighbouring threads of the same 32 x 8 tile (L1).  The arithmetic is eb_assemble, the same as
   * eval_eb_node: bit-identical E/B (GPU test: march == box).
   * Work items (32-plane chunk, i tile, j tile) are strided over a grid of fixed size because the box lives on the
   * device; pencils that the mask (spread_eb_mask below) leaves unmarked are skipped.
   * ------------------------------------------------------------------------------------------------ */
  #define MITHRA_MARCH_LOG2 5                   /* planes per work item of the march (four mask pencils)               */

  #ifndef MITHRA_MARCH_BARRIER
  #define MITHRA_MARCH_BARRIER 1
  #endif
  #ifndef MITHRA_MARCH_MINBLOCKS
  #define MITHRA_MARCH_MINBLOCKS 2                      /* 128 registers: all 30 loads of a node in flight before the arithmetic
							   (2.8 ms on FEL-LCLS; at 85 registers the scheduler sinks them: 3.5 ms) */
  #endif
  /* Work distribution: a work item is a tile of 32 x 8 node columns x 32 planes (four mask pencils per column).  With a
   * mask the marked (column, pencil) pairs of the tile are first COMPACTED into a list in shared memory, ordered by
   * pencil, then row, then column -- so that consecutive lanes still walk consecutive nodes of a row -- and the 256
   * threads take the list 256 units at a time: on a bunch with Gaussian tails (FEL-LCLS) only 39 % of the lanes of the
   * column-per-thread version had a marked pencil.  A unit is the 8 planes of one pencil: planes k-1, k of what is
   * differenced along z are loaded at its start (L1 hits when the pencil below belongs to the same column).           */
  template <bool SC>
  __global__ void __launch_bounds__(256, MITHRA_MARCH_MINBLOCKS)
  eval_eb_march (const FieldDev f, const double* __restrict__ anp1, const double* __restrict__ an,
		 float4* __restrict__ eb, const Box* __restrict__ boxp, const unsigned char* __restrict__ mask)
  {
    constexpr int L = MITHRA_MARCH_LOG2, LM = MITHRA_EB_CHUNK_LOG2, NS = 1 << (L - LM);
    __shared__ unsigned short unit[256 * NS];           /* (pencil << 8) | column                                    */
    __shared__ int wcount[NS][8], ubase[NS + 1];
    const Box b = *boxp;
    const int ni = b.hi[0] - b.lo[0] + 1, nj = b.hi[1] - b.lo[1] + 1;
    const int kfirst = max(b.lo[2], f.kb), klast = min(b.hi[2], f.np - 2);
    if (ni <= 0 || nj <= 0 || klast < kfirst) return;
    const int cfirst = kfirst >> L;
    const int njt = (nj + 31) >> 5, nit = (ni + 7) >> 3, nkc = (klast >> L) - cfirst + 1;
    const long nwork = (long) njt * nit * nkc;
    const int  tj = threadIdx.x & 31, ti = threadIdx.x >> 5;
    const long N1 = f.N1, Pp = f.Pp, cs = (long) f.np * Pp;
    const double* ax = an,   * ay = an   + cs, * az = an   + 2 * cs, * fn = an + 3 * cs;
    const double* px = anp1, * py = anp1 + cs, * pz = anp1 + 2 * cs;
    const int nch = (f.np + (1 << LM) - 1) >> LM;
    const long eP = 2L * f.P;

    for (long w = blockIdx.x; w < nwork; w += gridDim.x)
      {
	const int jt = (int) (w % njt), it = (int) ((w / njt) % nit), c = cfirst + (int) (w / ((long) njt * nit));
	const int i0 = b.lo[0] + (it << 3), j0 = b.lo[1] + (jt << 5);
	/* the pencils of this thread's column that are marked (no mask: all of them)                                */
	unsigned int on = 0u;
	{
	  const int j = j0 + tj, i = i0 + ti;
	  if (j <= b.hi[1] && i <= b.hi[0])
	    {
	      #pragma unroll
	      for (int s = 0; s < NS; s++)
		{ const int cm = (c << (L - LM)) + s; if (cm < nch && (!mask || mask[((long) cm * f.N0 + i) * N1 + j])) on |= 1u << s; }
	    }
	}
	/* compaction: per pencil the marked columns in (row, column) order                                          */
	unsigned int bal[NS];
	#pragma unroll
	for (int s = 0; s < NS; s++)
	  {
	    bal[s] = __ballot_sync(0xffffffffu, (on >> s) & 1u);
	    if (tj == 0) wcount[s][ti] = __popc(bal[s]);
	  }
	__syncthreads();
	if (threadIdx.x == 0)
	  {
	    int acc = 0;
	    #pragma unroll
	    for (int s = 0; s < NS; s++)
	      {
		ubase[s] = acc;
		for (int q = 0; q < 8; q++) { const int n = wcount[s][q]; wcount[s][q] = acc; acc += n; }
	      }
	    ubase[NS] = acc;
	  }
	__syncthreads();
	#pragma unroll
	for (int s = 0; s < NS; s++)
	  if ((on >> s) & 1u) unit[wcount[s][ti] + __popc(bal[s] & ((1u << tj) - 1u))] = (unsigned short) ((s << 8) | threadIdx.x);
	__syncthreads();
	const int nunits = ubase[NS];

	for (int u = threadIdx.x; u < nunits; u += 256)
	  {
	    const int code = unit[u], s = code >> 8, col = code & 255;
	    const int i = i0 + (col >> 5), j = j0 + (col & 31);
	    const int cm = (c << (L - LM)) + s;
	    const int ks = max(kfirst, cm << LM), ke = min(klast + 1, (cm + 1) << LM);
	    if (ks >= ke) continue;
	    /* the node in plane ks: one pointer per array, advanced by a plane per step (the compiler's own addressing of
	     * 26 loads through a recomputed 64-bit index was half of the kernel's instructions)                        */
	    const long m0 = (long) ks * Pp + (long) i * N1 + j;
	    const double* qax = ax + m0; const double* qay = ay + m0; const double* qaz = az + m0;
	    const double* qpx = px + m0; const double* qpy = py + m0; const double* qpz = pz + m0;
	    const double* qfn = fn + m0;
	    float4* qe = eb + 2 * ( (long) ks * f.P + (long) i * N1 + j );
	    /* planes k-1 (m), k (0), k+1 (p) of what is differenced along z                                            */
	    double axm = qax[-Pp], ax0 = qax[0], aym = qay[-Pp], ay0 = qay[0];
	    double pxm = qpx[-Pp], px0 = qpx[0], pym = qpy[-Pp], py0 = qpy[0];
	    double fm = 0.0, f0 = 0.0;
	    if (SC) { fm = qfn[-Pp]; f0 = qfn[0]; }
	    for (int k = ks; k < ke; k++, qax += Pp, qay += Pp, qaz += Pp, qpx += Pp, qpy += Pp, qpz += Pp, qfn += Pp, qe += eP)
	      {
		const double axp = qax[Pp], ayp = qay[Pp], pxp = qpx[Pp], pyp = qpy[Pp];
		const double q2 = qaz[0], p2 = qpz[0];
		double fp = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
		if (SC) { fp = qfn[Pp]; g0 = qfn[N1] - qfn[-N1]; g1 = qfn[1] - qfn[-1]; g2 = fp - fm; }
		const double azy = qaz[1 ] - qaz[-1 ], pzy = qpz[1 ] - qpz[-1 ];
		const double azx = qaz[N1] - qaz[-N1], pzx = qpz[N1] - qpz[-N1];
		const double ayx = qay[N1] - qay[-N1], pyx = qpy[N1] - qpy[-N1];
		const double axy = qax[1 ] - qax[-1 ], pxy = qpx[1 ] - qpx[-1 ];
		/* every load of the node is in flight before the arithmetic starts: left alone, the scheduler sinks each
		 * load to its first use to save registers and the kernel runs at the latency of one load after the other   */
		#if MITHRA_MARCH_BARRIER
		__syncwarp(__activemask());
		#endif
		const EB o = eb_assemble<SC>(f, px0, py0, p2, ax0, ay0, q2, g0, g1, g2,
					     azy, ayp - aym, pzy, pyp - pym,
					     axp - axm, azx, pxp - pxm, pzx,
					     ayx, axy, pyx, pxy);
		qe[0] = make_float4(o.e[0], o.e[1], o.e[2], 0.f);
		qe[1] = make_float4(o.b[0], o.b[1], o.b[2], 0.f);
		axm = ax0; ax0 = axp; aym = ay0; ay0 = ayp; pxm = px0; px0 = pxp; pym = py0; py0 = pyp;
		if (SC) { fm = f0; f0 = fp; }
	      }
	  }
	__syncthreads();                                   /* the list is rewritten by the next work item            */
      }
  }

  /* ------------------------------------------------------------------------------------------------
   * Node-pencil mask from the cell-pencil reach bytes (device_types.cuh "Reach mask"; the push and particle_box write
   * them).  The padded particle box is a bounding box: for a bunch with Gaussian tails (FEL-LCLS: sigma 7.5 cells,
   * truncation at 45, a particle per 25 cells) most of its nodes are out of every particle's reach.  Node pencil
   * (c, i, j) -- node column (i, j) x 8 planes -- is marked when some cell pencil (cc, ii, jj) holds a particle whose
   * reach covers it: the two node columns of its own cell per axis, one more column / the neighbouring plane chunk
   * where a particle of that pencil sits within c dt of the cell's face (the XLO .. ZHI bits).  eval_eb_march, the source
   * read of the next field update and the clear of J skip everything unmarked.
   * ------------------------------------------------------------------------------------------------ */
  /* Scatter form: the cell bytes are read four at a time (most words are zero), a marked cell pencil stores a 1 into the
   * node pencils it reaches -- 2 x 2 columns x 1 chunk for most particles, up to 4 x 4 x 3 -- and the node mask is cleared
   * beforehand (engine.cu).  The gather form (every node pencil of the box looking at its 48 neighbouring cell bytes) cost
   * 1.0 ms per step on FEL-LCLS once the pencils were 8 planes tall.                                                  */
  __global__ void __launch_bounds__(256)
  spread_eb_mask (const FieldDev f, const unsigned char* __restrict__ cells, unsigned char* __restrict__ nodes, long nbytes)
  {
    constexpr int L = MITHRA_EB_CHUNK_LOG2;
    const int nch = (f.np + (1 << L) - 1) >> L;
    /* the arrays are whole 16-byte words (engine.cu emask_bytes); sixteen cell bytes per load, nearly all of them zero:
     * the loop over the words is the whole cost of the kernel (32-bit counters; a slab has fewer than 2^31 cell pencils)  */
    const unsigned int nquads = (unsigned int) (nbytes >> 4);
    const uint4* cq = reinterpret_cast<const uint4*>(cells);
    const unsigned int P = (unsigned int) f.P, N1 = (unsigned int) f.N1;
    for (unsigned int w = blockIdx.x * blockDim.x + threadIdx.x; w < nquads; w += gridDim.x * blockDim.x)
      {
	const uint4 quad = __ldg(cq + w);
	if (!(quad.x | quad.y | quad.z | quad.w)) continue;
	/* (chunk, row, column) of the first of the sixteen cell pencils: one division per load, the rest by stepping      */
	const unsigned int t0 = 16u * w;
	int cc = (int) (t0 / P);
	const int r0 = (int) (t0 - (unsigned int) cc * P);
	int ii = r0 / (int) N1, jj = r0 - ii * (int) N1;
	#pragma unroll
	for (int h = 0; h < 4; h++)
	  {
	    unsigned int word = h == 0 ? quad.x : h == 1 ? quad.y : h == 2 ? quad.z : quad.w;
	    #pragma unroll
	    for (int q = 0; q < 4; q++, word >>= 8)
	      {
		const unsigned int v = word & 0xffu;
		if (v && cc < nch)
		  {
		    const int c0 = (v & REACH_ZLO) ? max(cc - 1, 0) : cc, c1 = (v & REACH_ZHI) ? min(cc + 1, nch - 1) : cc;
		    const int i0 = (v & REACH_XLO) ? max(ii - 1, 0) : ii, i1 = min((v & REACH_XHI) ? ii + 2 : ii + 1, f.N0 - 1);
		    const int j0 = (v & REACH_YLO) ? max(jj - 1, 0) : jj, j1 = min((v & REACH_YHI) ? jj + 2 : jj + 1, f.N1 - 1);
		    for (int c = c0; c <= c1; c++)
		      for (int i = i0; i <= i1; i++)
			{
			  unsigned char* row = nodes + ((long) c * f.N0 + i) * f.N1;
			  for (int j = j0; j <= j1; j++) row[j] = 1;
			}
		  }
		/* the next cell pencil                                                                                    */
		if (++jj == (int) N1) { jj = 0; if (++ii == f.N0) { ii = 0; ++cc; } }
	      }
	  }
      }
  }
}

#endif
