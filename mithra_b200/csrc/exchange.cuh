/* exchange.cuh -- z-slab exchange between the GPUs of one box over NVLink peer memory.
 *
 * Replaces the reference's blocking MPI ring:
 *   A / phi ghost planes         fdtd.cpp:728-738, fdtdSC.cpp:1003-1025      -> put_planes into the neighbour's array
 *   E / B ghost planes (float)   fdtd.cpp:742-799                             -> put_eb into the neighbour's array
 *   J / rho boundary merge       fdtd.cpp:201-210, fdtdSC.cpp:223-239         -> put_jmail + add_jmail
 *   particle migration           solver.cpp:1544-1568, 493-503                -> migrate_pack + put_outbox + unpack_inbox
 *
 * One handle = one slab = one GPU (one process per GPU under torchrun; several handles in one process, even on one
 * device, work the same way and are what the 1-GPU tests use).  Every slab exports a blob with the addresses of the
 * arrays its ring neighbours write into (CUDA IPC handles across processes, raw pointers inside one process).
 * Transfers are plain stores into mapped peer memory issued from the slab's own stream, followed by a sequence
 * number written to a flag in the neighbour's memory (__threadfence_system before it); the receiver's stream holds
 * a one-thread kernel that spins on the flag.  Puts are always enqueued before the matching waits, so no cycle can
 * form; a wait gives up after MITHRA_WAIT_TIMEOUT_NS and raises the slab's error flag instead of hanging the GPU.
 *
 * Write-after-read safety needs no extra handshake: a neighbour can only overwrite a ghost (or mailbox) of step n+1
 * after it has waited for this slab's put of step n+1, which this slab enqueues after every reader of step n.
 * The three rotating levels of A make the same argument hold for the potentials three steps apart.  The particle inboxes
 * are the exception -- the particle-only loop before the time origin has no other exchange to order them -- and exist
 * twice, one per parity of the hand-over: put(n + 2) follows this slab's put(n + 1), which follows its unpack(n).
 *
 * One process may drive several slabs (the host executable with --gpus larger than the number of devices, the one-GPU
 * tests): every handle owns two streams and some of their kernels spin on a neighbour's flag, so a process that
 * drives more than four slabs asks for 32 hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, set before the context exists:
 * host/main.cpp) -- with the default 8 the streams of more than four slabs on one device share queues and a spinning wait
 * can sit in front of the put it waits for.
 */
#ifndef MITHRA_EXCHANGE_CUH_
#define MITHRA_EXCHANGE_CUH_

#include <unistd.h>
#include <cstring>
#include <string>

#include "device_types.cuh"

namespace mithra
{
  /* a wait on a neighbour's flag gives up after this long and raises the slab's error flag: 10 s, or MITHRA_WAIT_TIMEOUT_S
   * seconds (a rank that stalls in long output I/O must not make its neighbours fail)                                */
  static inline unsigned long long wait_timeout_ns ()
  {
    static const unsigned long long t = [] { const char* e = getenv("MITHRA_WAIT_TIMEOUT_S"); const double s = e ? atof(e) : 10.0;
					     return (unsigned long long) ( ( s > 0.0 ? s : 10.0 ) * 1.0e9 ); } ();
    return t;
  }

  enum { XF_A_PREV = 0, XF_A_NEXT, XF_EB_PREV, XF_EB_NEXT, XF_J_PREV, XF_J_NEXT, XF_P_PREV, XF_P_NEXT, XF_COUNT = 16 };

  /* Header of the arena every slab owns (neighbours write into it).                                     */
  struct ArenaHeader
  {
    unsigned long long flag[XF_COUNT];     /* "data from prev / next of sequence s has arrived"             */
    unsigned int       in_count[2][2];     /* [parity][from prev / from next]: particles in that inbox          */
    Box                jbox_from_prev;     /* x-y-z extent (sender's internal numbering) of the J mail      */
    Box                jbox_from_next;
  };

  struct ArenaLayout
  {
    size_t header, jmail_prev, jmail_next, inbox_prev[2], inbox_next[2], bytes;   /* inboxes: one per parity of the hand-over */
    size_t plane_doubles;                  /* ncomp * Pp                                                    */
    unsigned int inbox_cap;
  };

  static inline ArenaLayout arena_layout (const FieldDev& f, unsigned int inbox_cap)
  {
    ArenaLayout L;
    L.plane_doubles = (size_t) f.ncomp * f.Pp;
    L.inbox_cap = inbox_cap;
    size_t o = 0;
    L.header = o;     o += (sizeof(ArenaHeader) + 255) / 256 * 256;
    L.jmail_prev = o; o += 1 * L.plane_doubles * sizeof(double);           /* their plane np-1 -> my plane kb     */
    L.jmail_next = o; o += 2 * L.plane_doubles * sizeof(double);           /* their planes 0,1 -> my np-3, np-2   */
    for (int b = 0; b < 2; b++)
      {
	L.inbox_prev[b] = o; o += (size_t) inbox_cap * 11 * sizeof(double);
	L.inbox_next[b] = o; o += (size_t) inbox_cap * 11 * sizeof(double);
      }
    L.bytes = (o + 255) / 256 * 256;
    return L;
  }

  /* What a ring neighbour needs to reach this slab.                                                      */
  struct ExchangeBlob
  {
    int    magic, pid, device, rank, size;
    int    np, kshift, kb;                 /* internal plane count etc. of the exporting slab               */
    int    ncomp; long Pp; unsigned int inbox_cap;
    void*  ptr[5];                         /* A[0], A[1], A[2], eb, arena                                   */
    cudaIpcMemHandle_t ipc[5];
  };
  #define MITHRA_BLOB_MAGIC 0x4d495448

  struct Peer
  {
    bool     present;                      /* a ring neighbour exists (size > 1)                            */
    bool     chain;                        /* it is also a field neighbour (no wrap for the potentials)     */
    bool     ipc;                          /* pointers were opened with cudaIpcOpenMemHandle                */
    int      np, kshift, kb;
    double*  A[3];
    float4*  eb;
    char*    arena;
  };

  struct Exchange
  {
    bool         connected;
    ArenaLayout  L;
    char*        arena;                    /* my arena (device)                                             */
    Peer         prev, next;
    unsigned long long seqA, seqEB, seqJ, seqP;
    int*         d_err;                    /* raised by a timed-out wait                                    */
    /* migration scratch */
    double*      outbox[2];                /* to prev / to next, [cap][11]                                  */
    unsigned int* d_cursor;                /* [0] to prev, [1] to next, [2] leavers                         */
    int*         d_leave;                  /* indices of the leavers                                        */
    int*         d_holes;                  /* scratch of fill_holes: holes | movers | marks                 */
    unsigned int* h_counts;                /* pinned: out_prev, out_next, in_prev, in_next, err             */
    Box*         d_planes_lo;              /* node boxes of the full planes evaluated for the E/B exchange  */
    Box*         d_planes_hi;
    std::string  error;
  };

  /* ---------------------------------------------------------------------------------------------------- */
  /* kernels                                                                                               */

  /* copy `nplanes` planes of every component from a local array into the neighbour's array                */
  __global__ void __launch_bounds__(256)
  put_planes (const double* __restrict__ src, double* __restrict__ dst, int ncomp, long Pp, int src_np, int dst_np,
	      int src_k, int dst_k, int nplanes)
  {
    const long per = (long) nplanes * Pp, tot = per * ncomp;
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	const int c = (int) (t / per); const long r = t - (long) c * per;
	dst[((long) c * dst_np + dst_k) * Pp + r] = src[((long) c * src_np + src_k) * Pp + r];
      }
  }

  /* E/B planes: [k][P][2] float4                                                                          */
  __global__ void __launch_bounds__(256)
  put_eb (const float4* __restrict__ src, float4* __restrict__ dst, long P, int src_k, int dst_k, int nplanes)
  {
    const long tot = (long) nplanes * P * 2;
    const float4* s = src + (long) src_k * P * 2; float4* d = dst + (long) dst_k * P * 2;
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x) d[t] = s[t];
  }

  /* J mail: planes [src_k, src_k + nplanes) of my J and my deposit box into the neighbour's mailbox       */
  __global__ void __launch_bounds__(256)
  put_jmail (const double* __restrict__ jn, double* __restrict__ mail, Box* __restrict__ mailbox, const Box* __restrict__ jbox,
	     int ncomp, long Pp, int np, int src_k, int nplanes)
  {
    const long per = (long) nplanes * Pp, tot = per * ncomp;
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	const int c = (int) (t / per); const long r = t - (long) c * per;
	mail[(long) c * per + r] = jn[((long) c * np + src_k) * Pp + r];
      }
    if (blockIdx.x == 0 && threadIdx.x == 0) *mailbox = *jbox;
  }

  /* add the mail into planes [dst_k, dst_k + nplanes) of my J where the sender deposited (its box, shifted by
   * dk = dst_k - sender's src_k), and grow my box accordingly                                               */
  __global__ void __launch_bounds__(256)
  add_jmail (double* __restrict__ jn, const double* __restrict__ mail, const Box* __restrict__ mailbox, Box* __restrict__ jbox,
	     int ncomp, long Pp, int N1, int np, int src_k, int dst_k, int nplanes)
  {
    const Box b = *mailbox;
    const int klo = max(b.lo[2], src_k), khi = min(b.hi[2], src_k + nplanes - 1);
    const int ni = b.hi[0] - b.lo[0] + 1, nj = b.hi[1] - b.lo[1] + 1, nk = khi - klo + 1;
    if (ni <= 0 || nj <= 0 || nk <= 0) return;
    const long per = (long) ni * nj * nk, tot = per * ncomp;
    const long mper = (long) nplanes * Pp;
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x)
      {
	const int c = (int) (t / per); long r = t - (long) c * per;
	const int k = klo + (int) (r / ((long) ni * nj)); r -= (long) (k - klo) * ni * nj;
	const int i = b.lo[0] + (int) (r / nj), j = b.lo[1] + (int) (r % nj);
	const long x = (long) i * N1 + j;
	jn[((long) c * np + (k - src_k + dst_k)) * Pp + x] += mail[(long) c * mper + (long) (k - src_k) * Pp + x];
      }
    if (blockIdx.x == 0 && threadIdx.x == 0)
      {
	atomicMin(&jbox->lo[0], b.lo[0]); atomicMin(&jbox->lo[1], b.lo[1]); atomicMin(&jbox->lo[2], klo - src_k + dst_k);
	atomicMax(&jbox->hi[0], b.hi[0]); atomicMax(&jbox->hi[1], b.hi[1]); atomicMax(&jbox->hi[2], khi - src_k + dst_k);
      }
  }

  __global__ void signal_flag (unsigned long long* flag, unsigned long long seq)
  {
    __threadfence_system();
    *((volatile unsigned long long*) flag) = seq;
    __threadfence_system();
  }

  /* err receives (first failure only) the flag index + 1 in the low byte and the awaited sequence above it  */
  __global__ void wait_flag (const unsigned long long* flag, unsigned long long seq, int* err, int id, unsigned long long timeout_ns)
  {
    unsigned long long t0, t1;
    asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*((volatile const unsigned long long*) flag) < seq)
      {
	asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t1));
	if (t1 - t0 > timeout_ns) { atomicCAS(err, 0, (int) (( seq << 8 ) | (unsigned) ( id + 1 ))); break; }
	__nanosleep(200);
      }
    __threadfence_system();
  }

  /* ---- particle migration ---------------------------------------------------------------------------- */

  __device__ __forceinline__ double xpmod (double a, double b) { double x = fmod(a, b); x += ( x < 0.0 ) ? b : 0.0; return x; }

  /* Leavers of this field step (solver.cpp:1544-1548, evaluated once per field step from the start-of-step
   * position rm): wrapped start position + displacement below zp0 -> previous slab, at or above zp1 -> next.    */
  __global__ void __launch_bounds__(256)
  migrate_pack (const __grid_constant__ BunchDev b, ParticlesDev P, long n, double* __restrict__ out_prev, double* __restrict__ out_next,
		unsigned int* __restrict__ cursor, int* __restrict__ leave, unsigned int cap, unsigned int leave_cap)
  {
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double z = P.r[2][t], zm = P.rm[2][t];
    const double zr = xpmod( zm - b.zmin, b.Lz ) + b.zmin + ( z - zm );
    int dir = -1;
    if      (zr <  b.zp0) dir = 0;
    else if (zr >= b.zp1) dir = 1;
    if (dir < 0) return;
    const unsigned int slot = atomicAdd(&cursor[dir], 1u);
    const unsigned int li   = atomicAdd(&cursor[2], 1u);
    if (li < leave_cap) leave[li] = (int) t;
    if (slot >= cap) return;                       /* overflow is reported by the host from the counters      */
    double* o = (dir == 0 ? out_prev : out_next) + (size_t) slot * 11;
    o[0] = P.q[t];
    o[1] = P.r[0][t];  o[2] = P.r[1][t];  o[3] = z;
    o[4] = P.rm[0][t]; o[5] = P.rm[1][t]; o[6] = zm;
    o[7] = P.gb[0][t]; o[8] = P.gb[1][t]; o[9] = P.gb[2][t];
    o[10] = P.e[t];
  }

  __global__ void __launch_bounds__(256)
  put_outbox (const double* __restrict__ out, const unsigned int* __restrict__ count, double* __restrict__ inbox,
	      unsigned int* __restrict__ in_count, unsigned int cap)
  {
    const unsigned int n = min(*count, cap);
    const long tot = (long) n * 11;
    for (long t = (long) blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long) gridDim.x * blockDim.x) inbox[t] = out[t];
    if (blockIdx.x == 0 && threadIdx.x == 0) *in_count = *count;
  }

  /* Close the gaps the nl leavers left in [0, n): survivors of the tail [n - nl, n) move into the holes below it.  */
  __global__ void __launch_bounds__(256)
  fill_holes (ParticlesDev P, long n, int nl, const int* __restrict__ leave, int* __restrict__ scratch)
  {
    const long n2 = n - nl;
    int* holes = scratch; int* movers = scratch + nl; int* mark = scratch + 2 * nl;
    __shared__ int nh, nm;
    for (int t = threadIdx.x; t < nl; t += blockDim.x) mark[t] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < nl; t += blockDim.x) if (leave[t] >= n2) mark[leave[t] - n2] = 1;
    __syncthreads();
    if (threadIdx.x == 0)
      {
	int a = 0, m = 0;
	for (int t = 0; t < nl; t++) { if (leave[t] < n2) holes[a++] = leave[t]; if (!mark[t]) movers[m++] = (int) (n2 + t); }
	nh = a; nm = m;
      }
    __syncthreads();
    const int cnt = min(nh, nm);                   /* equal by construction                                  */
    for (int t = threadIdx.x; t < cnt; t += blockDim.x)
      {
	const int d = holes[t], s = movers[t];
	P.q[d] = P.q[s]; P.e[d] = P.e[s]; P.id[d] = P.id[s];
	#pragma unroll
	for (int a = 0; a < 3; a++) { P.r[a][d] = P.r[a][s]; P.rm[a][d] = P.rm[a][s]; P.gb[a][d] = P.gb[a][s]; }
      }
  }

  /* Append the arrivals of one inbox at [n, n + cnt); they get the fresh upload indices id0, id0 + 1, ...     */
  __global__ void __launch_bounds__(256)
  unpack_inbox (ParticlesDev P, long n, const double* __restrict__ inbox, int cnt, unsigned int id0)
  {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const double* o = inbox + (size_t) t * 11;
    const long d = n + t;
    P.q[d] = o[0];
    P.r[0][d] = o[1];  P.r[1][d] = o[2];  P.r[2][d] = o[3];
    P.rm[0][d] = o[4]; P.rm[1][d] = o[5]; P.rm[2][d] = o[6];
    P.gb[0][d] = o[7]; P.gb[1][d] = o[8]; P.gb[2][d] = o[9];
    P.e[d] = o[10];
    P.id[d] = id0 + (unsigned int) t;
  }

  /* ---------------------------------------------------------------------------------------------------- */
  /* host side                                                                                             */

  #define XCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { x.error = std::string(#call) + ": " + cudaGetErrorString(e_); return 1; } } while (0)

  static inline ArenaHeader* hdr (char* arena) { return (ArenaHeader*) arena; }

  static inline int exchange_init (Exchange& x, const FieldDev& f, size_t pcap, cudaStream_t stream)
  {
    x.connected = false;
    x.arena = 0; x.d_err = 0; x.outbox[0] = x.outbox[1] = 0; x.d_cursor = 0; x.d_leave = 0; x.d_holes = 0; x.h_counts = 0;
    x.d_planes_lo = x.d_planes_hi = 0;
    memset(&x.prev, 0, sizeof(Peer)); memset(&x.next, 0, sizeof(Peer));
    x.seqA = x.seqEB = x.seqJ = x.seqP = 0;
    if (f.size <= 1) return 0;
    const unsigned int cap = (unsigned int) (pcap < 65536 ? pcap : 65536);
    x.L = arena_layout(f, cap);
    XCU(cudaMalloc(&x.arena, x.L.bytes));
    XCU(cudaMemsetAsync(x.arena, 0, x.L.bytes, stream));
    XCU(cudaMalloc(&x.d_err, sizeof(int))); XCU(cudaMemsetAsync(x.d_err, 0, sizeof(int), stream));
    for (int d = 0; d < 2; d++) XCU(cudaMalloc(&x.outbox[d], (size_t) cap * 11 * sizeof(double)));
    XCU(cudaMalloc(&x.d_cursor, 4 * sizeof(unsigned int))); XCU(cudaMemsetAsync(x.d_cursor, 0, 4 * sizeof(unsigned int), stream));
    XCU(cudaMalloc(&x.d_leave, (size_t) 2 * cap * sizeof(int)));
    XCU(cudaMalloc(&x.d_holes, (size_t) 6 * cap * sizeof(int)));
    XCU(cudaMallocHost(&x.h_counts, 8 * sizeof(unsigned int)));
    /* full planes whose E/B the neighbours need: kb -> prev; np-3, np-2 -> next                           */
    Box lo, hi;
    lo.lo[0] = 1; lo.hi[0] = f.N0 - 2; lo.lo[1] = 1; lo.hi[1] = f.N1 - 2; lo.lo[2] = f.kb; lo.hi[2] = f.kb;
    hi = lo; hi.lo[2] = f.np - 3; hi.hi[2] = f.np - 2;
    XCU(cudaMalloc(&x.d_planes_lo, sizeof(Box))); XCU(cudaMalloc(&x.d_planes_hi, sizeof(Box)));
    XCU(cudaMemcpyAsync(x.d_planes_lo, &lo, sizeof(Box), cudaMemcpyHostToDevice, stream));
    XCU(cudaMemcpyAsync(x.d_planes_hi, &hi, sizeof(Box), cudaMemcpyHostToDevice, stream));
    XCU(cudaStreamSynchronize(stream));
    return 0;
  }

  static inline void peer_close (Peer& p)
  {
    if (p.present && p.ipc)
      {
	for (int l = 0; l < 3; l++) if (p.A[l]) cudaIpcCloseMemHandle(p.A[l]);
	if (p.eb) cudaIpcCloseMemHandle(p.eb);
	if (p.arena) cudaIpcCloseMemHandle(p.arena);
      }
    memset(&p, 0, sizeof(Peer));
  }

  static inline void exchange_destroy (Exchange& x)
  {
    const bool same = (x.prev.present && x.next.present && x.prev.arena == x.next.arena);
    peer_close(x.prev);
    if (same) memset(&x.next, 0, sizeof(Peer)); else peer_close(x.next);
    cudaFree(x.arena); cudaFree(x.d_err); cudaFree(x.outbox[0]); cudaFree(x.outbox[1]); cudaFree(x.d_cursor);
    cudaFree(x.d_leave); cudaFree(x.d_holes); cudaFree(x.d_planes_lo); cudaFree(x.d_planes_hi);
    if (x.h_counts) cudaFreeHost(x.h_counts);
    x.arena = 0; x.connected = false;
  }

  static inline int exchange_export (Exchange& x, const FieldDev& f, int device, double* const* A, float4* eb, void* blob, size_t capacity, size_t* nbytes)
  {
    if (nbytes) *nbytes = sizeof(ExchangeBlob);
    if (!blob) return 0;
    if (f.size <= 1) { x.error = "a single slab has nothing to export"; return 1; }
    if (capacity < sizeof(ExchangeBlob)) { x.error = "blob buffer too small"; return 1; }
    ExchangeBlob b; memset(&b, 0, sizeof(b));
    b.magic = MITHRA_BLOB_MAGIC; b.pid = (int) getpid(); b.device = device; b.rank = f.rank; b.size = f.size;
    b.np = f.np; b.kshift = f.kshift; b.kb = f.kb; b.ncomp = f.ncomp; b.Pp = f.Pp; b.inbox_cap = x.L.inbox_cap;
    b.ptr[0] = A[0]; b.ptr[1] = A[1]; b.ptr[2] = A[2]; b.ptr[3] = eb; b.ptr[4] = x.arena;
    for (int i = 0; i < 5; i++)
      {
	cudaError_t e = cudaIpcGetMemHandle(&b.ipc[i], b.ptr[i]);
	if (e != cudaSuccess) { cudaGetLastError(); memset(&b.ipc[i], 0, sizeof(b.ipc[i])); }   /* in-process use still works */
      }
    memcpy(blob, &b, sizeof(b));
    return 0;
  }

  static inline int peer_open (Exchange& x, Peer& p, const ExchangeBlob& b, const FieldDev& f, int device, bool chain)
  {
    if (b.magic != MITHRA_BLOB_MAGIC) { x.error = "bad neighbour blob"; return 1; }
    if (b.ncomp != f.ncomp || b.Pp != f.Pp || b.size != f.size || b.inbox_cap != x.L.inbox_cap)
      { x.error = "neighbour blob describes a different mesh"; return 1; }
    p.present = true; p.chain = chain; p.np = b.np; p.kshift = b.kshift; p.kb = b.kb;
    void* q[5];
    if (b.pid == (int) getpid())
      {
	p.ipc = false;
	for (int i = 0; i < 5; i++) q[i] = b.ptr[i];
	if (b.device != device)
	  {
	    int can = 0; XCU(cudaDeviceCanAccessPeer(&can, device, b.device));
	    if (!can) { x.error = "no peer access between the two devices"; return 1; }
	    cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
	    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { x.error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return 1; }
	    cudaGetLastError();
	  }
      }
    else
      {
	p.ipc = true;
	for (int i = 0; i < 5; i++) XCU(cudaIpcOpenMemHandle(&q[i], b.ipc[i], cudaIpcMemLazyEnablePeerAccess));
      }
    p.A[0] = (double*) q[0]; p.A[1] = (double*) q[1]; p.A[2] = (double*) q[2]; p.eb = (float4*) q[3]; p.arena = (char*) q[4];
    return 0;
  }

  /* blobs of the ring neighbours rank-1 and rank+1 (mod size)                                            */
  static inline int exchange_connect (Exchange& x, const FieldDev& f, int device, const void* blob_prev, const void* blob_next)
  {
    if (f.size <= 1) { x.error = "a single slab has no neighbours"; return 1; }
    if (!blob_prev || !blob_next) { x.error = "both ring neighbours are required"; return 1; }
    ExchangeBlob bp, bn; memcpy(&bp, blob_prev, sizeof(bp)); memcpy(&bn, blob_next, sizeof(bn));
    if (bp.rank != (f.rank + f.size - 1) % f.size || bn.rank != (f.rank + 1) % f.size) { x.error = "neighbour blobs are not those of rank-1 / rank+1"; return 1; }
    if (peer_open(x, x.prev, bp, f, device, f.rank > 0)) return 1;
    if (bp.rank == bn.rank && bp.pid == bn.pid) { x.next = x.prev; x.next.chain = (f.rank < f.size - 1); }
    else if (peer_open(x, x.next, bn, f, device, f.rank < f.size - 1)) return 1;
    x.connected = true;
    return 0;
  }

  static inline int xgrid (long n, int cap) { long g = (n + 255) / 256; if (g < 1) g = 1; if (g > cap) g = cap; return (int) g; }

  /* A (and phi) ghost planes, fdtd.cpp:728-738: my plane kb -> prev's plane np-1; my plane np-2 -> next's plane kb-1. */
  static inline int exchange_potentials (Exchange& x, const FieldDev& f, double* anp1, int level, cudaStream_t s, int sms, unsigned long long* launches)
  {
    ++x.seqA;
    const long n = (long) f.ncomp * f.Pp;
    if (x.prev.chain)
      {
	put_planes<<<xgrid(n, sms * 2), 256, 0, s>>>(anp1, x.prev.A[level], f.ncomp, f.Pp, f.np, x.prev.np, f.kb, x.prev.np - 1, 1);
	signal_flag<<<1, 1, 0, s>>>(&hdr(x.prev.arena)->flag[XF_A_NEXT], x.seqA);
	*launches += 2;
      }
    if (x.next.chain)
      {
	put_planes<<<xgrid(n, sms * 2), 256, 0, s>>>(anp1, x.next.A[level], f.ncomp, f.Pp, f.np, x.next.np, f.np - 2, x.next.kb - 1, 1);
	signal_flag<<<1, 1, 0, s>>>(&hdr(x.next.arena)->flag[XF_A_PREV], x.seqA);
	*launches += 2;
      }
    if (x.prev.chain) { wait_flag<<<1, 1, 0, s>>>(&hdr(x.arena)->flag[XF_A_PREV], x.seqA, x.d_err, XF_A_PREV, wait_timeout_ns()); *launches += 1; }
    if (x.next.chain) { wait_flag<<<1, 1, 0, s>>>(&hdr(x.arena)->flag[XF_A_NEXT], x.seqA, x.d_err, XF_A_NEXT, wait_timeout_ns()); *launches += 1; }
    XCU(cudaGetLastError());
    return 0;
  }

  /* E/B ghost planes, fdtd.cpp:777-799 (+ one more plane below, see device_types.cuh): my plane kb -> prev's np-1;
   * my planes np-3, np-2 -> next's planes kb-2, kb-1.                                                       */
  static inline int exchange_eb (Exchange& x, const FieldDev& f, float4* eb, cudaStream_t s, int sms, unsigned long long* launches)
  {
    ++x.seqEB;
    if (x.prev.chain)
      {
	put_eb<<<xgrid(2L * f.P, sms * 2), 256, 0, s>>>(eb, x.prev.eb, f.P, f.kb, x.prev.np - 1, 1);
	signal_flag<<<1, 1, 0, s>>>(&hdr(x.prev.arena)->flag[XF_EB_NEXT], x.seqEB);
	*launches += 2;
      }
    if (x.next.chain)
      {
	put_eb<<<xgrid(4L * f.P, sms * 2), 256, 0, s>>>(eb, x.next.eb, f.P, f.np - 3, x.next.kb - 2, 2);
	signal_flag<<<1, 1, 0, s>>>(&hdr(x.next.arena)->flag[XF_EB_PREV], x.seqEB);
	*launches += 2;
      }
    if (x.prev.chain) { wait_flag<<<1, 1, 0, s>>>(&hdr(x.arena)->flag[XF_EB_PREV], x.seqEB, x.d_err, XF_EB_PREV, wait_timeout_ns()); *launches += 1; }
    if (x.next.chain) { wait_flag<<<1, 1, 0, s>>>(&hdr(x.arena)->flag[XF_EB_NEXT], x.seqEB, x.d_err, XF_EB_NEXT, wait_timeout_ns()); *launches += 1; }
    XCU(cudaGetLastError());
    return 0;
  }

  /* J (and rho) merge, fdtd.cpp:201-210: deposits on my ghost planes belong to the neighbours' stencils.
   * my planes kb-2, kb-1 -> prev adds into its np-3, np-2; my plane np-1 -> next adds into its plane kb.          */
  static inline int exchange_current (Exchange& x, const FieldDev& f, double* jn, Box* jbox, cudaStream_t s, int sms, unsigned long long* launches)
  {
    ++x.seqJ;
    const long n = (long) f.ncomp * f.Pp;
    if (x.prev.chain)
      {
	put_jmail<<<xgrid(2 * n, sms * 2), 256, 0, s>>>(jn, (double*) (x.prev.arena + x.L.jmail_next), &hdr(x.prev.arena)->jbox_from_next, jbox,
							  f.ncomp, f.Pp, f.np, f.kb - 2, 2);
	signal_flag<<<1, 1, 0, s>>>(&hdr(x.prev.arena)->flag[XF_J_NEXT], x.seqJ);
	*launches += 2;
      }
    if (x.next.chain)
      {
	put_jmail<<<xgrid(n, sms * 2), 256, 0, s>>>(jn, (double*) (x.next.arena + x.L.jmail_prev), &hdr(x.next.arena)->jbox_from_prev, jbox,
						      f.ncomp, f.Pp, f.np, f.np - 1, 1);
	signal_flag<<<1, 1, 0, s>>>(&hdr(x.next.arena)->flag[XF_J_PREV], x.seqJ);
	*launches += 2;
      }
    if (x.prev.chain)
      {
	wait_flag<<<1, 1, 0, s>>>(&hdr(x.arena)->flag[XF_J_PREV], x.seqJ, x.d_err, XF_J_PREV, wait_timeout_ns());
	/* the sender's plane np-1 is my plane kb                                                          */
	add_jmail<<<sms, 256, 0, s>>>(jn, (const double*) (x.arena + x.L.jmail_prev), &hdr(x.arena)->jbox_from_prev, jbox,
					f.ncomp, f.Pp, f.N1, f.np, x.prev.np - 1, f.kb, 1);
	*launches += 2;
      }
    if (x.next.chain)
      {
	wait_flag<<<1, 1, 0, s>>>(&hdr(x.arena)->flag[XF_J_NEXT], x.seqJ, x.d_err, XF_J_NEXT, wait_timeout_ns());
	/* the sender's planes kb-2, kb-1 are my planes np-3, np-2                                         */
	add_jmail<<<sms, 256, 0, s>>>(jn, (const double*) (x.arena + x.L.jmail_next), &hdr(x.arena)->jbox_from_next, jbox,
					f.ncomp, f.Pp, f.N1, f.np, x.next.kb - 2, f.np - 3, 2);
	*launches += 2;
      }
    XCU(cudaGetLastError());
    return 0;
  }

  /* Migration, first half: pack the leavers and put them into the ring neighbours' inboxes.               */
  static inline int migrate_begin (Exchange& x, const BunchDev& d_bd, ParticlesDev P, size_t pn, cudaStream_t s, int sms, unsigned long long* launches)
  {
    ++x.seqP;
    XCU(cudaMemsetAsync(x.d_cursor, 0, 4 * sizeof(unsigned int), s));
    if (pn > 0)
      {
	migrate_pack<<<(int) ((pn + 255) / 256), 256, 0, s>>>(d_bd, P, (long) pn, x.outbox[0], x.outbox[1], x.d_cursor, x.d_leave, x.L.inbox_cap, 2 * x.L.inbox_cap);
	*launches += 1;
      }
    /* to prev: it receives "from next"; to next: it receives "from prev"                                    */
    /* the inbox of this hand-over's parity: the neighbour may still be unpacking the other one (in the particle-only
     * loop before the time origin nothing else orders its unpack of hand-over n before this put of n + 1)           */
    const int par = (int) (x.seqP & 1ull);
    put_outbox<<<sms, 256, 0, s>>>(x.outbox[0], x.d_cursor + 0, (double*) (x.prev.arena + x.L.inbox_next[par]), &hdr(x.prev.arena)->in_count[par][1], x.L.inbox_cap);
    signal_flag<<<1, 1, 0, s>>>(&hdr(x.prev.arena)->flag[XF_P_NEXT], x.seqP);
    put_outbox<<<sms, 256, 0, s>>>(x.outbox[1], x.d_cursor + 1, (double*) (x.next.arena + x.L.inbox_prev[par]), &hdr(x.next.arena)->in_count[par][0], x.L.inbox_cap);
    signal_flag<<<1, 1, 0, s>>>(&hdr(x.next.arena)->flag[XF_P_PREV], x.seqP);
    *launches += 4;
    XCU(cudaGetLastError());
    return 0;
  }

  /* Migration, second half: wait for both inboxes, read the four counters (the one host synchronisation of a
   * multi-slab field step), close the holes and append the arrivals.  Updates pn.                          */
  static inline int migrate_end (Exchange& x, ParticlesDev P, size_t* pn, size_t pcap, unsigned int* next_id, cudaStream_t s, unsigned long long* launches)
  {
    wait_flag<<<1, 1, 0, s>>>(&hdr(x.arena)->flag[XF_P_PREV], x.seqP, x.d_err, XF_P_PREV, wait_timeout_ns());
    wait_flag<<<1, 1, 0, s>>>(&hdr(x.arena)->flag[XF_P_NEXT], x.seqP, x.d_err, XF_P_NEXT, wait_timeout_ns());
    *launches += 2;
    XCU(cudaMemcpyAsync(x.h_counts + 0, x.d_cursor, 3 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    const int par = (int) (x.seqP & 1ull);
    XCU(cudaMemcpyAsync(x.h_counts + 4, hdr(x.arena)->in_count[par], 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    XCU(cudaMemcpyAsync(x.h_counts + 6, x.d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    XCU(cudaStreamSynchronize(s));
    if (x.h_counts[6])
      {
	static const char* what[8] = { "A from prev", "A from next", "E/B from prev", "E/B from next", "J from prev", "J from next", "particles from prev", "particles from next" };
	const int id = (int) (x.h_counts[6] & 0xff) - 1;
	x.error = std::string("timed out waiting for a neighbouring slab (") + (id >= 0 && id < 8 ? what[id] : "?") + ", sequence " + std::to_string(x.h_counts[6] >> 8) + ")";
	return 1;
      }
    const unsigned int out_prev = x.h_counts[0], out_next = x.h_counts[1], nl = x.h_counts[2];
    const unsigned int in_prev = x.h_counts[4], in_next = x.h_counts[5];
    if (out_prev > x.L.inbox_cap || out_next > x.L.inbox_cap || in_prev > x.L.inbox_cap || in_next > x.L.inbox_cap)
      { x.error = "more particles crossed a slab boundary in one step than the migration buffers hold"; return 1; }
    size_t n = *pn;
    if (nl > 0)
      {
	fill_holes<<<1, 256, 0, s>>>(P, (long) n, (int) nl, x.d_leave, x.d_holes);
	*launches += 1;
	n -= nl;
      }
    if (n + in_prev + in_next > pcap) { x.error = "particle capacity exceeded by arrivals from the neighbouring slabs"; return 1; }
    if (in_prev > 0)
      {
	unpack_inbox<<<(in_prev + 255) / 256, 256, 0, s>>>(P, (long) n, (const double*) (x.arena + x.L.inbox_prev[par]), (int) in_prev, *next_id);
	n += in_prev; *next_id += in_prev; *launches += 1;
      }
    if (in_next > 0)
      {
	unpack_inbox<<<(in_next + 255) / 256, 256, 0, s>>>(P, (long) n, (const double*) (x.arena + x.L.inbox_next[par]), (int) in_next, *next_id);
	n += in_next; *next_id += in_next; *launches += 1;
      }
    XCU(cudaGetLastError());
    *pn = n;
    return 0;
  }
}

#endif
