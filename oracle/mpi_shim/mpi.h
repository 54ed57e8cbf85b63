/* mini-MPI: a header-only, fork()-based stand-in for the 25 MPI symbols the MITHRA reference uses.
 *
 * TEST INFRASTRUCTURE ONLY (lives under oracle/): it exists so that the UNMODIFIED reference sources under
 * /root/reference/src can be compiled here (the image has no mpic++ / mpi.h) into oracle/_ref/ and used
 *   (a) as the parity oracle (single rank), and
 *   (b) as the CPU baseline on all host cores (MINIMPI_NP=<ranks>).
 *
 * Semantics implemented (only what the reference needs, see SURVEY.md section 2.1 for the call sites):
 *   - MPI_Init forks MINIMPI_NP-1 children (default 1 rank => no fork); every ordered pair of ranks is
 *     connected by a pipe created before the fork; messages are {tag, nbytes, payload}.
 *   - Datatypes are encoded as their byte size (MPI_Type_contiguous(n,t) -> n*t).
 *   - Send is blocking-buffered (pipe capacity raised to 1 MiB); a rank sending to itself queues locally,
 *     because the reference sends migrating particles to itself when size == 1 (solver.cpp:1552-1564).
 *   - Recv/Probe match on (source, tag); non-matching messages are parked in a local queue.
 *   - Bcast / Reduce / Allreduce / Barrier are linear algorithms over the same pipes (reserved tags).
 *   - Reduce ops: SUM / MIN / MAX over double, float, int (by datatype size: 8 -> double, 4 -> int unless
 *     flagged float; the reference only reduces MPI_DOUBLE and MPI_INT).
 */
#ifndef MINIMPI_H_
#define MINIMPI_H_

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>
#include <unistd.h>
#include <fcntl.h>
#include <sys/wait.h>
#include <sys/types.h>

typedef int MPI_Datatype;
typedef int MPI_Comm;
typedef int MPI_Op;

struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int nbytes; };

#define MPI_COMM_WORLD 0
#define MPI_DOUBLE     8
#define MPI_FLOAT      (-4)   /* negative marks "floating 4-byte"; size is |t| */
#define MPI_INT        4
#define MPI_SUM        1
#define MPI_MIN        2
#define MPI_MAX        3
#define MPI_IN_PLACE   ((void*)(-1))
#define MPI_SUCCESS    0

namespace minimpi
{
  struct Msg { int tag; std::vector<char> data; };

  struct State
  {
    int rank, size;
    std::vector<int> rfd;                     /* rfd[src]  : read end of pipe src -> me                   */
    std::vector<int> wfd;                     /* wfd[dst]  : write end of pipe me -> dst                  */
    std::vector<std::deque<Msg> > parked;     /* parked[src]: messages read but not yet matched           */
    std::vector<pid_t> children;
    State () : rank(0), size(1) {}
  };

  inline State& st () { static State s; return s; }

  inline int tsize (MPI_Datatype t) { return t < 0 ? -t : t; }

  inline void xwrite (int fd, const void* p, size_t n)
  {
    const char* c = (const char*) p;
    while (n > 0) { ssize_t w = ::write(fd, c, n); if (w <= 0) { perror("minimpi write"); _exit(3); } c += w; n -= (size_t) w; }
  }

  inline void xread (int fd, void* p, size_t n)
  {
    char* c = (char*) p;
    while (n > 0) { ssize_t r = ::read(fd, c, n); if (r <= 0) { if (r < 0) perror("minimpi read"); _exit(4); } c += r; n -= (size_t) r; }
  }

  /* Pull the next message of the pipe src -> me into the parked queue. */
  inline void pull (int src)
  {
    State& s = st();
    int hdr[2];
    xread(s.rfd[src], hdr, sizeof(hdr));
    Msg m; m.tag = hdr[0]; m.data.resize((size_t) hdr[1]);
    if (hdr[1] > 0) xread(s.rfd[src], &m.data[0], (size_t) hdr[1]);
    s.parked[src].push_back(m);
  }

  /* Return the index in parked[src] of the first message with this tag, reading the pipe as needed. */
  inline size_t match (int src, int tag)
  {
    State& s = st();
    size_t i = 0;
    for (;;)
      {
	for (; i < s.parked[src].size(); i++)
	  if (s.parked[src][i].tag == tag) return i;
	if (src == s.rank) { fprintf(stderr, "minimpi: rank %d waits on itself for tag %d\n", s.rank, tag); _exit(5); }
	pull(src);
      }
  }

  inline void send (const void* buf, int nbytes, int dst, int tag)
  {
    State& s = st();
    if (dst == s.rank)
      {
	Msg m; m.tag = tag; m.data.assign((const char*) buf, (const char*) buf + nbytes);
	s.parked[dst].push_back(m);
	return;
      }
    int hdr[2] = { tag, nbytes };
    xwrite(s.wfd[dst], hdr, sizeof(hdr));
    if (nbytes > 0) xwrite(s.wfd[dst], buf, (size_t) nbytes);
  }

  inline int recv (void* buf, int maxbytes, int src, int tag)
  {
    State& s = st();
    size_t i = match(src, tag);
    Msg& m = s.parked[src][i];
    int n = (int) m.data.size();
    if (n > maxbytes) { fprintf(stderr, "minimpi: message truncated (%d > %d)\n", n, maxbytes); _exit(6); }
    if (n > 0) memcpy(buf, &m.data[0], (size_t) n);
    s.parked[src].erase(s.parked[src].begin() + (long) i);
    return n;
  }

  template <typename T>
  inline void combine (T* acc, const T* in, int count, MPI_Op op)
  {
    for (int i = 0; i < count; i++)
      {
	if      (op == MPI_SUM) acc[i] += in[i];
	else if (op == MPI_MIN) acc[i] = (in[i] < acc[i]) ? in[i] : acc[i];
	else if (op == MPI_MAX) acc[i] = (in[i] > acc[i]) ? in[i] : acc[i];
      }
  }

  inline void combineBytes (void* acc, const void* in, int count, MPI_Datatype t, MPI_Op op)
  {
    if      (t == MPI_DOUBLE) combine((double*) acc, (const double*) in, count, op);
    else if (t == MPI_FLOAT)  combine((float*)  acc, (const float*)  in, count, op);
    else if (t == MPI_INT)    combine((int*)    acc, (const int*)    in, count, op);
    else { fprintf(stderr, "minimpi: reduce on unsupported datatype %d\n", t); _exit(7); }
  }

  const int TAG_BCAST = 1000001, TAG_REDUCE = 1000002, TAG_BARRIER = 1000003;
}

inline int MPI_Init (int*, char***)
{
  using namespace minimpi;
  State& s = st();
  const char* e = getenv("MINIMPI_NP");
  int np = e ? atoi(e) : 1;
  if (np < 1) np = 1;
  s.size = np; s.rank = 0;
  s.parked.resize((size_t) np);
  s.rfd.assign((size_t) np, -1); s.wfd.assign((size_t) np, -1);
  if (np == 1) return MPI_SUCCESS;

  /* pipes[src][dst] */
  std::vector<std::vector<int> > pr((size_t) np, std::vector<int>((size_t) np, -1)), pw = pr;
  for (int a = 0; a < np; a++)
    for (int b = 0; b < np; b++)
      {
	if (a == b) continue;
	int fd[2];
	if (pipe(fd) != 0) { perror("minimpi pipe"); exit(2); }
#ifdef F_SETPIPE_SZ
	fcntl(fd[1], F_SETPIPE_SZ, 1 << 20);
#endif
	pr[a][b] = fd[0]; pw[a][b] = fd[1];
      }
  fflush(stdout); fflush(stderr);
  for (int r = 1; r < np; r++)
    {
      pid_t pid = fork();
      if (pid < 0) { perror("minimpi fork"); exit(2); }
      if (pid == 0) { s.rank = r; s.children.clear(); break; }
      s.children.push_back(pid);
    }
  for (int a = 0; a < np; a++)
    for (int b = 0; b < np; b++)
      {
	if (a == b) continue;
	if (b == s.rank) s.rfd[a] = pr[a][b]; else close(pr[a][b]);
	if (a == s.rank) s.wfd[b] = pw[a][b]; else close(pw[a][b]);
      }
  return MPI_SUCCESS;
}

inline int MPI_Finalize ()
{
  using namespace minimpi;
  State& s = st();
  fflush(stdout); fflush(stderr);
  if (s.size > 1 && s.rank != 0) _exit(0);
  for (size_t i = 0; i < s.children.size(); i++) { int status; waitpid(s.children[i], &status, 0); }
  return MPI_SUCCESS;
}

inline int MPI_Comm_rank (MPI_Comm, int* r) { *r = minimpi::st().rank; return MPI_SUCCESS; }
inline int MPI_Comm_size (MPI_Comm, int* n) { *n = minimpi::st().size; return MPI_SUCCESS; }

inline int MPI_Type_contiguous (int n, MPI_Datatype t, MPI_Datatype* out) { *out = n * minimpi::tsize(t); return MPI_SUCCESS; }
inline int MPI_Type_commit (MPI_Datatype*) { return MPI_SUCCESS; }

inline int MPI_Send (const void* buf, int count, MPI_Datatype t, int dst, int tag, MPI_Comm)
{ minimpi::send(buf, count * minimpi::tsize(t), dst, tag); return MPI_SUCCESS; }

inline int MPI_Recv (void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Status* status)
{
  int n = minimpi::recv(buf, count * minimpi::tsize(t), src, tag);
  if (status) { status->MPI_SOURCE = src; status->MPI_TAG = tag; status->nbytes = n; }
  return MPI_SUCCESS;
}

inline int MPI_Probe (int src, int tag, MPI_Comm, MPI_Status* status)
{
  size_t i = minimpi::match(src, tag);
  status->MPI_SOURCE = src; status->MPI_TAG = tag; status->nbytes = (int) minimpi::st().parked[src][i].data.size();
  return MPI_SUCCESS;
}

inline int MPI_Get_count (const MPI_Status* status, MPI_Datatype t, int* count)
{ *count = status->nbytes / minimpi::tsize(t); return MPI_SUCCESS; }

inline int MPI_Bcast (void* buf, int count, MPI_Datatype t, int root, MPI_Comm)
{
  using namespace minimpi;
  State& s = st();
  int nbytes = count * tsize(t);
  if (s.size == 1) return MPI_SUCCESS;
  if (s.rank == root) { for (int r = 0; r < s.size; r++) if (r != root) send(buf, nbytes, r, TAG_BCAST); }
  else recv(buf, nbytes, root, TAG_BCAST);
  return MPI_SUCCESS;
}

inline int MPI_Reduce (const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm)
{
  using namespace minimpi;
  State& s = st();
  int nbytes = count * tsize(t);
  if (s.rank == root)
    {
      if (sendbuf != MPI_IN_PLACE) memcpy(recvbuf, sendbuf, (size_t) nbytes);
      std::vector<char> tmp((size_t) nbytes);
      for (int r = 0; r < s.size; r++)
	{
	  if (r == root) continue;
	  recv(&tmp[0], nbytes, r, TAG_REDUCE);
	  combineBytes(recvbuf, &tmp[0], count, t, op);
	}
    }
  else
    send(sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf, nbytes, root, TAG_REDUCE);
  return MPI_SUCCESS;
}

inline int MPI_Allreduce (const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  MPI_Reduce(sendbuf, recvbuf, count, t, op, 0, c);
  MPI_Bcast(recvbuf, count, t, 0, c);
  return MPI_SUCCESS;
}

inline int MPI_Barrier (MPI_Comm c)
{
  int x = 0, y = 0;
  MPI_Allreduce(&x, &y, 1, MPI_INT, MPI_SUM, c);
  return MPI_SUCCESS;
}

#endif
