/* mithra_oracle.c -- CPU restatement of the MITHRA FDTD/PIC time-march.  TEST INFRASTRUCTURE ONLY.
 *
 * See mithra_oracle.h.  Every function cites the reference lines it restates (paths relative to
 * /root/reference).  Index-based loops replace the reference's pointer arithmetic; the floating-point
 * association order is the reference's, so on x86-64 (no FMA contraction) results are bit-identical to the
 * reference build for the field update and the index arithmetic.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "mithra_oracle.h"

#define PI_REF 3.1415926535          /* stdinclude.h:43 (truncated on purpose) */
#define EC_REF 1.602e-19             /* stdinclude.h:49 */
#define EM_REF 9.109e-31             /* stdinclude.h:50 */

struct Oracle
{
  MithraGpuParams p;
  long   P, nodes;
  double *anp1, *an, *anm1;
  double *fnp1, *fn, *fnm1;
  float  *en, *bn;
  unsigned char *pic;
  double *part; size_t npart, pcap;
  double time, timem1, timep1, time_bunch;
  unsigned int n_time, n_time_bunch;
  /* power */
  double *fdt;                /* [Nf][Nz*P][4] */
  double *ep_re, *ep_im;      /* [Nl][Nf] */
  double *rows; size_t nrows, rowcap;
  /* power map (power-visualization) */
  double *mfdt;               /* [Nf][P][4] */
  double *mep_re, *mep_im;    /* [Nf] */
  double *mpL;                /* [P] */
  /* screens */
  double *scr[MITHRA_MAX_SCREENS]; size_t scrn[MITHRA_MAX_SCREENS], scrcap[MITHRA_MAX_SCREENS];
};

/* ---------------------------------------------------------------------------------------------------- */

Oracle* oracle_create (const MithraGpuParams* p)
{
  Oracle* o = (Oracle*) calloc(1, sizeof(Oracle));
  o->p = *p;
  o->P = (long) p->N0 * p->N1;
  o->nodes = o->P * p->np;
  o->anp1 = (double*) calloc((size_t) o->nodes * 3, sizeof(double));
  o->an   = (double*) calloc((size_t) o->nodes * 3, sizeof(double));
  o->anm1 = (double*) calloc((size_t) o->nodes * 3, sizeof(double));
  if (p->space_charge)
    {
      o->fnp1 = (double*) calloc((size_t) o->nodes, sizeof(double));
      o->fn   = (double*) calloc((size_t) o->nodes, sizeof(double));
      o->fnm1 = (double*) calloc((size_t) o->nodes, sizeof(double));
    }
  o->en  = (float*) calloc((size_t) o->nodes * 3, sizeof(float));
  o->bn  = (float*) calloc((size_t) o->nodes * 3, sizeof(float));
  o->pic = (unsigned char*) calloc((size_t) o->nodes, 1);
  o->time = 0.0; o->timem1 = - p->dt; o->timep1 = p->dt; o->time_bunch = 0.0;
  if (p->power.enabled)
    {
      /* radiation.cpp:99, 110-117; here every plane is kept (single-slab oracle) */
      const MithraPower* w = &p->power;
      o->fdt   = (double*) calloc((size_t) w->Nf * w->N * o->P * 4, sizeof(double));
      o->ep_re = (double*) calloc((size_t) w->Nl * w->Nf, sizeof(double));
      o->ep_im = (double*) calloc((size_t) w->Nl * w->Nf, sizeof(double));
      for (int l = 0; l < w->Nl; l++)
	for (int j = 0; j < w->Nf; j++)
	  {
	    o->ep_re[l * w->Nf + j] = cos( w->w[l] * j * p->dt );
	    o->ep_im[l * w->Nf + j] = sin( w->w[l] * j * p->dt );
	  }
    }
  if (p->power_map.enabled)
    {
      /* radiation.cpp:279-314 */
      const MithraPowerMap* w = &p->power_map;
      o->mfdt   = (double*) calloc((size_t) w->Nf * o->P * 4, sizeof(double));
      o->mep_re = (double*) calloc((size_t) w->Nf, sizeof(double));
      o->mep_im = (double*) calloc((size_t) w->Nf, sizeof(double));
      o->mpL    = (double*) calloc((size_t) o->P, sizeof(double));
      for (int j = 0; j < w->Nf; j++)
	{
	  o->mep_re[j] = cos( w->w * j * p->dt );
	  o->mep_im[j] = sin( w->w * j * p->dt );
	}
    }
  return o;
}

void oracle_destroy (Oracle* o)
{
  if (!o) return;
  free(o->anp1); free(o->an); free(o->anm1); free(o->fnp1); free(o->fn); free(o->fnm1);
  free(o->en); free(o->bn); free(o->pic); free(o->part); free(o->fdt); free(o->ep_re); free(o->ep_im); free(o->rows); free(o->mfdt); free(o->mep_re); free(o->mep_im); free(o->mpL);
  for (int s = 0; s < MITHRA_MAX_SCREENS; s++) free(o->scr[s]);
  free(o);
}

double* oracle_anp1 (Oracle* o) { return o->anp1; }
double* oracle_an   (Oracle* o) { return o->an; }
double* oracle_anm1 (Oracle* o) { return o->anm1; }
double* oracle_fnp1 (Oracle* o) { return o->fnp1; }
double* oracle_fn   (Oracle* o) { return o->fn; }
double* oracle_fnm1 (Oracle* o) { return o->fnm1; }
float*  oracle_en   (Oracle* o) { return o->en; }
float*  oracle_bn   (Oracle* o) { return o->bn; }
unsigned char* oracle_pic (Oracle* o) { return o->pic; }

void oracle_set_particles (Oracle* o, const double* aos11, size_t n)
{
  if (n > o->pcap) { o->part = (double*) realloc(o->part, (n + 16) * 11 * sizeof(double)); o->pcap = n + 16; }
  if (n) memcpy(o->part, aos11, n * 11 * sizeof(double));
  o->npart = n;
}
size_t  oracle_num_particles (Oracle* o) { return o->npart; }
double* oracle_particles (Oracle* o) { return o->part; }
void oracle_set_time (Oracle* o, double time, double time_bunch, unsigned int n_time)
{ o->time = time; o->timem1 = time - o->p.dt; o->timep1 = time + o->p.dt; o->time_bunch = time_bunch; o->n_time = n_time; }
double oracle_time (Oracle* o) { return o->time; }
double oracle_time_bunch (Oracle* o) { return o->time_bunch; }

/* ---------------------------------------------------------------------------------------------------- */
/* Signal::self, classes.cpp:534-575 */

static double signal_self (const MithraSignal* g, double t, double phase)
{
  const double d = t - g->t0;
  if (fabs(d) > 10.0 * g->s) return 0.0;
  const double car = cos( 2 * PI_REF * g->f0 * d + g->cep + phase );
  switch (g->type)
    {
    case MITHRA_SIGNAL_NEUMANN:  return - car * 2.7724 * d / ( g->s * g->s ) * exp( -1.3863 * d * d / ( g->s * g->s ) );
    case MITHRA_SIGNAL_GAUSSIAN: { double u = d / g->s; return car * exp( -1.3863 * ( u * u ) ); }
    case MITHRA_SIGNAL_SECANT:   return car / cosh( d / g->s );
    case MITHRA_SIGNAL_FLATTOP:
    case MITHRA_SIGNAL_INVGAUSSIAN:
      {
	double env = 1.0;
	if (g->type == MITHRA_SIGNAL_INVGAUSSIAN)
	  {
	    double u0 = d / g->sigma_inv_g[0], u1 = d / g->sigma_inv_g[1];
	    env = pow( ( 1.0 + u0 * u0 ) * ( 1.0 + u1 * u1 ), 0.25 );
	  }
	if (d <= - g->s / 2.0)    { double u = ( d + g->s / 2.0 ) * g->f0 / g->nR; return (g->type == MITHRA_SIGNAL_FLATTOP) ? car * exp( - ( u * u ) ) : car * env * exp( - ( u * u ) ); }
	else if (d <= g->s / 2.0) { return (g->type == MITHRA_SIGNAL_FLATTOP) ? car : car * env; }
	else                      { double u = ( d - g->s / 2.0 ) * g->f0 / g->nR; return (g->type == MITHRA_SIGNAL_FLATTOP) ? car * exp( - ( u * u ) ) : car * env * exp( - ( u * u ) ); }
      }
    }
  return 0.0;
}

static double dot3 (const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3 (const double* a, const double* b, double* c)
{ c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; }
static double sq (double x) { return x * x; }

/* Seed::fields, classes.cpp:740-855 */
static void seed_fields (const MithraGpuParams* p, double px, double py, double pz, double time, double a[3])
{
  const MithraBeam* s = &p->seed;
  double rl[3], rv[3], yv[3];
  a[0] = a[1] = a[2] = 0.0;
  rl[0] = px; rl[1] = py;
  rl[2] = p->gamma * ( pz + p->beta * p->c0 * ( time + p->dt_shift ) );
  double tl = p->gamma * ( time + p->dt_shift + p->beta / p->c0 * pz );
  for (int c = 0; c < 3; c++) rv[c] = rl[c] - s->position[c];
  const double z = dot3(rv, s->direction);
  tl -= z / p->c0;
  double ph = 0.0;
  double ts = signal_self(&s->signal, tl, ph);
  if (s->seed_type == MITHRA_BEAM_PLANEWAVE)
    {
      if (!(fabs(ts) < 1.0e-6)) for (int c = 0; c < 3; c++) a[c] = s->amplitude * ts * s->polarization[c];
    }
  else if (s->seed_type == MITHRA_BEAM_PLANEWAVETRUNCATED)
    {
      const double x = dot3(rv, s->polarization);
      cross3(s->direction, s->polarization, yv);
      const double y = dot3(rv, yv);
      if (!(fabs(ts) < 1.0e-6 || fabs(x) > s->radius[0] || fabs(y) > s->radius[1]))
	for (int c = 0; c < 3; c++) a[c] = s->amplitude * ts * s->polarization[c];
    }
  else if (s->seed_type == MITHRA_BEAM_GAUSSIAN || s->seed_type == MITHRA_BEAM_SUPERGAUSSIAN)
    {
      if (!(fabs(ts) < 1.0e-6))
	{
	  const double x = dot3(rv, s->polarization);
	  cross3(s->direction, s->polarization, yv);
	  const double y = dot3(rv, yv);
	  const double l = p->c0 / s->signal.f0;
	  const double zRp = PI_REF * s->radius[0] * s->radius[0] / l, wrp = sqrt( 1.0 + z * z / ( zRp * zRp ) );
	  const double zRs = PI_REF * s->radius[1] * s->radius[1] / l, wrs = sqrt( 1.0 + z * z / ( zRs * zRs ) );
	  const int reps = (s->seed_type == MITHRA_BEAM_SUPERGAUSSIAN) ? ( 2 * s->order[0] + 1 ) * ( 2 * s->order[1] + 1 ) : 1;
	  for (int n = 0; n < reps; n++)       /* Q8: the super-gaussian loop ignores x0,y0 and accumulates */
	    {
	      ph = 0.5 * ( atan( z / zRp ) + atan( z / zRs ) - PI_REF ) - PI_REF * z / l * ( sq( x / ( zRp * wrp ) ) + sq( y / ( zRs * wrs ) ) );
	      ts = signal_self(&s->signal, tl, ph);
	      const double t = exp( - sq( x / ( s->radius[0] * wrp ) ) - sq( y / ( s->radius[1] * wrs ) ) ) / sqrt( wrs * wrp ) * s->amplitude;
	      for (int c = 0; c < 3; c++) a[c] += t * ts * s->polarization[c];
	    }
	}
    }
  a[2] *= p->gamma;
}

void oracle_seed_fields (Oracle* o, double x, double y, double z, double time, double a[3]) { seed_fields(&o->p, x, y, z, time, a); }

/* Solver::rc, solver.cpp:2302-2315 */
static void node_coord (const Oracle* o, int i, int j, int k, double r[3])
{
  r[0] = o->p.xmin + i * o->p.dx; r[1] = o->p.ymin + j * o->p.dy; r[2] = o->p.zmin + ( k + o->p.k0 ) * o->p.dz;
}

/* solver.cpp:828-839 */
void oracle_seed_initial (Oracle* o)
{
  const MithraGpuParams* p = &o->p;
  if (!p->seed_enabled) return;
  const int kb = (p->rank == 0) ? 2 : 0, ke = (p->rank == p->size - 1) ? p->np - 2 : p->np;
  for (int i = 2; i < p->N0 - 2; i++)
    for (int j = 2; j < p->N1 - 2; j++)
      for (int k = kb; k < ke; k++)
	{
	  const long m = o->P * k + (long) p->N1 * i + j;
	  double r[3]; node_coord(o, i, j, k, r);
	  seed_fields(p, r[0], r[1], r[2], o->time,   o->an   + 3 * m);
	  seed_fields(p, r[0], r[1], r[2], o->timem1, o->anm1 + 3 * m);
	}
}

/* ---------------------------------------------------------------------------------------------------- */
/* Field update.  The generic "component view": vp/v/vm point at component 0 of node 0 with stride `st`
 * doubles between nodes (3 for A, 1 for phi).                                                           */

typedef struct { double *vp; const double *v, *vm; int st; } View;

#define AT(arr, m) (arr)[(long) (m) * W->st]

/* AdvanceField::advanceBoundaryF/S, database.cpp:137-176 with the argument lists of fdtd.cpp:377-472 decoded:
 * s = face node, n = s + dn inward neighbour, d1/d2 = the two tangential node strides.                     */
static void face (const View* W, long s, long dn, long d1, long d2, const double* B)
{
  const long n = s + dn;
  AT(W->vp, s) = B[0] * ( AT(W->vm, s) + AT(W->vp, n) ) +
		 B[1] * AT(W->vm, n) +
		 B[2] * ( AT(W->v, s) + AT(W->v, n) ) +
		 B[3] * ( AT(W->v, n + d1) + AT(W->v, n - d1) + AT(W->v, s + d1) + AT(W->v, s - d1) ) +
		 B[4] * ( AT(W->v, n + d2) + AT(W->v, n - d2) + AT(W->v, s + d2) + AT(W->v, s - d2) );
}

/* AdvanceField::advanceEdgeF/S, database.cpp:179-229 with the argument lists of fdtd.cpp:480-628 decoded. */
static void edge (const View* W, long e, long du, long dv, long dw, const double* E)
{
  const long u = e + du, v = e + dv, d = e + du + dv;
  AT(W->vp, e) = E[0] * ( AT(W->vp, u) + AT(W->vm, v) ) +
		 E[1] * ( AT(W->vm, u) + AT(W->vp, v) ) +
		 E[2] * ( AT(W->vm, e) + AT(W->vp, d) ) +
		 E[3] * ( AT(W->v, e) + AT(W->v, u) + AT(W->v, v) + AT(W->v, d) ) +
		 E[4] * ( AT(W->v, e - dw) + AT(W->v, u - dw) + AT(W->v, v - dw) + AT(W->v, d - dw) +
			  AT(W->v, e + dw) + AT(W->v, u + dw) + AT(W->v, v + dw) + AT(W->v, d + dw) ) -
		 AT(W->vm, d);
}

/* AdvanceField::advanceCornerF/S, database.cpp:232-288 with fdtd.cpp:631-724. */
static void corner (const View* W, long m, long si, long sj, long sk, const double* h)
{
  const long nb[7] = { m + si, m + sj, m + sk, m + si + sj, m + si + sk, m + sj + sk, m + si + sj + sk };
  double s = AT(W->v, m) * h[16] + AT(W->vm, m) * h[8];
  for (int n = 0; n < 7; n++)
    {
      s = s + AT(W->vp, nb[n]) * h[1 + n];
      s = s + AT(W->v,  nb[n]) * h[16];
      s = s + AT(W->vm, nb[n]) * h[9 + n];
    }
  AT(W->vp, m) = - s / h[0];
}

static void update_component (Oracle* o, const View* W, double asrc)
{
  const MithraGpuParams* p = &o->p;
  const int N0 = p->N0, N1 = p->N1, np = p->np;
  const long P = o->P;
  const double a0 = p->a[0], a1 = p->a[1], a2 = p->a[2], a3 = p->a[3], al = p->alpha, be = p->beta_nsfd;

  /* interior: database.cpp:44-134 through fdtd.cpp:270-303; the source sits in the output array (fdtd.cpp:244) */
  for (int k = 1; k < np - 1; k++)
    for (int i = 1; i < N0 - 1; i++)
      for (int j = 1; j < N1 - 1; j++)
	{
	  const long m = P * k + (long) N1 * i + j;
	  const double src = AT(W->vp, m);
	  if (p->solver == MITHRA_SOLVER_NSFD)
	    AT(W->vp, m) = a0 * AT(W->v, m) - AT(W->vm, m) + al * (
		a1 * ( AT(W->v, m + N1) + AT(W->v, m - N1) + be * ( AT(W->v, m + N1 + P) + AT(W->v, m + N1 - P) + AT(W->v, m - N1 + P) + AT(W->v, m - N1 - P) ) ) +
		a2 * ( AT(W->v, m + 1 ) + AT(W->v, m - 1 ) + be * ( AT(W->v, m + 1  + P) + AT(W->v, m + 1  - P) + AT(W->v, m - 1  + P) + AT(W->v, m - 1  - P) ) ) ) +
		a3 * ( AT(W->v, m + P) + AT(W->v, m - P) ) +
		asrc * src;
	  else
	    AT(W->vp, m) = a0 * AT(W->v, m) - AT(W->vm, m) +
		a1 * ( AT(W->v, m + N1) + AT(W->v, m - N1) ) +
		a2 * ( AT(W->v, m + 1 ) + AT(W->v, m - 1 ) ) +
		a3 * ( AT(W->v, m + P) + AT(W->v, m - P) ) +
		asrc * src;
	}
}

/* TF/SF injection, fdtd.cpp:307-373 (vector potential only). */
static void seed_inject (Oracle* o)
{
  const MithraGpuParams* p = &o->p;
  if (!p->seed_enabled) return;
  const int N0 = p->N0, N1 = p->N1, np = p->np;
  const long P = o->P;
  const int KI = (p->rank == 0) ? 2 : 1, KF = (p->rank == p->size - 1) ? np - 2 : np - 1;
  double r[3], S[3];
  #define INJ(ii, jj, kk, si, sj, sk, coef, sign) do { \
      const long m_ = P * (kk) + (long) N1 * (ii) + (jj); \
      node_coord(o, (ii) + (si), (jj) + (sj), (kk) + (sk), r); seed_fields(p, r[0], r[1], r[2], o->time, S); \
      for (int c = 0; c < 3; c++) o->anp1[3 * m_ + c] = o->anp1[3 * m_ + c] sign (coef) * S[c]; } while (0)
  for (int j = 2; j < N1 - 2; j++)
    for (int k = KI; k < KF; k++)
      {
	INJ(1,      j, k, +1, 0, 0, p->a[1], -);
	INJ(2,      j, k, -1, 0, 0, p->a[1], +);
	INJ(N0 - 2, j, k, -1, 0, 0, p->a[1], -);
	INJ(N0 - 3, j, k, +1, 0, 0, p->a[1], +);
      }
  for (int i = 2; i < N0 - 2; i++)
    for (int k = KI; k < KF; k++)
      {
	INJ(i, 1,      k, 0, +1, 0, p->a[2], -);
	INJ(i, 2,      k, 0, -1, 0, p->a[2], +);
	INJ(i, N1 - 2, k, 0, -1, 0, p->a[2], -);
	INJ(i, N1 - 3, k, 0, +1, 0, p->a[2], +);
      }
  if (p->rank == 0)
    for (int i = 2; i < N0 - 2; i++)
      for (int j = 2; j < N1 - 2; j++)
	{
	  INJ(i, j, 1, 0, 0, +1, p->a[3], -);
	  INJ(i, j, 2, 0, 0, -1, p->a[3], +);
	}
  if (p->rank == p->size - 1)
    for (int i = 2; i < N0 - 2; i++)
      for (int j = 2; j < N1 - 2; j++)
	{
	  INJ(i, j, np - 2, 0, 0, -1, p->a[3], -);
	  INJ(i, j, np - 3, 0, 0, +1, p->a[3], +);
	}
  #undef INJ
}

static void boundaries (Oracle* o, const View* W)
{
  const MithraGpuParams* p = &o->p;
  const int N0 = p->N0, N1 = p->N1, np = p->np;
  const long P = o->P;
  const int zlo = (p->rank == 0), zhi = (p->rank == p->size - 1);

  /* faces, fdtd.cpp:377-472 */
  for (int j = 1; j < N1 - 1; j++) for (int k = 1; k < np - 1; k++) face(W, P * k + j, N1, 1, P, p->bB);
  for (int j = 1; j < N1 - 1; j++) for (int k = 1; k < np - 1; k++) face(W, P * k + (long) N1 * (N0 - 1) + j, -N1, 1, P, p->bB);
  for (int i = 1; i < N0 - 1; i++) for (int k = 1; k < np - 1; k++) face(W, P * k + (long) N1 * i, 1, N1, P, p->cB);
  for (int i = 1; i < N0 - 1; i++) for (int k = 1; k < np - 1; k++) face(W, P * k + (long) N1 * i + N1 - 1, -1, N1, P, p->cB);
  if (zlo) for (int i = 1; i < N0 - 1; i++) for (int j = 1; j < N1 - 1; j++) face(W, (long) N1 * i + j, P, N1, 1, p->dB);
  if (zhi) for (int i = 1; i < N0 - 1; i++) for (int j = 1; j < N1 - 1; j++) face(W, P * (np - 1) + (long) N1 * i + j, -P, N1, 1, p->dB);

  if (p->truncation_order != 2) return;

  /* edges, fdtd.cpp:480-628 */
  for (int k = 1; k < np - 1; k++)
    {
      edge(W, P * k,                                   N1,  1, P, p->eE);
      edge(W, P * k + (long) N1 * (N0 - 1),           -N1,  1, P, p->eE);
      edge(W, P * k + N1 - 1,                          N1, -1, P, p->eE);
      edge(W, P * k + (long) N1 * (N0 - 1) + N1 - 1,  -N1, -1, P, p->eE);
    }
  for (int i = 1; i < N0 - 1; i++)
    {
      if (zlo) { edge(W, (long) N1 * i, 1, P, N1, p->fE); edge(W, (long) N1 * i + N1 - 1, -1, P, N1, p->fE); }
      if (zhi) { edge(W, (long) N1 * i + P * (np - 1), 1, -P, N1, p->fE); edge(W, (long) N1 * i + P * (np - 1) + N1 - 1, -1, -P, N1, p->fE); }
    }
  for (int j = 1; j < N1 - 1; j++)
    {
      if (zlo) { edge(W, j, P, N1, 1, p->gE); edge(W, (long) N1 * (N0 - 1) + j, P, -N1, 1, p->gE); }
      if (zhi) { edge(W, P * (np - 1) + j, -P, N1, 1, p->gE); edge(W, P * (np - 1) + (long) N1 * (N0 - 1) + j, -P, -N1, 1, p->gE); }
    }

  /* corners, fdtd.cpp:631-724 */
  for (int q = 0; q < 8; q++)
    {
      const int ihi = q & 1, jhi = (q >> 1) & 1, khi = (q >> 2) & 1;
      if (!khi && !zlo) continue;
      if ( khi && !zhi) continue;
      const long m = (khi ? P * (np - 1) : 0) + (ihi ? (long) N1 * (N0 - 1) : 0) + (jhi ? N1 - 1 : 0);
      corner(W, m, ihi ? -N1 : N1, jhi ? -1 : 1, khi ? -P : P, p->hC);
    }
}

/* FdTd::fieldEvaluate fdtd.cpp:818-845 / FdTdSC::fieldEvaluate fdtdSC.cpp:1110-1141 */
void oracle_field_evaluate (Oracle* o, long m)
{
  const MithraGpuParams* p = &o->p;
  const long N1 = p->N1, P = o->P;
  const double mdt = - p->dt, dx2 = 2.0 * p->dx, dy2 = 2.0 * p->dy, dz2 = 2.0 * p->dz;
  const double *a = o->an, *ap = o->anp1;
  for (int c = 0; c < 3; c++)
    {
      float e = (float) ( ap[3 * m + c] / mdt );
      e = (float) ( (double) e - a[3 * m + c] / mdt );
      o->en[3 * m + c] = e;
    }
  if (p->space_charge)
    {
      const double* f = o->fn;
      o->en[3 * m    ] = (float) ( (double) o->en[3 * m    ] - ( f[m + N1] - f[m - N1] ) / dx2 );
      o->en[3 * m + 1] = (float) ( (double) o->en[3 * m + 1] - ( f[m + 1 ] - f[m - 1 ] ) / dy2 );
      o->en[3 * m + 2] = (float) ( (double) o->en[3 * m + 2] - ( f[m + P ] - f[m - P ] ) / dz2 );
    }
  #define C(arr, mm, c) (arr)[3 * (mm) + (c)]
  o->bn[3 * m] = (float) ( 0.5 * (
      ( C(a,  m + 1, 2) - C(a,  m - 1, 2) ) / dy2 - ( C(a,  m + P, 1) - C(a,  m - P, 1) ) / dz2 +
      ( C(ap, m + 1, 2) - C(ap, m - 1, 2) ) / dy2 - ( C(ap, m + P, 1) - C(ap, m - P, 1) ) / dz2 ) );
  o->bn[3 * m + 1] = (float) ( 0.5 * (
      ( C(a,  m + P, 0) - C(a,  m - P, 0) ) / dz2 - ( C(a,  m + N1, 2) - C(a,  m - N1, 2) ) / dx2 +
      ( C(ap, m + P, 0) - C(ap, m - P, 0) ) / dz2 - ( C(ap, m + N1, 2) - C(ap, m - N1, 2) ) / dx2 ) );
  o->bn[3 * m + 2] = (float) ( 0.5 * (
      ( C(a,  m + N1, 1) - C(a,  m - N1, 1) ) / dx2 - ( C(a,  m + 1, 0) - C(a,  m - 1, 0) ) / dy2 +
      ( C(ap, m + N1, 1) - C(ap, m - N1, 1) ) / dx2 - ( C(ap, m + 1, 0) - C(ap, m - 1, 0) ) / dy2 ) );
  #undef C
  o->pic[m] = 1;
}

/* FdTd::fieldSample fdtd.cpp:851-913 (FdTdSC::fieldSample likewise): interpolated et, bt, at of one sampling point
 * (moving-frame coordinates, already boosted and filtered like solver.cpp:862-905); out[9] = et[3], bt[3], at[3].       */
void oracle_field_sample (Oracle* o, const double* pos, double* out)
{
  const MithraGpuParams* p = &o->p;
  const long N1 = p->N1, P = o->P;
  double c1;
  const double dxr = modf( ( pos[0] - p->xmin ) / p->dx, &c1 ); const int i = (int) c1;
  const double dyr = modf( ( pos[1] - p->ymin ) / p->dy, &c1 ); const int j = (int) c1;
  const double dzr = modf( ( pos[2] - p->zmin ) / p->dz, &c1 ); const int k = (int) c1;
  const long m = ( k - p->k0 ) * P + i * N1 + j;
  const long node[8] = { m, m + N1, m + 1, m + N1 + 1, m + P, m + P + N1, m + P + 1, m + P + N1 + 1 };
  const double w[8] = {
    ( 1.0 - dxr ) * ( 1.0 - dyr ) * ( 1.0 - dzr ),         dxr   * ( 1.0 - dyr ) * ( 1.0 - dzr ),
    ( 1.0 - dxr ) *         dyr   * ( 1.0 - dzr ),         dxr   *         dyr   * ( 1.0 - dzr ),
    ( 1.0 - dxr ) * ( 1.0 - dyr ) *         dzr,           dxr   * ( 1.0 - dyr ) *         dzr,
    ( 1.0 - dxr ) *         dyr   *         dzr,           dxr   *         dyr   *         dzr };
  for (int q = 0; q < 8; q++) if (!o->pic[node[q]]) oracle_field_evaluate(o, node[q]);
  for (int c = 0; c < 3; c++)
    {
      /* FieldVector::mv then pmv (fieldvector.h:62-77): w * v, then += w * v                                       */
      double et = w[0] * o->en[3 * node[0] + c], bt = w[0] * o->bn[3 * node[0] + c], at = w[0] * o->an[3 * node[0] + c];
      for (int q = 1; q < 8; q++)
	{
	  et += w[q] * o->en[3 * node[q] + c];
	  bt += w[q] * o->bn[3 * node[q] + c];
	  at += w[q] * o->an[3 * node[q] + c];
	}
      out[c] = et; out[3 + c] = bt; out[6 + c] = at;
    }
}

/* FdTd::fieldUpdate fdtd.cpp:231-800 (single slab: no MPI exchange) */
void oracle_field_update (Oracle* o)
{
  const MithraGpuParams* p = &o->p;
  const long P = o->P;
  memset(o->pic, 0, (size_t) o->nodes);                                     /* fdtd.cpp:262-264 */

  View W;
  for (int c = 0; c < 3; c++)
    { W.vp = o->anp1 + c; W.v = o->an + c; W.vm = o->anm1 + c; W.st = 3; update_component(o, &W, p->a[4]); }
  if (p->space_charge)
    { W.vp = o->fnp1; W.v = o->fn; W.vm = o->fnm1; W.st = 1; update_component(o, &W, p->a[5]); }

  seed_inject(o);

  for (int c = 0; c < 3; c++)
    { W.vp = o->anp1 + c; W.v = o->an + c; W.vm = o->anm1 + c; W.st = 3; boundaries(o, &W); }
  if (p->space_charge)
    { W.vp = o->fnp1; W.v = o->fn; W.vm = o->fnm1; W.st = 1; boundaries(o, &W); }

  /* boundary-plane E/B, fdtd.cpp:742-774 */
  for (int i = 1; i < p->N0 - 1; i++)
    for (int j = 1; j < p->N1 - 1; j++)
      {
	long m = P + (long) p->N1 * i + j;
	oracle_field_evaluate(o, m);
	o->pic[m - P] = 1;
	if (p->rank == 0) for (int c = 0; c < 3; c++) { o->en[3 * (m - P) + c] = o->en[3 * m + c]; o->bn[3 * (m - P) + c] = o->bn[3 * m + c]; }
	m = P * ( p->np - 2 ) + (long) p->N1 * i + j;
	oracle_field_evaluate(o, m);
	o->pic[m + P] = 1;
	if (p->rank == p->size - 1) for (int c = 0; c < 3; c++) { o->en[3 * (m + P) + c] = o->en[3 * m + c]; o->bn[3 * (m + P) + c] = o->bn[3 * m + c]; }
      }
}

/* ---------------------------------------------------------------------------------------------------- */
/* Bunch update                                                                                          */

static double pmod (double a, double b) { double x = fmod(a, b); x += ( x < 0.0 ) ? b : 0.0; return x; }   /* stdinclude.cpp:88-93 */

/* beam.cc:79-496; standing-wave quirks as documented in beams.cuh */
static void beam_fields (const MithraBeam* s, double c0, const double* rv, double z, double tl, double tlm, double* eT, double* bT)
{
  double yv[3], ex[3] = { 0, 0, 0 }, by[3] = { 0, 0, 0 }, ez[3] = { 0, 0, 0 }, bz[3] = { 0, 0, 0 };
  double p0 = 0.0;
  for (int c = 0; c < 3; c++) eT[c] = bT[c] = 0.0;
  cross3(s->direction, s->polarization, yv);
  const int t = s->seed_type;
  const int standing = (t == MITHRA_BEAM_STANDINGPLANEWAVE || t == MITHRA_BEAM_STANDINGPLANEWAVETRUNCATED ||
			t == MITHRA_BEAM_STANDINGGAUSSIAN || t == MITHRA_BEAM_STANDINGSUPERGAUSSIAN);

  if (t == MITHRA_BEAM_PLANEWAVE || t == MITHRA_BEAM_PLANEWAVETRUNCATED)
    {
      const double ts = signal_self(&s->signal, tl, p0);
      if (fabs(ts) < 1.0e-6) return;
      if (t == MITHRA_BEAM_PLANEWAVETRUNCATED)
	{ const double x = dot3(rv, s->polarization), y = dot3(rv, yv); if (sq(x / s->radius[0]) + sq(y / s->radius[1]) > 1.0) return; }
      for (int c = 0; c < 3; c++) { eT[c] = s->amplitude * ts * s->polarization[c]; bT[c] = s->amplitude * ts / c0 * yv[c]; }
      return;
    }
  if (t == MITHRA_BEAM_STANDINGPLANEWAVE || t == MITHRA_BEAM_STANDINGPLANEWAVETRUNCATED)
    {
      if (t == MITHRA_BEAM_STANDINGPLANEWAVETRUNCATED)
	{ const double x = dot3(rv, s->polarization), y = dot3(rv, yv); if (sq(x / s->radius[0]) + sq(y / s->radius[1]) > 1.0) return; }
      const double ts = signal_self(&s->signal, tl, p0), tsm = signal_self(&s->signal, tlm, p0);
      const double tse = ts - tsm, tsb = ts + tsm;
      if (fabs(tse) < 1.0e-6 && fabs(tsb) < 1.0e-6) return;
      for (int c = 0; c < 3; c++) { eT[c] = s->amplitude * tse * s->polarization[c]; bT[c] = s->amplitude * tsb / c0 * yv[c]; }
      return;
    }

  const double x = dot3(rv, s->polarization), y = dot3(rv, yv);
  const double wrp = sqrt( 1.0 + z * z / ( s->zR[0] * s->zR[0] ) ), wrs = sqrt( 1.0 + z * z / ( s->zR[1] * s->zR[1] ) );
  const int super = (t == MITHRA_BEAM_SUPERGAUSSIAN || t == MITHRA_BEAM_STANDINGSUPERGAUSSIAN);
  if (!super)
    { if (fabs(x / wrp) > 4.0 * s->radius[0] || fabs(y / wrs) > 4.0 * s->radius[1]) return; }
  else
    { if ( ( fabs(x) - s->order[0] * s->radius[0] ) > 4.0 * s->radius[0] * wrp || ( fabs(y) - s->order[1] * s->radius[1] ) > 4.0 * s->radius[1] * wrs ) return; }
  {
    const double ts = signal_self(&s->signal, tl, p0);
    if (!standing) { if (fabs(ts) < 1.0e-6) return; }
    else { const double tsm = signal_self(&s->signal, tlm, p0); if (fabs(ts - tsm) < 1.0e-6 && fabs(ts + tsm) < 1.0e-6) return; }
  }
  const double atanP = atan( z / s->zR[0] ), atanS = atan( z / s->zR[1] );
  const int oi = super ? s->order[0] : 0, oj = super ? s->order[1] : 0;
  const double af = s->amplitude / sqrt( wrs * wrp );
  for (int i = -oi; i <= oi; i++)
    for (int j = -oj; j <= oj; j++)
      {
	double x0, y0, tt;
	if (!super)
	  {
	    x0 = x / wrp; y0 = y / wrs;
	    p0 = 0.5 * ( atanP + atanS ) - PI_REF * z / s->l * ( sq( x0 / s->zR[0] ) + sq( y0 / s->zR[1] ) );
	    tt = exp( - sq( x0 / s->radius[0] ) - sq( y0 / s->radius[1] ) ) / sqrt( wrs * wrp );
	    tt *= s->amplitude;
	  }
	else
	  {
	    x0 = ( x - i * s->radius[0] ) / wrp; y0 = ( y - j * s->radius[1] ) / wrs;
	    if (fabs(x0) > 4.0 * s->radius[0] || fabs(y0) > 4.0 * s->radius[1]) continue;
	    p0 = 0.5 * ( atanP + atanS ) - PI_REF * z / s->l * ( sq( x0 / s->zR[0] ) + sq( y / s->zR[1] ) );    /* Q5: y, not y0 */
	    tt = af * exp( - sq( x0 / s->radius[0] ) - sq( y0 / s->radius[1] ) );
	  }
	if (!standing)
	  {
	    double ts = signal_self(&s->signal, tl, p0 - PI_REF / 2.0);
	    for (int c = 0; c < 3; c++) { ex[c] += tt * ts * s->polarization[c]; by[c] += tt * ts / c0 * yv[c]; }
	    ts = signal_self(&s->signal, tl, p0 + atanP);
	    for (int c = 0; c < 3; c++) ez[c] += tt * ( - x0 / s->zR[0] ) * ts * s->direction[c];
	    ts = signal_self(&s->signal, tl, p0 + atanS);
	    if (!super) for (int c = 0; c < 3; c++) bz[c] += tt * ( - y0 / s->zR[1] ) / c0 * ts * s->direction[c];
	    else        for (int c = 0; c < 3; c++) bz[c] += tt * ( - y0 / s->zR[1] ) * ts / c0 * s->direction[c];
	  }
	else
	  {
	    double p1 = p0 - PI_REF / 2.0;
	    double ts = signal_self(&s->signal, tl, p1), tsm = signal_self(&s->signal, tlm, p1);
	    if (!super) for (int c = 0; c < 3; c++) { ex[c] += tt * ( ts - tsm ) * s->polarization[c]; by[c] += tt / c0 * ( ts + tsm ) * yv[c]; }
	    else        for (int c = 0; c < 3; c++) { ex[c] += tt * ( ts - tsm ) * s->polarization[c]; by[c] += tt * ( ts + tsm ) / c0 * yv[c]; }
	    p1 = p0 + atanP; ts = signal_self(&s->signal, tl, p1); tsm = signal_self(&s->signal, tlm, -p1);
	    for (int c = 0; c < 3; c++) ez[c] += tt * ( - x0 / s->zR[0] ) * ( ts - tsm ) * s->direction[c];
	    p1 = p0 + atanS; ts = signal_self(&s->signal, tl, p1); tsm = signal_self(&s->signal, tlm, -p1);
	    if (!super) for (int c = 0; c < 3; c++) bz[c] += tt * ( - y0 / s->zR[1] ) / c0 * ( ts - tsm ) * s->direction[c];
	    else        for (int c = 0; c < 3; c++) bz[c] += tt * ( - y0 / s->zR[1] ) * ( ts - tsm ) / c0 * s->direction[c];
	  }
      }
  for (int c = 0; c < 3; c++) { eT[c] = ex[c] + ez[c]; bT[c] = by[c] + bz[c]; }
}

/* Solver::undulatorField + staticUndulator (solver.cpp:1798-1880, beam.cc:14-76) and externalField (:1886-1947) */
static void analytic_fields (const Oracle* o, const double* r, double tb, double* et, double* bt)
{
  const MithraGpuParams* p = &o->p;
  for (int u = 0; u < p->n_undulators; u++)
    {
      const MithraUndulator* U = &p->undulator[u];
      const double b0 = ( U->lu != 0.0 ) ? EM_REF * p->c0 * 2 * PI_REF / U->lu * U->k / EC_REF : 0.0;
      const double ku = ( U->lu != 0.0 ) ? 2 * PI_REF / U->lu : 0.0;
      const double ct = cos( U->theta ), st = sin( U->theta );
      if (U->type == MITHRA_UNDULATOR_STATIC)
	{
	  const double lz = p->gamma * ( r[2] + p->beta * p->c0 * ( tb + p->dt_shift ) ) - U->rb;
	  const double ly = r[0] * ct + r[1] * st;
	  const double len = U->length * U->lu;
	  double d1, bz;
	  if (lz >= 0.0 && lz <= len)
	    {
	      d1 = b0 * cosh( ku * ly ) * sin( ku * lz ) * p->gamma;
	      bz = b0 * sinh( ku * ly ) * cos( ku * lz );
	    }
	  else if (lz < 0.0)
	    {
	      double sz = exp( - sq( ku * lz ) / 2.0 );
	      if (u > 0)
		{
		  const MithraUndulator* V = &p->undulator[u - 1];
		  const double r0 = V->rb + V->length * V->lu - U->rb;
		  if (lz < r0 || r0 == 0.0) sz = 0.0;
		  else sz *= 0.35875 + 0.48829 * cos( PI_REF * lz / r0 ) + 0.14128 * cos( 2.0 * PI_REF * lz / r0 ) + 0.01168 * cos( 3.0 * PI_REF * lz / r0 );
		}
	      d1 = b0 * cosh( ku * ly ) * sz * ku * lz * p->gamma;
	      bz = b0 * sinh( ku * ly ) * sz;
	    }
	  else
	    {
	      const double t0 = lz - len;
	      double sz = exp( - sq( ku * t0 ) / 2.0 );
	      if (u + 1 < p->n_undulators)
		{
		  const MithraUndulator* V = &p->undulator[u + 1];
		  const double r0 = V->rb - U->rb - U->length * U->lu;
		  if (t0 > r0 || r0 == 0.0) sz = 0.0;
		  else sz *= 0.35875 + 0.48829 * cos( PI_REF * t0 / r0 ) + 0.14128 * cos( 2.0 * PI_REF * t0 / r0 ) + 0.01168 * cos( 3.0 * PI_REF * t0 / r0 );
		}
	      d1 = b0 * cosh( ku * ly ) * sz * ku * t0 * p->gamma;
	      bz = b0 * sinh( ku * ly ) * sz;
	    }
	  bt[0] += d1 * ct; bt[1] += d1 * st; bt[2] += bz;
	  d1 *= p->c0 * p->beta;
	  et[1] += d1 * ct; et[0] += - d1 * st; et[2] += 0.0;
	}
      else
	{
	  double rl[3] = { r[0], r[1], p->gamma * ( r[2] + p->beta * p->c0 * ( tb + p->dt_shift ) ) };
	  const double t0 = p->gamma * ( tb + p->dt_shift + p->beta / p->c0 * r[2] );
	  double rv[3], eT[3], bT[3];
	  for (int c = 0; c < 3; c++) rv[c] = rl[c] - U->beam.position[c];
	  const double z = dot3(rv, U->beam.direction);
	  beam_fields(&U->beam, p->c0, rv, z, t0 - z / p->c0, t0 + z / p->c0, eT, bT);
	  bt[0] += p->gamma * ( bT[0] + p->beta / p->c0 * eT[1] );
	  bt[1] += p->gamma * ( bT[1] - p->beta / p->c0 * eT[0] );
	  bt[2] += bT[2];
	  et[0] += p->gamma * ( eT[0] - p->beta * p->c0 * bT[1] );
	  et[1] += p->gamma * ( eT[1] + p->beta * p->c0 * bT[0] );
	  et[2] += eT[2];
	}
    }
  if (p->n_ext_fields > 0)
    {
      double rl[3] = { r[0], r[1], p->gamma * ( r[2] + p->beta * p->c0 * ( tb + p->dt_shift ) ) };
      const double t0 = p->gamma * ( tb + p->dt_shift + p->beta / p->c0 * r[2] );
      for (int u = 0; u < p->n_ext_fields; u++)
	{
	  const MithraBeam* S = &p->ext_field[u];
	  double rv[3], eT[3], bT[3];
	  for (int c = 0; c < 3; c++) rv[c] = rl[c] - S->position[c];
	  const double z = dot3(rv, S->direction);
	  beam_fields(S, p->c0, rv, z, t0 - z / p->c0, t0 + z / p->c0, eT, bT);
	  bt[0] += p->gamma * ( bT[0] + p->beta / p->c0 * eT[1] );
	  bt[1] += p->gamma * ( bT[1] - p->beta / p->c0 * eT[0] );
	  bt[2] += bT[2];
	  et[0] += p->gamma * ( eT[0] - p->beta * p->c0 * bT[1] );
	  et[1] += p->gamma * ( eT[1] + p->beta * p->c0 * bT[0] );
	  et[2] += eT[2];
	}
    }
}

/* One sub-step of Solver::bunchUpdate, solver.cpp:1437-1549 (single slab: no migration). */
static void bunch_substep (Oracle* o, long* cells)
{
  const MithraGpuParams* p = &o->p;
  const long N1 = p->N1, P = o->P;
  for (size_t n = 0; n < o->npart; n++)
    {
      double* q = o->part + 11 * n;
      double* r = q + 1; double* gb = q + 7; double* e = q + 10;
      if (cells) cells[n] = -1;
      const double zr = pmod( r[2] - p->zmin, p->Lz ) + p->zmin;
      if ( ! ( ( zr >= p->zp[0] ) && ( zr < p->zp[1] ) ) ) continue;
      const int b1x = ( r[0] < p->xmax - p->dx && r[0] > p->xmin + p->dx );
      const int b1y = ( r[1] < p->ymax - p->dy && r[1] > p->ymin + p->dy );
      const int b1z = ( r[2] < p->zp[1] && r[2] >= p->zp[0] );
      double et[3] = { 0, 0, 0 }, bt[3] = { 0, 0, 0 };
      analytic_fields(o, r, o->time_bunch, et, bt);
      if (*e == 1.0)
	{
	  if (b1x && b1y && b1z)
	    {
	      double d1;
	      const double dxr = modf( ( r[0] - p->xmin ) / p->dx, &d1 ); const int i = (int) d1;
	      const double dyr = modf( ( r[1] - p->ymin ) / p->dy, &d1 ); const int j = (int) d1;
	      const double dzr = modf( ( r[2] - p->zmin ) / p->dz, &d1 ); const int k = (int) d1;
	      const long m = ( k - p->k0 ) * P + i * N1 + j;
	      if (cells) cells[n] = m;
	      const long off[8] = { 0, N1, 1, N1 + 1, P, P + N1, P + 1, P + N1 + 1 };
	      const double w[8] = {
		( 1.0 - dxr ) * ( 1.0 - dyr ) * ( 1.0 - dzr ),         dxr   * ( 1.0 - dyr ) * ( 1.0 - dzr ),
		( 1.0 - dxr ) *         dyr   * ( 1.0 - dzr ),         dxr   *         dyr   * ( 1.0 - dzr ),
		( 1.0 - dxr ) * ( 1.0 - dyr ) *         dzr,           dxr   * ( 1.0 - dyr ) *         dzr,
		( 1.0 - dxr ) *         dyr   *         dzr,           dxr   *         dyr   *         dzr };
	      if (!cells)
		{
		  for (int v = 0; v < 8; v++) if (!o->pic[m + off[v]]) oracle_field_evaluate(o, m + off[v]);
		  for (int v = 0; v < 8; v++) for (int c = 0; c < 3; c++) et[c] += w[v] * o->en[3 * (m + off[v]) + c];
		  for (int v = 0; v < 8; v++) for (int c = 0; c < 3; c++) bt[c] += w[v] * o->bn[3 * (m + off[v]) + c];
		}
	    }
	}
      else if (p->n_undulators > 0)
	{
	  if (!cells)
	    {
	      const double lz = p->gamma * ( r[2] + p->beta * p->c0 * ( o->time_bunch + p->dt_shift ) );
	      *e = ( lz > - p->undulator[0].dist ) ? 1.0 : 0.0;
	    }
	}
      else if (!cells) *e = 1.0;
      if (cells) continue;

      /* Boris, solver.cpp:1519-1541 */
      double gm[3], gp[3], gl[3], cr[3];
      for (int c = 0; c < 3; c++) gm[c] = gb[c] + p->r1 * et[c];
      cross3(gm, bt, cr);
      const double d1 = sqrt( 1.0 + ( gm[0] * gm[0] + gm[1] * gm[1] + gm[2] * gm[2] ) );
      for (int c = 0; c < 3; c++) gp[c] = p->r2 / d1 * cr[c] + gm[c];
      cross3(gp, bt, cr);
      const double f2 = 2.0 / ( d1 / p->r2 + p->r2 / d1 * ( bt[0] * bt[0] + bt[1] * bt[1] + bt[2] * bt[2] ) );
      for (int c = 0; c < 3; c++) gl[c] = f2 * cr[c] + gm[c];
      for (int c = 0; c < 3; c++) gb[c] = gl[c] + p->r1 * et[c];
      const double f3 = p->dtb / sqrt( 1.0 + ( gb[0] * gb[0] + gb[1] * gb[1] + gb[2] * gb[2] ) );
      for (int c = 0; c < 3; c++) r[c] += f3 * gb[c];
    }
}

void oracle_push_cells (Oracle* o, long* m_out) { bunch_substep(o, m_out); }

/* solver.cpp:1311-1321 */
void oracle_bunch_update (Oracle* o)
{
  for (size_t n = 0; n < o->npart; n++) { double* q = o->part + 11 * n; q[4] = q[1]; q[5] = q[2]; q[6] = q[3]; }
  for (int s = 0; s < o->p.n_update_bunch; s++)
    {
      bunch_substep(o, 0);
      o->time_bunch += o->p.dt_bunch; ++o->n_time_bunch;
    }
}

/* ---------------------------------------------------------------------------------------------------- */
/* Current deposition, fdtd.cpp:38-185 (+ rho fdtdSC.cpp:141-160)                                        */

static void scatter (Oracle* o, long m, double q, const double* mid, const double* jc)
{
  const MithraGpuParams* p = &o->p;
  const long N1 = p->N1, P = o->P;
  double c;
  const double dxp = modf( ( mid[0] - p->xmin ) / p->dx, &c ), dyp = modf( ( mid[1] - p->ymin ) / p->dy, &c ), dzp = modf( ( mid[2] - p->zmin ) / p->dz, &c );
  const double x1 = 1.0 - dxp, x2 = dxp, y1 = 1.0 - dyp, y2 = dyp, z1 = 1.0 - dzp, z2 = dzp;
  const long off[8] = { 0, N1, 1, N1 + 1, P, P + N1, P + 1, P + N1 + 1 };
  double* J = o->anp1;
  /* products are formed left to right exactly as written in the reference: q * 0.5 * w_a * w_b * j */
  J[3 * (m + off[0])    ] += q * 0.5 * y1 * z1 * jc[0]; J[3 * (m + off[1])    ] += q * 0.5 * y1 * z1 * jc[0];
  J[3 * (m + off[2])    ] += q * 0.5 * y2 * z1 * jc[0]; J[3 * (m + off[3])    ] += q * 0.5 * y2 * z1 * jc[0];
  J[3 * (m + off[4])    ] += q * 0.5 * y1 * z2 * jc[0]; J[3 * (m + off[5])    ] += q * 0.5 * y1 * z2 * jc[0];
  J[3 * (m + off[6])    ] += q * 0.5 * y2 * z2 * jc[0]; J[3 * (m + off[7])    ] += q * 0.5 * y2 * z2 * jc[0];
  J[3 * (m + off[0]) + 1] += q * 0.5 * x1 * z1 * jc[1]; J[3 * (m + off[1]) + 1] += q * 0.5 * x2 * z1 * jc[1];
  J[3 * (m + off[2]) + 1] += q * 0.5 * x1 * z1 * jc[1]; J[3 * (m + off[3]) + 1] += q * 0.5 * x2 * z1 * jc[1];
  J[3 * (m + off[4]) + 1] += q * 0.5 * x1 * z2 * jc[1]; J[3 * (m + off[5]) + 1] += q * 0.5 * x2 * z2 * jc[1];
  J[3 * (m + off[6]) + 1] += q * 0.5 * x1 * z2 * jc[1]; J[3 * (m + off[7]) + 1] += q * 0.5 * x2 * z2 * jc[1];
  J[3 * (m + off[0]) + 2] += q * 0.5 * x1 * y1 * jc[2]; J[3 * (m + off[1]) + 2] += q * 0.5 * x2 * y1 * jc[2];
  J[3 * (m + off[2]) + 2] += q * 0.5 * x1 * y2 * jc[2]; J[3 * (m + off[3]) + 2] += q * 0.5 * x2 * y2 * jc[2];
  J[3 * (m + off[4]) + 2] += q * 0.5 * x1 * y1 * jc[2]; J[3 * (m + off[5]) + 2] += q * 0.5 * x2 * y1 * jc[2];
  J[3 * (m + off[6]) + 2] += q * 0.5 * x1 * y2 * jc[2]; J[3 * (m + off[7]) + 2] += q * 0.5 * x2 * y2 * jc[2];
}

static int deposit_flags (const MithraGpuParams* p, const double* r)
{
  return ( r[0] < p->xmax - p->dx && r[0] > p->xmin + p->dx && r[1] < p->ymax - p->dy && r[1] > p->ymin + p->dy &&
	   r[2] < p->zp[1] && r[2] >= p->zp[0] );
}

void oracle_current_update (Oracle* o)
{
  const MithraGpuParams* p = &o->p;
  const long N1 = p->N1, P = o->P;
  const double d[3] = { p->dx, p->dy, p->dz }, mn[3] = { p->xmin, p->ymin, p->zmin };
  for (size_t n = 0; n < o->npart; n++)
    {
      const double* q = o->part + 11 * n;
      const double* rp = q + 1; const double* rm = q + 4;
      const int bp = deposit_flags(p, rp), bm = deposit_flags(p, rm);
      if (!(bp || bm)) continue;
      int ip[3], im[3]; double r[3], jcp[3], jcm[3], mid[3];
      for (int c = 0; c < 3; c++)
	{
	  ip[c] = (int) floor( ( rp[c] - mn[c] ) / d[c] );
	  im[c] = (int) floor( ( rm[c] - mn[c] ) / d[c] );
	  const int lo = im[c] < ip[c] ? im[c] : ip[c], hi = im[c] > ip[c] ? im[c] : ip[c];
	  r[c] = fmin( lo * d[c] + d[c] + mn[c], fmax( hi * d[c] + mn[c], 0.5 * ( rm[c] + rp[c] ) ) );
	  jcm[c] = r[c] - rm[c];
	  jcp[c] = rp[c] - r[c];
	}
      if (bp)
	{
	  const long m = P * ( ip[2] - p->k0 ) + N1 * ip[0] + ip[1];
	  for (int c = 0; c < 3; c++) mid[c] = 0.5 * ( rp[c] + r[c] );
	  scatter(o, m, q[0], mid, jcp);
	  if (p->space_charge)
	    {
	      double cc;
	      const double dxp = modf( ( rp[0] - p->xmin ) / p->dx, &cc ), dyp = modf( ( rp[1] - p->ymin ) / p->dy, &cc ), dzp = modf( ( rp[2] - p->zmin ) / p->dz, &cc );
	      const double x1 = 1.0 - dxp, x2 = dxp, y1 = 1.0 - dyp, y2 = dyp, z1 = 1.0 - dzp, z2 = dzp;
	      double* R = o->fnp1 + m;
	      R[0]          += q[0] * x1 * y1 * z1; R[N1]         += q[0] * x2 * y1 * z1;
	      R[1]          += q[0] * x1 * y2 * z1; R[N1 + 1]     += q[0] * x2 * y2 * z1;
	      R[P]          += q[0] * x1 * y1 * z2; R[P + N1]     += q[0] * x2 * y1 * z2;
	      R[P + 1]      += q[0] * x1 * y2 * z2; R[P + N1 + 1] += q[0] * x2 * y2 * z2;
	    }
	}
      if (bm)
	{
	  const long m = P * ( im[2] - p->k0 ) + N1 * im[0] + im[1];
	  for (int c = 0; c < 3; c++) mid[c] = 0.5 * ( rm[c] + r[c] );
	  scatter(o, m, q[0], mid, jcm);
	}
    }
}

void oracle_deposit_cells (Oracle* o, int* out)
{
  const MithraGpuParams* p = &o->p;
  const double d[3] = { p->dx, p->dy, p->dz }, mn[3] = { p->xmin, p->ymin, p->zmin };
  for (size_t n = 0; n < o->npart; n++)
    {
      const double* q = o->part + 11 * n;
      for (int c = 0; c < 3; c++)
	{
	  out[6 * n + c]     = (int) floor( ( q[1 + c] - mn[c] ) / d[c] );
	  out[6 * n + 3 + c] = (int) floor( ( q[4 + c] - mn[c] ) / d[c] );
	}
    }
}

void oracle_field_shift (Oracle* o)          /* fdtd.cpp:806-812, fdtdSC.cpp:1093-1104 */
{
  double* t = o->anm1; o->anm1 = o->an; o->an = o->anp1; o->anp1 = t;
  if (o->p.space_charge) { t = o->fnm1; o->fnm1 = o->fn; o->fn = o->fnp1; o->fnp1 = t; }
}

void oracle_current_reset (Oracle* o)        /* fdtd.cpp:23-32, fdtdSC.cpp:23-36 */
{
  memset(o->anp1, 0, (size_t) o->nodes * 3 * sizeof(double));
  if (o->p.space_charge) memset(o->fnp1, 0, (size_t) o->nodes * sizeof(double));
}

/* ---------------------------------------------------------------------------------------------------- */
/* Radiated power, radiation.cpp:127-232                                                                 */

void oracle_power_sample (Oracle* o)
{
  const MithraGpuParams* p = &o->p;
  const MithraPower* w = &p->power;
  if (!w->enabled) return;
  const long N1 = p->N1, P = o->P;
  const size_t width = (size_t) w->N * w->Nl;
  if (o->nrows == o->rowcap) { o->rowcap = o->rowcap ? 2 * o->rowcap : 1024; o->rows = (double*) realloc(o->rows, o->rowcap * width * sizeof(double)); }
  double* pL = o->rows + o->nrows * width;
  for (size_t t = 0; t < width; t++) pL[t] = 0.0;
  const unsigned int slot = o->n_time % (unsigned int) w->Nf;
  int kz = 0;
  for (int k = 0; k < w->N; k++)
    {
      if ( !( w->z[k] < p->zp[1] && w->z[k] >= p->zp[0] ) ) continue;
      double c;
      const double dzr = modf( ( w->z[k] - p->zmin ) / p->dz, &c );
      const int kk = (int) c;
      for (int i = 2; i < p->N0 - 2; i++)
	for (int j = 2; j < p->N1 - 2; j++)
	  {
	    const long mi = ( kk - p->k0 ) * P + i * N1 + j;
	    const long ni = kz * P + i * N1 + j;
	    if (!o->pic[mi])     oracle_field_evaluate(o, mi);
	    if (!o->pic[mi + P]) oracle_field_evaluate(o, mi + P);
	    const double et0 = ( 1.0 - dzr ) * o->en[3 * mi]     + dzr * o->en[3 * (mi + P)];
	    const double et1 = ( 1.0 - dzr ) * o->en[3 * mi + 1] + dzr * o->en[3 * (mi + P) + 1];
	    const double bt0 = ( 1.0 - dzr ) * o->bn[3 * mi]     + dzr * o->bn[3 * (mi + P)];
	    const double bt1 = ( 1.0 - dzr ) * o->bn[3 * mi + 1] + dzr * o->bn[3 * (mi + P) + 1];
	    double* f = o->fdt + ( (size_t) slot * w->N * P + ni ) * 4;
	    f[0] = p->gamma * ( et0 + p->c0 * p->beta * bt1 );
	    f[1] = p->gamma * ( et1 - p->c0 * p->beta * bt0 );
	    f[2] = p->gamma * ( bt0 - p->beta / p->c0 * et1 );
	    f[3] = p->gamma * ( bt1 + p->beta / p->c0 * et0 );
	    for (int l = 0; l < w->Nl; l++)
	      {
		double e1r = 0, e1i = 0, b1r = 0, b1i = 0, e2r = 0, e2i = 0, b2r = 0, b2i = 0;
		for (int m = 0; m < w->Nf; m++)
		  {
		    const double* g = o->fdt + ( (size_t) m * w->N * P + ni ) * 4;
		    const double cr = o->ep_re[l * w->Nf + m], ci = o->ep_im[l * w->Nf + m];
		    e1r += g[0] * cr; e1i += g[0] * ci;
		    b1r += g[3] * cr; b1i += g[3] * ( - ci );
		    e2r += g[1] * cr; e2i += g[1] * ci;
		    b2r += g[2] * cr; b2i += g[2] * ( - ci );
		  }
		pL[k * w->Nl + l] += w->pc * ( ( e1r * b1r - e1i * b1i ) - ( e2r * b2r - e2i * b2i ) );
	      }
	  }
      kz += 1;
    }
  o->nrows++;
}

/* Per-pixel power map, Solver::powerVisualize radiation.cpp:324-391 (the .vts writer :393-447 is the host's)  */
void oracle_power_visualize (Oracle* o)
{
  const MithraGpuParams* p = &o->p;
  const MithraPowerMap* w = &p->power_map;
  if (!w->enabled) return;
  if ( !( w->z < p->zp[1] && w->z >= p->zp[0] ) ) return;                /* rp_.Nz == 1, radiation.cpp:271 */
  const long N1 = p->N1, P = o->P;
  double c;
  const double dzr = modf( ( w->z - p->zmin ) / p->dz, &c );
  const int kk = (int) c - p->k0;
  const unsigned int slot = o->n_time % (unsigned int) w->Nf;
  for (int i = 1; i < p->N0 - 1; i++)
    for (int j = 1; j < p->N1 - 1; j++)
      {
	const long mi = kk * P + i * N1 + j;
	const long ni = i * N1 + j;
	if (!o->pic[mi])     oracle_field_evaluate(o, mi);
	if (!o->pic[mi + P]) oracle_field_evaluate(o, mi + P);
	const double et0 = ( 1.0 - dzr ) * o->en[3 * mi]     + dzr * o->en[3 * (mi + P)];
	const double et1 = ( 1.0 - dzr ) * o->en[3 * mi + 1] + dzr * o->en[3 * (mi + P) + 1];
	const double bt0 = ( 1.0 - dzr ) * o->bn[3 * mi]     + dzr * o->bn[3 * (mi + P)];
	const double bt1 = ( 1.0 - dzr ) * o->bn[3 * mi + 1] + dzr * o->bn[3 * (mi + P) + 1];
	double* f = o->mfdt + ( (size_t) slot * P + ni ) * 4;
	f[0] = p->gamma * ( et0 + p->c0 * p->beta * bt1 );
	f[1] = p->gamma * ( et1 - p->c0 * p->beta * bt0 );
	f[2] = p->gamma * ( bt0 - p->beta / p->c0 * et1 );
	f[3] = p->gamma * ( bt1 + p->beta / p->c0 * et0 );
	double e1r = 0, e1i = 0, b1r = 0, b1i = 0, e2r = 0, e2i = 0, b2r = 0, b2i = 0;
	for (int m = 0; m < w->Nf; m++)
	  {
	    const double* g = o->mfdt + ( (size_t) m * P + ni ) * 4;
	    const double cr = o->mep_re[m], ci = o->mep_im[m];
	    e1r += g[0] * cr; e1i += g[0] * ci;
	    b1r += g[3] * cr; b1i += g[3] * ( - ci );
	    e2r += g[1] * cr; e2i += g[1] * ci;
	    b2r += g[2] * cr; b2i += g[2] * ( - ci );
	  }
	o->mpL[ni] = w->pc * ( ( e1r * b1r - e1i * b1i ) - ( e2r * b2r - e2i * b2i ) );
      }
}

const double* oracle_power_map (Oracle* o) { return o->mpL; }

size_t        oracle_power_rows (Oracle* o) { return o->nrows; }
const double* oracle_power_data (Oracle* o) { return o->rows; }

/* ---------------------------------------------------------------------------------------------------- */
/* Screens, solver.cpp:2205-2257                                                                         */

void oracle_screen_profile (Oracle* o)
{
  const MithraGpuParams* p = &o->p;
  if (!p->screens.enabled) return;
  for (int s = 0; s < p->screens.N; s++)
    {
      const double lzs = p->screens.pos[s];
      for (size_t n = 0; n < o->npart; n++)
	{
	  const double* q = o->part + 11 * n;
	  const double* rp = q + 1; const double* rm = q + 4; const double* gb = q + 7;
	  const double zr = pmod( rp[2] - p->zmin, p->Lz ) + p->zmin;
	  if ( ! ( ( zr >= p->zp[0] ) && ( zr < p->zp[1] ) ) ) continue;
	  const double lzm = p->gamma * ( rm[2] + p->beta * p->c0 * ( o->time_bunch - p->dt + p->dt_shift ) );
	  if (lzm >= lzs) continue;
	  const double lzp = p->gamma * ( rp[2] + p->beta * p->c0 * ( o->time_bunch + p->dt_shift ) );
	  if (lzp <  lzs) continue;
	  if (o->scrn[s] == o->scrcap[s]) { o->scrcap[s] = o->scrcap[s] ? 2 * o->scrcap[s] : 256; o->scr[s] = (double*) realloc(o->scr[s], o->scrcap[s] * 6 * sizeof(double)); }
	  double* rec = o->scr[s] + 6 * o->scrn[s]++;
	  rec[0] = rm[0] + ( lzs - lzm ) / ( lzp - lzm ) * ( rp[0] - rm[0] );
	  rec[1] = rm[1] + ( lzs - lzm ) / ( lzp - lzm ) * ( rp[1] - rm[1] );
	  const double tm = p->gamma * ( o->time_bunch + p->dt_shift - p->dt + p->beta / p->c0 * rm[2] );
	  const double tp = p->gamma * ( o->time_bunch + p->dt_shift         + p->beta / p->c0 * rp[2] );
	  rec[2] = tm + ( lzs - lzm ) / ( lzp - lzm ) * ( tp - tm );
	  rec[3] = gb[0]; rec[4] = gb[1];
	  rec[5] = p->gamma * ( gb[2] + p->beta * sqrt( 1.0 + ( gb[0] * gb[0] + gb[1] * gb[1] + gb[2] * gb[2] ) ) );
	}
    }
}

size_t        oracle_screen_count (Oracle* o, int s) { return o->scrn[s]; }
const double* oracle_screen_data  (Oracle* o, int s) { return o->scr[s]; }

/* ---------------------------------------------------------------------------------------------------- */

void oracle_advance_time (Oracle* o)         /* solver.cpp:1396-1399 */
{ o->timem1 += o->p.dt; o->time += o->p.dt; o->timep1 += o->p.dt; ++o->n_time; }

/* body of the second while loop of Solver::solve, solver.cpp:1300-1399 */
void oracle_step (Oracle* o, int nsteps)
{
  for (int s = 0; s < nsteps; s++)
    {
      oracle_field_update(o);
      oracle_bunch_update(o);
      oracle_screen_profile(o);
      oracle_power_sample(o);
      oracle_power_visualize(o);
      oracle_field_shift(o);
      oracle_current_reset(o);
      oracle_current_update(o);
      oracle_advance_time(o);
    }
}
