/* beams.cuh -- analytic fields evaluated per particle per push (and the seed potential on the TF/SF shell).
 *
 * Device restatement of the reference's beam.cc (staticUndulator :14-76, planeWave ... standingSuperGaussianBeam
 * :79-496), Signal::self (classes.cpp:534-575) and Seed::fields (classes.cpp:740-855).  Operation order and
 * the truncated constant PI = 3.1415926535 (stdinclude.h:43) are kept; transcendental results differ from
 * glibc in the last ulp (CUDA libdevice), which is the only source of non-bitwise particle state.
 *
 * Reference quirks kept: the super-gaussian phase uses y instead of y0 (beam.cc:253);
 * the seed SUPERGAUSSIAN ignores x0,y0 and accumulates (classes.cpp:837-849).  Where the reference reads an
 * uninitialised variable (tlm for standing-wave *undulators*, solver.cpp:1837; ubp.l in
 * standingSuperGaussianBeam, beam.cc:466) the defined value is used instead (t0 + z/c0, the beam's own l).
 */
#ifndef MITHRA_BEAMS_CUH_
#define MITHRA_BEAMS_CUH_

#include "device_types.cuh"

namespace mithra
{
  #define MITHRA_PI 3.1415926535

  struct V3 { double x, y, z; };
  __device__ __forceinline__ V3     v3   (double a, double b, double c) { V3 r; r.x = a; r.y = b; r.z = c; return r; }
  __device__ __forceinline__ V3     v3a  (const double* p) { return v3(p[0], p[1], p[2]); }
  __device__ __forceinline__ double dot3 (const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
  /* cross product in the reference's order (fieldvector.h cross()) */
  __device__ __forceinline__ V3 cross3 (const V3& a, const V3& b)
  { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
  __device__ __forceinline__ V3 scale3 (double s, const V3& a) { return v3(s * a.x, s * a.y, s * a.z); }
  __device__ __forceinline__ V3 add3   (const V3& a, const V3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }

  /* cos() for the carrier phase 2 pi f0 (t - t0) + cep + phase, which reaches 1e7 rad for an ultraviolet seed.
   * CUDA's cos switches to a Payne-Hanek slow path above 1.05e5 rad; a three-constant Cody-Waite reduction by
   * 2 pi with explicit FMAs (exact products) reduces |x| < 1e9 to [-pi, pi] with an absolute error below 4e-16,
   * i.e. the same correctly-reduced argument to the last bit or two, at a fraction of the cost.             */
  __device__ __forceinline__ double cos_wide (double x)
  {
    const double ax = fabs(x);
    if (ax < 1.0e5 || !(ax < 1.0e9)) return cos(x);
    const double k = rint(x * 0.15915494309189535);
    double r = __fma_rn(-k, 6.283185307179586232, x);
    r = __fma_rn(-k, 2.4492935982947064e-16, r);
    r = __fma_rn(-k, -5.9895396194366794e-33, r);
    return cos(r);
  }

  /* Signal::self, classes.cpp:534-575 */
  __device__ inline double signal_self (const MithraSignal& g, double t, double phase)
  {
    const double PI = MITHRA_PI;
    const double d = t - g.t0;
    if (fabs(d) > 10.0 * g.s) return 0.0;
    switch (g.type)
      {
      case MITHRA_SIGNAL_NEUMANN:
	return - cos_wide( 2 * PI * g.f0 * d + g.cep + phase ) * 2.7724 * d / ( g.s * g.s ) * exp( -1.3863 * d * d / ( g.s * g.s ) );
      case MITHRA_SIGNAL_GAUSSIAN:
	{ const double u = d / g.s; return cos_wide( 2 * PI * g.f0 * d + g.cep + phase ) * exp( -1.3863 * ( u * u ) ); }
      case MITHRA_SIGNAL_SECANT:
	return cos_wide( 2 * PI * g.f0 * d + g.cep + phase ) / cosh( d / g.s );
      case MITHRA_SIGNAL_FLATTOP:
	if (d <= - g.s / 2.0)
	  { const double u = ( d + g.s / 2.0 ) * g.f0 / g.nR; return cos_wide( 2 * PI * g.f0 * d + g.cep + phase ) * exp( - ( u * u ) ); }
	else if (d <= g.s / 2.0)
	  return cos_wide( 2 * PI * g.f0 * d + g.cep + phase );
	else
	  { const double u = ( d - g.s / 2.0 ) * g.f0 / g.nR; return cos_wide( 2 * PI * g.f0 * d + g.cep + phase ) * exp( - ( u * u ) ); }
      case MITHRA_SIGNAL_INVGAUSSIAN:
	{
	  const double u0 = d / g.sigma_inv_g[0], u1 = d / g.sigma_inv_g[1];
	  const double env = pow( ( 1.0 + u0 * u0 ) * ( 1.0 + u1 * u1 ), 0.25 );
	  if (d <= - g.s / 2.0)
	    { const double u = ( d + g.s / 2.0 ) * g.f0 / g.nR; return cos_wide( 2 * PI * g.f0 * d + g.cep + phase ) * env * exp( - ( u * u ) ); }
	  else if (d <= g.s / 2.0)
	    return cos_wide( 2 * PI * g.f0 * d + g.cep + phase ) * env;
	  else
	    { const double u = ( d - g.s / 2.0 ) * g.f0 / g.nR; return cos_wide( 2 * PI * g.f0 * d + g.cep + phase ) * env * exp( - ( u * u ) ); }
	}
      }
    return 0.0;
  }

  __device__ __forceinline__ double sq (double x) { return x * x; }

  /* The eight lab-frame beams (beam.cc:79-496).  rv = r_lab - position, z = rv . direction, tl = t0 - z/c0,
   * tlm = t0 + z/c0.  Returns eT, bT (zero when the beam is negligible at this point).                  */
  __device__ inline void beam_fields (const MithraBeam& s, double c0, const V3& rv, double z, double tl, double tlm,
				      V3& eT, V3& bT)
  {
    const double PI = MITHRA_PI;
    const V3 dir = v3a(s.direction), pol = v3a(s.polarization);
    const V3 zero = v3(0.0, 0.0, 0.0);
    eT = zero; bT = zero;
    double p0 = 0.0;

    switch (s.seed_type)
      {
      case MITHRA_BEAM_PLANEWAVE:
      case MITHRA_BEAM_PLANEWAVETRUNCATED:
	{
	  const double ts = signal_self(s.signal, tl, p0);
	  if (fabs(ts) < 1.0e-6) return;
	  if (s.seed_type == MITHRA_BEAM_PLANEWAVETRUNCATED)
	    {
	      const double x = dot3(rv, pol);
	      const V3 yv = cross3(dir, pol);
	      const double y = dot3(rv, yv);
	      if (sq(x / s.radius[0]) + sq(y / s.radius[1]) > 1.0) return;
	    }
	  eT = scale3(s.amplitude * ts, pol);
	  bT = scale3(s.amplitude * ts / c0, cross3(dir, pol));
	  return;
	}

      case MITHRA_BEAM_STANDINGPLANEWAVE:
      case MITHRA_BEAM_STANDINGPLANEWAVETRUNCATED:
	{
	  if (s.seed_type == MITHRA_BEAM_STANDINGPLANEWAVETRUNCATED)
	    {
	      const double x = dot3(rv, pol);
	      const V3 yv = cross3(dir, pol);
	      const double y = dot3(rv, yv);
	      if (sq(x / s.radius[0]) + sq(y / s.radius[1]) > 1.0) return;
	    }
	  const double ts = signal_self(s.signal, tl, p0), tsm = signal_self(s.signal, tlm, p0);
	  const double tse = ts - tsm, tsb = ts + tsm;
	  if (fabs(tse) < 1.0e-6 && fabs(tsb) < 1.0e-6) return;
	  eT = scale3(s.amplitude * tse, pol);
	  bT = scale3(s.amplitude * tsb / c0, cross3(dir, pol));
	  return;
	}

      case MITHRA_BEAM_GAUSSIAN:
      case MITHRA_BEAM_STANDINGGAUSSIAN:
	{
	  const bool standing = (s.seed_type == MITHRA_BEAM_STANDINGGAUSSIAN);
	  const double x = dot3(rv, pol);
	  const V3 yv = cross3(dir, pol);
	  const double y = dot3(rv, yv);
	  const double wrp = sqrt( 1.0 + z * z / ( s.zR[0] * s.zR[0] ) );
	  const double wrs = sqrt( 1.0 + z * z / ( s.zR[1] * s.zR[1] ) );
	  const double x0 = x / wrp, y0 = y / wrs;
	  if (fabs(x0) > 4.0 * s.radius[0] || fabs(y0) > 4.0 * s.radius[1]) return;
	  if (!standing)
	    { const double ts = signal_self(s.signal, tl, p0); if (fabs(ts) < 1.0e-6) return; }
	  else
	    {
	      const double ts = signal_self(s.signal, tl, p0), tsm = signal_self(s.signal, tlm, p0);
	      if (fabs(ts - tsm) < 1.0e-6 && fabs(ts + tsm) < 1.0e-6) return;
	    }
	  const double atanP = atan( z / s.zR[0] ), atanS = atan( z / s.zR[1] );
	  p0 = 0.5 * ( atanP + atanS ) - PI * z / s.l * ( sq( x0 / s.zR[0] ) + sq( y0 / s.zR[1] ) );
	  double t = exp( - sq( x0 / s.radius[0] ) - sq( y0 / s.radius[1] ) ) / sqrt( wrs * wrp );
	  t *= s.amplitude;
	  V3 ex, by, ez, bz;
	  if (!standing)
	    {
	      double ts = signal_self(s.signal, tl, p0 - PI / 2.0);
	      ex = scale3(t * ts, pol);
	      by = scale3(t * ts / c0, yv);
	      ts = signal_self(s.signal, tl, p0 + atanP);
	      ez = scale3(t * ( - x0 / s.zR[0] ) * ts, dir);
	      ts = signal_self(s.signal, tl, p0 + atanS);
	      bz = scale3(t * ( - y0 / s.zR[1] ) / c0 * ts, dir);
	    }
	  else
	    {
	      double p1 = p0 - PI / 2.0;
	      double ts = signal_self(s.signal, tl, p1), tsm = signal_self(s.signal, tlm, p1);
	      ex = scale3(t * ( ts - tsm ), pol);
	      by = scale3(t / c0 * ( ts + tsm ), yv);
	      p1 = p0 + atanP;
	      ts = signal_self(s.signal, tl, p1); tsm = signal_self(s.signal, tlm, -p1);
	      ez = scale3(t * ( - x0 / s.zR[0] ) * ( ts - tsm ), dir);
	      p1 = p0 + atanS;
	      ts = signal_self(s.signal, tl, p1); tsm = signal_self(s.signal, tlm, -p1);
	      bz = scale3(t * ( - y0 / s.zR[1] ) / c0 * ( ts - tsm ), dir);
	    }
	  eT = add3(ex, ez); bT = add3(by, bz);
	  return;
	}

      case MITHRA_BEAM_SUPERGAUSSIAN:
      case MITHRA_BEAM_STANDINGSUPERGAUSSIAN:
	{
	  const bool standing = (s.seed_type == MITHRA_BEAM_STANDINGSUPERGAUSSIAN);
	  const double x = dot3(rv, pol);
	  const V3 yv = cross3(dir, pol);
	  const double y = dot3(rv, yv);
	  const double wrp = sqrt( 1.0 + z * z / ( s.zR[0] * s.zR[0] ) );
	  const double wrs = sqrt( 1.0 + z * z / ( s.zR[1] * s.zR[1] ) );
	  if ( ( fabs(x) - s.order[0] * s.radius[0] ) > 4.0 * s.radius[0] * wrp ||
	       ( fabs(y) - s.order[1] * s.radius[1] ) > 4.0 * s.radius[1] * wrs ) return;
	  if (!standing)
	    { const double ts = signal_self(s.signal, tl, p0); if (fabs(ts) < 1.0e-6) return; }
	  else
	    {
	      const double ts = signal_self(s.signal, tl, p0), tsm = signal_self(s.signal, tlm, p0);
	      if (fabs(ts - tsm) < 1.0e-6 && fabs(ts + tsm) < 1.0e-6) return;
	    }
	  const double atanP = atan( z / s.zR[0] ), atanS = atan( z / s.zR[1] );
	  const double af = s.amplitude / sqrt( wrs * wrp );
	  V3 ex = zero, by = zero, ez = zero, bz = zero;
	  for (int i = - s.order[0]; i <= s.order[0]; i++)
	    for (int j = - s.order[1]; j <= s.order[1]; j++)
	      {
		const double x0 = ( x - i * s.radius[0] ) / wrp;
		const double y0 = ( y - j * s.radius[1] ) / wrs;
		if (fabs(x0) > 4.0 * s.radius[0] || fabs(y0) > 4.0 * s.radius[1]) continue;
		p0 = 0.5 * ( atanP + atanS ) - PI * z / s.l * ( sq( x0 / s.zR[0] ) + sq( y / s.zR[1] ) );
		const double t = af * exp( - sq( x0 / s.radius[0] ) - sq( y0 / s.radius[1] ) );
		if (!standing)
		  {
		    double ts = signal_self(s.signal, tl, p0 - PI / 2.0);
		    ex = add3(ex, scale3(t * ts, pol));
		    by = add3(by, scale3(t * ts / c0, yv));
		    ts = signal_self(s.signal, tl, p0 + atanP);
		    ez = add3(ez, scale3(t * ( - x0 / s.zR[0] ) * ts, dir));
		    ts = signal_self(s.signal, tl, p0 + atanS);
		    bz = add3(bz, scale3(t * ( - y0 / s.zR[1] ) * ts / c0, dir));
		  }
		else
		  {
		    double p1 = p0 - PI / 2.0;
		    double ts = signal_self(s.signal, tl, p1), tsm = signal_self(s.signal, tlm, p1);
		    ex = add3(ex, scale3(t * ( ts - tsm ), pol));
		    by = add3(by, scale3(t * ( ts + tsm ) / c0, yv));
		    p1 = p0 + atanP;
		    ts = signal_self(s.signal, tl, p1); tsm = signal_self(s.signal, tlm, -p1);
		    ez = add3(ez, scale3(t * ( - x0 / s.zR[0] ) * ( ts - tsm ), dir));
		    p1 = p0 + atanS;
		    ts = signal_self(s.signal, tl, p1); tsm = signal_self(s.signal, tlm, -p1);
		    bz = add3(bz, scale3(t * ( - y0 / s.zR[1] ) * ( ts - tsm ) / c0, dir));
		  }
	      }
	  eT = add3(ex, ez); bT = add3(by, bz);
	  return;
	}
      }
  }

  /* Static undulator module, beam.cc:14-76.  lz = gamma (z + beta c0 (tb + dt_)) - rb, ly = x ct + y st.   */
  __device__ inline void static_undulator (const UndulatorDev& u, double gamma, double c0beta, double lz, double ly,
					   V3& et, V3& bt)
  {
    const double PI = MITHRA_PI;
    double d1, bz;
    if (lz >= 0.0 && lz <= u.len)
      {
	double sn, cs;
	sincos( u.ku * lz, &sn, &cs );                          /* one argument reduction for both (same values as sin, cos) */
	d1 = u.b0 * cosh( u.ku * ly ) * sn * gamma;
	bz = u.b0 * sinh( u.ku * ly ) * cs;
      }
    else if (lz < 0.0)
      {
	double sz = exp( - sq( u.ku * lz ) / 2.0 );
	if (u.has_prev)
	  {
	    const double r0 = u.r0_prev;
	    if (lz < r0 || r0 == 0.0) sz = 0.0;
	    else sz *= 0.35875 + 0.48829 * cos( PI * lz / r0 ) + 0.14128 * cos( 2.0 * PI * lz / r0 ) + 0.01168 * cos( 3.0 * PI * lz / r0 );
	  }
	d1 = u.b0 * cosh( u.ku * ly ) * sz * u.ku * lz * gamma;
	bz = u.b0 * sinh( u.ku * ly ) * sz;
      }
    else
      {
	const double t0 = lz - u.len;
	double sz = exp( - sq( u.ku * t0 ) / 2.0 );
	if (u.has_next)
	  {
	    const double r0 = u.r0_next;
	    if (t0 > r0 || r0 == 0.0) sz = 0.0;
	    else sz *= 0.35875 + 0.48829 * cos( PI * t0 / r0 ) + 0.14128 * cos( 2.0 * PI * t0 / r0 ) + 0.01168 * cos( 3.0 * PI * t0 / r0 );
	  }
	d1 = u.b0 * cosh( u.ku * ly ) * sz * u.ku * t0 * gamma;
	bz = u.b0 * sinh( u.ku * ly ) * sz;
      }
    bt.x += d1 * u.ct;
    bt.y += d1 * u.st;
    bt.z += bz;
    d1 *= c0beta;
    et.y +=   d1 * u.ct;
    et.x += - d1 * u.st;
    et.z += 0.0;
  }

  /* Seed::fields, classes.cpp:740-855: seed vector potential at a mesh node (moving frame) at `time`.    */
  /* Seed::fields is always  a = (sum of ni equal terms u * polarization), a_z *= gamma  (classes.cpp:783-852): the
   * scalar u carries all the space and time dependence.  seed_assemble rebuilds the vector with the reference's
   * roundings: scale3(u, pol), added ni times from zero for the super-Gaussian beam (quirk Q8), then a_z * gamma. */
  __device__ __forceinline__ int seed_terms (const MithraBeam& s)
  { return (s.seed_type == MITHRA_BEAM_SUPERGAUSSIAN) ? ( 2 * s.order[0] + 1 ) * ( 2 * s.order[1] + 1 ) : 1; }

  __device__ __forceinline__ double seed_assemble_comp (double u, double polc, int ni, bool supergaussian, bool zcomp, double gamma)
  {
    const double one = u * polc;
    double a = one;
    if (supergaussian) { a = 0.0; for (int n = 0; n < ni; n++) a = a + one; }
    if (zcomp) a *= gamma;
    return a;
  }

  __device__ __forceinline__ V3 seed_assemble (const MithraBeam& s, double gamma, double u)
  {
    const bool sg = (s.seed_type == MITHRA_BEAM_SUPERGAUSSIAN);
    const int  ni = seed_terms(s);
    return v3(seed_assemble_comp(u, s.polarization[0], ni, sg, false, gamma),
	      seed_assemble_comp(u, s.polarization[1], ni, sg, false, gamma),
	      seed_assemble_comp(u, s.polarization[2], ni, sg, true,  gamma));
  }

  /* the scalar u of Seed::fields at the point (px, py, pz) of the moving frame, any beam direction              */
  __device__ inline double seed_scalar (const MithraBeam& s, double c0, double gamma, double beta, double dt_shift,
					double px, double py, double pz, double time)
  {
    const double PI = MITHRA_PI;
    const V3 dir = v3a(s.direction), pol = v3a(s.polarization);
    const V3 rl = v3(px, py, gamma * ( pz + beta * c0 * ( time + dt_shift ) ));
    double tl = gamma * ( time + dt_shift + beta / c0 * pz );
    const V3 rv = v3(rl.x - s.position[0], rl.y - s.position[1], rl.z - s.position[2]);
    const double z = dot3(rv, dir);
    tl -= z / c0;
    double p = 0.0;
    double ts = signal_self(s.signal, tl, p);

    if (s.seed_type == MITHRA_BEAM_PLANEWAVE)
      {
	if (!(fabs(ts) < 1.0e-6)) return s.amplitude * ts;
      }
    else if (s.seed_type == MITHRA_BEAM_PLANEWAVETRUNCATED)
      {
	const double x = dot3(rv, pol);
	const double y = dot3(rv, cross3(dir, pol));
	if (!(fabs(ts) < 1.0e-6 || fabs(x) > s.radius[0] || fabs(y) > s.radius[1])) return s.amplitude * ts;
      }
    else if (s.seed_type == MITHRA_BEAM_GAUSSIAN || s.seed_type == MITHRA_BEAM_SUPERGAUSSIAN)
      {
	if (!(fabs(ts) < 1.0e-6))
	  {
	    const double x = dot3(rv, pol);
	    const double y = dot3(rv, cross3(dir, pol));
	    const double l = c0 / s.signal.f0;
	    const double zRp = PI * s.radius[0] * s.radius[0] / l;
	    const double wrp = sqrt( 1.0 + z * z / ( zRp * zRp ) );
	    const double zRs = PI * s.radius[1] * s.radius[1] / l;
	    const double wrs = sqrt( 1.0 + z * z / ( zRs * zRs ) );
	    /* every one of the ni terms of the reference's loop evaluates to this same value                         */
	    p  = 0.5 * ( atan( z / zRp ) + atan( z / zRs ) - PI ) - PI * z / l * ( sq( x / ( zRp * wrp ) ) + sq( y / ( zRs * wrs ) ) );
	    ts = signal_self(s.signal, tl, p);
	    const double t = exp( - sq( x / ( s.radius[0] * wrp ) ) - sq( y / ( s.radius[1] * wrs ) ) ) / sqrt( wrs * wrp ) * s.amplitude;
	    return t * ts;
	  }
      }
    return 0.0;
  }

  /* Seed::fields, classes.cpp:740-855 */
  __device__ inline V3 seed_fields (const MithraBeam& s, double c0, double gamma, double beta, double dt_shift,
				    double px, double py, double pz, double time)
  { return seed_assemble(s, gamma, seed_scalar(s, c0, gamma, beta, dt_shift, px, py, pz, time)); }
}

#endif
