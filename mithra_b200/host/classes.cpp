/* classes.cpp -- parameter model of a MITHRA job (see classes.h): defaults, normalisations, bunch generators.
 *
 * Everything here runs once on the host before the time-march.  The numbers it produces (unit vectors, super-gaussian
 * corrections, Halton positions and momenta of the macro-particles) enter the parity contract: they are computed with
 * the reference's formulas in the reference's operation order (cited per function) so that the parameter block and
 * the initial bunch handed to the GPU are bit-identical to what the reference's Solver::initialize() holds.
 */
#include "classes.h"

#include <algorithm>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sys/stat.h>

namespace MITHRA
{
  /* ---- small utilities -------------------------------------------------------------------------------- */

  void printmessage (std::string filename, unsigned int linenumber, std::string message)
  {
    /* stdinclude.h:77-108: "<ctime> ::: <file>:<line> ::: \t \t <message>"                                    */
    time_t raw; time(&raw);
    std::string ts = ctime(&raw);
    const size_t slash = filename.find_last_of('/');
    std::cout << ts.substr(0, ts.size() - 1) << " ::: " << (slash == std::string::npos ? filename : filename.substr(slash + 1))
	      << ":" << linenumber << " ::: \t \t " << message << std::endl;
  }

  void createDirectory (std::string filename, unsigned int rank)
  {
    const size_t slash = filename.find_last_of('/');
    if (slash == std::string::npos || rank != 0) return;
    const std::string path = filename.substr(0, slash);
    struct stat st;
    if (path.empty() || stat(path.c_str(), &st) == 0) return;
    if (mkdir(path.c_str(), S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH) == -1)
      {
	std::cout << "Could not create the directory " << path << ". Probably the given address does not exist." << std::endl;
	exit(1);
      }
  }

  /* Radical inverse of j+1 in the i-th prime base, returned as 1 - x (stdinclude.cpp:45-73); int arithmetic on
   * purpose: the reference's p0 overflows for long sequences in small bases and the overflow is part of its output */
  Double halton (unsigned int i, unsigned int j)
  {
    static const unsigned int prime[20] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71 };
    if (i > 20) { printmessage(__FILE__, __LINE__, " dimension can not be larger than 20. "); exit(1); }
    const int p = prime[i];
    int p0 = p, k = j + 1;
    Double x = 0.0;
    while (k > 0)
      {
	const int a = k % p;
	x  += a / (double) p0;
	k   = int (k / p);
	p0 *= p;
      }
    return 1.0 - x;
  }

  Double pmod (const Double& a, const Double& b)
  {
    Double x = fmod(a, b);
    x += ( x < 0.0 ) ? b : 0.0;
    return x;
  }

  /* ---- Mesh ------------------------------------------------------------------------------------------- */

  void Mesh::initialize ()
  {
    spaceCharge_ = false; optimizePosition_ = false; solver_ = NSFD; totalDist_ = 0.0; timeShift_ = 0.0; gamma_ = -1.0;
  }

  void Mesh::show ()
  {
    printmessage(__FILE__, __LINE__, " Length scale = " + stringify(lengthScale_) + ", time scale = " + stringify(timeScale_));
    printmessage(__FILE__, __LINE__, " Total simulation time = " + stringify(totalTime_) + ", truncation order = " + stringify(truncationOrder_));
    printmessage(__FILE__, __LINE__, std::string(" Space-charge = ") + (spaceCharge_ ? "true" : "false") +
		 ", solver = " + (solver_ == NSFD ? "non-standard finite-difference" : "finite-difference"));
  }

  /* ---- Bunch ------------------------------------------------------------------------------------------ */

  BunchInitialize::BunchInitialize ()
    : bunchType_(""), distribution_(""), generator_(""), numberOfParticles_(0), cloudCharge_(0.0), initialGamma_(0.0),
      initialBeta_(0.0), initialDirection_(0.0), betaVector_(0.0), numbers_(3, 0u), latticeConstants_(0.0), sigmaPosition_(0.0),
      sigmaGammaBeta_(0.0), tranTrun_(0.0), longTrun_(0.0), fileName_(""), bF_(0.0), bFP_(0.0), shotNoise_(false), lambda_(0.0)
  {}

  Bunch::Bunch ()
    : timeStep_(0.0), sampling_(false), directory_("./"), basename_(""), rhythm_(0.0), bunchVTK_(false), bunchVTKDirectory_("./"),
      bunchVTKBasename_(""), bunchVTKRhythm_(0.0), bunchProfile_(false), bunchProfileDirectory_("./"), bunchProfileBasename_(""),
      bunchProfileRhythm_(0.0), zu_(0.0), beta_(0.0)
  {}

  void Bunch::show ()
  {
    for (const BunchInitialize& b : bunchInit_)
      printmessage(__FILE__, __LINE__, " Bunch: type " + b.bunchType_ + ", " + stringify(b.numberOfParticles_) + " macro-particles, " +
		   stringify(b.cloudCharge_) + " electrons, gamma " + stringify(b.initialGamma_));
  }

  /* one charge equal to cloudCharge_ at the given position, classes.cpp:84-98                                  */
  void Bunch::initializeManual (BunchInitialize bunchInit, ChargeVector& chargeVector, Double (zp)[2], int rank, int size, int ia)
  {
    Charge charge;
    charge.q   = bunchInit.cloudCharge_;
    charge.rnp = bunchInit.position_[ia];
    charge.gb.mv( bunchInit.initialGamma_, bunchInit.betaVector_ );
    if ( ( charge.rnp[2] < zp[1] || rank == size - 1 ) && ( charge.rnp[2] >= zp[0] || rank == 0 ) )
      chargeVector.push_back(charge);
  }

  /* Gaussian transverse / uniform or Gaussian longitudinal ellipsoid from Halton (or rand) numbers, inserted in
   * groups of four particles a quarter of the bunching wavelength apart (quiet start), with optional bunching factor or
   * shot noise and, for the uniform profile, Gaussian tapers at both ends -- classes.cpp:104-298.                   */
  void Bunch::initializeEllipsoid (BunchInitialize bunchInit, ChargeVector& chargeVector, int rank, int size, int ia)
  {
    if ( bunchInit.numberOfParticles_ % 4 != 0 )
      {
	bunchInit.numberOfParticles_ += 4 - bunchInit.numberOfParticles_ % 4;
	printmessage(__FILE__, __LINE__, "Warning: The number of particles in the bunch is not a multiple of four. It is corrected to " +
		     stringify(bunchInit.numberOfParticles_));
      }

    const unsigned int Np = bunchInit.numberOfParticles_, Np0 = chargeVector.size();
    unsigned int       i;
    Charge             charge; charge.q = bunchInit.cloudCharge_ / Np;
    FieldVector        gb (0.0); gb.mv( bunchInit.initialGamma_, bunchInit.betaVector_ );
    FieldVector        r (0.0), t (0.0);
    Double             t0, zmin = 1e100, Ne, bF = 0.0, bFi;
    unsigned int       bmi;
    std::vector<Double> randomNumbers;

    /* groups of four only when an undulator defines a bunching wavelength                                        */
    const unsigned int ng = ( bunchInit.lambda_ == 0.0 ) ? 1 : 4;

    if ( bunchInit.bF_ > 2.0 || bunchInit.bF_ < 0.0 )
      { printmessage(__FILE__, __LINE__, "The bunching factor can not be larger than one or a negative value !!!"); exit(1); }

    if ( bunchInit.generator_ == "random" )
      {
	srand ( time(NULL) );
	randomNumbers.resize( Np / ng * 20, 0.0 );
	for (unsigned int ri = 0; ri < Np / ng * 20; ri++) randomNumbers[ri] = ( (double) rand() ) / RAND_MAX;
      }
    auto generate = [&] (unsigned int n, unsigned int m) -> Double {
      return ( bunchInit.generator_ == "random" ) ? randomNumbers[ n * 2 * Np / ng + m ] : halton(n, m); };

    const Double lam = bunchInit.lambda_;
    auto insertCharge = [&] (Charge q) {
      for (unsigned int ii = 0; ii < ng; ii++)
	{
	  if ( bunchInit.shotNoise_ )
	    {
	      bmi = int( ( charge.rnp[2] - zmin ) / lam );
	      bFi = bF * sqrt( - 2.0 * log( generate( 8 , bmi ) ) );
	      q.rnp[2]  = charge.rnp[2] - lam / 4 * ii;
	      q.rnp[2] -= lam / PI * bFi * sin( 2.0 * PI / lam * q.rnp[2] + 2.0 * PI * generate( 9 , bmi ) );
	    }
	  else if ( lam != 0.0 )
	    {
	      q.rnp[2]  = charge.rnp[2] - lam / 4 * ii;
	      q.rnp[2] -= lam / PI * bunchInit.bF_ * sin( 2.0 * PI / lam * q.rnp[2] + bunchInit.bFP_ * PI / 180.0 );
	    }
	  chargeVector.push_back(q);
	}
    };

    const Double sz = bunchInit.sigmaPosition_[2];
    const bool uniform = ( bunchInit.distribution_ == "uniform" ), gaussian = ( bunchInit.distribution_ == "gaussian" );
    auto badProfile = [] () { printmessage(__FILE__, __LINE__, "The longitudinal type is not correctly given to the code !!!"); exit(1); };
    /* number of body + taper samples, classes.cpp:203,268                                                        */
    auto taperEnd = [&] () { return unsigned( Np / ng * ( 1.0 + 2.0 * lam * sqrt( 2.0 * PI ) / ( 2.0 * sz ) ) ); };

    if ( bunchInit.shotNoise_ )
      {
	/* the lowest z of the bunch numbers the FEL buckets                                                      */
	for (i = 0; i < Np / ng; i++)
	  {
	    if      ( uniform )  zmin = std::min( ( 2.0 * generate(2, i + Np0) - 1.0 ) * sz , zmin );
	    else if ( gaussian ) zmin = std::min( sz * sqrt( - 2.0 * log( generate(2, i + Np0) ) ) * sin( 2.0 * PI * generate(3, i + Np0) ) , zmin );
	    else badProfile();
	  }
	if ( uniform )
	  for ( ; i < taperEnd(); i++)
	    {
	      t0  = 2.0 * lam * sqrt( - 2.0 * log( generate( 2, i + Np0 ) ) ) * sin( 2.0 * PI * generate( 3, i + Np0 ) );
	      t0 += ( t0 < 0.0 ) ? ( - sz ) : ( sz );
	      zmin = std::min( t0 , zmin );
	    }
	zmin = zmin + bunchInit.position_[ia][2];
	Ne = bunchInit.cloudCharge_ * lam / ( 2.0 * sz );
	bF = ( bunchInit.bF_ == 0.0 ) ? 1.0 / sqrt(Ne) : bunchInit.bF_;
	printmessage(__FILE__, __LINE__, "The standard deviation of the bunching factor for the shot noise implementation is set to " + stringify(bF));
      }

    /* Box-Muller pairs on Halton dimensions (0,1) position, (4,5) transverse and (6,7) longitudinal momentum       */
    auto transverse = [&] (unsigned int m) {
      r[0] = bunchInit.sigmaPosition_[0] * sqrt( - 2.0 * log( generate(0, m) ) ) * cos( 2.0 * PI * generate(1, m) );
      r[1] = bunchInit.sigmaPosition_[1] * sqrt( - 2.0 * log( generate(0, m) ) ) * sin( 2.0 * PI * generate(1, m) ); };
    auto momentum = [&] (unsigned int m) {
      t[0] = bunchInit.sigmaGammaBeta_[0] * sqrt( - 2.0 * log( generate(4, m) ) ) * cos( 2.0 * PI * generate(5, m) );
      t[1] = bunchInit.sigmaGammaBeta_[1] * sqrt( - 2.0 * log( generate(4, m) ) ) * sin( 2.0 * PI * generate(5, m) );
      t[2] = bunchInit.sigmaGammaBeta_[2] * sqrt( - 2.0 * log( generate(6, m) ) ) * cos( 2.0 * PI * generate(7, m) ); };
    auto accept = [&] () {
      if ( fabs(r[0]) < bunchInit.tranTrun_ && fabs(r[1]) < bunchInit.tranTrun_ && fabs(r[2]) < bunchInit.longTrun_ )
	{
	  charge.rnp  = bunchInit.position_[ia]; charge.rnp += r;
	  charge.gb   = gb;                      charge.gb  += t;
	  insertCharge(charge);
	} };

    for (i = rank; i < Np / ng; i += size)
      {
	transverse(i + Np0);
	if      ( uniform )  r[2] = ( 2.0 * generate(2, i + Np0) - 1.0 ) * sz;
	else if ( gaussian ) r[2] = sz * sqrt( - 2.0 * log( generate(2, i + Np0) ) ) * sin( 2.0 * PI * generate(3, i + Np0) );
	else badProfile();
	momentum(i + Np0);
	accept();
      }

    /* uniform profile: Gaussian tapers of two bunching wavelengths beyond both ends remove the coherent spontaneous
     * emission of the sharp edges                                                                                */
    if ( uniform )
      for ( ; i < taperEnd(); i += size)
	{
	  transverse(i + Np0);
	  r[2]  = 2.0 * lam * sqrt( - 2.0 * log( generate(2, i + Np0) ) ) * sin( 2.0 * PI * generate(3, i + Np0) );
	  r[2] += ( r[2] < 0.0 ) ? ( - sz ) : ( sz );
	  momentum(i + Np0);
	  accept();
	}
  }

  /* np particles per lattice point of an n0 x n1 x n2 crystal with a small Gaussian spread, classes.cpp:306-355   */
  void Bunch::initialize3DCrystal (BunchInitialize bunchInit, ChargeVector& chargeVector, Double (zp)[2], int rank, int size, int ia)
  {
    const unsigned int* n = &bunchInit.numbers_[0];
    if ( bunchInit.numberOfParticles_ % (n[0] * n[1] * n[2]) != 0 )
      { printmessage(__FILE__, __LINE__, "The number of the particles and their lattice numbers do not match !!!"); exit(1); }
    Charge      charge;
    FieldVector gb (0.0); gb.mv( bunchInit.initialGamma_, bunchInit.betaVector_ );
    const unsigned int np = bunchInit.numberOfParticles_ / (n[0] * n[1] * n[2]);
    const FieldVector& c = bunchInit.position_[ia], & a = bunchInit.latticeConstants_, & sp = bunchInit.sigmaPosition_, & sg = bunchInit.sigmaGammaBeta_;
    chargeVector.clear();
    for (unsigned int i = 0; i < n[0]; i++)
      for (unsigned int j = 0; j < n[1]; j++)
	for (unsigned int k = 0; k < n[2]; k++)
	  for (unsigned int l = 0; l < np; l++)
	    {
	      charge.q = bunchInit.cloudCharge_ / bunchInit.numberOfParticles_;
	      charge.rnp[0]  = c[0] + ( i + 1.0 - 0.5 * n[0] ) * a[0];
	      charge.rnp[1]  = c[1] + ( j + 1.0 - 0.5 * n[1] ) * a[1];
	      charge.rnp[2]  = c[2] + ( k + 1.0 - 0.5 * n[2] ) * a[2];
	      /* the Halton index is the x lattice index in the reference, for every coordinate                     */
	      charge.rnp[0] += 0.5 * sp[0] * sqrt( - 2.0 * log( halton(0,i) ) ) * sin( 2.0 * PI * halton(1,i) );
	      charge.rnp[1] += 0.5 * sp[1] * sqrt( - 2.0 * log( halton(2,i) ) ) * sin( 2.0 * PI * halton(3,i) );
	      charge.rnp[2] += 0.5 * sp[2] * sqrt( - 2.0 * log( halton(4,i) ) ) * sin( 2.0 * PI * halton(5,i) );
	      charge.gb     = gb;
	      charge.gb[0] += sg[0] * sqrt( - 2.0 * log( halton(6,i) ) ) * sin( 2.0 * PI * halton(7,i) );
	      charge.gb[1] += sg[1] * sqrt( - 2.0 * log( halton(8,i) ) ) * sin( 2.0 * PI * halton(9,i) );
	      charge.gb[2] += sg[2] * sqrt( - 2.0 * log( halton(10,i)) ) * sin( 2.0 * PI * halton(11,i));
	      if ( ( charge.rnp[2] < zp[1] || rank == size - 1 ) && ( charge.rnp[2] >= zp[0] || rank == 0 ) )
		chargeVector.push_back(charge);
	    }
  }

  /* six columns x y z gbx gby gbz per particle, dealt round-robin over the ranks, classes.cpp:363-417             */
  void Bunch::initializeFile (BunchInitialize bunchInit, ChargeVector& chargeVector, Double (zp)[2], int rank, int size, int ia)
  {
    Charge charge;
    int    saveRank = 0;
    bool   outside = false;
    chargeVector.clear();
    std::ifstream in ( bunchInit.fileName_.c_str() );
    charge.q = bunchInit.cloudCharge_ / bunchInit.numberOfParticles_;
    while (in.good())
      {
	in >> charge.rnp[0]; in >> charge.rnp[1]; in >> charge.rnp[2];
	in >> charge.gb[0];  in >> charge.gb[1];  in >> charge.gb[2];
	charge.rnp += bunchInit.position_[ia];
	if (saveRank == rank) chargeVector.push_back(charge);
	saveRank = ( saveRank == size - 1 ) ? 0 : saveRank + 1;
	if ( bunchInit.tranTrun_ > 0.0 && ( fabs(charge.rnp[0]) > bunchInit.tranTrun_ || fabs(charge.rnp[1]) > bunchInit.tranTrun_ ) ) outside = true;
      }
    if (outside)
      printmessage(__FILE__, __LINE__, "Warning: Some particle coordinates are out of the transverse truncation length for the bunch. The results may be inaccurate !!!");
    if ( size == 1 && bunchInit.numberOfParticles_ != chargeVector.size() )
      {
	printmessage(__FILE__, __LINE__, "The number of the particles and the file size do not match !!! The file contains " + stringify(chargeVector.size()) + " particles.");
	exit(1);
      }
  }

  /* ---- Signal ----------------------------------------------------------------------------------------- */

  Signal::Signal () : signalType_(GAUSSIAN), t0_(0.0), s_(0.0), f0_(1.0), cep_(0.0), nR_(1), sigmaInvG_(2, 0.0) {}

  /* offset, pulse length and wavelength are lengths in the job file; they become times / a frequency when
   * Solver::setSimulationParameters divides by c0 (solver.cpp:75-77) -- classes.cpp:487-532                      */
  void Signal::initialize (std::string type, Double l0, Double s, Double l, Double cep, unsigned int nR, std::vector<Double> sigmaInvG)
  {
    if      ( type == "neumann" )           signalType_ = NEUMANN;
    else if ( type == "gaussian" )          signalType_ = GAUSSIAN;
    else if ( type == "secant-hyperbolic" ) signalType_ = SECANT;
    else if ( type == "flat-top" )          signalType_ = FLATTOP;
    else if ( type == "inverse-gaussian" )  signalType_ = INVGAUSSIAN;
    else { std::cout << type << " is an unknown signal type for the given set of parameters." << std::endl; exit(1); }
    t0_  = l0;
    s_   = s;
    f0_  = 1 / l;
    cep_ = cep * PI / 180;
    nR_  = nR;
    sigmaInvG_ = sigmaInvG;
    if ( s_ == 0.0 )
      { printmessage(__FILE__, __LINE__, " Variance of signal is set to zero. This is not allowed because we divide through the variance. Exit!"); exit(1); }
    if ( signalType_ == INVGAUSSIAN && sigmaInvG_[0] * sigmaInvG_[1] == 0.0 )
      { printmessage(__FILE__, __LINE__, " sigma of the inverse-gaussian signal is set to zero. This is not allowed because we divide through the sigma value. Exit!"); exit(1); }
  }

  /* ---- Beam: Seed, optical Undulator, ExtField -------------------------------------------------------------- */

  Beam::Beam () : seedType_(PLANEWAVE), c0_(0.0), position_(0.0), direction_(0.0), polarization_(0.0), amplitude_(0.0), a0_(0.0),
		  radius_(2, 0.0), order_(2, 0), l_(0.0), zR_(2, 0.0) {}

  static SeedType beamType (const std::string& type)
  {
    static const char* names[8] = { "plane-wave", "truncated-plane-wave", "gaussian-beam", "super-gaussian-beam", "standing-plane-wave",
				    "standing-truncated-plane-wave", "standing-gaussian-beam", "standing-super-gaussian-beam" };
    for (int t = 0; t < 8; t++) if (type == names[t]) return (SeedType) t;
    std::cout << type << " is an unknown type." << std::endl; exit(1);
  }

  void Beam::initializeBeam (std::string type, std::vector<Double> position, std::vector<Double> direction, std::vector<Double> polarization,
			     Double a0, std::vector<Double> radius, std::vector<int> order, Signal signal, bool fourthRoot)
  {
    seedType_     = beamType(type);
    position_     = position;
    polarization_ = polarization;
    direction_    = direction;

    if ( direction_.norm2() == 0.0 )
      { printmessage(__FILE__, __LINE__, "The direction vector of the beam has zero length."); exit(1); }
    direction_ /= fourthRoot ? sqrt( direction_.norm() ) : sqrt( direction_.norm2() );
    if ( polarization_.norm2() == 0.0 )
      { printmessage(__FILE__, __LINE__, "The polarization vector of the beam has zero length."); exit(1); }
    polarization_ /= fourthRoot ? sqrt( polarization_.norm() ) : polarization_.norm();
    if ( fabs( polarization_ * direction_ ) > 1.0e-50 )
      { printmessage(__FILE__, __LINE__, "The polarization of the beam is not normal to its direction."); exit(1); }

    a0_     = a0;
    radius_ = radius;
    const bool gaussian = ( seedType_ == GAUSSIANBEAM || seedType_ == STANDINGGAUSSIANBEAM || seedType_ == SUPERGAUSSIANBEAM || seedType_ == STANDINGSUPERGAUSSIANBEAM );
    if ( gaussian && radius_[0] * radius_[1] == 0.0 )
      { printmessage(__FILE__, __LINE__, "One of the radii of the gaussian beam is set to zero."); exit(1); }
    signal_ = signal;

    if ( seedType_ == SUPERGAUSSIANBEAM || seedType_ == STANDINGSUPERGAUSSIANBEAM )
      {
	order_ = order;
	Double d1 = 0.0, d2 = 0.0;
	for ( int i = -order_[0]; i <= order_[0]; i++ ) d1 += exp(-i*i);
	for ( int i = -order_[1]; i <= order_[1]; i++ ) d2 += exp(-i*i);
	/* as shipped: the parallel radius is divided twice, the perpendicular one never (classes.cpp:733-734)        */
	radius_[0] /= order_[0] + sqrt( 1.0 - log(d1) );
	radius_[0] /= order_[0] + sqrt( 1.0 - log(d2) );
	a0_ /= d1 * d2;
      }
  }

  Seed::Seed () : beta_(0.0), gamma_(1.0), dt_(0.0), sampling_(false), samplingType_(ATPOINT), samplingDirectory_(""), samplingBasename_(""),
		  samplingRhythm_(0.0), samplingLineBegin_(0.0), samplingLineEnd_(0.0), samplingRes_(0), profile_(false), profileDirectory_(""),
		  profileBasename_(""), profileRhythm_(0.0) {}

  void Seed::initialize (std::string type, std::vector<Double> position, std::vector<Double> direction, std::vector<Double> polarization,
			 Double a0, std::vector<Double> radius, std::vector<int> order, Signal signal)
  { initializeBeam(type, position, direction, polarization, a0, radius, order, signal, false); }

  SamplingType Seed::samplingType (std::string s)
  {
    if (s == "at-point")  return ATPOINT;
    if (s == "over-line") return OVERLINE;
    std::cout << s << " is an unknown sampling type." << std::endl; exit(1);
  }
  SamplingType Seed::vtkType (std::string s)
  {
    if (s == "in-plane")   return INPLANE;
    if (s == "all-domain") return ALLDOMAIN;
    std::cout << s << " is an unknown vtk type." << std::endl; exit(1);
  }
  PlaneType Seed::planeType (std::string s)
  {
    if (s == "yz") return XNORMAL;
    if (s == "xz") return YNORMAL;
    if (s == "xy") return ZNORMAL;
    std::cout << s << " is an unknown vtk plane type." << std::endl; exit(1);
  }
  FieldType Seed::fieldType (std::string s)
  {
    static const char* names[10] = { "Ex", "Ey", "Ez", "Bx", "By", "Bz", "Ax", "Ay", "Az", "F" };
    for (int t = 0; t < 10; t++) if (s == names[t]) return (FieldType) t;
    std::cout << s << " is an unknown sampling field." << std::endl; exit(1);
  }

  Undulator::Undulator () : k_(0.0), lu_(0.0), rb_(0.0), length_(0), dist_(0.0), theta_(0.0), type_(STATIC) {}

  void Undulator::initialize (std::string type, std::vector<Double> position, std::vector<Double> direction, std::vector<Double> polarization,
			      Double a0, std::vector<Double> radius, Double wavelength, std::vector<int> order, Signal signal)
  {
    initializeBeam(type, position, direction, polarization, a0, radius, order, signal, false);
    lu_ = wavelength;                                    /* the undulator period of an optical undulator, classes.cpp:1073 */
  }

  ExtField::ExtField () : type_(EMWAVE) {}

  void ExtField::initialize (std::string type, std::vector<Double> position, std::vector<Double> direction, std::vector<Double> polarization,
			     Double a0, std::vector<Double> radius, Double wavelength, std::vector<int> order, Signal signal)
  { (void) wavelength; initializeBeam(type, position, direction, polarization, a0, radius, order, signal, true); }

  /* ---- FEL output ------------------------------------------------------------------------------------------- */

  FreeElectronLaser::RadiationSampling::RadiationSampling ()
    : sampling_(false), directory_(""), basename_(""), lineBegin_(0.0), lineEnd_(0.0), res_(0), samplingType_(ATPOINT),
      lambdaMin_(0.0), lambdaMax_(0.0), lambdaRes_(0) {}

  void FreeElectronLaser::RadiationSampling::samplingType (std::string s)
  {
    if      (s == "at-point")  samplingType_ = ATPOINT;
    else if (s == "over-line") samplingType_ = OVERLINE;
    else { std::cout << s << " is an unknown sampling type." << std::endl; exit(1); }
  }

  FreeElectronLaser::RadiationVisualization::RadiationVisualization ()
    : z_(0.0), sampling_(false), directory_(""), basename_(""), rhythm_(0.0), lambda_(0.0) {}

  FreeElectronLaser::ScreenProfile::ScreenProfile () : sampling_(false), directory_("./"), basename_(""), rhythm_(0.0) {}
}
