/* solver.cpp -- Solver / FdTd / FdTdSC (see solver.h).
 *
 * Host side of the time-march: Solver::initialize() derives, in FP64 and in the reference's operation order, every
 * scalar and table the device needs (cited per function), fills MithraGpuParams, splits the bunch over the z-slabs and
 * hands both to the library; solve() then runs the reference's loop (solver.cpp:1212-1418) whose method calls are
 * one-line forwards to the C ABI.  Power and screen files are written in the reference's byte format from the rows /
 * records the library returns (radiation.cpp:76-86, 222-230; solver.cpp:2175-2187, 2229-2252).
 */
#include "solver.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sys/time.h>

namespace MITHRA
{
  /* ========================================================================================================== */
  /* construction                                                                                                 */

  Solver::Solver (Mesh& mesh, Bunch& bunch, Seed& seed, std::vector<Undulator>& undulator, std::vector<ExtField>& extField,
		  std::vector<FreeElectronLaser>& FEL)
    : mesh_(mesh), bunch_(bunch), seed_(seed), undulator_(undulator), extField_(extField), FEL_(FEL),
      N0_(0), N1_(0), N2_(0), N1N0_(0), np_(0), k0_(0), rank_(0), size_(1),
      xmin_(0), xmax_(0), ymin_(0), ymax_(0), zmin_(0), zmax_(0),
      gamma_(1.0), beta_(0.0), dt_(0.0), timep1_(0.0), time_(0.0), timem1_(0.0), timeBunch_(0.0),
      nTime_(0), nTimeBunch_(0), Nc_(0), nUpdateBunch_(1.0), maxSteps_(-1), powerGroup_(-1), screenGroup_(-1), pmapGroup_(-1), bunchSampleFile_(0), fieldSampleFile_(0), sfCe_(0.0), sfCb_(0.0), sfCa_(0.0), spaceChargeSolver_(false)
  {
    zp_[0] = zp_[1] = 0.0;
    deviceBunch_ = 0;
    memset(&uf_, 0, sizeof(uf_)); memset(&uc_, 0, sizeof(uc_)); memset(&ub_, 0, sizeof(ub_));
    /* light speed, vacuum permeability and permittivity in the job's units, solver.cpp:59-61                    */
    c0_ = C0 / mesh_.lengthScale_ * mesh_.timeScale_;
    m0_ = MU_ZERO / mesh_.lengthScale_;
    e0_ = 1.0 / ( c0_ * c0_ * m0_ );
  }

  Solver::~Solver ()
  {
    for (MithraGpu* g : gpu_) mithra_gpu_destroy(g);
  }

  void Solver::check (int rc) const
  {
    if (rc == 0) return;
    printmessage(__FILE__, __LINE__, std::string("libmithra_gpu: ") + mithra_gpu_last_error());
    exit(1);
  }

  FdTd::FdTd (Mesh& mesh, Bunch& bunch, Seed& seed, std::vector<Undulator>& undulator, std::vector<ExtField>& extField,
	      std::vector<FreeElectronLaser>& FEL) : Solver(mesh, bunch, seed, undulator, extField, FEL) {}

  FdTdSC::FdTdSC (Mesh& mesh, Bunch& bunch, Seed& seed, std::vector<Undulator>& undulator, std::vector<ExtField>& extField,
		  std::vector<FreeElectronLaser>& FEL) : FdTd(mesh, bunch, seed, undulator, extField, FEL) { spaceChargeSolver_ = true; }

  /* ========================================================================================================== */
  /* initialisation chain                                                                                         */

  /* order of solver.cpp:547-595 */
  void Solver::initialize ()
  {
    setSimulationParameters();
    initializeBunch();
    timeBunch_ = time_;
    lorentzBoostMesh();
    initializeMesh();
    lorentzBoostBunch();
    initializeField();
    if ( seed_.sampling_ ) initializeSeedSampling();
    initializeSeedVTK();
    if ( seed_.profile_ ) initializeSeedProfile();
    initializeBunchUpdate();
    initializePowerSample();
    initializePowerVisualize();
    initializeScreenProfile();
    shiftBackInTime();
  }

  /* unit conversion of the beams, frame gamma, undulator sorting, bunching wavelength -- solver.cpp:68-214      */
  void Solver::setSimulationParameters ()
  {
    auto toSolverUnits = [&] (Beam& b) {
      b.c0_          = c0_;
      b.signal_.t0_ /= c0_;
      b.signal_.f0_ *= c0_;
      b.signal_.s_  /= c0_;
      b.l_           = c0_ / b.signal_.f0_;
      b.zR_.assign(2, 0.0);
      b.zR_[0]       = PI * b.radius_[0] * b.radius_[0] / b.l_;
      b.zR_[1]       = PI * b.radius_[1] * b.radius_[1] / b.l_; };
    toSolverUnits(seed_);
    for (Undulator& u : undulator_) toSolverUnits(u);
    for (ExtField&  e : extField_)  toSolverUnits(e);

    /* the seed amplitude is that of the vector potential, the others are field amplitudes (solver.cpp:119-123)   */
    seed_.amplitude_ = seed_.a0_ * EM * c0_ / EC;
    for (Undulator& u : undulator_) u.amplitude_ = u.a0_ * EM * c0_ * 2 * PI * u.signal_.f0_ / EC;
    for (ExtField&  e : extField_)  e.amplitude_ = e.a0_ * EM * c0_ * 2 * PI * e.signal_.f0_ / EC;

    Double gamma = 0.0;
    for (BunchInitialize& b : bunch_.bunchInit_)
      {
	if ( b.bunchType_ == "file" )  computeFileGamma(b);
	if ( b.bunchType_ == "other" ) printmessage(__FILE__, __LINE__, "Bunch mean gamma and direction are given by an external program. ");
	gamma += b.initialGamma_ / bunch_.bunchInit_.size();
      }

    /* range of the bunch's longitudinal gamma inside the undulators (solver.cpp:138-165)                         */
    Double gmin = 1.0e100, gmax = -1.0e100, g = 0.0;
    for (Undulator& u : undulator_)
      {
	const Double k = ( u.type_ == STATIC ) ? u.k_ : u.a0_;
	g    = gamma / sqrt( 1.0 + k * k / 2.0 );
	gmin = ( gmin < g ) ? gmin : g;
	/* an optical pulse that is not flat-top leaves the electrons at their full gamma outside the pulse         */
	if ( u.type_ != STATIC && u.signal_.signalType_ != FLATTOP ) g = gamma;
	gmax = ( gmax > g ) ? gmax : g;
      }
    if ( mesh_.gamma_ == -1.0 ) gamma_ = ( undulator_.size() == 0 ) ? gamma : ( gmin + gmax ) / 2.0;
    else                        gamma_ = mesh_.gamma_;

    beta_ = sqrt( 1.0 - 1.0 / ( gamma_ * gamma_ ) );
    for (Undulator& u : undulator_) if ( u.type_ == OPTICAL ) u.lu_ /= ( 1 + beta_ );

    std::sort(undulator_.begin(), undulator_.end(), undulatorCompare);
    for (std::vector<Undulator>::reverse_iterator u = undulator_.rbegin(); u != undulator_.rend(); u++) u->rb_ -= undulator_[0].rb_;

    for (BunchInitialize& b : bunch_.bunchInit_)
      {
	b.initialBeta_ = sqrt( 1.0 - 1.0 / pow( b.initialGamma_ , 2 ) );
	b.betaVector_.mv( b.initialBeta_, b.initialDirection_ );
	b.lambda_ = ( undulator_.size() > 0 ) ? undulator_[0].lu_ / ( 2.0 * gamma_ * gamma_ ) * b.betaVector_[2] / beta_ : 0.0;
	printmessage(__FILE__, __LINE__, "Modulation wavelength of the bunch outside the undulator is set to " + stringify(b.lambda_));
      }
    seed_.beta_  = beta_;
    seed_.gamma_ = gamma_;
  }

  /* solver.cpp:508-539 */
  void Solver::computeFileGamma (BunchInitialize& b)
  {
    Double ignore;
    FieldVector gb (0.0);
    b.initialGamma_ = 0.0;
    b.initialDirection_ = FieldVector(0.0);
    std::ifstream in ( b.fileName_.c_str() );
    while (in.good())
      {
	in >> ignore; in >> ignore; in >> ignore;
	in >> gb[0]; in >> gb[1]; in >> gb[2];
	b.initialGamma_ += std::sqrt( 1 + gb.norm2() );
	b.initialDirection_ += gb;
      }
    b.initialGamma_ /= b.numberOfParticles_;
    b.initialDirection_ /= b.initialDirection_.norm();
    printmessage(__FILE__, __LINE__, "Computed average gamma from file is " + stringify(b.initialGamma_));
  }

  /* solver.cpp:220-257 */
  void Solver::lorentzBoostMesh ()
  {
    mesh_.meshLength_[2]     *= gamma_;
    mesh_.meshResolution_[2] *= gamma_;
    mesh_.meshCenter_[2]     *= gamma_;
    mesh_.totalTime_         /= gamma_;
    mesh_.timeShift_         /= gamma_;
    std::vector<Double>& d = mesh_.meshResolution_;
    if ( mesh_.solver_ == NSFD )
      {
	/* the transverse cells are enlarged until the NSFD stability margin holds, then dt = dz / c               */
	const Double t = 1.0 / sqrt( pow( d[2] / d[0], 2.0 ) + pow( d[2] / d[1], 2.0 ) );
	if ( t < 1.02 )
	  {
	    d[0] *= 1.02 / t;
	    d[1] *= 1.02 / t;
	    printmessage(__FILE__, __LINE__, "Transverse discretization is set to " + stringify(d[0]) + " x " + stringify(d[1]));
	  }
	mesh_.timeStep_ = d[2] / c0_;
      }
    else
      mesh_.timeStep_ = 0.98 / ( c0_ * sqrt( 1.0 / pow(d[0], 2.0) + 1.0 / pow(d[1], 2.0) + 1.0 / pow(d[2], 2.0) ) );
    printmessage(__FILE__, __LINE__, "Time step for the field update is set to " + stringify(mesh_.timeStep_ * gamma_));
  }

  /* node counts, slab extents (one slab per GPU), mesh borders -- solver.cpp:601-689                             */
  void Solver::initializeMesh ()
  {
    N0_ = (int) ( mesh_.meshLength_[0] / mesh_.meshResolution_[0] ) + 2;
    N1_ = (int) ( mesh_.meshLength_[1] / mesh_.meshResolution_[1] ) + 2;
    N2_ = (int) ( mesh_.meshLength_[2] / mesh_.meshResolution_[2] ) + 2;
    N1N0_ = N1_ * N0_;
    mesh_.meshLength_[0] = ( N0_ - 1 ) * mesh_.meshResolution_[0];
    mesh_.meshLength_[1] = ( N1_ - 1 ) * mesh_.meshResolution_[1];
    mesh_.meshLength_[2] = ( N2_ - 1 ) * mesh_.meshResolution_[2];

    xmin_ = mesh_.meshCenter_[0] - mesh_.meshLength_[0] / 2.0;
    xmax_ = mesh_.meshCenter_[0] + mesh_.meshLength_[0] / 2.0;
    ymin_ = mesh_.meshCenter_[1] - mesh_.meshLength_[1] / 2.0;
    ymax_ = mesh_.meshCenter_[1] + mesh_.meshLength_[1] / 2.0;
    zmin_ = mesh_.meshCenter_[2] - mesh_.meshLength_[2] / 2.0;
    zmax_ = mesh_.meshCenter_[2] + mesh_.meshLength_[2] / 2.0;

    if ( size_ > 1 && N2_ / size_ < 6 )
      { printmessage(__FILE__, __LINE__, "The mesh has too few planes along z for " + stringify(size_) + " slabs."); exit(1); }
    slabNp_.assign(size_, 0); slabK0_.assign(size_, 0); slabZp0_.assign(size_, 0.0); slabZp1_.assign(size_, 0.0);
    for (int r = 0; r < size_; r++)
      {
	/* np_, k0_ of rank r (solver.cpp:619-641): two planes shared with each neighbour                          */
	int np, k0;
	if      ( size_ == 1 )      { np = N2_;                                          k0 = 0; }
	else if ( r == 0 )          { np = N2_ / size_ + 1;                              k0 = 0; }
	else if ( r == size_ - 1 )  { np = N2_ - ( size_ - 1 ) * ( N2_ / size_ ) + 1;    k0 = ( size_ - 1 ) * ( N2_ / size_ ) - 1; }
	else                        { np = N2_ / size_ + 2;                              k0 = r * ( N2_ / size_ ) - 1; }
	slabNp_[r] = np; slabK0_[r] = k0;
	/* ownership interval: z of local plane 0 and of local plane np-2 (np-1 on the last slab), solver.cpp:677-680 */
	slabZp0_[r] = zmin_ + ( 0 + k0 ) * mesh_.meshResolution_[2];
	slabZp1_[r] = zmin_ + ( ( np - ( ( r == size_ - 1 ) ? 1 : 2 ) ) + k0 ) * mesh_.meshResolution_[2];
      }
    np_ = slabNp_[0]; k0_ = slabK0_[0];
    /* the process as a whole owns [z(0), z(N2-1)), what a single rank of the reference owns                       */
    zp_[0] = slabZp0_[0]; zp_[1] = slabZp1_[size_ - 1];

    timep1_ =  mesh_.timeStep_;
    time_   =  0.0;
    timem1_ = -mesh_.timeStep_;
  }

  FieldVector Solver::rc (const long int& m)
  {
    const unsigned int k = m / N1N0_, i = ( m % N1N0_ ) / N1_, j = m - ( N1N0_ * k + N1_ * i );
    FieldVector v;
    v[0] = xmin_ + i          * mesh_.meshResolution_[0];
    v[1] = ymin_ + j          * mesh_.meshResolution_[1];
    v[2] = zmin_ + (k + k0_)  * mesh_.meshResolution_[2];
    return v;
  }

  bool Solver::particleInProcessor (const Double& z)
  {
    const Double zr = pmod( z - zmin_ , mesh_.meshLength_[2] ) + zmin_;
    return ( ( zr >= zp_[0] ) && ( zr < zp_[1] ) );
  }

  /* solver.cpp:1130-1178; the process generates the whole bunch (rank 0 of 1)                                    */
  bool Solver::wantDeviceBunch () const
  {
    if ( getenv("MITHRA_HOST_BUNCH") || bunch_.bunchInit_.size() != 1 ) return false;
    const BunchInitialize& b = bunch_.bunchInit_[0];
    if ( b.bunchType_ != "ellipsoid" || b.generator_ == "random" || b.shotNoise_ || b.position_.size() > 1 ) return false;
    if ( b.distribution_ != "uniform" && b.distribution_ != "gaussian" ) return false;
    /* what would need sums over the bunch in the reference's (sequential) order stays on the host                   */
    if ( mesh_.optimizePosition_ || mesh_.totalDist_ > 0.0 || mesh_.timeShift_ != 0.0 ) return false;
    if ( b.numberOfParticles_ < ( 1u << 20 ) && !getenv("MITHRA_DEVICE_BUNCH") ) return false;
    return mithra_gpu_device_count() > 0;
  }

  void Solver::initializeBunch ()
  {
    printmessage(__FILE__, __LINE__, "[[[ Initializing the bunch and prepare the charge vector ");
    deviceBunch_ = 0;
    if ( wantDeviceBunch() )
      {
	BunchInitialize& b = bunch_.bunchInit_[0];
	if ( b.position_.size() == 0 ) b.position_.push_back( FieldVector(0.0) );
	if ( b.numberOfParticles_ % 4 != 0 )                     /* classes.cpp:107-113 */
	  {
	    b.numberOfParticles_ += 4 - b.numberOfParticles_ % 4;
	    printmessage(__FILE__, __LINE__, "Warning: The number of particles in the bunch is not a multiple of four. It is corrected to " + stringify(b.numberOfParticles_));
	  }
	MithraBunchEllipsoid e; memset(&e, 0, sizeof(e));
	e.number_of_particles = b.numberOfParticles_; e.index_offset = 0;
	e.cloud_charge = b.cloudCharge_; e.initial_gamma = b.initialGamma_;
	for (int c = 0; c < 3; c++)
	  { e.beta_vector[c] = b.betaVector_[c]; e.position[c] = b.position_[0][c]; e.sigma_position[c] = b.sigmaPosition_[c]; e.sigma_gamma_beta[c] = b.sigmaGammaBeta_[c]; }
	e.tran_trun = b.tranTrun_; e.long_trun = b.longTrun_; e.lambda = b.lambda_; e.bunching_factor = b.bF_; e.bunching_phase = b.bFP_;
	e.distribution = ( b.distribution_ == "uniform" ) ? 0 : 1;
	e.device = -1;
	size_t n = 0;
	check(mithra_gpu_bunch_generate(&e, &deviceBunch_, &n));
	printmessage(__FILE__, __LINE__, "The bunch (" + stringify(n) + " macro-particles) is generated on the device. ]]]");
	return;
      }
    std::list<Charge> qv;
    for (BunchInitialize& b : bunch_.bunchInit_)
      {
	qv.clear();
	if ( b.position_.size() == 0 ) b.position_.push_back( FieldVector(0.0) );
	for (unsigned int ia = 0; ia < b.position_.size(); ia++)
	  {
	    if      ( b.bunchType_ == "manual" )     bunch_.initializeManual   (b, qv, zp_, 0, 1, ia);
	    else if ( b.bunchType_ == "ellipsoid" )  bunch_.initializeEllipsoid(b, qv, 0, 1, ia);
	    else if ( b.bunchType_ == "3D-crystal" ) bunch_.initialize3DCrystal(b, qv, zp_, 0, 1, ia);
	    else if ( b.bunchType_ == "file" )       bunch_.initializeFile     (b, qv, zp_, 0, 1, ia);
	  }
	if ( b.bunchType_ == "other" ) printmessage(__FILE__, __LINE__, "The charge vector has been filled in by an external program. ");
	chargeVectorn_.splice(chargeVectorn_.end(), qv);
      }
    printmessage(__FILE__, __LINE__, "The bunch is initialized and the charge vector is prepared. ]]]");
  }

  /* bunch time step, boost of the particles, time origin dt_, ballistic back-projection -- solver.cpp:263-423     */
  void Solver::lorentzBoostBunch ()
  {
    bunch_.timeStep_ /= gamma_;
    if ( bunch_.timeStep_ == 0 ) bunch_.timeStep_ = mesh_.timeStep_;
    else                         bunch_.timeStep_ = mesh_.timeStep_ / ceil( mesh_.timeStep_ / bunch_.timeStep_ );
    nUpdateBunch_ = mesh_.timeStep_ / bunch_.timeStep_;
    printmessage(__FILE__, __LINE__, "Time step for the bunch update is set to " + stringify(bunch_.timeStep_ * gamma_));

    bunch_.rhythm_ /= gamma_; bunch_.bunchVTKRhythm_ /= gamma_; bunch_.bunchProfileRhythm_ /= gamma_;
    for (Double& t : bunch_.bunchProfileTime_) t /= gamma_;

    Double zmaxG = -1.0e100;
    for (Charge& q : chargeVectorn_)
      {
	const Double g  = std::sqrt( 1.0 + q.gb.norm2() );
	const Double bz = q.gb[2] / g;
	q.rnp[2] *= gamma_;
	q.gb[2]   = gamma_ * g * ( bz - beta_ );
	zmaxG     = std::max( zmaxG , q.rnp[2] );
      }
    if ( deviceBunch_ ) check(mithra_gpu_bunch_boost(deviceBunch_, gamma_, beta_, &zmaxG));

    /* at bunch time zero the bunch head is undulator_[0].dist_ (lab frame) in front of the entrance              */
    if ( undulator_.size() > 0 )
      {
	Undulator& u0 = undulator_[0];
	const Double nl = ( u0.type_ == STATIC ) ? 2.0 : 5.0 * u0.signal_.nR_;
	if      ( u0.dist_ == 0.0 )         u0.dist_ = nl * u0.lu_;
	else if ( u0.dist_ < nl * u0.lu_ )  printmessage(__FILE__, __LINE__, "Warning: the undulator is set very close to the bunch, the results may be inaccurate.");
	dt_ = - 1.0 / ( beta_ * u0.c0_ ) * ( zmaxG + u0.dist_ / gamma_ );
	printmessage(__FILE__, __LINE__, "Initial distance from bunch head to undulator is " + stringify(u0.dist_));
      }
    seed_.dt_    = dt_;
    bunch_.zu_   = zmaxG;
    bunch_.beta_ = beta_;

    /* the bunch properties are given at the start point: move the particles back along straight lines             */
    for (Charge& q : chargeVectorn_)
      {
	const Double g = std::sqrt( 1.0 + q.gb.norm2() );
	q.rnp[0] += q.gb[0] / g * ( q.rnp[2] - bunch_.zu_ ) * beta_;
	q.rnp[1] += q.gb[1] / g * ( q.rnp[2] - bunch_.zu_ ) * beta_;
	q.rnp[2] += q.gb[2] / g * ( q.rnp[2] - bunch_.zu_ ) * beta_;
      }
    if ( deviceBunch_ ) check(mithra_gpu_bunch_backproject(deviceBunch_, bunch_.zu_, beta_));

    if ( mesh_.optimizePosition_ && undulator_.size() > 0 )
      {
	Double zG = 0.0, bzG = 0.0;
	for (Charge& q : chargeVectorn_) { zG += q.rnp[2]; bzG += q.gb[2] / std::sqrt( 1 + q.gb.norm2() ); }
	const unsigned int NqG = chargeVectorn_.size();
	zG /= NqG; bzG /= NqG;
	const Double shift = bzG * ( zmaxG + undulator_[0].dist_ / gamma_ - zG ) / ( bzG + beta_ ) + zG;
	zmaxG     -= shift;
	bunch_.zu_ = zmaxG;
	dt_        = - 1.0 / ( beta_ * undulator_[0].c0_ ) * ( zmaxG + undulator_[0].dist_ / gamma_ );
	seed_.dt_  = dt_;
	for (Charge& q : chargeVectorn_) q.rnp[2] -= shift;
	printmessage(__FILE__, __LINE__, "The bunch center is shifted back by " + stringify(shift) + " .");
      }

    distributeParticles(chargeVectorn_);
    Nc_ = chargeVectorn_.size();
    if ( deviceBunch_ ) { size_t n = 0; check(mithra_gpu_bunch_download(deviceBunch_, 0, 0, &n)); Nc_ = n; }
    printmessage(__FILE__, __LINE__, "The total number of macro-particles is equal to " + stringify(Nc_) + " .");

    if ( mesh_.totalDist_ > 0.0 )
      {
	double Lu = 0.0;
	for (Undulator& u : undulator_) Lu += u.lu_ * u.length_ / gamma_;
	const double zEnd = mesh_.totalDist_ / gamma_;
	double zMin = 1e100, bz = 0;
	for (Charge& q : chargeVectorn_) { zMin = std::min(zMin, q.rnp[2]); bz += q.gb[2] / std::sqrt( 1 + q.gb.norm2() ); }
	bz /= chargeVectorn_.size();
	mesh_.totalTime_ = 1 / ( c0_ * ( bz + beta_ ) ) * ( zEnd - beta_ * c0_ * dt_ - zMin + bz / beta_ * Lu );
	printmessage(__FILE__, __LINE__, "The total time to simulate has been set to " + stringify(mesh_.totalTime_ * gamma_) + " .");
      }
  }

  /* A single rank of the reference keeps every particle whose wrapped z lies inside the mesh; one that does not is
   * re-queued behind the others with rnm and e reset (solver.cpp:429-487).  The split over the slabs happens at
   * upload time (attachGpu).                                                                                     */
  void Solver::distributeParticles (std::list<Charge>& chargeVector)
  {
    std::list<Charge> moved;
    for (std::list<Charge>::iterator it = chargeVector.begin(); it != chargeVector.end(); )
      {
	if ( particleInProcessor( it->rnp[2] ) ) { ++it; continue; }
	Charge c; c.q = it->q; c.rnp = it->rnp; c.gb = it->gb;
	moved.push_back(c);
	it = chargeVector.erase(it);
      }
    for (Charge& c : moved) if ( particleInProcessor( c.rnp[2] ) ) chargeVector.push_back(c);
  }

  /* every coefficient of the field update, solver.cpp:695-824                                                    */
  void Solver::initializeField ()
  {
    uc_.dx = mesh_.meshResolution_[0]; uc_.dy = mesh_.meshResolution_[1]; uc_.dz = mesh_.meshResolution_[2];
    uc_.dv = - m0_ * EC / mesh_.timeStep_ / ( uc_.dx * uc_.dy * uc_.dz );
    uc_.rc = - EC / e0_ /                   ( uc_.dx * uc_.dy * uc_.dz );

    UpdateField& u = uf_;
    u.dt = mesh_.timeStep_; u.dx = mesh_.meshResolution_[0]; u.dy = mesh_.meshResolution_[1]; u.dz = mesh_.meshResolution_[2];
    u.dx2 = 2.0 * u.dx; u.dy2 = 2.0 * u.dy; u.dz2 = 2.0 * u.dz;
    const Double c = c0_, dt = u.dt, dx = u.dx, dy = u.dy, dz = u.dz;

    /* non-standard finite difference weights (doc MITHRA_FDTDPIC.tex:244-277)                                     */
    const Double beta  = ( 1.0 + 0.02 / ( pow(dz/dx,2.0) + pow(dz/dy,2.0) ) ) / 4.0;
    const Double alpha = 1.0 - 2.0 * beta;
    u.alpha = alpha;
    u.beta  = beta / alpha;
    u.a[0] = 2.0 * ( 1.0 - alpha * pow(c*dt/dx,2) - alpha * pow(c*dt/dy,2) - pow(c*dt/dz,2) );
    u.a[1] = pow(c*dt/dx,2.0);
    u.a[2] = pow(c*dt/dy,2.0);
    u.a[3] = pow(c*dt/dz,2.0) - 2.0 * ( beta * pow(c*dt/dx,2) + beta * pow(c*dt/dy,2) );
    u.a[4] = pow(c*dt,2.0) * uc_.dv;
    u.a[5] = pow(c*dt,2.0) * uc_.rc;

    /* faces: first / second order absorbing condition at normal incidence (alpha1 = alpha2 = 0)                   */
    const Double alpha1 = 0.0, alpha2 = 0.0;
    const Double p = ( 1.0 + cos(alpha1) * cos(alpha2) ) / ( cos(alpha1) + cos(alpha2) );
    const Double q = - 1.0 / ( cos(alpha1) + cos(alpha2) );
    const Double o = mesh_.truncationOrder_ - 1.0;
    auto face = [&] (Double* B, Double dn, Double dt1, Double dt2) {
      const Double d = 1.0 / ( 2.0 * dt * dn ) + p / ( 2.0 * c * dt * dt );
      B[0] = (   1.0 / ( 2.0 * dt * dn ) - p / ( 2.0 * c * dt * dt ) ) / d;
      B[1] = ( - 1.0 / ( 2.0 * dt * dn ) - p / ( 2.0 * c * dt * dt ) ) / d;
      B[2] = (   p / ( c * dt * dt ) + q * o * ( c / ( dt1 * dt1 ) + c / ( dt2 * dt2 ) ) ) / d;
      B[3] = - q * o * ( c / ( 2.0 * dt1 * dt1 ) ) / d ;
      B[4] = - q * o * ( c / ( 2.0 * dt2 * dt2 ) ) / d ; };
    face(u.bB, dx, dy, dz);
    face(u.cB, dy, dx, dz);
    face(u.dB, dz, dx, dy);

    /* edges along w between the faces normal to u and v: (da, db) = (d_v, d_u) in the reference's naming            */
    auto edge = [&] (Double* E, Double da, Double db, Double dw) {
      const Double d = ( 1.0 / da + 1.0 / db ) / ( 4.0 * dt ) + 3.0 / ( 8.0 * c * dt * dt );
      E[0] = ( - ( 1.0 / da - 1.0 / db ) / ( 4.0 * dt ) - 3.0 / ( 8.0 * c * dt * dt ) ) / d;
      E[1] = (   ( 1.0 / da - 1.0 / db ) / ( 4.0 * dt ) - 3.0 / ( 8.0 * c * dt * dt ) ) / d;
      E[2] = (   ( 1.0 / da + 1.0 / db ) / ( 4.0 * dt ) - 3.0 / ( 8.0 * c * dt * dt ) ) / d;
      E[3] = ( 3.0 / ( 4.0 * c * dt * dt ) - c / ( 4.0 * dw * dw ) ) / d;
      E[4] = c / ( 8.0 * dw * dw ) / d; };
    edge(u.eE, dy, dx, dz);
    edge(u.fE, dz, dy, dx);
    edge(u.gE, dx, dz, dy);

    /* corners: signs of (1/dx, 1/dy, 1/dz) for hC[0..7]; hC[8+n] mirrors hC[n]                                    */
    static const int sgn[8][3] = { {-1,-1,-1}, {1,-1,-1}, {-1,1,-1}, {-1,-1,1}, {1,1,-1}, {1,-1,1}, {-1,1,1}, {1,1,1} };
    for (int n = 0; n < 8; n++)
      {
	const Double s = ( sgn[n][0] * 1.0 / dx + sgn[n][1] * 1.0 / dy + sgn[n][2] * 1.0 / dz );
	u.hC[n]     =   s / ( 8.0 * dt ) - 1.0 / ( 4.0 * c * dt * dt );
	u.hC[8 + n] = - s / ( 8.0 * dt ) - 1.0 / ( 4.0 * c * dt * dt );
      }
    u.hC[16] = 1.0 / ( 2.0 * c * dt * dt );
    /* the seed's initial condition inside the total-field box (solver.cpp:828-839) is set on the device by
     * mithra_gpu_seed_initial in attachGpu()                                                                      */
  }

  /* Solver::initializeSeedSampling, solver.cpp:848-931: boost of the rhythm and of the points, the points of an
   * over-line request, the filter to the mesh, the output file (single-rank naming: <base>-0.txt holds every point,
   * whatever the number of slabs) and the unit factors.                                                            */
  void Solver::initializeSeedSampling ()
  {
    printmessage(__FILE__, __LINE__, " ::: Initializing the field sampling data");
    if ( seed_.samplingRhythm_ == 0 )
      { printmessage(__FILE__, __LINE__, "The sampling rhythm of the field is zero although sampling is activated !!!"); exit(1); }
    seed_.samplingRhythm_ /= gamma_;
    for (unsigned i = 0; i < seed_.samplingPosition_.size(); i++) seed_.samplingPosition_[i][2] *= gamma_;
    seed_.samplingLineBegin_[2] *= gamma_;
    seed_.samplingLineEnd_  [2] *= gamma_;
    if ( seed_.samplingType_ == OVERLINE )
      {
	FieldVector l = seed_.samplingLineEnd_;
	const Double n = seed_.samplingRes_;
	for (int d = 0; d < 3; d++) { l[d] -= seed_.samplingLineBegin_[d]; l[d] /= n; }
	for (unsigned i = 0; i < n; i++)
	  {
	    FieldVector position;
	    position[0] = seed_.samplingLineBegin_[0] + i * l[0];
	    position[1] = seed_.samplingLineBegin_[1] + i * l[1];
	    position[2] = seed_.samplingLineBegin_[2] + i * l[2];
	    seed_.samplingPosition_.push_back(position);
	  }
      }
    /* the reference reads ub_.dx, dy, dz here before initializeBunchUpdate has set them (solver.cpp:891-893 against
     * :1056): zero-initialised members, so the margin of the test is zero                                            */
    std::vector<FieldVector> kept;
    for (unsigned int n = 0; n < seed_.samplingPosition_.size(); ++n)
      {
	const FieldVector& q = seed_.samplingPosition_[n];
	if ( q[0] < xmax_ - ub_.dx && q[0] > xmin_ + ub_.dx && q[1] < ymax_ - ub_.dy && q[1] > ymin_ + ub_.dy &&
	     q[2] < zmax_ - ub_.dz && q[2] > zmin_ + ub_.dz )
	  kept.push_back(q);                                       /* every slab of this process: zmin_ <= z < zmax_       */
	else
	  printmessage(__FILE__, __LINE__, "The sampling point does not reside in the grid. No data is saved.");
      }
    seed_.samplingPosition_ = kept;
    if ( !kept.empty() )
      {
	std::string name = "";
	if ( seed_.samplingBasename_.compare(0, 1, "/") != 0 ) name = seed_.samplingDirectory_;
	name += seed_.samplingBasename_ + "-" + stringify(0) + ".txt";
	createDirectory(name, 0);
	fieldSampleFile_ = new std::ofstream(name.c_str(), std::ios::trunc);
      }
    sfCe_ = mesh_.lengthScale_ / pow( mesh_.timeScale_, 2 );
    sfCb_ = 1.0 / ( mesh_.lengthScale_ * mesh_.timeScale_ );
    sfCa_ = 1.0 / mesh_.timeScale_;
    printmessage(__FILE__, __LINE__, " The field sampling data are initialized. :::");
  }

  /* FdTd::fieldSample, fdtd.cpp:851-950 (FdTdSC::fieldSample writes the same columns): one line per call -- the time,
   * then per point its coordinates and the requested fields; the interpolated et, bt, at come from the slab that
   * holds the point (mithra_gpu_field_sample), the lab-frame combinations are the reference's expressions.          */
  void FdTd::fieldSample ()
  {
    const size_t N = seed_.samplingPosition_.size();
    if ( N == 0 || !fieldSampleFile_ ) return;
    std::vector<double> pos(3 * N), val(9 * N, 0.0), part(9 * N);
    std::vector<unsigned char> mine(N), have(N, 0);
    for (size_t n = 0; n < N; n++) for (int d = 0; d < 3; d++) pos[3 * n + d] = seed_.samplingPosition_[n][d];
    for (MithraGpu* g : gpu_)
      {
	check(mithra_gpu_field_sample(g, pos.data(), N, part.data(), mine.data()));
	for (size_t n = 0; n < N; n++)
	  if ( mine[n] && !have[n] ) { have[n] = 1; for (int q = 0; q < 9; q++) val[9 * n + q] = part[9 * n + q]; }
      }
    std::ofstream& f = *fieldSampleFile_;
    f.setf(std::ios::scientific);
    f.precision(4);
    f << time_ * gamma_ << "\t";
    for (size_t n = 0; n < N; n++)
      {
	const double* et = &val[9 * n]; const double* bt = et + 3; const double* at = et + 6;
	f << pos[3 * n] << "\t" << pos[3 * n + 1] << "\t" << pos[3 * n + 2] << "\t";
	for (unsigned int i = 0; i < seed_.samplingField_.size(); i++)
	  {
	    const FieldType t = seed_.samplingField_[i];
	    if      ( t == Ex ) f << ( gamma_ * et[0] + c0_ * sqrt( pow(gamma_, 2) - 1 ) * bt[1] ) * sfCe_ << "\t";
	    else if ( t == Ey ) f << ( gamma_ * et[1] - c0_ * sqrt( pow(gamma_, 2) - 1 ) * bt[0] ) * sfCe_ << "\t";
	    else if ( t == Ez ) f << et[2] * sfCe_ << "\t";
	    else if ( t == Bx ) f << ( gamma_ * bt[0] - sqrt( pow(gamma_, 2) - 1 ) / c0_ * et[1] ) * sfCb_ << "\t";
	    else if ( t == By ) f << ( gamma_ * bt[1] + sqrt( pow(gamma_, 2) - 1 ) / c0_ * et[0] ) * sfCb_ << "\t";
	    else if ( t == Bz ) f << bt[2] * sfCb_ << "\t";
	    else if ( t == Ax ) f << at[0] * sfCa_ << "\t";
	    else if ( t == Ay ) f << at[1] * sfCa_ << "\t";
	    else if ( t == Az ) f << at[2] * sfCa_ << "\t";
	  }
      }
    f << std::endl;
  }

  /* Solver::initializeSeedVTK, solver.cpp:938-1016: boost of rhythm and plane position, output directory, and the
   * test that the plane lies in the mesh (with the reference's still-zero ub_.dx, dy, dz)                         */
  void Solver::initializeSeedVTK ()
  {
    for (unsigned int i = 0; i < seed_.vtk_.size(); i++)
      {
	Seed::vtk& v = seed_.vtk_[i];
	if ( !v.sample_ ) continue;
	if ( v.rhythm_ == 0 )
	  { printmessage(__FILE__, __LINE__, "The visualization rhythm of the field is zero although visualization is activated !!!"); exit(1); }
	v.rhythm_      /= gamma_;
	v.position_[2] *= gamma_;
	if ( v.basename_.compare(0, 1, "/") != 0 ) v.basename_ = v.directory_ + v.basename_;
	createDirectory(v.basename_, 0);
	bool outside = false;
	if ( v.type_ == INPLANE )
	  {
	    if      ( v.plane_ == XNORMAL ) outside = ( v.position_[0] > xmax_ - ub_.dx || v.position_[0] < xmin_ + ub_.dx );
	    else if ( v.plane_ == YNORMAL ) outside = ( v.position_[1] > ymax_ - ub_.dy || v.position_[1] < ymin_ + ub_.dy );
	    else if ( v.plane_ == ZNORMAL ) outside = ( v.position_[2] > zmax_ - ub_.dz || v.position_[2] < zmin_ + ub_.dz );
	  }
	if ( outside )
	  {
	    printmessage(__FILE__, __LINE__, "The plane does not reside in the grid. No data is saved.");
	    seed_.vtk_.erase( seed_.vtk_.begin() + i );          /* like the reference, the loop index moves on regardless */
	  }
      }
  }

  /* en_[m], bn_[m], (*an_)[m] of global nodes (i, j, k) from the slab that holds each as one of its own planes        */
  void FdTd::nodeValues (const std::vector<int>& ijk, std::vector<double>& val)
  {
    const size_t n = ijk.size() / 3;
    val.assign(9 * n, 0.0);
    std::vector<double> part(9 * n);
    std::vector<unsigned char> mine(n), have(n, 0);
    std::vector<int> loc(ijk);
    for (size_t r = 0; r < gpu_.size(); r++)
      {
	for (size_t t = 0; t < n; t++) loc[3 * t + 2] = ijk[3 * t + 2] - slabK0_[r];
	check(mithra_gpu_field_nodes(gpu_[r], loc.data(), n, part.data(), mine.data()));
	for (size_t t = 0; t < n; t++)
	  if ( mine[t] && !have[t] ) { have[t] = 1; for (int q = 0; q < 9; q++) val[9 * t + q] = part[9 * t + q]; }
      }
  }

  namespace
  {
    /* column of the 9 node values a FieldType selects: en 0-2, bn 3-5, an 6-8                                       */
    inline int fieldColumn (FieldType t)
    { switch (t) { case Ex: return 0; case Ey: return 1; case Ez: return 2; case Bx: return 3; case By: return 4; case Bz: return 5;
		   case Ax: return 6; case Ay: return 7; case Az: return 8; default: return -1; } }
    /* en_, bn_ are floats in the reference: float x double products, summed in double (fdtd.cpp:1157-1165)          */
    inline double blend (const double* a, const double* b, int col, Double d)
    { return col < 0 ? 0.0 : a[col] * ( 1.0 - d ) + b[col] * d; }
  }

  void FdTd::fieldVisualizeInPlane (unsigned int ivtk)
  {
    if      ( seed_.vtk_[ivtk].plane_ == XNORMAL ) fieldVisualizeInPlaneXNormal(ivtk);
    else if ( seed_.vtk_[ivtk].plane_ == YNORMAL ) fieldVisualizeInPlaneYNormal(ivtk);
    else if ( seed_.vtk_[ivtk].plane_ == ZNORMAL ) fieldVisualizeInPlaneZNormal(ivtk);
  }

  /* the x- and y-normal planes share everything but the roles of i and j: one body, `xn` selects                      */
  static void writeSidePlane (FdTd& s, unsigned int ivtk, bool xn)
  {
    const Seed::vtk& V = s.seed_.vtk_[ivtk];
    const int N0 = s.N0_, N1 = s.N1_, N2 = s.N2_, NT = xn ? N1 : N0;          /* NT: nodes along the in-plane transverse axis */
    Double c;
    const Double dr = xn ? modf( ( V.position_[0] - s.xmin_ ) / s.mesh_.meshResolution_[0], &c )
			 : modf( ( V.position_[1] - s.ymin_ ) / s.mesh_.meshResolution_[1], &c );
    const int fixed = (int) c;
    const size_t nf = V.field_.size();

    /* the values: v[n][l], n = k NT + t, zero where the reference leaves them (t = 0, NT-1)                           */
    std::vector<int> ijk; ijk.reserve((size_t) 6 * N2 * NT);
    for (int k = 0; k < N2; k++)
      for (int t = 1; t < NT - 1; t++)
	{
	  const int i = xn ? fixed : t, j = xn ? t : fixed;
	  ijk.push_back(i); ijk.push_back(j); ijk.push_back(k);
	  ijk.push_back(xn ? i + 1 : i); ijk.push_back(xn ? j : j + 1); ijk.push_back(k);
	}
    std::vector<double> val;
    s.nodeValues(ijk, val);
    std::vector<double> v((size_t) N2 * NT * nf, 0.0);
    size_t q = 0;
    for (int k = 0; k < N2; k++)
      for (int t = 1; t < NT - 1; t++, q += 2)
	for (size_t l = 0; l < nf; l++)
	  v[((size_t) k * NT + t) * nf + l] = blend(&val[9 * q], &val[9 * (q + 1)], fieldColumn(V.field_[l]), dr);

    std::string name = V.basename_ + "-p" + stringify(0) + "-" + stringify(s.nTime_) + ".vts";
    {
      std::ofstream f(name.c_str(), std::ios::trunc);
      f.setf(std::ios::scientific);
      f.precision(4);
      f << "<?xml version=\"1.0\"?>" << std::endl;
      f << "<VTKFile type=\"StructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" compressor=\"vtkZLibDataCompressor\">" << std::endl;
      if ( xn )
	{
	  f << "<StructuredGrid WholeExtent=\"0  0  0 " << N1 - 1 << " " << 0 << " " << N2 - 1 << "\">" << std::endl;
	  f << "<Piece Extent=\" 0 0 0 " << N1 - 1 << " " << 0 << " " << N2 - 1 << "\">" << std::endl;
	}
      else
	{
	  f << "<StructuredGrid WholeExtent=\"0 " << N0 - 1 << " 0 0 " << 0 << " " << N2 - 1 << "\">" << std::endl;
	  f << "<Piece Extent=\"0 " << N0 - 1 << " 0 0 " << 0 << " " << N2 - 1 << "\">" << std::endl;
	}
      f << "<Points>" << std::endl;
      f << "<DataArray type = \"Float64\" NumberOfComponents=\"3\" format=\"ascii\">" << std::endl;
      for (int k = 0; k < N2; k++)
	for (int t = 0; t < NT; t++)
	  {
	    const long int m = (long int) k * s.N1N0_ + ( xn ? fixed : t ) * N1 + ( xn ? t : fixed );
	    const FieldVector r1 = s.rc(m), r2 = s.rc(m + ( xn ? N1 : 1 ));
	    f << r1[0] * ( 1.0 - dr ) + r2[0] * dr << " " << r1[1] << " " << r1[2] << std::endl;
	  }
      f << "</DataArray>" << std::endl;
      f << "</Points>" << std::endl;
      f << "<CellData>" << std::endl;
      f << "</CellData>" << std::endl;
      f << "<PointData Vectors = \"field\">" << std::endl;
      f << "<DataArray type=\"Float64\" Name=\"field\" NumberOfComponents=\"" << nf << "\" format=\"ascii\">" << std::endl;
      for (int k = 0; k < N2; k++)
	for (int t = 0; t < NT; t++)
	  {
	    const size_t n = (size_t) k * NT + t;
	    f << v[n * nf];
	    for (size_t l = 1; l < nf; l++) f << " " << v[n * nf + l];
	    f << std::endl;
	  }
      f << "</DataArray>" << std::endl;
      f << "</PointData>" << std::endl;
      f << "</Piece>" << std::endl;
      f << "</StructuredGrid>" << std::endl;
      f << "</VTKFile>" << std::endl;
    }
    /* the file that ties the pieces together: one piece, the single-rank layout                                      */
    name = V.basename_ + "-" + stringify(s.nTime_) + ".pvts";
    const std::string base = V.basename_.substr(V.basename_.find_last_of("/") + 1);
    std::ofstream f(name.c_str(), std::ios::trunc);
    f << "<?xml version=\"1.0\"?>" << std::endl;
    f << "<VTKFile type=\"PStructuredGrid\" version=\"0.1\" >" << std::endl;
    if ( xn ) f << "<PStructuredGrid WholeExtent=\" 0 0 0 " << N1 - 1 << " 0 " << N2 - 1 << "\" GhostLevel = \"0\" >" << std::endl;
    else      f << "<PStructuredGrid WholeExtent=\"0 " << N0 - 1 << " 0 0 0 " << N2 - 1 << "\" GhostLevel = \"0\" >" << std::endl;
    f << "<PPoints>" << std::endl;
    f << "<DataArray type = \"Float64\" NumberOfComponents=\"3\" format=\"ascii\" />" << std::endl;
    f << "</PPoints>" << std::endl;
    f << "<PPointData>" << std::endl;
    f << "<DataArray type=\"Float64\" NumberOfComponents=\"" << nf << "\" Name=\"field\" format=\"ascii\" />" << std::endl;
    f << "</PPointData>" << std::endl;
    const std::string piece = base + "-p" + stringify(0) + "-" + stringify(s.nTime_) + ".vts";
    if ( xn ) f << "<Piece Extent=\"0 0 0 " << N1 - 1 << " " << 0 << " " << N2 - 2 + 1 << "\"" << " Source=\"" << piece << "\" />" << std::endl;
    else      f << "<Piece Extent=\"0 " << N0 - 1 << " 0 0 " << 0 << " " << N2 - 2 + 1 << "\"" << " Source=\"" << piece << "\" />" << std::endl;
    f << "</PStructuredGrid>" << std::endl;
    f << "</VTKFile>" << std::endl;
  }

  void FdTd::fieldVisualizeInPlaneXNormal (unsigned int ivtk) { writeSidePlane(*this, ivtk, true); }
  void FdTd::fieldVisualizeInPlaneYNormal (unsigned int ivtk) { writeSidePlane(*this, ivtk, false); }

  /* fdtd.cpp:1452-1540: one piece, no .pvts; the x coordinate of the points is blended like the reference does       */
  void FdTd::fieldVisualizeInPlaneZNormal (unsigned int ivtk)
  {
    const Seed::vtk& V = seed_.vtk_[ivtk];
    Double c;
    const Double dzr = modf( ( V.position_[2] - zmin_ ) / mesh_.meshResolution_[2], &c );
    const int k = (int) c;
    const size_t nf = V.field_.size();
    std::vector<int> ijk; ijk.reserve((size_t) 6 * N0_ * N1_);
    for (int j = 1; j < N1_ - 1; j++)
      for (int i = 1; i < N0_ - 1; i++)
	{ ijk.push_back(i); ijk.push_back(j); ijk.push_back(k); ijk.push_back(i); ijk.push_back(j); ijk.push_back(k + 1); }
    std::vector<double> val;
    nodeValues(ijk, val);
    std::vector<double> v((size_t) N1N0_ * nf, 0.0);
    size_t q = 0;
    for (int j = 1; j < N1_ - 1; j++)
      for (int i = 1; i < N0_ - 1; i++, q += 2)
	for (size_t l = 0; l < nf; l++)
	  v[((size_t) i * N1_ + j) * nf + l] = blend(&val[9 * q], &val[9 * (q + 1)], fieldColumn(V.field_[l]), dzr);

    const std::string name = V.basename_ + "-p" + stringify(0) + "-" + stringify(nTime_) + ".vts";
    std::ofstream f(name.c_str(), std::ios::trunc);
    f.setf(std::ios::scientific);
    f.precision(4);
    f << "<?xml version=\"1.0\"?>" << std::endl;
    f << "<VTKFile type=\"StructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" compressor=\"vtkZLibDataCompressor\">" << std::endl;
    f << "<StructuredGrid WholeExtent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " " << 0 << " " << 0 << "\">" << std::endl;
    f << "<Piece Extent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " " << 0 << " " << 0 << "\">" << std::endl;
    f << "<Points>" << std::endl;
    f << "<DataArray type = \"Float64\" NumberOfComponents=\"3\" format=\"ascii\">" << std::endl;
    for (int j = 0; j < N1_; j++)
      for (int i = 0; i < N0_; i++)
	{
	  const long int m = (long int) k * N1N0_ + i * N1_ + j;
	  const FieldVector r1 = rc(m), r2 = rc(m + N1N0_);
	  f << r1[0] * ( 1.0 - dzr ) + r2[0] * dzr << " " << r1[1] << " " << r1[2] << std::endl;
	}
    f << "</DataArray>" << std::endl;
    f << "</Points>" << std::endl;
    f << "<CellData>" << std::endl;
    f << "</CellData>" << std::endl;
    f << "<PointData Vectors = \"field\">" << std::endl;
    f << "<DataArray type=\"Float64\" Name=\"field\" NumberOfComponents=\"" << nf << "\" format=\"ascii\">" << std::endl;
    for (int j = 0; j < N1_; j++)
      for (int i = 0; i < N0_; i++)
	{
	  const size_t n = (size_t) i * N1_ + j;
	  f << v[n * nf];
	  for (size_t l = 1; l < nf; l++) f << " " << v[n * nf + l];
	  f << std::endl;
	}
    f << "</DataArray>" << std::endl;
    f << "</PointData>" << std::endl;
    f << "</Piece>" << std::endl;
    f << "</StructuredGrid>" << std::endl;
    f << "</VTKFile>" << std::endl;
  }

  /* FdTd::fieldVisualizeAllDomain, fdtd.cpp:956-1105: the requested fields on every node of the mesh as one ASCII .vts piece
   * (single-rank naming, "-p0-") plus the .pvts.  The reference evaluates E/B on every node with 1 <= i <= N0-2,
   * 1 <= j <= N1-2 at this moment (its lazy flags are bypassed) and leaves the transverse boundary nodes at zero: the
   * same here, node by node through mithra_gpu_field_nodes.  On the two end planes of the mesh the reference's
   * fieldEvaluate reads beyond its arrays; here they carry the E/B of their interior neighbour plane, like everywhere
   * else in this build (fdtd.cpp:754-773) -- the A columns are the reference's on every node.                         */
  void FdTd::fieldVisualizeAllDomain (unsigned int ivtk)
  {
    const Seed::vtk& V = seed_.vtk_[ivtk];
    const size_t nf = V.field_.size();
    const std::string name = V.basename_ + "-p" + stringify(0) + "-" + stringify(nTime_) + ".vts";
    std::ofstream f(name.c_str(), std::ios::trunc);
    f.setf(std::ios::scientific);
    f.precision(4);
    f << "<?xml version=\"1.0\"?>" << std::endl;
    f << "<VTKFile type=\"StructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" compressor=\"vtkZLibDataCompressor\">" << std::endl;
    f << "<StructuredGrid WholeExtent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " " << 0 << " " << N2_ - 2 + 1 << "\">" << std::endl;
    f << "<Piece Extent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " " << 0 << " " << N2_ - 2 + 1 << "\">" << std::endl;
    f << "<Points>" << std::endl;
    f << "<DataArray type = \"Float64\" NumberOfComponents=\"3\" format=\"ascii\">" << std::endl;
    for (int k = 0; k < N2_; k++)
      for (int j = 0; j < N1_; j++)
	for (int i = 0; i < N0_; i++)
	  {
	    const FieldVector r = rc((long int) k * N1N0_ + i * N1_ + j);
	    f << r[0] << " " << r[1] << " " << r[2] << std::endl;
	  }
    f << "</DataArray>" << std::endl;
    f << "</Points>" << std::endl;
    f << "<CellData>" << std::endl;
    f << "</CellData>" << std::endl;
    f << "<PointData Vectors = \"field\">" << std::endl;
    f << "<DataArray type=\"Float64\" Name=\"field\" NumberOfComponents=\"" << nf << "\" format=\"ascii\">" << std::endl;
    /* a few planes at a time: nine doubles per node come back from the device                                        */
    const int chunk = std::max(1, (int) ( 2000000L / N1N0_ ));
    std::vector<int> ijk; std::vector<double> val;
    for (int k0 = 0; k0 < N2_; k0 += chunk)
      {
	const int k1 = std::min(N2_, k0 + chunk);
	ijk.clear();
	for (int k = k0; k < k1; k++)
	  for (int j = 0; j < N1_; j++)
	    for (int i = 0; i < N0_; i++) { ijk.push_back(i); ijk.push_back(j); ijk.push_back(k); }
	nodeValues(ijk, val);
	size_t q = 0;
	for (int k = k0; k < k1; k++)
	  for (int j = 0; j < N1_; j++)
	    for (int i = 0; i < N0_; i++, q++)
	      {
		/* vf_.v stays zero on the transverse boundary, A included (the loop of fdtd.cpp:969-971 skips it)         */
		const bool inner = ( i >= 1 && i <= N0_ - 2 && j >= 1 && j <= N1_ - 2 );
		for (size_t l = 0; l < nf; l++)
		  {
		    const int col = fieldColumn(V.field_[l]);
		    const double v = ( inner && col >= 0 ) ? ( col < 6 ? (double) (float) val[9 * q + col] : val[9 * q + col] ) : 0.0;
		    f << ( l ? " " : "" ) << v;
		  }
		f << std::endl;
	      }
      }
    f << "</DataArray>" << std::endl;
    f << "</PointData>" << std::endl;
    f << "</Piece>" << std::endl;
    f << "</StructuredGrid>" << std::endl;
    f << "</VTKFile>" << std::endl;
    f.close();

    const std::string pname = V.basename_ + "-" + stringify(nTime_) + ".pvts";
    const std::string base = V.basename_.substr(V.basename_.find_last_of("/") + 1);
    std::ofstream g(pname.c_str(), std::ios::trunc);
    g << "<?xml version=\"1.0\"?>" << std::endl;
    g << "<VTKFile type=\"PStructuredGrid\" version=\"0.1\" >" << std::endl;
    g << "<PStructuredGrid WholeExtent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " 0 " << N2_ - 1 << "\" GhostLevel = \"0\" >" << std::endl;
    g << "<PPoints>" << std::endl;
    g << "<DataArray type = \"Float64\" NumberOfComponents=\"3\" format=\"ascii\" />" << std::endl;
    g << "</PPoints>" << std::endl;
    g << "<PPointData>" << std::endl;
    g << "<DataArray type=\"Float64\" NumberOfComponents=\"" << nf << "\" Name=\"field\" format=\"ascii\" />" << std::endl;
    g << "</PPointData>" << std::endl;
    g << "<Piece Extent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " " << 0 << " " << N2_ - 2 + 1 << "\"" << " Source=\""
      << base + "-p" + stringify(0) + "-" + stringify(nTime_) + ".vts" << "\" />" << std::endl;
    g << "</PStructuredGrid>" << std::endl;
    g << "</VTKFile>" << std::endl;
  }

  /* Solver::initializeSeedProfile, solver.cpp:1022-1044 */
  void Solver::initializeSeedProfile ()
  {
    if ( seed_.profileRhythm_ == 0 && seed_.profileTime_.size() == 0 )
      { printmessage(__FILE__, __LINE__, "The profiling rhythm of the field is zero and no time is set although profiling of the field is activated !!!"); exit(1); }
    seed_.profileRhythm_ /= gamma_;
    for (unsigned i = 0; i < seed_.profileTime_.size(); i++) seed_.profileTime_[i] /= gamma_;
    if ( seed_.profileBasename_.compare(0, 1, "/") != 0 ) seed_.profileBasename_ = seed_.profileDirectory_ + seed_.profileBasename_;
    createDirectory(seed_.profileBasename_, 0);
  }

  /* FdTd::fieldProfile, fdtd.cpp:1546-1594: x y z and the requested fields of every node, i outermost, one text file
   * ("-p0-").  The reference prints en_ / bn_ as its lazy evaluation last left them -- fresh only on the nodes a particle
   * or a sampler touched in this step, stale or zero elsewhere; this build prints the E/B of THIS step on every node
   * (zero on the transverse boundary, where the reference never evaluates either).  The A columns are the reference's.  */
  void FdTd::fieldProfile ()
  {
    const std::string name = seed_.profileBasename_ + "-p" + stringify(0) + "-" + stringify(nTime_) + ".txt";
    std::ofstream f(name.c_str(), std::ios::trunc);
    f.setf(std::ios::scientific);
    f.precision(4);
    const size_t nf = seed_.profileField_.size();
    std::vector<int> ijk; std::vector<double> val;
    const int chunk = std::max(1, (int) ( 2000000L / ( (long) N1_ * N2_ ) ));          /* rows i at a time          */
    for (int i0 = 0; i0 < N0_; i0 += chunk)
      {
	const int i1 = std::min(N0_, i0 + chunk);
	ijk.clear();
	for (int i = i0; i < i1; i++)
	  for (int j = 0; j < N1_; j++)
	    for (int k = 0; k < N2_; k++) { ijk.push_back(i); ijk.push_back(j); ijk.push_back(k); }
	nodeValues(ijk, val);
	size_t q = 0;
	for (int i = i0; i < i1; i++)
	  for (int j = 0; j < N1_; j++)
	    for (int k = 0; k < N2_; k++, q++)
	      {
		const FieldVector r = rc((long int) k * N1N0_ + i * N1_ + j);
		f << r[0] << "\t" << r[1] << "\t" << r[2] << "\t";
		for (size_t l = 0; l < nf; l++)
		  {
		    const int col = fieldColumn(seed_.profileField_[l]);
		    if ( col < 0 ) continue;
		    if ( col < 6 ) f << (float) val[9 * q + col] << "\t"; else f << val[9 * q + col] << "\t";
		  }
		f << std::endl;
	      }
      }
  }

  /* solver.cpp:1050-1059 */
  void Solver::initializeBunchUpdate ()
  {
    ub_.dt  = mesh_.timeStep_;
    ub_.dtb = c0_ * bunch_.timeStep_;
    ub_.dx  = mesh_.meshResolution_[0]; ub_.dy = mesh_.meshResolution_[1]; ub_.dz = mesh_.meshResolution_[2];
    ub_.r1  = - EC / ( EM * c0_ ) * bunch_.timeStep_ / 2.0;
    ub_.r2  = - EC / EM * bunch_.timeStep_ / 2.0;
    /* solver.cpp:1062-1124: files and checks of the bunch samplers                                                */
    if ( bunch_.sampling_ )
      {
	std::string name = "";
	if ( bunch_.basename_.compare(0, 1, "/") != 0 ) name = bunch_.directory_;
	name += bunch_.basename_ + ".txt";
	createDirectory(name, 0);
	bunchSampleFile_ = new std::ofstream(name.c_str(), std::ios::trunc);
	if ( bunch_.rhythm_ == 0 )
	  { printmessage(__FILE__, __LINE__, "The sampling rhythm of the bunch is zero although sampling is activated !!!"); exit(1); }
      }
    if ( bunch_.bunchProfile_ )
      {
	if ( bunch_.bunchProfileBasename_.compare(0, 1, "/") != 0 ) bunch_.bunchProfileBasename_ = bunch_.bunchProfileDirectory_ + bunch_.bunchProfileBasename_;
	createDirectory(bunch_.bunchProfileBasename_, 0);
      }
    if ( bunch_.bunchVTK_ )
      {
	if ( bunch_.bunchVTKBasename_.compare(0, 1, "/") != 0 ) bunch_.bunchVTKBasename_ = bunch_.bunchVTKDirectory_ + bunch_.bunchVTKBasename_;
	createDirectory(bunch_.bunchVTKBasename_, 0);
	if ( bunch_.bunchVTKRhythm_ == 0 )
	  { printmessage(__FILE__, __LINE__, "The visualization rhythm of the bunch is zero although visualization is activated !!!"); exit(1); }
      }
  }

  /* planes, wavelengths, window length and prefactor of the power sampling; opens the files -- radiation.cpp:18-121 */
  void Solver::initializePowerSample ()
  {
    rp_.clear(); rp_.resize(FEL_.size());
    for (unsigned int jf = 0; jf < FEL_.size(); jf++)
      {
	FreeElectronLaser::RadiationSampling& R = FEL_[jf].radiationPower_;
	if (!R.sampling_) continue;
	if ( undulator_.size() == 0 ) { printmessage(__FILE__, __LINE__, "Radiation power sampling needs an undulator."); exit(1); }
	for (Double& z : R.z_) z *= gamma_;
	R.lineBegin_ *= gamma_;
	R.lineEnd_   *= gamma_;
	Double dl = fabs( R.lineEnd_ - R.lineBegin_ ) / R.res_;
	if ( R.samplingType_ == OVERLINE )
	  for (Double l = 0.0; fabs(l) < fabs( R.lineEnd_ - R.lineBegin_ ); l += dl) R.z_.push_back( R.lineBegin_ + l );
	rp_[jf].N = R.z_.size();
	dl = ( R.lambdaMax_ - R.lambdaMin_ ) / R.lambdaRes_;
	for (Double rw = R.lambdaMin_; rw < R.lambdaMax_; rw += dl) R.lambda_.push_back(rw);
	rp_[jf].Nl = R.lambda_.size();
	rp_[jf].Nf = 0;
	rp_[jf].file.resize(rp_[jf].Nl);
	rp_[jf].w.resize(rp_[jf].Nl);
	for (unsigned int i = 0; i < rp_[jf].Nl; i++)
	  {
	    std::string name = "";
	    if ( R.basename_.compare(0, 1, "/") != 0 ) name = R.directory_;
	    name += R.basename_ + "-" + stringify(i) + ".txt";
	    createDirectory(name, 0);
	    rp_[jf].file[i] = new std::ofstream(name.c_str(), std::ios::trunc);
	    rp_[jf].file[i]->setf(std::ios::scientific);
	    rp_[jf].file[i]->precision(15);
	    rp_[jf].file[i]->width(40);
	    /* three radiation periods of this harmonic in the moving frame                                            */
	    const Double dt = undulator_[0].lu_ / R.lambda_[i] / ( gamma_ * c0_ );
	    rp_[jf].Nf   = ( unsigned( 3.0 * dt / mesh_.timeStep_ ) > rp_[jf].Nf ) ? unsigned( 3.0 * dt / mesh_.timeStep_ ) : rp_[jf].Nf;
	    rp_[jf].w[i] = 2 * PI / dt;
	  }
	rp_[jf].pc = 2.0 * mesh_.meshResolution_[0] * mesh_.meshResolution_[1] / ( m0_ * rp_[jf].Nf * rp_[jf].Nf ) * pow(mesh_.lengthScale_,2) / pow(mesh_.timeScale_,3);
	if ( powerGroup_ < 0 ) powerGroup_ = jf;
	else printmessage(__FILE__, __LINE__, "Note: only the first radiation-power group is sampled by this build.");
      }
  }

  /* plane, window length and prefactor of the power-visualization group -- radiation.cpp:238-318                */
  void Solver::initializePowerVisualize ()
  {
    for (unsigned int jf = 0; jf < FEL_.size(); jf++)
      {
	FreeElectronLaser::RadiationVisualization& V = FEL_[jf].vtkPower_;
	if (!V.sampling_) continue;
	if ( V.rhythm_ == 0 )
	  { printmessage(__FILE__, __LINE__, "The power visualization rhythm of the field is zero although power visualization is activated !!!"); exit(1); }
	if ( undulator_.size() == 0 ) { printmessage(__FILE__, __LINE__, "Radiation power visualization needs an undulator."); exit(1); }
	V.rhythm_ /= gamma_;
	V.z_      *= gamma_;
	if ( V.basename_.compare(0, 1, "/") != 0 ) V.basename_ = V.directory_ + V.basename_;
	createDirectory(V.basename_, 0);
	rp_[jf].N = 1; rp_[jf].Nl = 1;
	rp_[jf].w.resize(1);
	/* three radiation periods of the harmonic in the moving frame                                               */
	const Double dt = undulator_[0].lu_ / V.lambda_ / ( gamma_ * c0_ );
	rp_[jf].Nf   = unsigned( 3.0 * dt / mesh_.timeStep_ );
	rp_[jf].w[0] = 2 * PI / dt;
	rp_[jf].pc   = 2.0 * mesh_.meshResolution_[0] * mesh_.meshResolution_[1] / ( m0_ * rp_[jf].Nf * rp_[jf].Nf ) * pow(mesh_.lengthScale_,2) / pow(mesh_.timeScale_,3);
	if ( pmapGroup_ < 0 ) pmapGroup_ = jf;
	else printmessage(__FILE__, __LINE__, "Note: only the first power-visualization group is evaluated by this build.");
      }
  }

  /* solver.cpp:2145-2200 */
  void Solver::initializeScreenProfile ()
  {
    scrp_.clear(); scrp_.resize(FEL_.size());
    for (unsigned int jf = 0; jf < FEL_.size(); jf++)
      {
	FreeElectronLaser::ScreenProfile& S = FEL_[jf].screenProfile_;
	if (!S.sampling_) continue;
	if ( S.basename_.compare(0, 1, "/") != 0 ) S.basename_ = S.directory_ + S.basename_;
	if ( S.rhythm_ > 0.0 && undulator_.size() > 0 )
	  {
	    /* screens every rhythm_ up to the end of the last module (the reference dereferences end() here, Q10)   */
	    const Undulator& last = undulator_.back();
	    for (Double z = 0.0; z < last.rb_ + last.length_ * last.lu_; z += S.rhythm_) S.pos_.push_back(z);
	  }
	if ( S.pos_.size() == 0 )
	  { printmessage(__FILE__, __LINE__, "No position is set for the screen although the screen sampling is activated !!!"); exit(1); }
	createDirectory(S.basename_, 0);
	std::sort(S.pos_.begin(), S.pos_.end());
	scrp_[jf].fileNames.resize(S.pos_.size());
	scrp_[jf].files.resize(S.pos_.size());
	for (unsigned int i = 0; i < S.pos_.size(); i++)
	  {
	    scrp_[jf].fileNames[i] = S.basename_ + "-p" + stringify(rank_) + "-screen" + stringify(i) + ".txt";
	    scrp_[jf].files[i] = new std::ofstream(scrp_[jf].fileNames[i].c_str(), std::ios::trunc);
	    scrp_[jf].files[i]->setf(std::ios::scientific);
	    scrp_[jf].files[i]->precision(15);
	    scrp_[jf].files[i]->width(40);
	  }
	if ( screenGroup_ < 0 ) screenGroup_ = jf;
	else printmessage(__FILE__, __LINE__, "Note: only the first bunch-profile-lab-frame group is recorded by this build.");
      }
  }

  /* solver.cpp:1184-1206 */
  void Solver::shiftBackInTime ()
  {
    if ( mesh_.timeShift_ == 0.0 ) return;
    timem1_ -= mesh_.timeShift_; time_ -= mesh_.timeShift_; timep1_ -= mesh_.timeShift_; timeBunch_ -= mesh_.timeShift_;
    for (Charge& q : chargeVectorn_)
      {
	const Double t = c0_ * mesh_.timeShift_ / std::sqrt( 1.0 + q.gb.norm2() );
	q.rnp.mmv( t , q.gb );
      }
    distributeParticles(chargeVectorn_);
  }

  /* ========================================================================================================== */
  /* hand-over to the library                                                                                     */

  static void fillBeam (MithraBeam& d, const Beam& b)
  {
    d.seed_type = (int) b.seedType_;
    for (int c = 0; c < 3; c++) { d.position[c] = b.position_[c]; d.direction[c] = b.direction_[c]; d.polarization[c] = b.polarization_[c]; }
    d.amplitude = b.amplitude_;
    d.radius[0] = b.radius_[0]; d.radius[1] = b.radius_[1];
    d.l = b.l_;
    d.zR[0] = b.zR_[0]; d.zR[1] = b.zR_[1];
    d.order[0] = b.order_[0]; d.order[1] = b.order_[1];
    d.signal.type = (int) b.signal_.signalType_;
    d.signal.t0 = b.signal_.t0_; d.signal.s = b.signal_.s_; d.signal.f0 = b.signal_.f0_; d.signal.cep = b.signal_.cep_;
    d.signal.nR = (int) b.signal_.nR_;
    d.signal.sigma_inv_g[0] = b.signal_.sigmaInvG_.size() > 0 ? b.signal_.sigmaInvG_[0] : 0.0;
    d.signal.sigma_inv_g[1] = b.signal_.sigmaInvG_.size() > 1 ? b.signal_.sigmaInvG_[1] : 0.0;
  }

  void Solver::fillParams (MithraGpuParams& p, int slab) const
  {
    memset(&p, 0, sizeof(p));
    p.abi_version = MITHRA_GPU_ABI_VERSION;
    p.N0 = N0_; p.N1 = N1_; p.N2 = N2_; p.np = slabNp_[slab]; p.k0 = slabK0_[slab]; p.rank = slab; p.size = size_;
    p.dx = mesh_.meshResolution_[0]; p.dy = mesh_.meshResolution_[1]; p.dz = mesh_.meshResolution_[2]; p.dt = mesh_.timeStep_;
    p.xmin = xmin_; p.xmax = xmax_; p.ymin = ymin_; p.ymax = ymax_; p.zmin = zmin_; p.zmax = zmax_;
    p.zp[0] = slabZp0_[slab]; p.zp[1] = slabZp1_[slab];
    p.Lz = mesh_.meshLength_[2];
    p.solver = (int) mesh_.solver_; p.space_charge = mesh_.spaceCharge_ ? 1 : 0; p.truncation_order = mesh_.truncationOrder_;
    memcpy(p.a, uf_.a, sizeof(p.a)); p.alpha = uf_.alpha; p.beta_nsfd = uf_.beta;
    memcpy(p.bB, uf_.bB, sizeof(p.bB)); memcpy(p.cB, uf_.cB, sizeof(p.cB)); memcpy(p.dB, uf_.dB, sizeof(p.dB));
    memcpy(p.eE, uf_.eE, sizeof(p.eE)); memcpy(p.fE, uf_.fE, sizeof(p.fE)); memcpy(p.gE, uf_.gE, sizeof(p.gE));
    memcpy(p.hC, uf_.hC, sizeof(p.hC));
    p.c0 = c0_; p.gamma = gamma_; p.beta = beta_; p.dt_shift = dt_;
    p.dt_bunch = bunch_.timeStep_;
    p.n_update_bunch = 0;
    for (Double t = 0.0; t < nUpdateBunch_; t += 1.0) ++p.n_update_bunch;            /* trip count of the loop at solver.cpp:1316 */
    p.r1 = ub_.r1; p.r2 = ub_.r2; p.dtb = ub_.dtb;

    if ( undulator_.size() > MITHRA_MAX_UNDULATORS || extField_.size() > MITHRA_MAX_EXTFIELDS )
      { printmessage(__FILE__, __LINE__, "Too many undulator modules / external fields for the GPU parameter block."); exit(1); }
    p.n_undulators = undulator_.size();
    for (size_t u = 0; u < undulator_.size(); u++)
      {
	const Undulator& U = undulator_[u];
	MithraUndulator& D = p.undulator[u];
	D.type = (int) U.type_; D.k = U.k_; D.lu = U.lu_; D.rb = U.rb_; D.theta = U.theta_; D.length = U.length_; D.dist = U.dist_;
	fillBeam(D.beam, U);
      }
    p.n_ext_fields = extField_.size();
    for (size_t u = 0; u < extField_.size(); u++) fillBeam(p.ext_field[u], extField_[u]);
    p.seed_enabled = ( fabs(seed_.amplitude_) > 1.0e-50 ) ? 1 : 0;             /* fdtd.cpp:307 */
    fillBeam(p.seed, seed_);

    if ( powerGroup_ >= 0 )
      {
	const FreeElectronLaser::RadiationSampling& R = FEL_[powerGroup_].radiationPower_;
	const SampleRadiationPower& S = rp_[powerGroup_];
	if ( S.N > MITHRA_MAX_POWER_PLANES || S.Nl > MITHRA_MAX_POWER_LAMBDAS )
	  { printmessage(__FILE__, __LINE__, "Too many power planes / frequencies for the GPU parameter block."); exit(1); }
	p.power.enabled = 1; p.power.N = S.N; p.power.Nl = S.Nl; p.power.Nf = S.Nf; p.power.pc = S.pc;
	for (unsigned i = 0; i < S.N; i++)  p.power.z[i] = R.z_[i];
	for (unsigned i = 0; i < S.Nl; i++) p.power.w[i] = S.w[i];
      }
    if ( pmapGroup_ >= 0 )
      {
	const SampleRadiationPower& S = rp_[pmapGroup_];
	p.power_map.enabled = 1; p.power_map.Nf = S.Nf; p.power_map.z = FEL_[pmapGroup_].vtkPower_.z_;
	p.power_map.w = S.w[0]; p.power_map.pc = S.pc;
      }
    if ( screenGroup_ >= 0 )
      {
	const std::vector<Double>& pos = FEL_[screenGroup_].screenProfile_.pos_;
	if ( pos.size() > MITHRA_MAX_SCREENS ) { printmessage(__FILE__, __LINE__, "Too many screens for the GPU parameter block."); exit(1); }
	p.screens.enabled = 1; p.screens.N = pos.size();
	for (size_t i = 0; i < pos.size(); i++) p.screens.pos[i] = pos[i];
      }
    p.device = -1;
  }

  void Solver::attachGpu ()
  {
    const int ndev = mithra_gpu_device_count();
    if ( ndev < 1 ) { printmessage(__FILE__, __LINE__, "No CUDA device: the time-march of this build runs on sm_100 GPUs only."); exit(1); }

    /* the bunch of every slab: the particles it owns, in list order (solver.cpp:1440-1441)                         */
    std::vector<std::vector<double>> rows(size_);
    for (const Charge& q : chargeVectorn_)
      {
	const Double zr = pmod( q.rnp[2] - zmin_ , mesh_.meshLength_[2] ) + zmin_;
	int owner = -1;
	for (int r = 0; r < size_; r++) if ( zr >= slabZp0_[r] && zr < slabZp1_[r] ) owner = r;
	if ( owner < 0 ) continue;
	std::vector<double>& v = rows[owner];
	v.push_back(q.q);
	for (int d = 0; d < 3; d++) v.push_back(q.rnp[d]);
	for (int d = 0; d < 3; d++) v.push_back(q.rnm[d]);
	for (int d = 0; d < 3; d++) v.push_back(q.gb[d]);
	v.push_back(q.e);
      }

    gpu_.assign(size_, (MithraGpu*) 0);
    for (int r = 0; r < size_; r++)
      {
	MithraGpuParams p; fillParams(p, r);
	p.device = r % ndev;
	const size_t n = rows[r].size() / 11;
	/* room for the particles that migrate in: the whole bunch fits on any slab                                   */
	p.max_particles = ( size_ == 1 && !deviceBunch_ ) ? n + 1024 : (size_t) Nc_ + 1024;
	p.max_screen_records = std::max<size_t>(Nc_, 4096);
	check(mithra_gpu_create(&p, &gpu_[r]));
      }
    if ( size_ > 1 )
      {
	std::vector<std::vector<char>> blob(size_);
	for (int r = 0; r < size_; r++)
	  {
	    size_t nb = 0; check(mithra_gpu_ipc_export(gpu_[r], 0, 0, &nb));
	    blob[r].resize(nb); check(mithra_gpu_ipc_export(gpu_[r], blob[r].data(), nb, &nb));
	  }
	for (int r = 0; r < size_; r++)
	  check(mithra_gpu_ipc_connect(gpu_[r], blob[(r + size_ - 1) % size_].data(), blob[(r + 1) % size_].data()));
      }
    for (int r = 0; r < size_; r++)
      {
	check(mithra_gpu_set_time(gpu_[r], time_, timeBunch_, nTime_));
	if ( deviceBunch_ ) check(mithra_gpu_upload_particles_device(gpu_[r], deviceBunch_));      /* distributeParticles on the device */
	else                check(mithra_gpu_upload_particles(gpu_[r], rows[r].data(), rows[r].size() / 11));
	check(mithra_gpu_seed_initial(gpu_[r]));
      }
    if ( deviceBunch_ ) { mithra_gpu_bunch_destroy(deviceBunch_); deviceBunch_ = 0; }
  }

  /* ========================================================================================================== */
  /* the time march                                                                                               */

  void FdTd::fieldUpdate ()        { for (MithraGpu* g : gpu_) check(mithra_gpu_field_update(g)); }          /* fdtd.cpp:231-800  */
  void FdTd::fieldShift ()         { for (MithraGpu* g : gpu_) check(mithra_gpu_field_shift(g)); }           /* fdtd.cpp:806-812  */
  void FdTd::currentReset ()       { for (MithraGpu* g : gpu_) check(mithra_gpu_current_reset(g)); }         /* fdtd.cpp:23-32    */
  void FdTd::currentUpdate ()      { for (MithraGpu* g : gpu_) check(mithra_gpu_current_update(g)); }        /* fdtd.cpp:38-185   */
  void FdTd::currentCommunicate ()                                                                           /* fdtd.cpp:191-225  */
  {
    for (MithraGpu* g : gpu_) check(mithra_gpu_current_communicate(g));
    /* particle hand-over between the slabs (solver.cpp:1544-1568, 493-503), once per field step after the deposit  */
    for (MithraGpu* g : gpu_) check(mithra_gpu_migrate_begin(g));
    for (MithraGpu* g : gpu_) check(mithra_gpu_migrate_end(g));
  }

  /* rnm = rnp and the nUpdateBunch_ sub-steps of Solver::bunchUpdate in one launch per slab, solver.cpp:1311-1321 */
  void Solver::bunchUpdate ()
  {
    for (MithraGpu* g : gpu_) check(mithra_gpu_bunch_update(g));
    for (Double t = 0.0; t < nUpdateBunch_; t += 1.0) { timeBunch_ += bunch_.timeStep_; ++nTimeBunch_; }
  }

  /* Solver::bunchSample, solver.cpp:1582-1641: the sums come from the device (one reduction per slab, added in slab
   * order like the MPI_Reduce of :1611-1615), the line is written as the reference writes it                        */
  void Solver::bunchSample ()
  {
    double T[13]; for (int q = 0; q < 13; q++) T[q] = 0.0;
    for (MithraGpu* g : gpu_)
      {
	double s[13];
	check(mithra_gpu_bunch_moments(g, s));
	for (int q = 0; q < 13; q++) T[q] += s[q];
      }
    const Double qT = T[0];
    Double rT[3], r2T[3], gbT[3], gb2T[3];
    for (int l = 0; l < 3; l++) { rT[l] = T[1 + l] / qT; r2T[l] = T[4 + l] / qT; gbT[l] = T[7 + l] / qT; gb2T[l] = T[10 + l] / qT; }
    std::ofstream& f = *bunchSampleFile_;
    f.setf(std::ios::scientific);
    f.precision(4);
    f << timeBunch_ << "\t";
    f << rT[0]  << "\t" << rT[1]  << "\t" << rT[2]  << "\t";
    f << gbT[0] << "\t" << gbT[1] << "\t" << gbT[2] << "\t";
    f << sqrt( r2T[0]  - rT[0]  * rT[0]  ) << "\t";
    f << sqrt( r2T[1]  - rT[1]  * rT[1]  ) << "\t";
    f << sqrt( r2T[2]  - rT[2]  * rT[2]  ) << "\t";
    f << sqrt( gb2T[0] - gbT[0] * gbT[0] ) << "\t";
    f << sqrt( gb2T[1] - gbT[1] * gbT[1] ) << "\t";
    f << sqrt( gb2T[2] - gbT[2] * gbT[2] ) << std::endl;
  }

  /* Solver::bunchProfile, solver.cpp:1763-1792: the particle list of every slab in the reference's order, one file
   * (the reference writes one per rank; with one process there is one, "-p0-")                                       */
  void Solver::bunchProfile ()
  {
    const std::string name = bunch_.bunchProfileBasename_ + "-p" + stringify(0) + "-" + stringify(nTime_) + ".txt";
    std::ofstream f(name.c_str(), std::ios::trunc);
    f.setf(std::ios::scientific);
    f.precision(15);
    f.width(40);
    f << time_ * gamma_ << std::endl;
    std::vector<double> rows;
    for (MithraGpu* g : gpu_)
      {
	size_t n = 0;
	check(mithra_gpu_num_particles(g, &n));
	if ( n == 0 ) continue;
	rows.resize(n * 11);
	check(mithra_gpu_download_particles(g, rows.data(), n, &n));
	for (size_t i = 0; i < n; i++)
	  {
	    const double* q = &rows[11 * i];
	    /* one slab: the reference's ownership test; several: the list of a slab is what it owns                     */
	    if ( gpu_.size() == 1 && !particleInProcessor(q[3]) ) continue;
	    f << q[0] << "\t" << q[1] << "\t" << q[2] << "\t" << q[3] << "\t" << q[7] << "\t" << q[8] << "\t" << q[9] << std::endl;
	  }
      }
    f.close();
  }

  /* Solver::bunchVisualize, solver.cpp:1647-1757: the particle cloud as one ASCII .vtu piece (single-rank naming,
   * "-p0-") with (q, lab-frame gamma, gamma x 0.512 MeV) per particle, plus the .pvtu that lists it                  */
  void Solver::bunchVisualize ()
  {
    std::vector<double> all;
    for (MithraGpu* g : gpu_)
      {
	size_t n = 0;
	check(mithra_gpu_num_particles(g, &n));
	if ( n == 0 ) continue;
	std::vector<double> rows(n * 11);
	check(mithra_gpu_download_particles(g, rows.data(), n, &n));
	for (size_t i = 0; i < n; i++)
	  {
	    if ( gpu_.size() == 1 && !particleInProcessor(rows[11 * i + 3]) ) continue;
	    all.insert(all.end(), rows.begin() + 11 * i, rows.begin() + 11 * i + 11);
	  }
      }
    const size_t N = all.size() / 11;
    std::string name = bunch_.bunchVTKBasename_ + "-p" + stringify(0) + "-" + stringify(nTimeBunch_) + ".vtu";
    {
      std::ofstream f(name.c_str(), std::ios::trunc);
      f.setf(std::ios::scientific);
      f.precision(4);
      f << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">" << std::endl;
      f << "<UnstructuredGrid>" << std::endl;
      f << "<Piece NumberOfPoints=\"" << N + 1 << "\" NumberOfCells=\"" << 1 << "\">" << std::endl;
      f << "<Points>" << std::endl;
      f << "<DataArray type = \"Float64\" NumberOfComponents=\"3\" format=\"ascii\">" << std::endl;
      for (size_t i = 0; i < N; i++) f << all[11 * i + 1] << " " << all[11 * i + 2] << " " << all[11 * i + 3] << std::endl;
      f << xmin_ << " " << ymin_ << " " << zmin_ << std::endl;
      f << "</DataArray>" << std::endl;
      f << "</Points>" << std::endl;
      f << "<Cells>" << std::endl;
      f << "<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">" << std::endl;
      for (size_t i = 0; i < N + 1; ++i) f << i << " ";
      f << std::endl;
      f << "</DataArray>" << std::endl;
      f << "<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">" << std::endl;
      f << N + 1 << std::endl;
      f << "</DataArray>" << std::endl;
      f << "<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">" << std::endl;
      f << 2 << std::endl;
      f << "</DataArray>" << std::endl;
      f << "</Cells>" << std::endl;
      f << "<PointData Vectors = \"charge\">" << std::endl;
      f << "<DataArray type=\"Float64\" Name=\"charge\" NumberOfComponents=\"3\" format=\"ascii\">" << std::endl;
      for (size_t i = 0; i < N; i++)
	{
	  const double* q = &all[11 * i];
	  const Double gamma = sqrt( 1.0 + ( q[7] * q[7] + q[8] * q[8] + q[9] * q[9] ) );
	  const Double beta  = q[9] / gamma;
	  f << q[0] << " " << gamma * gamma_ * ( 1.0 + beta_ * beta ) << " " << gamma * gamma_ * ( 1.0 + beta_ * beta ) * 0.512 << std::endl;
	}
      f << 0.0 << " " << 0.0 << " " << 0.0 << std::endl;
      f << 0.0 << " " << 0.0 << " " << 0.0 << std::endl;
      f << "</DataArray>" << std::endl;
      f << "</PointData>" << std::endl;
      f << "</Piece>" << std::endl;
      f << "</UnstructuredGrid>" << std::endl;
      f << "</VTKFile>" << std::endl;
    }
    name = bunch_.bunchVTKBasename_ + "-" + stringify(nTimeBunch_) + ".pvtu";
    std::ofstream f(name.c_str(), std::ios::trunc);
    const size_t found = bunch_.bunchVTKBasename_.find_last_of("/");
    const std::string base = bunch_.bunchVTKBasename_.substr(found + 1);
    f << "<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">" << std::endl;
    f << "<PUnstructuredGrid> GhostLevel = \"0\"" << std::endl;
    f << "<PPoints>" << std::endl;
    f << "<PDataArray type = \"Float64\" NumberOfComponents=\"3\" format=\"ascii\" />" << std::endl;
    f << "</PPoints>" << std::endl;
    f << "<PPointData>" << std::endl;
    f << "<PDataArray type=\"Float64\" Name=\"charge\" NumberOfComponents=\"3\" format=\"ascii\" />" << std::endl;
    f << "</PPointData>" << std::endl;
    f << "<Piece  Source=\"" << base + "-p" + stringify(0) + "-" + stringify(nTimeBunch_) + ".vtu" << "\"/>" << std::endl;
    f << "</PUnstructuredGrid>" << std::endl;
    f << "</VTKFile>" << std::endl;
  }

  void Solver::screenProfile () { if ( screenGroup_ >= 0 ) for (MithraGpu* g : gpu_) check(mithra_gpu_screen_profile(g)); }

  void Solver::powerSample ()
  {
    if ( powerGroup_ < 0 ) return;
    for (MithraGpu* g : gpu_) check(mithra_gpu_power_sample(g));
    powerTimes_.push_back(timeBunch_);
  }

  /* Solver::powerVisualize, radiation.cpp:324-450: the library updates the per-pixel map every step; at the rhythm
   * the slab that holds the plane hands it over and the .vts file is written in the reference's format (:393-447)  */
  void Solver::powerVisualize ()
  {
    if ( pmapGroup_ < 0 ) return;
    for (MithraGpu* g : gpu_) check(mithra_gpu_power_visualize(g));
    const FreeElectronLaser::RadiationVisualization& V = FEL_[pmapGroup_].vtkPower_;
    if ( !( fmod(time_, V.rhythm_) < mesh_.timeStep_ ) ) return;
    std::vector<double> pL((size_t) N1N0_, 0.0);
    bool have = false;
    for (MithraGpu* g : gpu_)
      {
	int mine = 0;
	check(mithra_gpu_fetch_power_map(g, pL.data(), pL.size(), &mine));
	if ( mine ) { have = true; break; }
      }
    if ( !have ) return;
    Double c;
    const Double dzr = modf( ( V.z_ - zmin_ ) / mesh_.meshResolution_[2], &c );
    const long int k = (long int) c;                            /* global plane index (k0_ of slab 0 is 0)           */
    const std::string name = V.basename_ + "-" + stringify(nTime_) + ".vts";
    std::ofstream f(name.c_str(), std::ios::trunc);
    f.setf(std::ios::scientific);
    f.precision(4);
    f << "<?xml version=\"1.0\"?>" << std::endl;
    f << "<VTKFile type=\"StructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" compressor=\"vtkZLibDataCompressor\">" << std::endl;
    f << "<StructuredGrid WholeExtent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " " << 0 << " " << 0 << "\">" << std::endl;
    f << "<Piece Extent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " " << 0 << " " << 0 << "\">" << std::endl;
    f << "<Points>" << std::endl;
    f << "<DataArray type = \"Float64\" NumberOfComponents=\"3\" format=\"ascii\">" << std::endl;
    for (int j = 0; j < N1_; j++)
      for (int i = 0; i < N0_; i++)
	{
	  const long int m = k * N1N0_ + i * N1_ + j;
	  const FieldVector r1 = rc(m), r2 = rc(m + N1N0_);
	  f << r1[0] * ( 1.0 - dzr ) + r2[0] * dzr << " " << r1[1] << " " << r1[2] << std::endl;
	}
    f << "</DataArray>" << std::endl;
    f << "</Points>" << std::endl;
    f << "<CellData>" << std::endl;
    f << "</CellData>" << std::endl;
    f << "<PointData Vectors = \"power\">" << std::endl;
    f << "<DataArray type=\"Float64\" Name=\"power\" NumberOfComponents=\"" << 1 << "\" format=\"ascii\">" << std::endl;
    for (int j = 0; j < N1_; j++)
      for (int i = 0; i < N0_; i++)
	f << pL[(size_t) i * N1_ + j] << std::endl;
    f << "</DataArray>" << std::endl;
    f << "</PointData>" << std::endl;
    f << "</Piece>" << std::endl;
    f << "</StructuredGrid>" << std::endl;
    f << "</VTKFile>" << std::endl;
    f.close();
  }

  /* write what the library has collected since the last call: power lines (radiation.cpp:222-230), screen records
   * (solver.cpp:2229-2252)                                                                                        */
  void Solver::flushOutputs ()
  {
    if ( powerGroup_ >= 0 && !powerTimes_.empty() )
      {
	const SampleRadiationPower& S = rp_[powerGroup_];
	const std::vector<Double>& z = FEL_[powerGroup_].radiationPower_.z_;
	const size_t w = (size_t) S.N * S.Nl, nrows = powerTimes_.size();
	std::vector<double> sum(nrows * w, 0.0), part(nrows * w);
	for (MithraGpu* g : gpu_)
	  {
	    size_t got = 0;
	    check(mithra_gpu_fetch_power(g, part.data(), nrows, &got));
	    if ( got != nrows ) { printmessage(__FILE__, __LINE__, "power rows out of step with the host loop"); exit(1); }
	    /* every plane is sampled by exactly one slab, the others return zeros (the MPI_Allreduce of :218)         */
	    for (size_t i = 0; i < nrows * w; i++) sum[i] += part[i];
	  }
	for (size_t r = 0; r < nrows; r++)
	  for (unsigned l = 0; l < S.Nl; l++)
	    {
	      std::ofstream& f = *S.file[l];
	      for (unsigned k = 0; k < S.N; ++k)
		f << gamma_ * ( z[k] + beta_ * c0_ * ( powerTimes_[r] + dt_ ) ) << "\t" << sum[r * w + k * S.Nl + l] << "\t";
	      f << std::endl;
	    }
	powerTimes_.clear();
      }
    if ( screenGroup_ >= 0 )
      {
	const size_t ns = FEL_[screenGroup_].screenProfile_.pos_.size();
	std::vector<double> rec;
	for (size_t s = 0; s < ns; s++)
	  for (MithraGpu* g : gpu_)
	    {
	      size_t n = 0;
	      check(mithra_gpu_fetch_screen(g, (int) s, 0, 0, &n));
	      if ( n == 0 ) continue;
	      rec.resize(n * 6);
	      check(mithra_gpu_fetch_screen(g, (int) s, rec.data(), n, &n));
	      std::ofstream& f = *scrp_[screenGroup_].files[s];
	      for (size_t i = 0; i < n; i++)
		f << rec[6 * i] << "\t" << rec[6 * i + 1] << "\t" << rec[6 * i + 2] << "\t" << rec[6 * i + 3] << "\t" << rec[6 * i + 4] << "\t" << rec[6 * i + 5] << std::endl;
	    }
      }
  }

  void Solver::finalize ()
  {
    for (MithraGpu* g : gpu_) check(mithra_gpu_synchronize(g));
    flushOutputs();
    if ( bunchSampleFile_ ) bunchSampleFile_->close();
    if ( fieldSampleFile_ ) fieldSampleFile_->close();
    for (SampleRadiationPower& S : rp_) for (std::ofstream* f : S.file) if (f) f->close();
    for (SampleScreenProfile& S : scrp_) for (std::ofstream* f : S.files) if (f) f->close();
  }

  /* solver.cpp:1212-1418 */
  void Solver::solve ()
  {
    initialize();
    attachGpu();

    timeval t0, t1;
    gettimeofday(&t0, NULL);
    const unsigned int flushEvery = 512;
    long steps = 0;
    auto advance = [&] () {
      for (MithraGpu* g : gpu_) check(mithra_gpu_advance_time(g));
      timem1_ += mesh_.timeStep_; time_ += mesh_.timeStep_; timep1_ += mesh_.timeStep_; ++nTime_; ++steps;
      if ( nTime_ % flushEvery == 0 ) flushOutputs(); };

    /* particles only, up to the time origin (initial-time-back-shift), solver.cpp:1232-1291                        */
    while ( time_ < 0.0 && ( maxSteps_ < 0 || steps < maxSteps_ ) )
      {
	bunchUpdate();
	screenProfile();
	/* the bunch samplers of the first loop, solver.cpp:1253-1270                                                  */
	if ( bunch_.sampling_ && fmod(time_ + mesh_.timeShift_, bunch_.rhythm_) < mesh_.timeStep_ && ( time_ + mesh_.timeShift_ > 0.0 ) ) bunchSample();
	if ( bunch_.bunchVTK_ && fmod(time_ + mesh_.timeShift_, bunch_.bunchVTKRhythm_) < mesh_.timeStep_ && ( time_ + mesh_.timeShift_ > 0.0 ) ) bunchVisualize();
	if ( bunch_.bunchProfile_ )
	  {
	    for (unsigned int i = 0; i < bunch_.bunchProfileTime_.size(); i++)
	      if ( time_ - bunch_.bunchProfileTime_[i] < mesh_.timeStep_ && time_ > bunch_.bunchProfileTime_[i] ) bunchProfile();
	    if ( fmod(time_ + mesh_.timeShift_, bunch_.bunchProfileRhythm_) < mesh_.timeStep_ && ( time_ + mesh_.timeShift_ > 0.0 ) && ( bunch_.bunchProfileRhythm_ != 0.0 ) )
	      bunchProfile();
	  }
	for (MithraGpu* g : gpu_) check(mithra_gpu_migrate_begin(g));
	for (MithraGpu* g : gpu_) check(mithra_gpu_migrate_end(g));
	advance();
      }

    gettimeofday(&t0, NULL);
    unsigned int nStart = nTime_;
    /* MITHRA_HOST_TIMING_SKIP=W: the clock of the "Time march" line below starts after W field steps (bench.py's warm-up)   */
    const long timingSkip = getenv("MITHRA_HOST_TIMING_SKIP") ? atol(getenv("MITHRA_HOST_TIMING_SKIP")) : 0;
    const long stepsAtStart = steps;
    while ( time_ < mesh_.totalTime_ && ( maxSteps_ < 0 || steps < maxSteps_ ) )
      {
	if ( timingSkip > 0 && steps - stepsAtStart == timingSkip )
	  {
	    for (MithraGpu* g : gpu_) check(mithra_gpu_synchronize(g));
	    gettimeofday(&t0, NULL);
	    nStart = nTime_;
	  }
	/* A step without a rhythm-gated bunch output is exactly mithra_gpu_step: the same calls in the same order, with
	 * the library free to run its housekeeping beside the particle kernels and to test the screens inside the push.
	 * One process driving several slabs keeps the call-by-call loop (every slab must enqueue its sends before the
	 * first one waits for its neighbours).                                                                          */
	bool gated = false;
	{
	  const Double tb = time_ + mesh_.timeShift_;
	  if ( bunch_.sampling_ && fmod(tb, bunch_.rhythm_) < mesh_.timeStep_ && tb > 0.0 ) gated = true;
	  if ( bunch_.bunchVTK_ && fmod(tb, bunch_.bunchVTKRhythm_) < mesh_.timeStep_ && tb > 0.0 ) gated = true;
	  if ( bunch_.bunchProfile_ )
	    {
	      for (unsigned int i = 0; i < bunch_.bunchProfileTime_.size(); i++)
		if ( time_ - bunch_.bunchProfileTime_[i] < mesh_.timeStep_ && time_ > bunch_.bunchProfileTime_[i] ) gated = true;
	      if ( fmod(tb, bunch_.bunchProfileRhythm_) < mesh_.timeStep_ && tb > 0.0 && bunch_.bunchProfileRhythm_ != 0.0 ) gated = true;
	    }
	  if ( pmapGroup_ >= 0 ) gated = true;                  /* the power map is fetched from inside powerVisualize()    */
	  if ( seed_.sampling_ && fmod(time_, seed_.samplingRhythm_) < mesh_.timeStep_ && time_ > 0.0 ) gated = true;
	  for (unsigned int i = 0; i < seed_.vtk_.size(); i++)
	    if ( seed_.vtk_[i].sample_ && fmod(time_, seed_.vtk_[i].rhythm_) < mesh_.timeStep_ && time_ > 0.0 ) gated = true;
	  if ( seed_.profile_ )
	    {
	      for (unsigned int i = 0; i < seed_.profileTime_.size(); i++)
		if ( time_ - seed_.profileTime_[i] < mesh_.timeStep_ && time_ > seed_.profileTime_[i] ) gated = true;
	      if ( fmod(time_, seed_.profileRhythm_) < mesh_.timeStep_ && time_ > 0.0 && seed_.profileRhythm_ != 0 ) gated = true;
	    }
	}
	if ( gpu_.size() == 1 && !gated && !getenv("MITHRA_HOST_CALL_BY_CALL") )
	  {
	    check(mithra_gpu_step(gpu_[0], 1));
	    for (Double t = 0.0; t < nUpdateBunch_; t += 1.0) { timeBunch_ += bunch_.timeStep_; ++nTimeBunch_; }
	    if ( powerGroup_ >= 0 ) powerTimes_.push_back(timeBunch_);
	    timem1_ += mesh_.timeStep_; time_ += mesh_.timeStep_; timep1_ += mesh_.timeStep_; ++nTime_; ++steps;
	    if ( nTime_ % flushEvery == 0 ) flushOutputs();
	  }
	else
	  {
	fieldUpdate();
	bunchUpdate();
	recycleParticles();
	/* rhythm-gated field sampling, solver.cpp:1326-1328                                                           */
	if ( seed_.sampling_ && fmod(time_, seed_.samplingRhythm_) < mesh_.timeStep_ && time_ > 0.0 ) fieldSample();
	/* rhythm-gated field visualisation, solver.cpp:1332-1340                                                      */
	for (unsigned int i = 0; i < seed_.vtk_.size(); i++)
	  if ( seed_.vtk_[i].sample_ && fmod(time_, seed_.vtk_[i].rhythm_) < mesh_.timeStep_ && time_ > 0.0 )
	    {
	      if      ( seed_.vtk_[i].type_ == ALLDOMAIN ) fieldVisualizeAllDomain(i);
	      else if ( seed_.vtk_[i].type_ == INPLANE   ) fieldVisualizeInPlane(i);
	    }
	/* field profile at the given times and at the rhythm, solver.cpp:1344-1351                                    */
	if ( seed_.profile_ )
	  {
	    for (unsigned int i = 0; i < seed_.profileTime_.size(); i++)
	      if ( time_ - seed_.profileTime_[i] < mesh_.timeStep_ && time_ > seed_.profileTime_[i] ) fieldProfile();
	    if ( fmod(time_, seed_.profileRhythm_) < mesh_.timeStep_ && time_ > 0.0 && seed_.profileRhythm_ != 0 ) fieldProfile();
	  }
	/* rhythm-gated bunch samplers, solver.cpp:1352-1371                                                           */
	if ( bunch_.sampling_ && fmod(time_ + mesh_.timeShift_, bunch_.rhythm_) < mesh_.timeStep_ && ( time_ + mesh_.timeShift_ > 0.0 ) ) bunchSample();
	if ( bunch_.bunchVTK_ && fmod(time_ + mesh_.timeShift_, bunch_.bunchVTKRhythm_) < mesh_.timeStep_ && ( time_ + mesh_.timeShift_ > 0.0 ) ) bunchVisualize();
	if ( bunch_.bunchProfile_ )
	  {
	    for (unsigned int i = 0; i < bunch_.bunchProfileTime_.size(); i++)
	      if ( time_ - bunch_.bunchProfileTime_[i] < mesh_.timeStep_ && time_ > bunch_.bunchProfileTime_[i] ) bunchProfile();
	    if ( fmod(time_ + mesh_.timeShift_, bunch_.bunchProfileRhythm_) < mesh_.timeStep_ && ( time_ + mesh_.timeShift_ > 0.0 ) && ( bunch_.bunchProfileRhythm_ != 0.0 ) )
	      bunchProfile();
	  }
	screenProfile();
	powerSample();
	powerVisualize();
	fieldShift();
	currentReset();
	currentUpdate();
	currentCommunicate();
	advance();
	  }

	if ( int( time_ / mesh_.totalTime_ * 1000.0 ) != int( timem1_ / mesh_.totalTime_ * 1000.0 ) )
	  {
	    for (MithraGpu* g : gpu_) check(mithra_gpu_synchronize(g));
	    gettimeofday(&t1, NULL);
	    const Double dT = ( t1.tv_usec - t0.tv_usec ) / 1.0e6 + ( t1.tv_sec - t0.tv_sec );
	    printmessage(__FILE__, __LINE__, " Percentage of the total simulation completed (%)      = " + stringify( time_ / mesh_.totalTime_ * 100.0 ));
	    printmessage(__FILE__, __LINE__, " Average calculation time for each time step (s) = " + stringify( dT / (double) ( nTime_ - nStart ) ));
	    printmessage(__FILE__, __LINE__, " Estimated remaining time (min)                  = " + stringify( ( mesh_.totalTime_ / time_ - 1 ) * dT / 60 ));
	  }
      }
    /* the march of this call in one line (with --steps no 0.1 % mark may have been passed; bench.py reads it)           */
    if ( nTime_ > nStart )
      {
	for (MithraGpu* g : gpu_) check(mithra_gpu_synchronize(g));
	gettimeofday(&t1, NULL);
	const Double dT = ( t1.tv_usec - t0.tv_usec ) / 1.0e6 + ( t1.tv_sec - t0.tv_sec );
	printmessage(__FILE__, __LINE__, " Time march: " + stringify(nTime_ - nStart) + " field steps in " + stringify(dT) +
		     " s; average calculation time for each time step (s) = " + stringify( dT / (double) ( nTime_ - nStart ) ));
      }
    finalize();
  }

  /* ========================================================================================================== */
  /* record file with the results of initialize(), same names as oracle/ref_dump.cpp's meta record                 */

  namespace
  {
    struct Writer
    {
      std::ofstream f;
      Writer (const std::string& fn) : f(fn.c_str(), std::ios::binary | std::ios::trunc) {}
      void raw (const char* name, int32_t type, const void* p, int64_t n, size_t item)
      {
	char key[48]; memset(key, 0, sizeof(key)); strncpy(key, name, 47);
	f.write(key, 48); f.write((const char*) &type, 4); f.write((const char*) &n, 8); f.write((const char*) p, n * item);
      }
      void f64 (const std::string& name, const double* p, int64_t n) { raw(name.c_str(), 0, p, n, 8); }
      void d (const std::string& name, double v) { f64(name, &v, 1); }
      void i (const std::string& name, int v) { int32_t x = v; raw(name.c_str(), 2, &x, 1, 4); }
    };

    void dumpBeam (Writer& w, const std::string& key, const Beam& b)
    {
      const double o[18] = { (double) b.seedType_, b.position_[0], b.position_[1], b.position_[2], b.direction_[0], b.direction_[1], b.direction_[2],
			     b.polarization_[0], b.polarization_[1], b.polarization_[2], b.amplitude_, b.radius_[0], b.radius_[1], b.l_, b.zR_[0], b.zR_[1],
			     (double) b.order_[0], (double) b.order_[1] };
      w.f64(key + "beam", o, 18);
      const double g[8] = { (double) b.signal_.signalType_, b.signal_.t0_, b.signal_.s_, b.signal_.f0_, (double) b.signal_.nR_, b.signal_.cep_,
			    b.signal_.sigmaInvG_.size() > 0 ? b.signal_.sigmaInvG_[0] : 0.0, b.signal_.sigmaInvG_.size() > 1 ? b.signal_.sigmaInvG_[1] : 0.0 };
      w.f64(key + "sig", g, 8);
    }
  }

  void Solver::dumpParams (const std::string& prefix)
  {
    Writer w(prefix + ".meta.bin");
    w.i("N0", N0_); w.i("N1", N1_); w.i("N2", N2_); w.i("np", np_); w.i("k0", k0_); w.i("rank", 0); w.i("size", size_);
    w.i("spaceCharge", mesh_.spaceCharge_ ? 1 : 0); w.i("solver", (int) mesh_.solver_); w.i("truncationOrder", mesh_.truncationOrder_);
    w.d("dx", mesh_.meshResolution_[0]); w.d("dy", mesh_.meshResolution_[1]); w.d("dz", mesh_.meshResolution_[2]);
    w.d("Lx", mesh_.meshLength_[0]); w.d("Ly", mesh_.meshLength_[1]); w.d("Lz", mesh_.meshLength_[2]);
    w.d("dt", mesh_.timeStep_); w.d("dtBunch", bunch_.timeStep_); w.d("nUpdateBunch", nUpdateBunch_);
    w.d("totalTime", mesh_.totalTime_); w.d("timeShift", mesh_.timeShift_);
    w.d("xmin", xmin_); w.d("xmax", xmax_); w.d("ymin", ymin_); w.d("ymax", ymax_); w.d("zmin", zmin_); w.d("zmax", zmax_);
    const double zp[2] = { slabZp0_[0], slabZp1_[0] }; w.f64("zp", zp, 2);
    w.d("gamma", gamma_); w.d("beta", beta_); w.d("dtShift", dt_); w.d("c0", c0_); w.d("m0", m0_); w.d("e0", e0_);
    w.f64("a", uf_.a, 6); w.d("alpha", uf_.alpha); w.d("betaNSFD", uf_.beta);
    w.f64("bB", uf_.bB, 5); w.f64("cB", uf_.cB, 5); w.f64("dB", uf_.dB, 5); w.f64("eE", uf_.eE, 5); w.f64("fE", uf_.fE, 5); w.f64("gE", uf_.gE, 5);
    w.f64("hC", uf_.hC, 17);
    w.d("dv", uc_.dv); w.d("rc", uc_.rc); w.d("r1", ub_.r1); w.d("r2", ub_.r2); w.d("dtb", ub_.dtb);
    w.d("seedAmplitude", seed_.amplitude_);
    dumpBeam(w, "seed.", seed_);
    w.i("nExtFields", (int) extField_.size());
    for (size_t u = 0; u < extField_.size(); u++) dumpBeam(w, "ext" + stringify(u) + ".", extField_[u]);
    w.i("nUndulators", (int) undulator_.size());
    for (size_t u = 0; u < undulator_.size(); u++)
      {
	const Undulator& U = undulator_[u];
	const double v[8] = { U.k_, U.lu_, U.rb_, (double) U.length_, U.dist_, U.theta_, (double) U.type_, (double) U.seedType_ };
	w.f64("und" + stringify(u) + ".static", v, 8);
	dumpBeam(w, "und" + stringify(u) + ".", U);
      }
    w.i("nFEL", (int) FEL_.size());
    for (size_t jf = 0; jf < FEL_.size(); jf++)
      {
	if (!FEL_[jf].radiationPower_.sampling_) continue;
	const std::string k = "power" + stringify(jf) + ".";
	w.i(k + "N", rp_[jf].N); w.i(k + "Nl", rp_[jf].Nl); w.i(k + "Nf", rp_[jf].Nf); w.d(k + "pc", rp_[jf].pc);
	w.f64(k + "z", FEL_[jf].radiationPower_.z_.data(), FEL_[jf].radiationPower_.z_.size());
	w.f64(k + "w", rp_[jf].w.data(), rp_[jf].w.size());
      }
    for (size_t jf = 0; jf < FEL_.size(); jf++)
      {
	if (!FEL_[jf].vtkPower_.sampling_) continue;
	const std::string k = "pmap" + stringify(jf) + ".";
	w.i(k + "Nf", rp_[jf].Nf); w.d(k + "pc", rp_[jf].pc); w.d(k + "z", FEL_[jf].vtkPower_.z_);
	w.d(k + "w", rp_[jf].w.empty() ? 0.0 : rp_[jf].w[0]); w.d(k + "rhythm", FEL_[jf].vtkPower_.rhythm_);
      }
    for (size_t jf = 0; jf < FEL_.size(); jf++)
      if (FEL_[jf].screenProfile_.sampling_)
	w.f64("screen" + stringify(jf) + ".pos", FEL_[jf].screenProfile_.pos_.data(), FEL_[jf].screenProfile_.pos_.size());
    w.d("time", time_); w.d("timem1", timem1_); w.d("timep1", timep1_); w.d("timeBunch", timeBunch_);
    w.i("nTime", (int) nTime_); w.i("nTimeBunch", (int) nTimeBunch_);
    std::vector<double> rows; rows.reserve(chargeVectorn_.size() * 11);
    if ( deviceBunch_ )
      {
	size_t n = 0; check(mithra_gpu_bunch_download(deviceBunch_, 0, 0, &n));
	rows.resize(n * 11); check(mithra_gpu_bunch_download(deviceBunch_, rows.data(), n, &n));
      }
    for (const Charge& q : chargeVectorn_)
      {
	rows.push_back(q.q);
	for (int d = 0; d < 3; d++) rows.push_back(q.rnp[d]);
	for (int d = 0; d < 3; d++) rows.push_back(q.rnm[d]);
	for (int d = 0; d < 3; d++) rows.push_back(q.gb[d]);
	rows.push_back(q.e);
      }
    w.f64("particles", rows.data(), rows.size());
    /* the parameter block of every slab exactly as mithra_gpu_create receives it                                   */
    for (int r = 0; r < size_; r++)
      {
	MithraGpuParams p; fillParams(p, r);
	w.raw(("params" + stringify(r)).c_str(), 3, &p, sizeof(p), 1);
      }
  }
}
