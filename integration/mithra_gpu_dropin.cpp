/* mithra_gpu_dropin.cpp -- INTEGRATION.md option B made real: the file a maintainer of the reference adds to route the
 * time march through libmithra_gpu.so.  It is compiled against the UNMODIFIED reference headers and linked with the
 * UNMODIFIED reference objects (mithra.cpp's main, readdata, datainput, classes, database, stdinclude, solver.cpp with its
 * own initialize() and solve() loop, radiation.cpp with its own initializePowerSample, fdtd.cpp and fdtdSC.cpp with their
 * constructors, fieldEvaluate and field writers) by `make -C oracle ref_gpu`; it OVERRIDES
 *   - the five time-march virtuals of FdTd / FdTdSC (solver.h:139-178): fieldUpdate, fieldShift, currentReset,
 *     currentUpdate, currentCommunicate call the C ABI of include/mithra_gpu.h;
 *   - the four non-virtual Solver methods of the loop that touch particles or fields -- bunchUpdate
 *     (solver.cpp:1424-1576), screenProfile (:2205-2257), powerSample and powerVisualize (radiation.cpp:127-450; the power
 *     map is accumulated on the device and written here in the reference's .vts layout).
 * The build weakens those symbols in copies of the reference objects (objcopy --weaken-symbol), the linker takes these.
 * Nothing of the reference is edited or copied.
 *
 * The rhythm-gated outputs stay the reference's OWN code working on its own host arrays, which the stub refreshes from the
 * device in a field step in which one of them is due (the loop's own conditions, evaluated at the end of the step's last
 * bunchUpdate call, a few lines before the loop evaluates them):
 *   - bunchSample, bunchVisualize, bunchProfile (solver.cpp:1582-1792) on chargeVectorn_  <- mithra_gpu_download_particles
 *     (the order of the upload);
 *   - fieldSample, fieldVisualize*, fieldProfile (fdtd.cpp:851-1594, fdtdSC.cpp) through the reference's lazy fieldEvaluate
 *     on anp1_ / an_ (/ fnp1_ / fn_)  <- mithra_gpu_download_fields, pic_ cleared, the end planes given the E/B of their inner
 *     neighbours as the tail of the reference's fieldUpdate does (fdtd.cpp:742-773).
 *
 * This is test infrastructure of the repository (it proves the drop-in claim against the reference's own main and
 * loop); the product is the library.
 */
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "fdtd.h"
#include "fdtdSC.h"
#include "../include/mithra_gpu.h"

namespace MITHRA
{
  namespace
  {
    MithraGpu* gpu = 0;
    int        subStep = 0;                        /* calls of bunchUpdate within the current field step (solver.cpp:1316) */

    void check (int rc)
    {
      if (rc) { printmessage(std::string(__FILE__), __LINE__, std::string("GPU time march: ") + mithra_gpu_last_error()); exit(1); }
    }

    void refuse (const char* what)
    {
      printmessage(std::string(__FILE__), __LINE__, std::string(what) + " is not routed by this stub (mithra_b200/host writes it over the same ABI).");
      exit(1);
    }

    template <class B>
    void fillBeam (MithraBeam& d, const B& b)
    {
      d.seed_type = (int) b.seedType_;
      for (int c = 0; c < 3; c++) { d.position[c] = b.position_[c]; d.direction[c] = b.direction_[c]; d.polarization[c] = b.polarization_[c]; }
      d.amplitude = b.amplitude_;
      d.radius[0] = b.radius_.size() > 0 ? b.radius_[0] : 0.0; d.radius[1] = b.radius_.size() > 1 ? b.radius_[1] : 0.0;
      d.l = b.l_;
      d.zR[0] = b.zR_.size() > 0 ? b.zR_[0] : 0.0; d.zR[1] = b.zR_.size() > 1 ? b.zR_[1] : 0.0;
      d.order[0] = b.order_.size() > 0 ? b.order_[0] : 0; d.order[1] = b.order_.size() > 1 ? b.order_[1] : 0;
      d.signal.type = (int) b.signal_.signalType_;
      d.signal.t0 = b.signal_.t0_; d.signal.s = b.signal_.s_; d.signal.f0 = b.signal_.f0_; d.signal.cep = b.signal_.cep_;
      d.signal.nR = (int) b.signal_.nR_;
      d.signal.sigma_inv_g[0] = b.signal_.sigmaInvG_.size() > 0 ? b.signal_.sigmaInvG_[0] : 0.0;
      d.signal.sigma_inv_g[1] = b.signal_.sigmaInvG_.size() > 1 ? b.signal_.sigmaInvG_[1] : 0.0;
    }

    /* Create the device solver from the state Solver::initialize() left behind: every scalar and table of the parameter
     * block is a member of the reference's Solver (all public, solver.h:23-345), the potentials are an_ / anm1_ (with the
     * seed already in them, solver.cpp:828-839), the bunch is chargeVectorn_.                                          */
    void attach (Solver& s)
    {
      if (gpu) return;
      if (s.size_ != 1) refuse("A run with more than one MPI rank");

      MithraGpuParams p; memset(&p, 0, sizeof(p));
      p.abi_version = MITHRA_GPU_ABI_VERSION;
      p.N0 = s.N0_; p.N1 = s.N1_; p.N2 = s.N2_; p.np = s.np_; p.k0 = s.k0_; p.rank = 0; p.size = 1;             /* solver.cpp:610-641 */
      p.dx = s.mesh_.meshResolution_[0]; p.dy = s.mesh_.meshResolution_[1]; p.dz = s.mesh_.meshResolution_[2]; p.dt = s.mesh_.timeStep_;
      p.xmin = s.xmin_; p.xmax = s.xmax_; p.ymin = s.ymin_; p.ymax = s.ymax_; p.zmin = s.zmin_; p.zmax = s.zmax_; /* :661-666 */
      p.zp[0] = s.zp_[0]; p.zp[1] = s.zp_[1]; p.Lz = s.mesh_.meshLength_[2];                                       /* :677-680 */
      p.solver = (int) s.mesh_.solver_; p.space_charge = s.mesh_.spaceCharge_ ? 1 : 0; p.truncation_order = s.mesh_.truncationOrder_;
      memcpy(p.a, s.uf_.a, sizeof(p.a)); p.alpha = s.uf_.af.alpha_; p.beta_nsfd = s.uf_.af.beta_;                 /* :729-739 */
      memcpy(p.bB, s.uf_.bB, sizeof(p.bB)); memcpy(p.cB, s.uf_.cB, sizeof(p.cB)); memcpy(p.dB, s.uf_.dB, sizeof(p.dB));
      memcpy(p.eE, s.uf_.eE, sizeof(p.eE)); memcpy(p.fE, s.uf_.fE, sizeof(p.fE)); memcpy(p.gE, s.uf_.gE, sizeof(p.gE));
      memcpy(p.hC, s.uf_.hC, sizeof(p.hC));                                                                       /* :745-824 */
      p.c0 = s.c0_; p.gamma = s.gamma_; p.beta = s.beta_; p.dt_shift = s.dt_;
      p.dt_bunch = s.bunch_.timeStep_;
      p.n_update_bunch = 0;
      for (Double t = 0.0; t < s.nUpdateBunch_; t += 1.0) ++p.n_update_bunch;                                     /* the loop of :1316 */
      p.r1 = s.ub_.r1; p.r2 = s.ub_.r2; p.dtb = s.ub_.dtb;                                                        /* :1053-1059 */

      if (s.undulator_.size() > MITHRA_MAX_UNDULATORS || s.extField_.size() > MITHRA_MAX_EXTFIELDS) refuse("That many undulator modules / external fields");
      p.n_undulators = (int) s.undulator_.size();
      for (size_t u = 0; u < s.undulator_.size(); u++)
	{
	  const Undulator& U = s.undulator_[u];
	  MithraUndulator& D = p.undulator[u];
	  D.type = (int) U.type_; D.k = U.k_; D.lu = U.lu_; D.rb = U.rb_; D.theta = U.theta_; D.length = U.length_; D.dist = U.dist_;
	  fillBeam(D.beam, U);
	}
      p.n_ext_fields = (int) s.extField_.size();
      for (size_t u = 0; u < s.extField_.size(); u++) fillBeam(p.ext_field[u], s.extField_[u]);
      p.seed_enabled = ( fabs(s.seed_.amplitude_) > 1.0e-50 ) ? 1 : 0;                                            /* fdtd.cpp:307 */
      fillBeam(p.seed, s.seed_);

      for (unsigned jf = 0; jf < s.FEL_.size(); jf++)
	{
	  if (s.FEL_[jf].vtkPower_.sampling_ && s.rp_[jf].Nz == 1)                                               /* radiation.cpp:238-318, :338 */
	    {
	      if (p.power_map.enabled) refuse("A second power-visualization group");
	      const SampleRadiationPower& S = s.rp_[jf];
	      p.power_map.enabled = 1; p.power_map.Nf = S.Nf; p.power_map.z = s.FEL_[jf].vtkPower_.z_; p.power_map.w = S.w[0]; p.power_map.pc = S.pc;
	    }
	  if (s.FEL_[jf].radiationPower_.sampling_)
	    {
	      if (p.power.enabled) refuse("A second radiation-power group");
	      const SampleRadiationPower& S = s.rp_[jf];                                                          /* radiation.cpp:18-121 */
	      if (S.N > MITHRA_MAX_POWER_PLANES || S.Nl > MITHRA_MAX_POWER_LAMBDAS) refuse("That many power planes / frequencies");
	      p.power.enabled = 1; p.power.N = S.N; p.power.Nl = S.Nl; p.power.Nf = S.Nf; p.power.pc = S.pc;
	      for (unsigned i = 0; i < S.N; i++)  p.power.z[i] = s.FEL_[jf].radiationPower_.z_[i];
	      for (unsigned i = 0; i < S.Nl; i++) p.power.w[i] = S.w[i];
	    }
	  if (s.FEL_[jf].screenProfile_.sampling_)
	    {
	      if (p.screens.enabled) refuse("A second screen group");
	      const std::vector<Double>& pos = s.FEL_[jf].screenProfile_.pos_;
	      if (pos.size() > MITHRA_MAX_SCREENS) refuse("That many screens");
	      p.screens.enabled = 1; p.screens.N = (int) pos.size();
	      for (size_t i = 0; i < pos.size(); i++) p.screens.pos[i] = pos[i];
	    }
	}
      p.device = -1;
      p.max_particles = s.chargeVectorn_.size() + 1024;
      p.max_screen_records = s.chargeVectorn_.size() + 4096;
      check(mithra_gpu_create(&p, &gpu));
      check(mithra_gpu_set_time(gpu, s.time_, s.timeBunch_, s.nTime_));

      /* FieldVector<double> is double[3] (fieldvector.h:22), std::vector<Double> for phi                            */
      const bool sc = s.mesh_.spaceCharge_;
      check(mithra_gpu_upload_fields(gpu, &(*s.an_)[0][0], &(*s.anm1_)[0][0], 0, sc ? &(*s.fn_)[0] : 0, sc ? &(*s.fnm1_)[0] : 0, 0));
      std::vector<double> q; q.reserve(11 * s.chargeVectorn_.size());                                             /* Charge, stdinclude.h:130-144 */
      for (auto it = s.chargeVectorn_.begin(); it != s.chargeVectorn_.end(); ++it)
	{
	  q.push_back(it->q);
	  for (int d = 0; d < 3; d++) q.push_back(it->rnp[d]);
	  for (int d = 0; d < 3; d++) q.push_back(it->rnm[d]);
	  for (int d = 0; d < 3; d++) q.push_back(it->gb[d]);
	  q.push_back(it->e);
	}
      check(mithra_gpu_upload_particles(gpu, q.empty() ? 0 : &q[0], s.chargeVectorn_.size()));
    }

    /* Will the loop call bunchSample / bunchVisualize / bunchProfile in this field step?  Its own conditions
     * (solver.cpp:1253-1270 in the particle-only loop, :1356-1371 in the main loop), evaluated a few lines earlier.   */
    bool bunchOutputDue (Solver& s)
    {
      const Double tb = s.time_ + s.mesh_.timeShift_, dt = s.mesh_.timeStep_;
      if ( s.bunch_.sampling_ && fmod(tb, s.bunch_.rhythm_) < dt && tb > 0.0 ) return true;
      if ( s.bunch_.bunchVTK_ && fmod(tb, s.bunch_.bunchVTKRhythm_) < dt && tb > 0.0 ) return true;
      if ( s.bunch_.bunchProfile_ )
	{
	  for (unsigned int i = 0; i < s.bunch_.bunchProfileTime_.size(); i++)
	    if ( s.time_ - s.bunch_.bunchProfileTime_[i] < dt && s.time_ > s.bunch_.bunchProfileTime_[i] ) return true;
	  if ( fmod(tb, s.bunch_.bunchProfileRhythm_) < dt && tb > 0.0 && s.bunch_.bunchProfileRhythm_ != 0.0 ) return true;
	}
      return false;
    }

    /* the bunch of the device back into chargeVectorn_: mithra_gpu_download_particles returns the reference's order (the
     * order of the upload), a single slab neither gains nor loses particles                                          */
    void refreshBunch (Solver& s)
    {
      size_t n = 0;
      check(mithra_gpu_num_particles(gpu, &n));
      if (n != s.chargeVectorn_.size()) refuse("A bunch whose size changed on the device");
      if (n == 0) return;
      std::vector<double> q(11 * n);
      check(mithra_gpu_download_particles(gpu, &q[0], n, &n));
      const double* r = &q[0];
      for (auto it = s.chargeVectorn_.begin(); it != s.chargeVectorn_.end(); ++it, r += 11)
	{
	  it->q = r[0];
	  for (int d = 0; d < 3; d++) { it->rnp[d] = r[1 + d]; it->rnm[d] = r[4 + d]; it->gb[d] = r[7 + d]; }
	  it->e = r[10];
	}
    }

    /* Will the loop call fieldSample / fieldVisualize* / fieldProfile in this field step (solver.cpp:1326-1351)?        */
    bool fieldOutputDue (Solver& s)
    {
      const Double dt = s.mesh_.timeStep_;
      if ( s.seed_.sampling_ && fmod(s.time_, s.seed_.samplingRhythm_) < dt && s.time_ > 0.0 ) return true;
      for (unsigned int i = 0; i < s.seed_.vtk_.size(); i++)
	if ( s.seed_.vtk_[i].sample_ && fmod(s.time_, s.seed_.vtk_[i].rhythm_) < dt && s.time_ > 0.0 ) return true;
      if ( s.seed_.profile_ )
	{
	  for (unsigned int i = 0; i < s.seed_.profileTime_.size(); i++)
	    if ( s.time_ - s.seed_.profileTime_[i] < dt && s.time_ > s.seed_.profileTime_[i] ) return true;
	  if ( fmod(s.time_, s.seed_.profileRhythm_) < dt && s.time_ > 0.0 && s.seed_.profileRhythm_ != 0 ) return true;
	}
      return false;
    }

    /* The potentials of the device (A^{n+1} just computed, A^n) back into the reference's host arrays, for its writers and
     * the lazy fieldEvaluate they call: no node evaluated yet, except that the two end planes carry the E/B of the planes
     * next to them, which the reference sets up at the end of its fieldUpdate (fdtd.cpp:742-773; one rank: both ends).   */
    void refreshFields (Solver& s)
    {
      const bool sc = s.mesh_.spaceCharge_;
      check(mithra_gpu_download_fields(gpu, &(*s.anp1_)[0][0], &(*s.an_)[0][0], &(*s.anm1_)[0][0],
				       sc ? &(*s.fnp1_)[0] : 0, sc ? &(*s.fn_)[0] : 0, sc ? &(*s.fnm1_)[0] : 0));
      /* the raw views the reference's fieldUpdate sets at the start of every step (fdtd.cpp:241-246, fdtdSC.cpp:270-280)  */
      s.uf_.anp1 = &(*s.anp1_)[0][0]; s.uf_.an = &(*s.an_)[0][0]; s.uf_.anm1 = &(*s.anm1_)[0][0]; s.uf_.jn = s.uf_.anp1;
      s.uf_.en = &s.en_[0][0]; s.uf_.bn = &s.bn_[0][0];
      if (sc) { s.uf_.fnp1 = &(*s.fnp1_)[0]; s.uf_.fn = &(*s.fn_)[0]; s.uf_.fnm1 = &(*s.fnm1_)[0]; s.uf_.rn = s.uf_.fnp1; }
      s.pic_.assign(s.pic_.size(), false);
      const long P = s.N1N0_;
      for (int i = 1; i < s.N0_ - 1; i++)
	for (int j = 1; j < s.N1_ - 1; j++)
	  {
	    const long lo = P + (long) s.N1_ * i + j, hi = P * ( s.np_ - 2 ) + (long) s.N1_ * i + j;
	    s.fieldEvaluate(lo); s.pic_[lo - P] = true; s.en_[lo - P] = s.en_[lo]; s.bn_[lo - P] = s.bn_[lo];
	    s.fieldEvaluate(hi); s.pic_[hi + P] = true; s.en_[hi + P] = s.en_[hi]; s.bn_[hi + P] = s.bn_[hi];
	  }
    }

    /* the reference's loop keeps the clocks (solver.cpp:1396-1399, :1318-1319); the library follows them                */
    void follow (Solver& s)
    {
      attach(s);
      if (subStep == 0) check(mithra_gpu_set_time(gpu, s.time_, s.timeBunch_, s.nTime_));
    }
  }

  /* ---- Solver: the four methods of the loop that are not virtual ------------------------------------------------- */

  /* solve() calls this nUpdateBunch_ times per field step and advances timeBunch_ after each call; the library runs all
   * the sub-steps of a field step in one launch (rnm = rnp included), on the first of those calls                   */
  void Solver::bunchUpdate ()
  {
    follow(*this);
    if (subStep == 0) check(mithra_gpu_bunch_update(gpu));
    int trips = 0;
    for (Double t = 0.0; t < nUpdateBunch_; t += 1.0) ++trips;
    if (++subStep >= trips)
      {
	subStep = 0;
	/* the last sub-step of the field step: the reference's own bunch writers come next in the loop                  */
	if (bunchOutputDue(*this)) refreshBunch(*this);
	if (fieldOutputDue(*this)) refreshFields(*this);
      }
  }

  void Solver::screenProfile ()
  {
    follow(*this);
    check(mithra_gpu_screen_profile(gpu));
    for (unsigned jf = 0; jf < FEL_.size(); jf++)
      {
	if (!FEL_[jf].screenProfile_.sampling_) continue;
	for (unsigned i = 0; i < FEL_[jf].screenProfile_.pos_.size(); i++)
	  {
	    size_t n = 0;
	    check(mithra_gpu_fetch_screen(gpu, (int) i, 0, 0, &n));
	    if (n == 0) continue;
	    std::vector<double> rec(6 * n);
	    check(mithra_gpu_fetch_screen(gpu, (int) i, &rec[0], n, &n));
	    for (size_t r = 0; r < n; r++)                                                                       /* solver.cpp:2229-2252 */
	      {
		for (int c = 0; c < 5; c++) *scrp_[jf].files[i] << rec[6 * r + c] << "\t";
		*scrp_[jf].files[i] << rec[6 * r + 5] << std::endl;
	      }
	  }
      }
  }

  void Solver::powerSample ()
  {
    follow(*this);
    for (unsigned jf = 0; jf < FEL_.size(); jf++)
      {
	if (!FEL_[jf].radiationPower_.sampling_) continue;
	check(mithra_gpu_power_sample(gpu));
	size_t n = 0;
	check(mithra_gpu_fetch_power(gpu, &rp_[jf].pG[0], 1, &n));
	for (unsigned l = 0; l < rp_[jf].Nl; l++)                                                               /* radiation.cpp:222-230 */
	  {
	    for (unsigned k = 0; k < rp_[jf].N; ++k)
	      *(rp_[jf].file[l]) << gamma_ * ( FEL_[jf].radiationPower_.z_[k] + beta_ * c0_ * ( timeBunch_ + dt_ ) ) << "\t" << rp_[jf].pG[k * rp_[jf].Nl + l] << "\t";
	    *(rp_[jf].file[l]) << std::endl;
	  }
      }
  }

  /* The per-pixel DFT window advances on the device in every step; at the rhythm the map comes back (pL[i N1 + j],
   * radiation.cpp:388) and is written in the reference's .vts layout (radiation.cpp:393-447): the points of the plane
   * between the two node planes around it, j outermost, then one power value per point in the same order.           */
  void Solver::powerVisualize ()
  {
    follow(*this);
    for (unsigned jf = 0; jf < FEL_.size(); jf++)
      {
	const FreeElectronLaser::RadiationVisualization& V = FEL_[jf].vtkPower_;
	if ( !( V.sampling_ && rp_[jf].Nz == 1 ) ) continue;
	check(mithra_gpu_power_visualize(gpu));
	if ( !( fmod(time_, V.rhythm_) < mesh_.timeStep_ ) ) continue;
	std::vector<double> map((size_t) N1N0_, 0.0);
	int mine = 0;
	check(mithra_gpu_fetch_power_map(gpu, &map[0], map.size(), &mine));
	Double whole;
	const Double frac  = modf( ( V.z_ - zmin_ ) / mesh_.meshResolution_[2], &whole );
	const long   plane = (long) whole - k0_;
	std::ofstream out((V.basename_ + "-" + stringify(nTime_) + VTS_FILE_SUFFIX).c_str(), std::ios::trunc);
	out.setf(std::ios::scientific);
	out.precision(4);
	out << "<?xml version=\"1.0\"?>" << std::endl
	    << "<VTKFile type=\"StructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" compressor=\"vtkZLibDataCompressor\">" << std::endl
	    << "<StructuredGrid WholeExtent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " " << 0 << " " << 0 << "\">" << std::endl
	    << "<Piece Extent=\"0 " << N0_ - 1 << " 0 " << N1_ - 1 << " " << 0 << " " << 0 << "\">" << std::endl
	    << "<Points>" << std::endl
	    << "<DataArray type = \"Float64\" NumberOfComponents=\"3\" format=\"ascii\">" << std::endl;
	for (int j = 0; j < N1_; j++)
	  for (int i = 0; i < N0_; i++)
	    {
	      const long node = plane * N1N0_ + (long) i * N1_ + j;
	      FieldVector<Double> lo = rc(node), hi = rc(node + N1N0_);
	      out << lo[0] * ( 1.0 - frac ) + hi[0] * frac << " " << lo[1] << " " << lo[2] << std::endl;
	    }
	out << "</DataArray>" << std::endl << "</Points>" << std::endl
	    << "<CellData>" << std::endl << "</CellData>" << std::endl
	    << "<PointData Vectors = \"power\">" << std::endl
	    << "<DataArray type=\"Float64\" Name=\"power\" NumberOfComponents=\"" << 1 << "\" format=\"ascii\">" << std::endl;
	for (int j = 0; j < N1_; j++)
	  for (int i = 0; i < N0_; i++)
	    out << map[(size_t) i * N1_ + j] << std::endl;
	out << "</DataArray>" << std::endl << "</PointData>" << std::endl
	    << "</Piece>" << std::endl << "</StructuredGrid>" << std::endl << "</VTKFile>" << std::endl;
	out.close();
      }
  }

  /* ---- FdTd / FdTdSC: the five time-march virtuals (the rest of fdtd.cpp / fdtdSC.cpp stays the reference's) --------- */

  #define MITHRA_GPU_FIELD_SOLVER(CLASS)                                                                                          \
    void CLASS::fieldUpdate ()        { follow(*this); check(mithra_gpu_field_update(gpu)); }                                    \
    void CLASS::fieldShift ()         { check(mithra_gpu_field_shift(gpu)); }                                                    \
    void CLASS::currentReset ()       { check(mithra_gpu_current_reset(gpu)); }                                                  \
    void CLASS::currentUpdate ()      { check(mithra_gpu_current_update(gpu)); }                                                 \
    void CLASS::currentCommunicate () { check(mithra_gpu_current_communicate(gpu)); }

  MITHRA_GPU_FIELD_SOLVER(FdTd)
  MITHRA_GPU_FIELD_SOLVER(FdTdSC)
}
