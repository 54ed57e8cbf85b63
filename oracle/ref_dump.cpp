/* ref_dump : state-dump driver around the UNMODIFIED MITHRA reference (test infrastructure, oracle/ only).
 *
 * It is compiled together with the reference sources where they lie (/root/reference/src, see
 * oracle/Makefile) into oracle/_ref/ref_dump. It builds the reference objects exactly like the reference
 * main() does (mithra.cpp:48-90), calls Solver::initialize() and then replays the body of Solver::solve()
 * (solver.cpp:1232-1414) by calling the reference's own public methods in the reference's order, writing
 * the solver state to disk at chosen points. No reference file is modified or copied; every Solver member
 * is public (solver.h:23-345), which is what makes this possible.
 *
 * usage: ref_dump <job-file> <out-prefix> <nsteps> [--full-at a,b,c] [--phases-at s] [--quiet] [--bench W]
 *   <nsteps>          number of field steps of the second while loop to run (0 = only initialise)
 *   --full-at LIST    write <prefix>.full<step>.bin holding the state at the START of field step <step>
 *                     (step 0 = right after initialize()); step == nsteps is allowed (= final state)
 *   --no-fields       ... without the field arrays (times and particles only)
 *   --init-only       write <prefix>.meta.bin and <prefix>.full0.bin (times and particles, no fields) right after
 *                     Solver::initialize() and stop -- before the particle-only loop of a job with an
 *                     initial-time-back-shift (tools/check_shipped_jobs.py); with several mini-MPI ranks every rank
 *                     writes <prefix>.rank<r>.bin instead: its slab (np, k0, zp) and its particles
 *   --phases-at s     additionally write <prefix>.phase<s>.bin with the intermediate arrays of step s
 *   --bench W         CPU-baseline mode (any number of mini-MPI ranks, MINIMPI_NP): no state dumps; W untimed
 *                     warm-up steps, then <nsteps> field steps timed between two MPI_Barriers; rank 0 prints one
 *                     line "BENCH {json}" with the wall seconds, the global node count and the particle count
 *   always written:   <prefix>.meta.bin (scalars, coefficient tables) and <prefix>.power.bin (pG per step)
 *
 * Record format (little endian): repeated { char name[48]; int32 dtype (0=f64,1=f32,2=i32,3=u8);
 * int64 count; payload }.
 */

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "classes.h"
#include "database.h"
#include "datainput.h"
#include "fdtd.h"
#include "fdtdSC.h"
#include "fieldvector.h"
#include "readdata.h"
#include "solver.h"

using namespace MITHRA;

namespace
{
  struct Writer
  {
    FILE* f;
    explicit Writer (const std::string& name) { f = fopen(name.c_str(), "wb"); if (!f) { perror(name.c_str()); exit(2); } }
    ~Writer () { if (f) fclose(f); }

    void raw (const char* name, int32_t dtype, int64_t count, const void* data, size_t elem)
    {
      char nm[48]; memset(nm, 0, sizeof(nm)); strncpy(nm, name, 47);
      fwrite(nm, 1, 48, f); fwrite(&dtype, 4, 1, f); fwrite(&count, 8, 1, f);
      if (count > 0) fwrite(data, elem, (size_t) count, f);
    }
    void f64 (const char* n, const double* p, int64_t c) { raw(n, 0, c, p, 8); }
    void f32 (const char* n, const float* p,  int64_t c) { raw(n, 1, c, p, 4); }
    void i32 (const char* n, const int* p,    int64_t c) { raw(n, 2, c, p, 4); }
    void u8  (const char* n, const unsigned char* p, int64_t c) { raw(n, 3, c, p, 1); }
    void d   (const char* n, double v) { f64(n, &v, 1); }
    void i   (const char* n, int v)    { i32(n, &v, 1); }
  };

  std::vector<double> particles (Solver& s)
  {
    std::vector<double> v; v.reserve(s.chargeVectorn_.size() * 11);
    for (auto it = s.chargeVectorn_.begin(); it != s.chargeVectorn_.end(); ++it)
      {
	v.push_back(it->q);
	for (int c = 0; c < 3; c++) v.push_back(it->rnp[c]);
	for (int c = 0; c < 3; c++) v.push_back(it->rnm[c]);
	for (int c = 0; c < 3; c++) v.push_back(it->gb[c]);
	v.push_back(it->e);
      }
    return v;
  }

  void dumpFields (Writer& w, Solver& s, bool sc, const char* tagNp1)
  {
    const int64_t n = (int64_t) s.N1N0_ * s.np_;
    w.f64("an",   &(*s.an_)  [0][0], 3 * n);
    w.f64("anm1", &(*s.anm1_)[0][0], 3 * n);
    w.f64(tagNp1, &(*s.anp1_)[0][0], 3 * n);
    if (sc)
      {
	std::string t = std::string(tagNp1) == "anp1" ? "fnp1" : "rho";
	w.f64("fn",   &(*s.fn_)  [0], n);
	w.f64("fnm1", &(*s.fnm1_)[0], n);
	w.f64(t.c_str(), &(*s.fnp1_)[0], n);
      }
  }

  void dumpEB (Writer& w, Solver& s)
  {
    const int64_t n = (int64_t) s.N1N0_ * s.np_;
    w.f32("en", &s.en_[0][0], 3 * n);
    w.f32("bn", &s.bn_[0][0], 3 * n);
    std::vector<unsigned char> pic((size_t) n);
    for (int64_t i = 0; i < n; i++) pic[(size_t) i] = s.pic_[(size_t) i] ? 1 : 0;
    w.u8("pic", &pic[0], n);
  }

  void dumpTimes (Writer& w, Solver& s)
  {
    w.d("time", s.time_); w.d("timem1", s.timem1_); w.d("timep1", s.timep1_);
    w.d("timeBunch", s.timeBunch_); w.i("nTime", (int) s.nTime_); w.i("nTimeBunch", (int) s.nTimeBunch_);
  }

  bool noFields = false;           /* --no-fields: the full dumps hold times and particles only              */

  void dumpFull (const std::string& prefix, int step, Solver& s, bool sc)
  {
    std::ostringstream nm; nm << prefix << ".full" << step << ".bin";
    Writer w(nm.str());
    w.i("step", step);
    dumpTimes(w, s);
    if (!noFields) dumpFields(w, s, sc, "jn");      /* at the start of a step anp1_ holds the deposited current      */
    std::vector<double> p = particles(s);
    w.f64("particles", p.empty() ? 0 : &p[0], (int64_t) p.size());
  }

  /* Seed, Undulator and ExtField carry the same beam members (classes.h:208-262, 303-352, 383-427). */
  template <class T>
  void dumpBeam (Writer& w, const std::string& key, const T& b)
  {
    double o[18] = { (double) b.seedType_, b.position_[0], b.position_[1], b.position_[2], b.direction_[0], b.direction_[1], b.direction_[2],
		     b.polarization_[0], b.polarization_[1], b.polarization_[2], b.amplitude_,
		     b.radius_.size() > 0 ? b.radius_[0] : 0.0, b.radius_.size() > 1 ? b.radius_[1] : 0.0, b.l_,
		     b.zR_.size() > 0 ? b.zR_[0] : 0.0, b.zR_.size() > 1 ? b.zR_[1] : 0.0,
		     b.order_.size() > 0 ? (double) b.order_[0] : 0.0, b.order_.size() > 1 ? (double) b.order_[1] : 0.0 };
    w.f64((key + "beam").c_str(), o, 18);
    double g[8] = { (double) b.signal_.signalType_, b.signal_.t0_, b.signal_.s_, b.signal_.f0_, (double) b.signal_.nR_, b.signal_.cep_,
		    b.signal_.sigmaInvG_.size() > 0 ? b.signal_.sigmaInvG_[0] : 0.0, b.signal_.sigmaInvG_.size() > 1 ? b.signal_.sigmaInvG_[1] : 0.0 };
    w.f64((key + "sig").c_str(), g, 8);
  }

  std::set<int> parseList (const char* a)
  {
    std::set<int> out; std::stringstream ss(a); std::string tok;
    while (std::getline(ss, tok, ',')) if (!tok.empty()) out.insert(atoi(tok.c_str()));
    return out;
  }
}

int main (int argc, char* argv[])
{
  MPI_Init(&argc, &argv);

  if (argc < 4) { fprintf(stderr, "usage: ref_dump <job> <out-prefix> <nsteps> [--full-at a,b] [--phases-at s] [--quiet]\n"); return 2; }
  const std::string prefix = argv[2];
  const int nsteps = atoi(argv[3]);
  std::set<int> fullAt; int phasesAt = -1; bool quiet = false; int benchWarm = -1; bool initOnly = false;
  for (int a = 4; a < argc; a++)
    {
      if      (!strcmp(argv[a], "--full-at")   && a + 1 < argc) fullAt = parseList(argv[++a]);
      else if (!strcmp(argv[a], "--phases-at") && a + 1 < argc) phasesAt = atoi(argv[++a]);
      else if (!strcmp(argv[a], "--quiet")) quiet = true;
      else if (!strcmp(argv[a], "--no-fields")) noFields = true;
      else if (!strcmp(argv[a], "--init-only")) initOnly = true;
      else if (!strcmp(argv[a], "--bench")     && a + 1 < argc) benchWarm = atoi(argv[++a]);
    }

  std::streambuf* coutBuf = std::cout.rdbuf();
  std::ostringstream sink;
  if (quiet) std::cout.rdbuf(sink.rdbuf());

  /* Same construction sequence as the reference main(), mithra.cpp:48-90. */
  std::list<std::string> jobFile = read_file(argv[1]);
  cleanJobFile(jobFile);
  Mesh mesh; mesh.initialize();
  Bunch bunch; Seed seed;
  std::vector<Undulator> undulator; std::vector<ExtField> extField; std::vector<FreeElectronLaser> FEL;
  ParseDarius parser(jobFile, mesh, bunch, seed, undulator, extField, FEL);
  parser.setJobParameters();

  Solver* sp;
  const bool sc = mesh.spaceCharge_;
  if (sc) sp = new FdTdSC(mesh, bunch, seed, undulator, extField, FEL);
  else    sp = new FdTd  (mesh, bunch, seed, undulator, extField, FEL);
  Solver& s = *sp;

  s.initialize();

  /* ---- meta ------------------------------------------------------------------------------------- */
  if (s.rank_ == 0)
  {
    Writer w(prefix + ".meta.bin");
    w.i("N0", s.N0_); w.i("N1", s.N1_); w.i("N2", s.N2_); w.i("np", s.np_); w.i("k0", s.k0_);
    w.i("rank", s.rank_); w.i("size", s.size_);
    w.i("spaceCharge", sc ? 1 : 0); w.i("solver", (int) mesh.solver_); w.i("truncationOrder", (int) mesh.truncationOrder_);
    w.d("dx", mesh.meshResolution_[0]); w.d("dy", mesh.meshResolution_[1]); w.d("dz", mesh.meshResolution_[2]);
    w.d("Lx", mesh.meshLength_[0]); w.d("Ly", mesh.meshLength_[1]); w.d("Lz", mesh.meshLength_[2]);
    w.d("dt", mesh.timeStep_); w.d("dtBunch", bunch.timeStep_); w.d("nUpdateBunch", s.nUpdateBunch_);
    w.d("totalTime", mesh.totalTime_); w.d("timeShift", mesh.timeShift_);
    w.d("xmin", s.xmin_); w.d("xmax", s.xmax_); w.d("ymin", s.ymin_); w.d("ymax", s.ymax_); w.d("zmin", s.zmin_); w.d("zmax", s.zmax_);
    w.f64("zp", s.zp_, 2);
    w.d("gamma", s.gamma_); w.d("beta", s.beta_); w.d("dtShift", s.dt_);
    w.d("c0", s.c0_); w.d("m0", s.m0_); w.d("e0", s.e0_);
    w.f64("a", s.uf_.a, 6); w.d("alpha", s.uf_.af.alpha_); w.d("betaNSFD", s.uf_.af.beta_);
    w.f64("bB", s.uf_.bB, 5); w.f64("cB", s.uf_.cB, 5); w.f64("dB", s.uf_.dB, 5);
    w.f64("eE", s.uf_.eE, 5); w.f64("fE", s.uf_.fE, 5); w.f64("gE", s.uf_.gE, 5);
    w.f64("hC", s.uf_.hC, 17);
    w.d("dv", s.uc_.dv); w.d("rc", s.uc_.rc);
    w.d("r1", s.ub_.r1); w.d("r2", s.ub_.r2); w.d("dtb", s.ub_.dtb);
    w.d("seedAmplitude", seed.amplitude_);
    dumpBeam(w, "seed.", seed);
    w.i("nExtFields", (int) extField.size());
    for (size_t u = 0; u < extField.size(); u++)
      { std::ostringstream k; k << "ext" << u << "."; dumpBeam(w, k.str(), extField[u]); }
    w.i("nUndulators", (int) undulator.size());
    for (size_t u = 0; u < undulator.size(); u++)
      {
	std::ostringstream k; k << "und" << u << ".";
	const Undulator& U = undulator[u];
	double v[8] = { U.k_, U.lu_, U.rb_, (double) U.length_, U.dist_, U.theta_, (double) U.type_, (double) U.seedType_ };
	w.f64((k.str() + "static").c_str(), v, 8);
	double o[16] = { U.position_[0], U.position_[1], U.position_[2], U.direction_[0], U.direction_[1], U.direction_[2],
			 U.polarization_[0], U.polarization_[1], U.polarization_[2], U.amplitude_, U.a0_,
			 U.radius_.size() > 0 ? U.radius_[0] : 0.0, U.radius_.size() > 1 ? U.radius_[1] : 0.0, U.l_,
			 U.zR_.size() > 0 ? U.zR_[0] : 0.0, U.zR_.size() > 1 ? U.zR_[1] : 0.0 };
	w.f64((k.str() + "optical").c_str(), o, 16);
	double g[6] = { (double) U.signal_.signalType_, U.signal_.t0_, U.signal_.s_, U.signal_.f0_, (double) U.signal_.nR_, U.signal_.cep_ };
	w.f64((k.str() + "signal").c_str(), g, 6);
	dumpBeam(w, k.str(), U);
      }
    w.i("nFEL", (int) FEL.size());
    for (size_t jf = 0; jf < FEL.size(); jf++)
      {
	if (!FEL[jf].radiationPower_.sampling_) continue;
	std::ostringstream k; k << "power" << jf << ".";
	w.i((k.str() + "N").c_str(), (int) s.rp_[jf].N); w.i((k.str() + "Nl").c_str(), (int) s.rp_[jf].Nl);
	w.i((k.str() + "Nf").c_str(), (int) s.rp_[jf].Nf); w.d((k.str() + "pc").c_str(), s.rp_[jf].pc);
	w.f64((k.str() + "z").c_str(), &FEL[jf].radiationPower_.z_[0], (int64_t) FEL[jf].radiationPower_.z_.size());
	w.f64((k.str() + "w").c_str(), &s.rp_[jf].w[0], (int64_t) s.rp_[jf].w.size());
      }
    for (size_t jf = 0; jf < FEL.size(); jf++)
      {
	/* power-visualization group: Solver::initializePowerVisualize, radiation.cpp:238-318                      */
	if (!FEL[jf].vtkPower_.sampling_) continue;
	std::ostringstream k; k << "pmap" << jf << ".";
	w.i((k.str() + "Nf").c_str(), (int) s.rp_[jf].Nf); w.d((k.str() + "pc").c_str(), s.rp_[jf].pc);
	w.d((k.str() + "z").c_str(), FEL[jf].vtkPower_.z_); w.d((k.str() + "w").c_str(), s.rp_[jf].w.empty() ? 0.0 : s.rp_[jf].w[0]);
	w.d((k.str() + "rhythm").c_str(), FEL[jf].vtkPower_.rhythm_);
      }
    for (size_t jf = 0; jf < FEL.size(); jf++)
      {
	if (!FEL[jf].screenProfile_.sampling_) continue;
	std::ostringstream k; k << "screen" << jf << ".pos";
	w.f64(k.str().c_str(), &FEL[jf].screenProfile_.pos_[0], (int64_t) FEL[jf].screenProfile_.pos_.size());
      }
    dumpTimes(w, s);
  }

  if (initOnly)
    {
      /* the state right after Solver::initialize(), before either loop of solve(): times and particles            */
      noFields = true;
      if (s.size_ == 1) dumpFull(prefix, 0, s, sc);
      else
	{
	  /* several ranks (MINIMPI_NP): every rank writes its slab of the partition (solver.cpp:619-641) and the particles
	   * distributeParticles left it with (solver.cpp:429-487) to <prefix>.rank<r>.bin                                */
	  std::ostringstream nm; nm << prefix << ".rank" << s.rank_ << ".bin";
	  Writer w(nm.str());
	  w.i("rank", s.rank_); w.i("size", s.size_); w.i("np", s.np_); w.i("k0", s.k0_);
	  w.f64("zp", s.zp_, 2);
	  std::vector<double> p = particles(s);
	  w.f64("particles", p.empty() ? 0 : &p[0], (int64_t) p.size());
	}
      MPI_Finalize();
      return 0;
    }

  std::vector<double> powerSeries;

  /* ---- first loop of solve(): particles only while time_ < 0 (solver.cpp:1232-1291) -------------- */
  while (s.time_ < 0.0)
    {
      for (auto iter = s.chargeVectorn_.begin(); iter != s.chargeVectorn_.end(); iter++) iter->rnm = iter->rnp;
      for (Double t = 0.0; t < s.nUpdateBunch_; t += 1.0) { s.bunchUpdate(); s.timeBunch_ += bunch.timeStep_; ++s.nTimeBunch_; }
      s.screenProfile();
      /* the bunch samplers of the first loop, solver.cpp:1253-1270                                   */
      if (bunch.sampling_ && fmod(s.time_ + mesh.timeShift_, bunch.rhythm_) < mesh.timeStep_ && (s.time_ + mesh.timeShift_ > 0.0)) s.bunchSample();
      if (bunch.bunchVTK_ && fmod(s.time_ + mesh.timeShift_, bunch.bunchVTKRhythm_) < mesh.timeStep_ && (s.time_ + mesh.timeShift_ > 0.0)) s.bunchVisualize();
      if (bunch.bunchProfile_)
	{
	  for (unsigned int i = 0; i < bunch.bunchProfileTime_.size(); i++)
	    if (s.time_ - bunch.bunchProfileTime_[i] < mesh.timeStep_ && s.time_ > bunch.bunchProfileTime_[i]) s.bunchProfile();
	  if (fmod(s.time_ + mesh.timeShift_, bunch.bunchProfileRhythm_) < mesh.timeStep_ && (s.time_ + mesh.timeShift_ > 0.0) && (bunch.bunchProfileRhythm_ != 0.0)) s.bunchProfile();
	}
      s.recycleParticles();
      s.timem1_ += mesh.timeStep_; s.time_ += mesh.timeStep_; s.timep1_ += mesh.timeStep_; ++s.nTime_;
    }

  /* ---- second loop of solve(): solver.cpp:1300-1414 --------------------------------------------- */
  const bool bench = benchWarm >= 0;
  std::chrono::steady_clock::time_point tBegin;
  long nPushes = 0;
  if (bench && benchWarm == 0) { MPI_Barrier(MPI_COMM_WORLD); tBegin = std::chrono::steady_clock::now(); }
  const int nLoop = bench ? nsteps + benchWarm : nsteps;
  for (int step = 0; step < nLoop; step++)
    {
      if (bench && benchWarm > 0 && step == benchWarm) { MPI_Barrier(MPI_COMM_WORLD); tBegin = std::chrono::steady_clock::now(); nPushes = 0; }
      if (bench) nPushes += (long) s.chargeVectorn_.size() * (long) s.nUpdateBunch_;
      if (fullAt.count(step)) dumpFull(prefix, step, s, sc);
      Writer* ph = 0;
      if (step == phasesAt)
	{
	  std::ostringstream nm; nm << prefix << ".phase" << step << ".bin";
	  ph = new Writer(nm.str());
	  ph->i("step", step); dumpTimes(*ph, s);
	}

      s.fieldUpdate();
      if (ph)
	{
	  const int64_t n = (int64_t) s.N1N0_ * s.np_;
	  ph->f64("anp1_after_fieldUpdate", &(*s.anp1_)[0][0], 3 * n);
	  if (sc) ph->f64("fnp1_after_fieldUpdate", &(*s.fnp1_)[0], n);
	}

      for (auto iter = s.chargeVectorn_.begin(); iter != s.chargeVectorn_.end(); iter++) iter->rnm = iter->rnp;
      for (Double t = 0.0; t < s.nUpdateBunch_; t += 1.0) { s.bunchUpdate(); s.timeBunch_ += bunch.timeStep_; ++s.nTimeBunch_; }
      s.recycleParticles();
      if (ph)
	{
	  std::vector<double> p = particles(s);
	  ph->f64("particles_after_push", p.empty() ? 0 : &p[0], (int64_t) p.size());
	}

      if (seed.sampling_ && fmod(s.time_, seed.samplingRhythm_) < mesh.timeStep_ && s.time_ > 0.0) s.fieldSample();
      for (unsigned int i = 0; i < seed.vtk_.size(); i++)
	if (seed.vtk_[i].sample_ && fmod(s.time_, seed.vtk_[i].rhythm_) < mesh.timeStep_ && s.time_ > 0.0)
	  {
	    if      (seed.vtk_[i].type_ == ALLDOMAIN) s.fieldVisualizeAllDomain(i);
	    else if (seed.vtk_[i].type_ == INPLANE)   s.fieldVisualizeInPlane(i);
	  }
      if (seed.profile_)
	{
	  for (unsigned int i = 0; i < seed.profileTime_.size(); i++)
	    if (s.time_ - seed.profileTime_[i] < mesh.timeStep_ && s.time_ > seed.profileTime_[i]) s.fieldProfile();
	  if (fmod(s.time_, seed.profileRhythm_) < mesh.timeStep_ && s.time_ > 0.0 && seed.profileRhythm_ != 0) s.fieldProfile();
	}
      if (bunch.sampling_ && fmod(s.time_ + mesh.timeShift_, bunch.rhythm_) < mesh.timeStep_ && (s.time_ + mesh.timeShift_ > 0.0)) s.bunchSample();
      if (bunch.bunchVTK_ && fmod(s.time_ + mesh.timeShift_, bunch.bunchVTKRhythm_) < mesh.timeStep_ && (s.time_ + mesh.timeShift_ > 0.0)) s.bunchVisualize();
      if (bunch.bunchProfile_)
	{
	  for (unsigned int i = 0; i < bunch.bunchProfileTime_.size(); i++)
	    if (s.time_ - bunch.bunchProfileTime_[i] < mesh.timeStep_ && s.time_ > bunch.bunchProfileTime_[i]) s.bunchProfile();
	  if (fmod(s.time_ + mesh.timeShift_, bunch.bunchProfileRhythm_) < mesh.timeStep_ && (s.time_ + mesh.timeShift_ > 0.0) && (bunch.bunchProfileRhythm_ != 0.0)) s.bunchProfile();
	}

      s.screenProfile();
      s.powerSample(); s.powerVisualize();
      s.energySample();

      for (size_t jf = 0; jf < FEL.size(); jf++)
	if (FEL[jf].radiationPower_.sampling_)
	  for (size_t q = 0; q < s.rp_[jf].pG.size(); q++) powerSeries.push_back(s.rp_[jf].pG[q]);

      if (ph) dumpEB(*ph, s);

      s.fieldShift();
      s.currentReset();
      s.currentUpdate();
      s.currentCommunicate();
      if (ph)
	{
	  const int64_t n = (int64_t) s.N1N0_ * s.np_;
	  ph->f64("jn_after_deposit", &(*s.anp1_)[0][0], 3 * n);
	  if (sc) ph->f64("rho_after_deposit", &(*s.fnp1_)[0], n);
	  delete ph;
	}

      s.timem1_ += mesh.timeStep_; s.time_ += mesh.timeStep_; s.timep1_ += mesh.timeStep_; ++s.nTime_;
    }
  if (bench)
    {
      MPI_Barrier(MPI_COMM_WORLD);
      const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - tBegin).count();
      double mine[2] = { (double) nPushes, (double) s.chargeVectorn_.size() }, all[2] = { 0.0, 0.0 };
      MPI_Reduce(mine, all, 2, MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD);
      if (s.rank_ == 0)
	{
	  if (quiet) std::cout.rdbuf(coutBuf);
	  printf("BENCH {\"seconds\": %.6f, \"steps\": %d, \"warmup\": %d, \"ranks\": %d, \"N0\": %d, \"N1\": %d, \"N2\": %d, "
		 "\"pushes\": %.0f, \"particles\": %.0f, \"sub_steps\": %d, \"space_charge\": %d}\n",
		 sec, nsteps, benchWarm, s.size_, s.N0_, s.N1_, s.N2_, all[0], all[1], (int) s.nUpdateBunch_, sc ? 1 : 0);
	  fflush(stdout);
	  if (quiet) std::cout.rdbuf(sink.rdbuf());
	}
    }
  if (fullAt.count(nsteps)) dumpFull(prefix, nsteps, s, sc);

  if (s.rank_ == 0)
  {
    Writer w(prefix + ".power.bin");
    w.i("nsteps", nsteps);
    w.f64("pG", powerSeries.empty() ? 0 : &powerSeries[0], (int64_t) powerSeries.size());
    /* the per-pixel map of the last powerVisualize call (rp_[jf].pL, radiation.cpp:388)                         */
    for (size_t jf = 0; jf < FEL.size(); jf++)
      if (FEL[jf].vtkPower_.sampling_ && s.rp_[jf].Nz == 1 && !s.rp_[jf].pL.empty())
	{
	  std::ostringstream k; k << "pmap" << jf;
	  w.f64(k.str().c_str(), &s.rp_[jf].pL[0], (int64_t) s.rp_[jf].pL.size());
	}
  }

  s.finalize();
  if (quiet) std::cout.rdbuf(coutBuf);
  MPI_Finalize();
  return 0;
}
