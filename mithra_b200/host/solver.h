/* solver.h -- Solver / FdTd / FdTdSC: the reference's class surface over the B200 library.
 *
 * Constructor signature, method names and public members follow src/solver.h:23-345, src/fdtd.h:18-66 and
 * src/fdtdSC.h:18-66.  Solver::initialize() runs the reference's initialisation chain on the host in FP64
 * (solver.cpp:547-595); every method of the time-march forwards to one entry point of include/mithra_gpu.h.
 * There is no CPU implementation of the march behind these classes.
 *
 * One process drives all GPUs of the box: size_ is the number of z-slabs (= GPUs in use), the per-slab quantities the
 * reference keeps per MPI rank (np_, k0_, zp_) are vectors here, and the scalar members hold slab 0's values so that
 * code reading solver.np_ etc. in a single-GPU run sees the reference's single-rank numbers.
 */
#ifndef MITHRA_B200_SOLVER_H_
#define MITHRA_B200_SOLVER_H_

#include <fstream>
#include <string>
#include <vector>

#include "classes.h"
#include "../../include/mithra_gpu.h"

namespace MITHRA
{
  /* coefficient tables of the field update, database.h:178-213 */
  struct UpdateField
  {
    Double dt, dx, dy, dz, dx2, dy2, dz2;
    Double a[6], alpha, beta;                      /* the reference keeps alpha / beta in uf_.af (AdvanceField)      */
    Double bB[5], cB[5], dB[5], eE[5], fE[5], gE[5], hC[17];
  };
  struct UpdateCurrent { Double dx, dy, dz, dv, rc; };          /* database.h:322-336                             */
  struct UpdateBunch   { Double dt, dtb, dx, dy, dz, r1, r2; }; /* database.h:286-295                             */

  struct SampleRadiationPower                                    /* database.h:339-354                             */
  {
    unsigned int                N, Nl, Nf;
    Double                      pc;
    std::vector<Double>         w;
    std::vector<std::ofstream*> file;
    SampleRadiationPower () : N(0), Nl(0), Nf(0), pc(0.0) {}
  };
  struct SampleScreenProfile { std::vector<std::string> fileNames; std::vector<std::ofstream*> files; };

  class Solver
  {
  public:
    Solver (Mesh& mesh, Bunch& bunch, Seed& seed, std::vector<Undulator>& undulator, std::vector<ExtField>& extField,
	    std::vector<FreeElectronLaser>& FEL);
    virtual ~Solver ();

    /* ---- initialisation chain, solver.cpp:68-1206 ------------------------------------------------------------ */
    void initialize ();
    void setSimulationParameters ();
    void lorentzBoostMesh ();
    void lorentzBoostBunch ();
    void distributeParticles (std::list<Charge>& chargeVector);
    void computeFileGamma (BunchInitialize& bunchInit);
    void initializeMesh ();
    void initializeSeedSampling ();                            /* solver.cpp:848-931 */
    void initializeSeedVTK ();
    void initializeSeedProfile ();                 /* solver.cpp:1022-1044 */                                 /* solver.cpp:938-1016 */
    void initializeField ();
    void initializeBunchUpdate ();
    void initializeBunch ();
    void initializePowerSample ();
    void initializePowerVisualize ();
    void initializeScreenProfile ();
    void shiftBackInTime ();

    /* ---- the time march, solver.cpp:1212-1576 ---------------------------------------------------------------- */
    void solve ();
    void bunchUpdate ();                           /* all nUpdateBunch_ sub-steps of one field step + rnm = rnp      */
    void recycleParticles () {}                    /* ownership is the slab's particle list; migration does the rest */
    void screenProfile ();
    void powerSample ();
    void finalize ();

    /* rhythm-gated writers of the reference that are outside the hot path: accepted, not run */
    void bunchSample ();
    void bunchVisualize ();                        /* solver.cpp:1647-1757: .vtu / .pvtu of the particle cloud          */
    void bunchProfile ();
    void powerVisualize ();
    void energySample () {}                        /* unreachable in the reference too: its parser stores a radiation-energy group in
						      radiationPower_ (datainput.cpp:706-717), radiationEnergy_.sampling_ stays false */

    /* ---- the thirteen virtuals, solver.h:139-178 ------------------------------------------------------------ */
    virtual void currentReset () = 0;
    virtual void currentUpdate () = 0;
    virtual void currentCommunicate () = 0;
    virtual void fieldUpdate () = 0;
    virtual void fieldShift () = 0;
    virtual void fieldEvaluate (long int m) = 0;
    virtual void fieldSample () = 0;
    virtual void fieldVisualizeAllDomain (unsigned int ivtk) = 0;
    virtual void fieldVisualizeInPlane (unsigned int ivtk) = 0;
    virtual void fieldVisualizeInPlaneXNormal (unsigned int ivtk) = 0;
    virtual void fieldVisualizeInPlaneYNormal (unsigned int ivtk) = 0;
    virtual void fieldVisualizeInPlaneZNormal (unsigned int ivtk) = 0;
    virtual void fieldProfile () = 0;

    /* ---- helpers, solver.cpp:2276-2315 ---------------------------------------------------------------------- */
    static bool undulatorCompare (Undulator i, Undulator j) { return i.rb_ < j.rb_; }
    Double      interp (Double x0, Double x1, Double y0, Double y1, Double x) { return y0 + ( x - x0 ) / ( x1 - x0 ) * ( y1 - y0 ); }
    bool        particleInProcessor (const Double& z);
    FieldVector rc (const long int& m);

    /* ---- this build ----------------------------------------------------------------------------------------- */
    void setNumberOfGpus (int n)    { size_ = n < 1 ? 1 : n; }
    void setMaxSteps (long n)       { maxSteps_ = n; }
    void fillParams (MithraGpuParams& p, int slab) const;      /* parameter block of one slab                      */
    void attachGpu ();                                         /* create the handles, upload the initial state     */
    void dumpParams (const std::string& prefix);               /* record file in oracle/ref_dump's meta format     */
    void flushOutputs ();
    void check (int rc) const;

    Mesh&                           mesh_;
    Bunch&                          bunch_;
    Seed&                           seed_;
    std::vector<Undulator>&         undulator_;
    std::vector<ExtField>&          extField_;
    std::vector<FreeElectronLaser>& FEL_;

    int          N0_, N1_, N2_, N1N0_, np_, k0_;
    int          rank_, size_;
    Double       xmin_, xmax_, ymin_, ymax_, zmin_, zmax_, zp_[2];
    std::vector<int>    slabNp_, slabK0_;
    std::vector<Double> slabZp0_, slabZp1_;
    Double       c0_, m0_, e0_;
    Double       gamma_, beta_, dt_;
    Double       timep1_, time_, timem1_, timeBunch_;
    unsigned int nTime_, nTimeBunch_, Nc_;
    Double       nUpdateBunch_;
    UpdateField   uf_;
    UpdateCurrent uc_;
    UpdateBunch   ub_;
    std::vector<SampleRadiationPower> rp_;
    std::vector<SampleScreenProfile>  scrp_;
    ChargeVector chargeVectorn_;
    /* Large Halton ellipsoids are generated, boosted and back-projected on the device and handed to the slabs there
     * (mithra_gpu_bunch_*): chargeVectorn_ then stays empty.  MITHRA_DEVICE_BUNCH=1 forces it for any size,
     * MITHRA_HOST_BUNCH=1 keeps the host path (bit-identical to the reference's list; the device list is to 1e-15). */
    MithraGpuBunch* deviceBunch_;
    bool wantDeviceBunch () const;

    std::vector<MithraGpu*> gpu_;                  /* one handle per slab / GPU                                      */

  protected:
    long                maxSteps_;
    int                 powerGroup_, screenGroup_;     /* FEL_ entries the C ABI's single power / screen group mirror */
    int                 pmapGroup_;                    /* ... and its single power-visualization group               */
    std::ofstream*      bunchSampleFile_;              /* sb_.file, solver.cpp:1073                                  */
    std::ofstream*      fieldSampleFile_;              /* sf_.file, solver.cpp:918                                   */
    Double              sfCe_, sfCb_, sfCa_;           /* sf_.Ce, Cb, Ca, solver.cpp:922-924                         */
    std::vector<Double> powerTimes_;                   /* timeBunch_ of the sampled steps not yet written             */
    bool                spaceChargeSolver_;
  };

  class FdTd : public Solver
  {
  public:
    FdTd (Mesh& mesh, Bunch& bunch, Seed& seed, std::vector<Undulator>& undulator, std::vector<ExtField>& extField,
	  std::vector<FreeElectronLaser>& FEL);
    void currentReset ();
    void currentUpdate ();
    void currentCommunicate ();
    void fieldUpdate ();
    void fieldShift ();
    void fieldEvaluate (long int) {}               /* E, B are evaluated eagerly on the device (eval_eb_box)         */
    void fieldSample ();                           /* fdtd.cpp:851-950: values from mithra_gpu_field_sample, the reference's line */
    void fieldVisualizeAllDomain (unsigned int ivtk);            /* fdtd.cpp:956-1105 */
    void fieldVisualizeInPlane (unsigned int ivtk);              /* fdtd.cpp:1111-1121 */
    void fieldVisualizeInPlaneXNormal (unsigned int ivtk);       /* fdtd.cpp:1128-1285 */
    void fieldVisualizeInPlaneYNormal (unsigned int ivtk);       /* fdtd.cpp:1292-1447 */
    void fieldVisualizeInPlaneZNormal (unsigned int ivtk);       /* fdtd.cpp:1452-1540 */
    void nodeValues (const std::vector<int>& ijk, std::vector<double>& val);   /* en_, bn_, an_ at global nodes */
    void fieldProfile ();                                        /* fdtd.cpp:1546-1594 */
  };

  /* identical forwarding: the library switches to the A + phi kernels when MithraGpuParams.space_charge is set     */
  class FdTdSC : public FdTd
  {
  public:
    FdTdSC (Mesh& mesh, Bunch& bunch, Seed& seed, std::vector<Undulator>& undulator, std::vector<ExtField>& extField,
	    std::vector<FreeElectronLaser>& FEL);
  };
}

#endif
