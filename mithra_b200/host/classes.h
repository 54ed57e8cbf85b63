/* classes.h -- parameter model of a MITHRA job: Mesh, Bunch, Signal, Seed, Undulator, ExtField, FreeElectronLaser.
 *
 * Same class and member names as the reference (src/classes.h:20-578, src/database.h:22-75, src/stdinclude.h:18-144)
 * for everything the time-march and its initialisation read, so code written against the reference's structs
 * compiles against these; members that only feed out-of-scope writers (VTK visualisation, field profiles) are parsed
 * and kept but nothing consumes them.  Seed, Undulator and ExtField share one `Beam` base here: in the reference they
 * are three classes repeating the same members (classes.h:208-262, 303-352, 383-427).
 */
#ifndef MITHRA_B200_CLASSES_H_
#define MITHRA_B200_CLASSES_H_

#include <cmath>
#include <list>
#include <string>
#include <vector>

#include "jobfile.h"

namespace MITHRA
{
  /* enums with the reference's values (stdinclude.h:18-40): they cross the C ABI as ints                  */
  enum SignalType    { NEUMANN, GAUSSIAN, SECANT, FLATTOP, INVGAUSSIAN };
  enum SeedType      { PLANEWAVE, PLANEWAVETRUNCATED, GAUSSIANBEAM, SUPERGAUSSIANBEAM,
		       STANDINGPLANEWAVE, STANDINGPLANEWAVETRUNCATED, STANDINGGAUSSIANBEAM, STANDINGSUPERGAUSSIANBEAM };
  enum ExtFieldType  { EMWAVE };
  enum SamplingType  { ATPOINT, OVERLINE, INPLANE, ALLDOMAIN };
  enum PlaneType     { XNORMAL, YNORMAL, ZNORMAL };
  enum FieldType     { Ex, Ey, Ez, Bx, By, Bz, Ax, Ay, Az, F };
  enum UndulatorType { STATIC, OPTICAL };
  enum SolverType    { FD, NSFD };

  /* constants with the reference's (truncated) digits, stdinclude.h:43-52 -- parity is against the code as shipped */
  const Double PI           = 3.1415926535;
  const Double EPSILON_ZERO = 8.85418782e-12;
  const Double MU_ZERO      = 4.0 * PI * 1.0e-7;
  const Double C0           = 1.0 / sqrt(EPSILON_ZERO * MU_ZERO);
  const Double EC           = 1.602e-19;
  const Double EM           = 9.109e-31;

  /* FieldVector<double> of the reference (fieldvector.h:22-204): three doubles, same algebra, same operation order */
  struct FieldVector
  {
    Double a[3];
    FieldVector (Double v = 0.0) { a[0] = a[1] = a[2] = v; }
    Double&       operator[] (unsigned n)       { return a[n]; }
    const Double& operator[] (unsigned n) const { return a[n]; }
    Double norm2 () const { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
    Double norm  () const { return sqrt(norm2()); }
    void mv  (Double y, const FieldVector& x) { a[0] = y * x[0]; a[1] = y * x[1]; a[2] = y * x[2]; }
    void mmv (Double y, const FieldVector& x) { a[0] -= y * x[0]; a[1] -= y * x[1]; a[2] -= y * x[2]; }
    FieldVector& operator=  (const std::vector<Double>& y) { for (unsigned i = 0; i < 3 && i < y.size(); i++) a[i] = y[i]; return *this; }
    FieldVector& operator+= (const FieldVector& y) { a[0] += y[0]; a[1] += y[1]; a[2] += y[2]; return *this; }
    FieldVector& operator/= (Double y) { a[0] /= y; a[1] /= y; a[2] /= y; return *this; }
  };
  inline Double operator* (const FieldVector& x, const FieldVector& y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }

  /* One macro-particle, stdinclude.h:130-144: exactly the 11 doubles of the C ABI's particle rows          */
  struct Charge
  {
    Double      q;
    FieldVector rnp, rnm, gb;
    Double      e;
    Charge () : q(0.0), rnp(0.0), rnm(0.0), gb(0.0), e(0.0) {}
  };
  typedef std::list<Charge> ChargeVector;

  Double halton (unsigned int i, unsigned int j);                      /* stdinclude.cpp:45-73                 */
  Double pmod   (const Double& a, const Double& b);                    /* stdinclude.cpp:88-93                 */
  void   createDirectory (std::string filename, unsigned int rank);    /* stdinclude.cpp:26-36                 */
  void   printmessage (std::string filename, unsigned int linenumber, std::string message);
  template <class T> std::string stringify (T v);

  struct Mesh
  {
    Double              lengthScale_, timeScale_;
    std::vector<Double> meshLength_, meshResolution_, meshCenter_;
    Double              totalTime_, totalDist_, timeStep_;
    int                 truncationOrder_;
    bool                spaceCharge_, optimizePosition_;
    SolverType          solver_;
    Double              timeShift_, gamma_;
    void initialize ();                                                /* classes.cpp:38-46                    */
    void show ();
    Mesh () : lengthScale_(1.0), timeScale_(1.0), totalTime_(0.0), totalDist_(0.0), timeStep_(0.0), truncationOrder_(2) { initialize(); }
  };

  struct BunchInitialize                                               /* database.h:22-75                     */
  {
    std::string               bunchType_, distribution_, generator_;
    unsigned int              numberOfParticles_;
    Double                    cloudCharge_, initialGamma_, initialBeta_;
    FieldVector               initialDirection_, betaVector_;
    std::vector<FieldVector>  position_;
    std::vector<unsigned int> numbers_;
    FieldVector               latticeConstants_, sigmaPosition_, sigmaGammaBeta_;
    Double                    tranTrun_, longTrun_;
    std::string               fileName_;
    Double                    bF_, bFP_;
    bool                      shotNoise_;
    Double                    lambda_;
    BunchInitialize ();
  };

  struct Bunch                                                         /* classes.h:63-140                     */
  {
    std::vector<BunchInitialize> bunchInit_;
    Double      timeStep_;
    bool        sampling_;       std::string directory_, basename_;                           Double rhythm_;
    bool        bunchVTK_;       std::string bunchVTKDirectory_, bunchVTKBasename_;           Double bunchVTKRhythm_;
    bool        bunchProfile_;   std::string bunchProfileDirectory_, bunchProfileBasename_;   std::vector<Double> bunchProfileTime_; Double bunchProfileRhythm_;
    Double      zu_, beta_;
    Bunch ();
    /* generators, classes.cpp:84-417; rank / size select every size-th particle like the reference           */
    void initializeManual    (BunchInitialize bunchInit, ChargeVector& chargeVector, Double (zp)[2], int rank, int size, int ia);
    void initializeEllipsoid (BunchInitialize bunchInit, ChargeVector& chargeVector, int rank, int size, int ia);
    void initialize3DCrystal (BunchInitialize bunchInit, ChargeVector& chargeVector, Double (zp)[2], int rank, int size, int ia);
    void initializeFile      (BunchInitialize bunchInit, ChargeVector& chargeVector, Double (zp)[2], int rank, int size, int ia);
    void show ();
  };

  struct Signal                                                        /* classes.h:143-180, classes.cpp:475-575 */
  {
    SignalType          signalType_;
    Double              t0_, s_, f0_, cep_;
    unsigned int        nR_;
    std::vector<Double> sigmaInvG_;
    Signal ();
    void initialize (std::string type, Double l0, Double s, Double l, Double cep, unsigned int nR, std::vector<Double> sigmaInvG);
  };

  /* what Seed, Undulator (optical) and ExtField have in common                                            */
  struct Beam
  {
    SeedType            seedType_;
    Double              c0_;
    FieldVector         position_, direction_, polarization_;
    Double              amplitude_, a0_;
    std::vector<Double> radius_;
    Signal              signal_;
    std::vector<int>    order_;
    Double              l_;
    std::vector<Double> zR_;
    Beam ();
  protected:
    /* common body of Seed::initialize (classes.cpp:663-737), Undulator::initialize (:1007-1100) and
     * ExtField::initialize (:1163-1253); fourthRoot reproduces ExtField's normalisation by sqrt(norm()) (:1197,1210) */
    void initializeBeam (std::string type, std::vector<Double> position, std::vector<Double> direction, std::vector<Double> polarization,
			 Double a0, std::vector<Double> radius, std::vector<int> order, Signal signal, bool fourthRoot);
  };

  struct Seed : public Beam
  {
    Double beta_, gamma_, dt_;
    /* field sampling / visualisation / profile requests (parsed; their writers are out of scope here)        */
    bool sampling_; SamplingType samplingType_; std::vector<FieldType> samplingField_; std::string samplingDirectory_, samplingBasename_;
    Double samplingRhythm_; std::vector<FieldVector> samplingPosition_; FieldVector samplingLineBegin_, samplingLineEnd_; unsigned int samplingRes_;
    struct vtk { bool sample_; std::string directory_, basename_; SamplingType type_; PlaneType plane_; std::vector<FieldType> field_; Double rhythm_; FieldVector position_;
                 vtk () : sample_(false), type_(ALLDOMAIN), plane_(ZNORMAL), rhythm_(0.0), position_(0.0) {} };
    std::vector<vtk> vtk_;
    bool profile_; std::vector<FieldType> profileField_; std::string profileDirectory_, profileBasename_; std::vector<Double> profileTime_; Double profileRhythm_;
    Seed ();
    void initialize (std::string type, std::vector<Double> position, std::vector<Double> direction, std::vector<Double> polarization,
		     Double a0, std::vector<Double> radius, std::vector<int> order, Signal signal);
    SamplingType samplingType (std::string s);
    SamplingType vtkType (std::string s);
    PlaneType    planeType (std::string s);
    FieldType    fieldType (std::string s);
  };

  struct Undulator : public Beam
  {
    Double        k_, lu_, rb_;
    unsigned int  length_;
    Double        dist_, theta_;
    UndulatorType type_;
    Undulator ();
    void initialize (std::string type, std::vector<Double> position, std::vector<Double> direction, std::vector<Double> polarization,
		     Double a0, std::vector<Double> radius, Double wavelength, std::vector<int> order, Signal signal);
  };

  struct ExtField : public Beam
  {
    ExtFieldType type_;
    ExtField ();
    void initialize (std::string type, std::vector<Double> position, std::vector<Double> direction, std::vector<Double> polarization,
		     Double a0, std::vector<Double> radius, Double wavelength, std::vector<int> order, Signal signal);
  };

  struct FreeElectronLaser                                             /* classes.h:430-578                    */
  {
    struct RadiationSampling
    {
      std::vector<Double> z_;
      bool                sampling_;
      std::string         directory_, basename_;
      Double              lineBegin_, lineEnd_;
      unsigned int        res_;
      SamplingType        samplingType_;
      std::vector<Double> lambda_;
      Double              lambdaMin_, lambdaMax_;
      unsigned int        lambdaRes_;
      void samplingType (std::string s);
      RadiationSampling ();
    };
    struct RadiationVisualization
    {
      Double z_; bool sampling_; std::string directory_, basename_; Double rhythm_, lambda_;
      RadiationVisualization ();
    };
    struct ScreenProfile
    {
      bool sampling_; std::string directory_, basename_; std::vector<Double> pos_; Double rhythm_;
      ScreenProfile ();
    };
    RadiationSampling      radiationPower_, radiationEnergy_;
    RadiationVisualization vtkPower_, vtkEnergy_;
    ScreenProfile          screenProfile_;
  };
}

#include <sstream>
namespace MITHRA
{
  template <class T> std::string stringify (T v) { std::ostringstream o; o << v; return o.str(); }
}

#endif
