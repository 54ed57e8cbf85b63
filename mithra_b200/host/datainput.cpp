/* datainput.cpp -- see datainput.h.  Key names, value types and the defaults of omitted keys follow
 * src/datainput.cpp:49-751; the dispatch is table-driven instead of the reference's if-else chains.              */
#include "datainput.h"

#include <cstdlib>
#include <iostream>

namespace MITHRA
{
  ParseDarius::ParseDarius (std::list<std::string>& jobFile, Mesh& mesh, Bunch& bunch, Seed& seed, std::vector<Undulator>& undulator,
			    std::vector<ExtField>& extField, std::vector<FreeElectronLaser>& FEL)
    : jobFile_(jobFile), mesh_(mesh), bunch_(bunch), seed_(seed), undulator_(undulator), extField_(extField), FEL_(FEL)
  {}

  void ParseDarius::setJobParameters ()
  {
    Iter iter = jobFile_.begin();
    while (iter != jobFile_.end())
      {
	if      (*iter == "MESH")           readMesh(iter);
	else if (*iter == "BUNCH")          readBunch(iter);
	else if (*iter == "FIELD")          readField(iter);
	else if (*iter == "UNDULATOR")      readUndulator(iter);
	else if (*iter == "EXTERNAL-FIELD") readExtField(iter);
	else if (*iter == "FEL-OUTPUT")     readFEL(iter);
	else { std::cout << (*iter) << " is not a defined group." << std::endl; exit(1); }
	++iter;
      }
  }

  /* on entry iter is at the (sub-)group name, on exit at its closing brace                                      */
  void ParseDarius::block (Iter& iter, const char* what, const Keys& keys, const char* groupName)
  {
    ++iter;
    if (iter == jobFile_.end() || *iter != "{") { std::cout << "The " << what << " directory is empty" << std::endl; exit(1); }
    /* the reference reads a block with do { ... } while (*iter != "}") (datainput.cpp:153-209 and every other block): the first
     * line is taken as a key before the brace is looked at, so an EMPTY block stops with "} is not defined in ... group."      */
    ++iter;
    do
      {
	if (iter == jobFile_.end()) break;
	const Keys::const_iterator k = keys.find(parameterName(*iter));
	if (k == keys.end()) { std::cout << parameterName(*iter) << " is not defined in " << groupName << " group." << std::endl; exit(1); }
	k->second(*iter);
	++iter;
      }
    while (iter != jobFile_.end() && *iter != "}");
    if (iter == jobFile_.end()) { std::cout << "The " << what << " directory is not closed" << std::endl; exit(1); }
  }

  void ParseDarius::group (Iter& iter, const char* what, const std::map<std::string, std::function<void (Iter&)>>& subs, bool strict, const char* unknownIn)
  {
    ++iter;
    if (iter == jobFile_.end() || *iter != "{") { std::cout << "The " << what << " directory is empty" << std::endl; exit(1); }
    bool first = true;
    for (++iter; iter != jobFile_.end(); ++iter)
      {
	/* a strict group is a do { ... } while (*iter != "}") in the reference too (BUNCH, datainput.cpp:142-283): empty, it
	 * stops at the brace with the message of its last branch; the other groups are left at the brace (the reference
	 * runs past it there)                                                                                          */
	if (*iter == "}" && !(strict && first)) break;
	first = false;
	const auto s = subs.find(*iter);
	if (s != subs.end()) { s->second(iter); continue; }
	if (strict) { std::cout << parameterName(*iter) << " is not defined in the " << unknownIn << " group." << std::endl; exit(1); }
	/* the reference steps over names it does not know in these groups (datainput.cpp:421,556,626,748)           */
      }
    if (iter == jobFile_.end()) { std::cout << "The " << what << " directory is not closed" << std::endl; exit(1); }
  }

  static Double lengthUnit (const std::string& line)
  {
    static const std::map<std::string, Double> u = { {"METER", 1.0}, {"DECIMETER", 1.0e-1}, {"CENTIMETER", 1.0e-2}, {"MILLIMETER", 1.0e-3},
						     {"MICROMETER", 1.0e-6}, {"NANOMETER", 1.0e-9}, {"ANGSTROM", 1.0e-10} };
    const auto f = u.find(stringValue(line));
    return f != u.end() ? f->second : doubleValue(line);
  }

  static Double timeUnit (const std::string& line)
  {
    static const std::map<std::string, Double> u = { {"SECOND", 1.0}, {"MILLISECOND", 1.0e-3}, {"MICROSECOND", 1.0e-6}, {"NANOSECOND", 1.0e-9},
						     {"PICOSECOND", 1.0e-12}, {"FEMTOSECOND", 1.0e-15}, {"ATTOSECOND", 1.0e-18} };
    const auto f = u.find(stringValue(line));
    return f != u.end() ? f->second : doubleValue(line);
  }

  /* MESH, datainput.cpp:49-134 */
  void ParseDarius::readMesh (Iter& iter)
  {
    Mesh& m = mesh_;
    const Keys keys = {
      { "length-scale",            [&] (const std::string& l) { m.lengthScale_ = lengthUnit(l); } },
      { "time-scale",              [&] (const std::string& l) { m.timeScale_ = timeUnit(l); } },
      { "mesh-lengths",            [&] (const std::string& l) { m.meshLength_ = vectorDoubleValue(l); } },
      { "mesh-resolution",         [&] (const std::string& l) { m.meshResolution_ = vectorDoubleValue(l); } },
      { "mesh-center",             [&] (const std::string& l) { m.meshCenter_ = vectorDoubleValue(l); } },
      { "total-time",              [&] (const std::string& l) { m.totalTime_ = doubleValue(l); } },
      { "total-distance",          [&] (const std::string& l) { m.totalDist_ = doubleValue(l); } },
      { "bunch-time-step",         [&] (const std::string& l) { bunch_.timeStep_ = doubleValue(l); } },
      { "mesh-truncation-order",   [&] (const std::string& l) {
	  m.truncationOrder_ = intValue(l);
	  if (m.truncationOrder_ != 1 && m.truncationOrder_ != 2)
	    { printmessage(__FILE__, __LINE__, "Mesh truncation order can not be different from one or two."); exit(1); } } },
      { "space-charge",            [&] (const std::string& l) { m.spaceCharge_ = boolValue(l); } },
      { "optimize-bunch-position", [&] (const std::string& l) { m.optimizePosition_ = boolValue(l); } },
      { "initial-time-back-shift", [&] (const std::string& l) {
	  m.timeShift_ = doubleValue(l);
	  if (m.timeShift_ < 0.0) { printmessage(__FILE__, __LINE__, "The shift back in time should always be positive."); exit(1); } } },
      { "solver",                  [&] (const std::string& l) {
	  const std::string s = stringValue(l);
	  if      (s == "FD")   m.solver_ = FD;
	  else if (s == "NSFD") m.solver_ = NSFD;
	  else { printmessage(__FILE__, __LINE__, "The solver type is not among the accepted solvers."); exit(1); } } },
      { "lorentz-factor",          [&] (const std::string& l) { m.gamma_ = doubleValue(l); } },
    };
    block(iter, "solver", keys, "solver");
  }

  /* BUNCH, datainput.cpp:137-272 */
  void ParseDarius::readBunch (Iter& iter)
  {
    const std::map<std::string, std::function<void (Iter&)>> subs = {
      { "bunch-initialization", [&] (Iter& it) {
	  BunchInitialize b;
	  const Keys keys = {
	    { "type",                    [&] (const std::string& l) { b.bunchType_ = stringValue(l); } },
	    { "distribution",            [&] (const std::string& l) { b.distribution_ = stringValue(l); } },
	    { "generator",               [&] (const std::string& l) {
		b.generator_ = stringValue(l);
		if (b.generator_ != "halton" && b.generator_ != "random")
		  { printmessage(__FILE__, __LINE__, "The inserted generator is not accepted !!!"); exit(1); } } },
	    { "charge",                  [&] (const std::string& l) { b.cloudCharge_ = doubleValue(l); } },
	    { "number-of-particles",     [&] (const std::string& l) { b.numberOfParticles_ = intValue(l); } },
	    { "gamma",                   [&] (const std::string& l) { b.initialGamma_ = doubleValue(l); } },
	    { "direction",               [&] (const std::string& l) { b.initialDirection_ = vectorDoubleValue(l); } },
	    { "position",                [&] (const std::string& l) { FieldVector p (0.0); p = vectorDoubleValue(l); b.position_.push_back(p); } },
	    { "numbers",                 [&] (const std::string& l) { b.numbers_ = vectorIntValue(l); } },
	    { "lattice-constants",       [&] (const std::string& l) { b.latticeConstants_ = vectorDoubleValue(l); } },
	    { "sigma-position",          [&] (const std::string& l) { b.sigmaPosition_ = vectorDoubleValue(l); } },
	    { "sigma-momentum",          [&] (const std::string& l) { b.sigmaGammaBeta_ = vectorDoubleValue(l); } },
	    { "transverse-truncation",   [&] (const std::string& l) { b.tranTrun_ = doubleValue(l); } },
	    { "longitudinal-truncation", [&] (const std::string& l) { b.longTrun_ = doubleValue(l); } },
	    { "file-name",               [&] (const std::string& l) { b.fileName_ = stringValue(l); } },
	    { "bunching-factor",         [&] (const std::string& l) { b.bF_ = doubleValue(l); } },
	    { "bunching-factor-phase",   [&] (const std::string& l) { b.bFP_ = doubleValue(l); } },
	    { "shot-noise",              [&] (const std::string& l) { b.shotNoise_ = boolValue(l); } },
	  };
	  block(it, "bunch-initialization", keys, "bunch-initialization");
	  bunch_.bunchInit_.push_back(b); } },
      { "bunch-sampling", [&] (Iter& it) {
	  const Keys keys = {
	    { "sample",    [&] (const std::string& l) { bunch_.sampling_ = boolValue(l); } },
	    { "directory", [&] (const std::string& l) { bunch_.directory_ = stringValue(l); } },
	    { "base-name", [&] (const std::string& l) { bunch_.basename_ = stringValue(l); } },
	    { "rhythm",    [&] (const std::string& l) { bunch_.rhythm_ = doubleValue(l); } },
	  };
	  block(it, "bunch-sampling", keys, "the bunch-sampling"); } },
      { "bunch-visualization", [&] (Iter& it) {
	  const Keys keys = {
	    { "sample",    [&] (const std::string& l) { bunch_.bunchVTK_ = boolValue(l); } },
	    { "directory", [&] (const std::string& l) { bunch_.bunchVTKDirectory_ = stringValue(l); } },
	    { "base-name", [&] (const std::string& l) { bunch_.bunchVTKBasename_ = stringValue(l); } },
	    { "rhythm",    [&] (const std::string& l) { bunch_.bunchVTKRhythm_ = doubleValue(l); } },
	  };
	  block(it, "bunch-visualization", keys, "the bunch-visualization"); } },
      { "bunch-profile", [&] (Iter& it) {
	  const Keys keys = {
	    { "sample",    [&] (const std::string& l) { bunch_.bunchProfile_ = boolValue(l); } },
	    { "directory", [&] (const std::string& l) { bunch_.bunchProfileDirectory_ = stringValue(l); } },
	    { "base-name", [&] (const std::string& l) { bunch_.bunchProfileBasename_ = stringValue(l); } },
	    { "time",      [&] (const std::string& l) { bunch_.bunchProfileTime_.push_back(doubleValue(l)); } },
	    { "rhythm",    [&] (const std::string& l) { bunch_.bunchProfileRhythm_ = doubleValue(l); } },
	  };
	  block(it, "bunch-profile", keys, "the bunch-profile"); } },
    };
    group(iter, "bunch", subs, true, "bunch-profile");
  }

  /* the keys shared by field-initialization, optical-undulator and electromagnetic-wave                          */
  struct BeamKeys
  {
    std::string         type, signalType;
    std::vector<Double> position, direction, polarization, sigmaInvG, radius;
    Double              a0, offset, pulseLength, wavelength, cep;
    unsigned int        nR;
    std::vector<int>    order;
    BeamKeys () : position(3, 0.0), direction(3, 0.0), polarization(3, 0.0), sigmaInvG(2, 0.0), radius(2, 0.0),
		  a0(0.0), offset(0.0), pulseLength(0.0), wavelength(0.0), cep(0.0), nR(2), order(2, 0) {}

    void add (std::map<std::string, std::function<void (const std::string&)>>& keys, const char* typeKey)
    {
      keys[typeKey]                  = [this] (const std::string& l) { type = stringValue(l); };
      keys["position"]               = [this] (const std::string& l) { position = vectorDoubleValue(l); };
      keys["direction"]              = [this] (const std::string& l) { direction = vectorDoubleValue(l); };
      keys["polarization"]           = [this] (const std::string& l) { polarization = vectorDoubleValue(l); };
      keys["strength-parameter"]     = [this] (const std::string& l) { a0 = doubleValue(l); };
      keys["radius-parallel"]        = [this] (const std::string& l) { radius[0] = doubleValue(l); };
      keys["radius-perpendicular"]   = [this] (const std::string& l) { radius[1] = doubleValue(l); };
      keys["order-parallel"]         = [this] (const std::string& l) { order[0] = doubleValue(l); };
      keys["order-perpendicular"]    = [this] (const std::string& l) { order[1] = doubleValue(l); };
      keys["signal-type"]            = [this] (const std::string& l) { signalType = stringValue(l); };
      keys["offset"]                 = [this] (const std::string& l) { offset = doubleValue(l); };
      keys["pulse-length"]           = [this] (const std::string& l) { pulseLength = doubleValue(l); };
      keys["wavelength"]             = [this] (const std::string& l) { wavelength = doubleValue(l); };
      keys["rising-cycles"]          = [this] (const std::string& l) { nR = intValue(l); };
      keys["CEP"]                    = [this] (const std::string& l) { cep = doubleValue(l); };
      keys["sigma-inverse-gaussian"] = [this] (const std::string& l) { sigmaInvG = vectorDoubleValue(l); };
    }

    Signal signal () const
    {
      static const char* known[5] = { "neumann", "gaussian", "secant-hyperbolic", "flat-top", "inverse-gaussian" };
      bool ok = false;
      for (int i = 0; i < 5; i++) ok = ok || signalType == known[i];
      if (!ok) { std::cout << signalType << " is an unknown signal type." << std::endl; exit(1); }
      Signal s; s.initialize(signalType, offset, pulseLength, wavelength, cep, nR, sigmaInvG);
      return s;
    }
  };

  /* FIELD, datainput.cpp:275-424 */
  void ParseDarius::readField (Iter& iter)
  {
    const std::map<std::string, std::function<void (Iter&)>> subs = {
      { "field-initialization", [&] (Iter& it) {
	  BeamKeys b; Keys keys; b.add(keys, "type");
	  block(it, "seed-initialization", keys, "seed-initialization");
	  seed_.initialize(b.type, b.position, b.direction, b.polarization, b.a0, b.radius, b.order, b.signal()); } },
      { "field-sampling", [&] (Iter& it) {
	  const Keys keys = {
	    { "sample",           [&] (const std::string& l) { seed_.sampling_ = boolValue(l); } },
	    { "type",             [&] (const std::string& l) { seed_.samplingType_ = seed_.samplingType(stringValue(l)); } },
	    { "field",            [&] (const std::string& l) { seed_.samplingField_.push_back(seed_.fieldType(stringValue(l))); } },
	    { "directory",        [&] (const std::string& l) { seed_.samplingDirectory_ = stringValue(l); } },
	    { "base-name",        [&] (const std::string& l) { seed_.samplingBasename_ = stringValue(l); } },
	    { "rhythm",           [&] (const std::string& l) { seed_.samplingRhythm_ = doubleValue(l); } },
	    { "position",         [&] (const std::string& l) { FieldVector p; p = vectorDoubleValue(l); seed_.samplingPosition_.push_back(p); } },
	    { "line-begin",       [&] (const std::string& l) { seed_.samplingLineBegin_ = vectorDoubleValue(l); } },
	    { "line-end",         [&] (const std::string& l) { seed_.samplingLineEnd_ = vectorDoubleValue(l); } },
	    { "number-of-points", [&] (const std::string& l) { seed_.samplingRes_ = intValue(l); } },
	  };
	  block(it, "seed-sampling", keys, "seed-sampling"); } },
      { "field-visualization", [&] (Iter& it) {
	  seed_.vtk_.resize(seed_.vtk_.size() + 1);
	  Seed::vtk& v = seed_.vtk_.back();
	  const Keys keys = {
	    { "sample",    [&] (const std::string& l) { v.sample_ = boolValue(l); } },
	    { "directory", [&] (const std::string& l) { v.directory_ = stringValue(l); } },
	    { "type",      [&] (const std::string& l) { v.type_ = seed_.vtkType(stringValue(l)); } },
	    { "plane",     [&] (const std::string& l) { v.plane_ = seed_.planeType(stringValue(l)); } },
	    { "base-name", [&] (const std::string& l) { v.basename_ = stringValue(l); } },
	    { "field",     [&] (const std::string& l) { v.field_.push_back(seed_.fieldType(stringValue(l))); } },
	    { "rhythm",    [&] (const std::string& l) { v.rhythm_ = doubleValue(l); } },
	    { "position",  [&] (const std::string& l) { v.position_ = vectorDoubleValue(l); } },
	  };
	  block(it, "seed-visualization", keys, "seed-visualization"); } },
      { "field-profile", [&] (Iter& it) {
	  const Keys keys = {
	    { "sample",    [&] (const std::string& l) { seed_.profile_ = boolValue(l); } },
	    { "directory", [&] (const std::string& l) { seed_.profileDirectory_ = stringValue(l); } },
	    { "base-name", [&] (const std::string& l) { seed_.profileBasename_ = stringValue(l); } },
	    { "time",      [&] (const std::string& l) { seed_.profileTime_.push_back(doubleValue(l)); } },
	    { "field",     [&] (const std::string& l) { seed_.profileField_.push_back(seed_.fieldType(stringValue(l))); } },
	    { "rhythm",    [&] (const std::string& l) { seed_.profileRhythm_ = doubleValue(l); } },
	  };
	  block(it, "seed-profile", keys, "seed-profile"); } },
    };
    group(iter, "seed", subs, false, "");
  }

  /* UNDULATOR, datainput.cpp:427-559 */
  void ParseDarius::readUndulator (Iter& iter)
  {
    const std::map<std::string, std::function<void (Iter&)>> subs = {
      { "static-undulator", [&] (Iter& it) {
	  Undulator u; u.type_ = STATIC;
	  const Keys keys = {
	    { "undulator-parameter",    [&] (const std::string& l) { u.k_ = doubleValue(l); } },
	    { "period",                 [&] (const std::string& l) { u.lu_ = doubleValue(l); } },
	    { "polarization-angle",     [&] (const std::string& l) { u.theta_ = PI / 180.0 * doubleValue(l); } },
	    { "length",                 [&] (const std::string& l) { u.length_ = intValue(l); } },
	    { "distance-to-bunch-head", [&] (const std::string& l) { u.dist_ = doubleValue(l); } },
	    { "offset",                 [&] (const std::string& l) { u.rb_ = doubleValue(l); } },
	  };
	  block(it, "static undulator", keys, "the static-undulator");
	  undulator_.push_back(u); } },
      { "static-undulator-array", [&] (Iter& it) {
	  Undulator u;
	  Double k = u.k_, lu = u.lu_, theta = u.theta_, g = 0.0, t = 0.0, d = 0.0;
	  unsigned int len = u.length_, N = 1;
	  const Keys keys = {
	    { "undulator-parameter",    [&] (const std::string& l) { k = doubleValue(l); } },
	    { "period",                 [&] (const std::string& l) { lu = doubleValue(l); } },
	    { "polarization-angle",     [&] (const std::string& l) { theta = PI / 180.0 * doubleValue(l); } },
	    { "length",                 [&] (const std::string& l) { len = intValue(l); } },
	    { "gap",                    [&] (const std::string& l) { g = doubleValue(l); } },
	    { "number",                 [&] (const std::string& l) { N = intValue(l); } },
	    { "tapering-parameter",     [&] (const std::string& l) { t = doubleValue(l); } },
	    { "distance-to-bunch-head", [&] (const std::string& l) { d = doubleValue(l); } },
	  };
	  block(it, "static undulator", keys, "the static-undulator-array");
	  /* N modules, K tapered linearly, separated by the gap (datainput.cpp:487-501)                            */
	  for (unsigned int i = 0; i < N; i++)
	    {
	      u.type_ = STATIC; u.k_ = k + i * t; u.lu_ = lu; u.theta_ = theta; u.length_ = len;
	      u.rb_ = i * ( len * lu + g ); u.dist_ = d;
	      undulator_.push_back(u);
	    } } },
      { "optical-undulator", [&] (Iter& it) {
	  Undulator u; u.type_ = OPTICAL;
	  BeamKeys b; Keys keys; b.add(keys, "beam-type");
	  keys["distance-to-bunch-head"] = [&] (const std::string& l) { u.dist_ = doubleValue(l); };
	  block(it, "optical undulator", keys, "the optical-undulator");
	  u.initialize(b.type, b.position, b.direction, b.polarization, b.a0, b.radius, b.wavelength, b.order, b.signal());
	  undulator_.push_back(u); } },
    };
    group(iter, "undulator", subs, false, "");
  }

  /* EXTERNAL-FIELD, datainput.cpp:562-629 */
  void ParseDarius::readExtField (Iter& iter)
  {
    /* one ExtField object serves every wave of the group (datainput.cpp:568): what a wave does not set -- order_ outside
     * the super-gaussian beams -- is what the previous wave left there                                                  */
    ExtField e;
    const std::map<std::string, std::function<void (Iter&)>> subs = {
      { "electromagnetic-wave", [&] (Iter& it) {
	  e.type_ = EMWAVE;
	  BeamKeys b; Keys keys; b.add(keys, "beam-type");
	  Signal d;                                     /* defaults of omitted keys, datainput.cpp:585-589               */
	  b.a0 = e.a0_; b.offset = d.t0_; b.pulseLength = d.s_; b.wavelength = 1 / d.f0_; b.cep = d.cep_;
	  block(it, "electromagnetic-field", keys, "the electromagnetic external field");
	  e.initialize(b.type, b.position, b.direction, b.polarization, b.a0, b.radius, b.wavelength, b.order, b.signal());
	  extField_.push_back(e); } },
    };
    group(iter, "EXTERNAL-FIELD", subs, false, "");
  }

  /* FEL-OUTPUT, datainput.cpp:632-751: every sub-group is its own FreeElectronLaser entry                         */
  void ParseDarius::readFEL (Iter& iter)
  {
    auto radiation = [&] (Iter& it, bool energy) {
      FreeElectronLaser F;
      FreeElectronLaser::RadiationSampling& r = F.radiationPower_;   /* radiation-energy writes here too, as shipped (:706-717) */
      const Keys keys = {
	{ energy ? "distance-from-bunch" : "plane-position",                   [&] (const std::string& l) { r.z_.push_back(doubleValue(l)); } },
	{ "sample",                                                            [&] (const std::string& l) { r.sampling_ = boolValue(l); } },
	{ "directory",                                                         [&] (const std::string& l) { r.directory_ = stringValue(l); } },
	{ "base-name",                                                         [&] (const std::string& l) { r.basename_ = stringValue(l); } },
	{ "line-begin",                                                        [&] (const std::string& l) { r.lineBegin_ = doubleValue(l); } },
	{ "line-end",                                                          [&] (const std::string& l) { r.lineEnd_ = doubleValue(l); } },
	{ energy ? "resolution" : "number-of-points",                          [&] (const std::string& l) { r.res_ = energy ? (unsigned int) doubleValue(l) : intValue(l); } },
	{ energy ? "normalized-wavelength" : "normalized-frequency",           [&] (const std::string& l) { r.lambda_.push_back(doubleValue(l)); } },
	{ energy ? "minimum-normalized-wavelength" : "minimum-normalized-frequency", [&] (const std::string& l) { r.lambdaMin_ = doubleValue(l); } },
	{ energy ? "maximum-normalized-wavelength" : "maximum-normalized-frequency", [&] (const std::string& l) { r.lambdaMax_ = doubleValue(l); } },
	{ energy ? "normalized-wavelength-resolution" : "number-of-frequency-points", [&] (const std::string& l) { r.lambdaRes_ = energy ? (unsigned int) doubleValue(l) : intValue(l); } },
	{ "type",                                                              [&] (const std::string& l) { r.samplingType(stringValue(l)); } },
      };
      block(it, energy ? "radiation-energy" : "radiation-power", keys, "radiation-power");
      FEL_.push_back(F); };

    const std::map<std::string, std::function<void (Iter&)>> subs = {
      { "radiation-power",  [&] (Iter& it) { radiation(it, false); } },
      { "radiation-energy", [&] (Iter& it) { radiation(it, true); } },
      { "power-visualization", [&] (Iter& it) {
	  FreeElectronLaser F;
	  FreeElectronLaser::RadiationVisualization& v = F.vtkPower_;
	  const Keys keys = {
	    { "sample",               [&] (const std::string& l) { v.sampling_ = boolValue(l); } },
	    { "directory",            [&] (const std::string& l) { v.directory_ = stringValue(l); } },
	    { "base-name",            [&] (const std::string& l) { v.basename_ = stringValue(l); } },
	    { "plane-position",       [&] (const std::string& l) { v.z_ = doubleValue(l); } },
	    { "rhythm",               [&] (const std::string& l) { v.rhythm_ = doubleValue(l); } },
	    { "normalized-frequency", [&] (const std::string& l) { v.lambda_ = doubleValue(l); } },
	  };
	  block(it, "power-visualization", keys, "power-visualization");
	  FEL_.push_back(F); } },
      { "bunch-profile-lab-frame", [&] (Iter& it) {
	  FreeElectronLaser F;
	  FreeElectronLaser::ScreenProfile& s = F.screenProfile_;
	  const Keys keys = {
	    { "sample",    [&] (const std::string& l) { s.sampling_ = boolValue(l); } },
	    { "directory", [&] (const std::string& l) { s.directory_ = stringValue(l); } },
	    { "base-name", [&] (const std::string& l) { s.basename_ = stringValue(l); } },
	    { "position",  [&] (const std::string& l) { s.pos_.push_back(doubleValue(l)); } },
	    { "rhythm",    [&] (const std::string& l) { s.rhythm_ = doubleValue(l); } },
	  };
	  block(it, "bunch-profile-lab-frame", keys, "the bunch-profile-lab-frame");
	  FEL_.push_back(F); } },
    };
    group(iter, "FEL-OUTPUT", subs, false, "");
  }
}
