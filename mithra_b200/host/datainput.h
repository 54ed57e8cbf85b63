/* datainput.h -- ParseDarius: job-file groups -> parameter classes.
 *
 * Same class name, constructor and entry point as the reference (src/datainput.h:20-62); accepts the same six top-level
 * groups and the same keys (src/datainput.cpp:18-751, documented in doc/MITHRA_UI/MITHRA_UI.tex:49-1142) and reacts to
 * an unknown group / key like the reference: a line on stdout and exit(1).                                    */
#ifndef MITHRA_B200_DATAINPUT_H_
#define MITHRA_B200_DATAINPUT_H_

#include <functional>
#include <list>
#include <map>
#include <string>
#include <vector>

#include "classes.h"

namespace MITHRA
{
  class ParseDarius
  {
  public:
    ParseDarius (std::list<std::string>& jobFile, Mesh& mesh, Bunch& bunch, Seed& seed, std::vector<Undulator>& undulator,
		 std::vector<ExtField>& extField, std::vector<FreeElectronLaser>& FEL);
    void setJobParameters ();

  private:
    typedef std::list<std::string>::iterator                               Iter;
    typedef std::map<std::string, std::function<void (const std::string&)>> Keys;

    /* walk "{ key=value ... }" after a (sub-)group name, dispatching every line on its key                       */
    void block (Iter& iter, const char* what, const Keys& keys, const char* group);
    /* walk "{ sub-group { ... } ... }", dispatching on the sub-group names                                      */
    void group (Iter& iter, const char* what, const std::map<std::string, std::function<void (Iter&)>>& subs, bool strict, const char* unknownIn);

    void readMesh      (Iter& iter);
    void readBunch     (Iter& iter);
    void readField     (Iter& iter);
    void readUndulator (Iter& iter);
    void readExtField  (Iter& iter);
    void readFEL       (Iter& iter);

    std::list<std::string>&         jobFile_;
    Mesh&                           mesh_;
    Bunch&                          bunch_;
    Seed&                           seed_;
    std::vector<Undulator>&         undulator_;
    std::vector<ExtField>&          extField_;
    std::vector<FreeElectronLaser>& FEL_;
  };
}

#endif
