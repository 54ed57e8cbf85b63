/* jobfile.h -- reader of MITHRA job files (the grammar of the prj job files).
 *
 * Keeps the interface of the reference's readdata.h (src/readdata.h:18-47, implementation src/readdata.cpp:10-159):
 * the job file becomes a list of lines with comments (#...) and ALL blanks and tabs removed; a line is either a group
 * name, "{", "}" or "key=value"; values are strings, numbers (atof), booleans (true/false) or vectors "(a,b,c)".
 */
#ifndef MITHRA_B200_JOBFILE_H_
#define MITHRA_B200_JOBFILE_H_

#include <list>
#include <string>
#include <vector>

namespace MITHRA
{
  typedef double Double;

  std::list<std::string> read_file (char const* filename);
  void                   cleanJobFile (std::list<std::string>& jobFile);

  std::string               parameterName     (std::string line);
  std::string               stringValue       (std::string line);
  Double                    doubleValue       (std::string line);
  int                       intValue          (std::string line);
  bool                      boolValue         (std::string line);
  std::vector<Double>       vectorDoubleValue (std::string line);
  std::vector<unsigned int> vectorIntValue    (std::string line);
}

#endif
