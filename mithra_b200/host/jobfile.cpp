/* jobfile.cpp -- see jobfile.h.  Behaviour follows src/readdata.cpp:10-159 (what is accepted, what a value means,
 * "print and exit(1)" on errors); the implementation is this project's own.                                  */
#include "jobfile.h"

#include <cstdlib>
#include <fstream>
#include <iostream>

namespace MITHRA
{
  std::list<std::string> read_file (char const* filename)
  {
    std::ifstream in(filename);
    if (!in.is_open()) { std::cout << "Unable to open file" << std::endl; exit(1); }
    std::list<std::string> lines;
    std::string line;
    while (std::getline(in, line)) lines.push_back(line);
    return lines;
  }

  void cleanJobFile (std::list<std::string>& jobFile)
  {
    std::list<std::string> kept;
    for (const std::string& raw : jobFile)
      {
	std::string s;
	for (char c : raw)
	  {
	    if (c == '#') break;
	    if (c != ' ' && c != '\t' && c != '\r') s.push_back(c);
	  }
	if (!s.empty()) kept.push_back(s);
      }
    jobFile.swap(kept);
  }

  static std::string afterEqual (const std::string& line)
  {
    const size_t p = line.find('=');
    return (p == std::string::npos) ? line : line.substr(p + 1);
  }

  std::string parameterName (std::string line) { return line.substr(0, line.find('=')); }

  std::string stringValue (std::string line)
  {
    std::string v = afterEqual(line);
    if (v.size() >= 2 && v[0] == '"') { v.erase(0, 1); const size_t q = v.find('"'); if (q != std::string::npos) v.erase(q, 1); }
    return v;
  }

  Double doubleValue (std::string line) { return std::atof(afterEqual(line).c_str()); }

  int intValue (std::string line) { return (int) std::atof(afterEqual(line).c_str()); }

  bool boolValue (std::string line)
  {
    const std::string v = afterEqual(line);
    if (v == "true")  return true;
    if (v == "false") return false;
    std::cout << "boolValue(std::string line) got unexpected input. Input should be \"true\" or \"false\" " << std::endl;
    exit(1);
  }

  /* "(a,b,c)" -> the strings between the brackets and commas                                              */
  static std::vector<std::string> vectorItems (const std::string& line)
  {
    std::string v = afterEqual(line);
    if (!v.empty()) v.erase(0, 1);                       /* "("                                            */
    if (!v.empty()) v.erase(v.size() - 1, 1);            /* ")"                                            */
    std::vector<std::string> items;
    size_t a = 0;
    while (true)
      {
	const size_t c = v.find(',', a);
	items.push_back(v.substr(a, c == std::string::npos ? std::string::npos : c - a));
	if (c == std::string::npos) break;
	a = c + 1;
      }
    return items;
  }

  std::vector<Double> vectorDoubleValue (std::string line)
  {
    std::vector<Double> out;
    for (const std::string& s : vectorItems(line)) out.push_back(std::atof(s.c_str()));
    return out;
  }

  std::vector<unsigned int> vectorIntValue (std::string line)
  {
    std::vector<unsigned int> out;
    for (const std::string& s : vectorItems(line)) out.push_back((unsigned int) std::atoi(s.c_str()));
    return out;
  }
}
