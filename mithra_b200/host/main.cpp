/* main.cpp -- mithra_b200 <job-file> [--gpus N] [--steps K] [--dump-params PREFIX]
 *
 * Same sequence as the reference's main() (src/mithra.cpp:32-102): read and clean the job file, build the parameter
 * objects on the stack, parse, pick FdTdSC or FdTd on `space-charge`, solve(), print the wall time.  MPI is gone: one
 * process drives every GPU it is given (--gpus, default 1) as z-slabs.
 *   --steps K             stop after K field steps (tests)
 *   --dump-params PREFIX  run Solver::initialize() only and write PREFIX.meta.bin (no GPU needed): what the parity test
 *                         compares bit for bit with the reference's own initialize()
 */
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sys/time.h>

#include "datainput.h"
#include "solver.h"

using namespace MITHRA;

int main (int argc, char* argv[])
{
  if (argc < 2) { std::cout << "usage: mithra_b200 <job-file> [--gpus N] [--steps K] [--dump-params PREFIX]" << std::endl; return 1; }
  int gpus = 1; long steps = -1; std::string dump;
  for (int a = 2; a < argc; a++)
    {
      if      (!strcmp(argv[a], "--gpus")        && a + 1 < argc) gpus = atoi(argv[++a]);
      else if (!strcmp(argv[a], "--steps")       && a + 1 < argc) steps = atol(argv[++a]);
      else if (!strcmp(argv[a], "--dump-params") && a + 1 < argc) dump = argv[++a];
      else { std::cout << argv[a] << " is not a known option." << std::endl; return 1; }
    }


  timeval t0, t1;
  gettimeofday(&t0, NULL);

  std::list<std::string> jobFile = read_file(argv[1]);
  cleanJobFile(jobFile);

  Mesh mesh; mesh.initialize();
  Bunch bunch;
  Seed seed;
  std::vector<Undulator> undulator;
  std::vector<ExtField> extField;
  std::vector<FreeElectronLaser> FEL;

  ParseDarius parser (jobFile, mesh, bunch, seed, undulator, extField, FEL);
  parser.setJobParameters();
  mesh.show(); bunch.show();

  /* more than four slabs in this process: two streams each, some kernels spin on flags (csrc/exchange.cuh) -- ask for
   * more hardware queues than the default 8 before the context exists (it makes context creation slower: only then)  */
  if ( gpus > 4 ) setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);

  Solver* solver = mesh.spaceCharge_ ? (Solver*) new FdTdSC (mesh, bunch, seed, undulator, extField, FEL)
				     : (Solver*) new FdTd   (mesh, bunch, seed, undulator, extField, FEL);
  solver->setNumberOfGpus(gpus);
  solver->setMaxSteps(steps);

  if (!dump.empty())
    {
      solver->initialize();
      solver->dumpParams(dump);
      delete solver;
      return 0;
    }

  solver->solve();

  gettimeofday(&t1, NULL);
  const double dT = ( t1.tv_usec - t0.tv_usec ) / 1.0e6 + ( t1.tv_sec - t0.tv_sec );
  printmessage(__FILE__, __LINE__, "Total simulation time = " + stringify(dT) + " s, " + stringify(solver->nTime_) + " field steps.");
  delete solver;
  return 0;
}
