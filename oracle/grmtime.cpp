// Oracle-side timing driver (TEST / BASELINE INFRASTRUCTURE): runs the reference's own hot path —
// `new Genotype(file)` (genotype.cpp:110 -> readBEDFile :548) and `new Kernel(genotype)` (kernel.cpp:61:
// normalizeGenotypes, two pdsyrk_, elementWiseDivision) — exactly as loadGRMUsingOptions does
// (auxiliar.cpp:604-605), from the reference's unmodified objects, and prints wall-clock phase times with
// clock_gettime (the reference's own stopwatch has 1 s resolution, misc.cpp:210-252).
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <omp.h>

#include "communicator.h"
#include "genotype.h"
#include "global.h"
#include "kernel.h"
#include "matrix.h"
#include "misc.h"
#include "options.h"

Communicator* communicator;
Options options;
Misc misc;

static double now() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char** argv) {
  communicator = new Communicator(argc, argv);
  options = Options(argc, argv);
  misc.changeOutputs(std::cout, options.outFile + ".log");
  communicator->nDefaultBlockRows = options.defaultBlockSize;
  communicator->nDefaultBlockCols = options.defaultBlockSize;
  const double t0 = now();
  Genotype* g = new Genotype(options.genotypeFile);
  const double t1 = now();
  Kernel* k = new Kernel(g);
  const double t2 = now();
  // a checksum so the work cannot be elided and the run can be sanity-checked
  double tr = 0.;
  const int n = k->kernel->nGlobRows;
  for (int i = 0; i < n; ++i) tr += k->kernel->m[(size_t)i * k->kernel->nRows + i];
  printf("GRMTIME n %d m %d threads %d load_s %.6f compute_s %.6f trace %.9g\n", g->nIndividuals, g->nSNPs,
         omp_get_max_threads(), t1 - t0, t2 - t1, tr);
  fflush(stdout);
  delete k;
  delete g;
  return 0;
}
