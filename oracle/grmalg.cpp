// Oracle-side algebra driver (TEST INFRASTRUCTURE): the reference's own Kernel::addKernels on GRM files the
// reference wrote, from the reference's unmodified objects:
//     grmalg add|sub <grm1> <grm2> --out <result>
// add:  result = grm1 + grm2                       (kernel.cpp:1705-1859 with factors 1., 1.)
// sub:  result = grm1 - grm2, the LOCO operation   (`globalGRM->addKernels(1., grm, -1., regionalGRM)`,
//       singlereml.cpp:298): both kernels are denormalised (S = A o N), S and N are combined with pdgeadd_, and the
//       result is divided again.
// The result is written with Kernel::writeKernel (kernel.cpp:881 -> writeGRM :893).
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "communicator.h"
#include "global.h"
#include "kernel.h"
#include "matrix.h"
#include "misc.h"
#include "options.h"

Communicator* communicator;
Options options;
Misc misc;

int main(int argc, char** argv) {
  if (argc < 6 || (strcmp(argv[1], "add") && strcmp(argv[1], "sub"))) {
    fprintf(stderr, "usage: grmalg add|sub <grm1> <grm2> --out <result>\n");
    return 2;
  }
  const bool sub = !strcmp(argv[1], "sub");
  const std::string f1 = argv[2], f2 = argv[3];
  // the remaining arguments (`--out <result>`) go through the reference's own option parser
  char* rest[3] = {argv[0], argv[4], argv[5]};
  communicator = new Communicator(3, rest);
  options = Options(3, rest);
  misc.changeOutputs(std::cout, options.outFile + ".log");
  communicator->nDefaultBlockRows = options.defaultBlockSize;
  communicator->nDefaultBlockCols = options.defaultBlockSize;
  Kernel* k1 = new Kernel(f1);
  Kernel* k2 = new Kernel(f2);
  Kernel* r = new Kernel();
  r->addKernels(1., k1, sub ? -1. : 1., k2);
  r->writeKernel(options.outFile);
  delete r;
  delete k2;
  delete k1;
  return 0;
}
