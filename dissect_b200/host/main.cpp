// dgrm_make_grm — the reference's `dissect --make-grm` dispatch (main.cpp:117-120 -> Analysis::makeGRM,
// analysis.cpp:43-66) on top of the host mirror: load, Kernel ctor (GPU), sanitizeKernel, writeKernel.
#include <cstdio>

#include "host.h"

using namespace dgrm_host;

int main(int argc, char** argv) {
  try {
    options.parseOptions(argc, argv);
    communicator->nDefaultBlockRows = communicator->nDefaultBlockCols = options.defaultBlockSize;
    if (!options.makeGRM) misc.error("Error: only --make-grm is implemented by this adapter.", 0);
    if (options.genotypeFile.empty()) misc.error("Error: No genotype file(s) specified.", 0);
    Genotype* genotype = new Genotype(options.genotypeFile);
    Kernel* grm = new Kernel(genotype);
    delete genotype;
    const bool sanitized = grm->sanitizeKernel();
    if (sanitized) grm->writeGRM(options.outFile);
    else *misc.message << "Sorry, a problem was happened while computing the GRM. Please, check the logs." << std::endl;
    delete grm;
    return 0;
  } catch (const DissectError&) {
    return 1;  // misc.error -> MPI_Abort(-1) in the reference (misc.cpp:169)
  }
}
