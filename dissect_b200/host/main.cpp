// dgrm_make_grm — the reference's `dissect --make-grm` dispatch (main.cpp:117-120 -> Analysis::makeGRM,
// analysis.cpp:43-66) on top of the host mirror: load, Kernel ctor (GPU), sanitizeKernel, writeKernel.
#include <cstdio>

#include "host.h"

using namespace dgrm_host;

int main(int argc, char** argv) {
  try {
    options.lazyHostMatrices = true;   // --make-grm only writes files: kernel / N stay on the device (--gpu-eager-host undoes it)
    options.parseOptions(argc, argv);
    communicator->nDefaultBlockRows = communicator->nDefaultBlockCols = options.defaultBlockSize;
    if (!options.makeGRM) misc.error("Error: only --make-grm is implemented by this adapter.", 0);
    if (options.genotypeFile.empty() && options.genotypeListFile.empty() && options.bgenGenotypeFile.empty())
      misc.error("Error: No genotype file(s) specified.", 0);
    if (!options.regionalAnalysis) {
      Kernel* grm = loadGRMUsingOptions();   // analysis.cpp:45-47
      const bool sanitized = grm->sanitizeKernel();
      if (sanitized) grm->writeGRM(options.outFile);
      else *misc.message << "Sorry, a problem was happened while computing the GRM. Please, check the logs." << std::endl;
      delete grm;
    } else {
      // analysis.cpp:67-110: one GRM per SNP group -> <out>.<group>.grm.*.  The packed cohort is uploaded once.
      Genotype* genotype = new Genotype(options.genotypeFile);
      genotype->groupSNPs(options.regionBy);
      genotype->loadResident();
      const std::string baseOutFile = options.outFile;
      Kernel* global = NULL;
      if (options.writeLOCO) {  // singlereml.cpp:242: the global GRM, kept on the device for global - regional
        global = new Kernel(genotype);
        if (dgrm_save(genotype->engine(), DGRM_SAVE_COPY) != DGRM_OK) misc.error(dgrm_last_error(genotype->engine()), 0);
      }
      for (auto it = genotype->groupedSNPs.begin(); it != genotype->groupedSNPs.end(); ++it) {
        const std::string group = it->first;
        options.outFile = baseOutFile + "." + group;
        Genotype* regionalGenotype = new Genotype();
        genotype->genotypeOfSNPsGroup(group, regionalGenotype);
        Kernel* grm = new Kernel(regionalGenotype);
        if (global) {  // singlereml.cpp:298: globalGRM->addKernels(1., grm, -1., regionalGRM), before any filtering
          Kernel* loco = grm->combinedWithSaved(global, -1.);
          if (loco->sanitizeKernel()) loco->writeGRM(baseOutFile + ".loco." + group);
          else *misc.message << "WARNING: Sorry, a problem was happened while computing the LOCO GRM on region " << group << ". Please, check the logs." << std::endl;
          delete loco;
        }
        const bool sanitized = grm->sanitizeKernel();
        if (sanitized) grm->writeGRM(options.outFile);
        else *misc.message << "WARNING: Sorry, a problem was happened while computing the GRM on region " << group << ". Please, check the logs." << std::endl;
        delete grm;
        delete regionalGenotype;
      }
      options.outFile = baseOutFile;
      delete global;
      delete genotype;
    }
    return 0;
  } catch (const DissectError&) {
    return 1;  // misc.error -> MPI_Abort(-1) in the reference (misc.cpp:169)
  }
}
