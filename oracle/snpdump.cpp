// Oracle-side dump driver (TEST INFRASTRUCTURE): links against the reference's own, unmodified objects
// (everything but main.o) and prints what `struct SNP` (genotype.h:56-72) holds on the root rank after
// `new Genotype(file)` — the per-SNP integer counts and fp64 statistics that `--make-grm` never writes
// to disk.  With DGRM_DUMP_Z=<path> it also runs the reference's normalizeGenotypes()
// (genotype.cpp:888) and writes the standardised m x n matrix and the missings matrix as raw doubles
// (row-major by SNP).  Arguments are the reference's own (`--make-grm --bfile X [--extract F] [--keep F]
// --out Y`) and are parsed by the reference's Options class.
#include <cstdio>
#include <cstdlib>
#include <omp.h>
#include <iostream>

#include "communicator.h"
#include "genotype.h"
#include "global.h"
#include "matrix.h"
#include "misc.h"
#include "options.h"

Communicator* communicator;
Options options;
Misc misc;

int main(int argc, char** argv) {
  communicator = new Communicator(argc, argv);
  options = Options(argc, argv);
  misc.changeOutputs(std::cout, options.outFile + ".log");
  communicator->nDefaultBlockRows = options.defaultBlockSize;
  communicator->nDefaultBlockCols = options.defaultBlockSize;

  Genotype* g = new Genotype(options.genotypeFile);
  const std::string path = options.outFile + ".snpdump";
  FILE* f = fopen(path.c_str(), "w");
  if (!f) { fprintf(stderr, "snpdump: cannot write %s\n", path.c_str()); return 2; }
  fprintf(f, "# nSNPs %d nIndividuals %d\n", g->nSNPs, g->nIndividuals);
  fprintf(f, "# name f0 f1 f2 f3 nNonMissing p1 p2 standardDev\n");
  for (int i = 0; i < g->nSNPs; ++i) {
    const SNP& s = g->SNPs[i];
    fprintf(f, "%s %d %d %d %d %d %a %a %a\n", s.name.c_str(), s.frequencies[0], s.frequencies[1], s.frequencies[2],
            s.frequencies[3], s.nNonMissing, s.p1, s.p2, s.standardDev);
  }
  fclose(f);

  const char* zpath = getenv("DGRM_DUMP_Z");
  if (zpath) {
    g->normalizeGenotypes();
    FILE* z = fopen(zpath, "wb");
    if (!z) { fprintf(stderr, "snpdump: cannot write %s\n", zpath); return 2; }
    const int m = g->genotypes->nRows, n = g->genotypes->nCols;  // 1 x 1 grid: local == global
    for (int pass = 0; pass < 2; ++pass) {
      const double* src = pass == 0 ? g->genotypes->m : g->missings->m;
      for (int r = 0; r < m; ++r)
        for (int c = 0; c < n; ++c) fwrite(&src[(size_t)c * m + r], sizeof(double), 1, z);
    }
    fclose(z);
  }
  delete g;
  delete communicator;
  return 0;
}
