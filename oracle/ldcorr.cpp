// Oracle-side LD driver (TEST INFRASTRUCTURE): the SNP x SNP correlation step of the reference's GWAS
// (GWAS::getLessSignificantCorrelatedSNPs, gwas.cpp:1156-1172; same in igwas.cpp:979-995) executed by the
// reference's own, unmodified objects:
//     Genotype(file) -> normalizeGenotypes()                         (gwas.cpp:149)
//     correlations  = genotypes * genotypes^T                        (Matrix::multiply -> pdsyrk_('L','N'), matrix.cpp:2677-2683)
//     normalization = missings  * missings^T
//     correlations->elementWiseDivision(normalization)               (matrix.cpp:4091-4098)
//     correlations->getGlobalIndexOutsideRange(-t, t, rows, cols)    (matrix.cpp:4265-4330)
// Arguments are the reference's own (`--bfile X [--extract F] [--keep F] --out Y`, parsed by its Options class);
// the threshold (options.correlatedSNPsThreshold in the reference, which its parser only accepts together with
// a recursive GWAS analysis, options.cpp:814-817) comes from the environment variable LDCORR_THRESHOLD.  Writes <out>.ldcorr: int64 m, then the m x m matrix (column-major doubles; only
// the lower triangle is defined, as pdsyrk_ leaves it), then int64 count and count (row, col) int32 pairs.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>
#include <stdint.h>

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
  const double threshold = getenv("LDCORR_THRESHOLD") ? atof(getenv("LDCORR_THRESHOLD")) : 0.5;

  Genotype* genotypes = new Genotype(options.genotypeFile);
  genotypes->normalizeGenotypes();
  // gwas.cpp:1163-1172, verbatim call sequence
  Matrix* correlations = new Matrix(MATRIX_DEFAULT_DISTRIBUTION);
  Matrix* normalization = new Matrix(MATRIX_DEFAULT_DISTRIBUTION);
  correlations->multiply(genotypes->genotypes, 'N', genotypes->genotypes, 'T');
  normalization->multiply(genotypes->missings, 'N', genotypes->missings, 'T');
  correlations->elementWiseDivision(normalization);
  delete normalization;
  std::vector<int> idxSNPs1;
  std::vector<int> idxSNPs2;
  correlations->getGlobalIndexOutsideRange(-threshold, threshold, idxSNPs1, idxSNPs2);

  const std::string path = options.outFile + ".ldcorr";
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { fprintf(stderr, "ldcorr: cannot write %s\n", path.c_str()); return 2; }
  const int64_t m = correlations->nGlobRows;   // 1 x 1 grid: local == global, column-major
  fwrite(&m, sizeof m, 1, f);
  fwrite(correlations->m, sizeof(double), (size_t)m * m, f);
  const int64_t cnt = (int64_t)idxSNPs1.size();
  fwrite(&cnt, sizeof cnt, 1, f);
  for (int64_t i = 0; i < cnt; ++i) {
    const int32_t rc[2] = {idxSNPs1[i], idxSNPs2[i]};
    fwrite(rc, sizeof(int32_t), 2, f);
  }
  fclose(f);
  delete correlations;
  delete genotypes;
  delete communicator;
  return 0;
}
