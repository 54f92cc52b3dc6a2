// Host-side mirror of the reference's interface for the `--make-grm` path, on top of the C ABI
// (include/dgrm.h).  Same class names, member names and argument meaning as the reference so that the rest
// of DISSECT (reml, multireml, gwas, pca) would use the result unchanged:
//   Options      options.h:  the flag subset of SURVEY.md §2.1 (options.cpp:268-1183)
//   Communicator communicator.h: process-grid description only (no MPI here; one engine per rank)
//   Misc         misc.h: message stream + error() (misc.cpp:151-170)
//   Matrix       matrix.h:79-963: cyclicDistribution local piece, numroc/local2global, symmetric/uplo
//   Genotype     genotype.h:109-399: .bim/.fam/.bed metadata, --extract/--keep, per-SNP `SNP` records
//   Kernel       kernel.h:49-379: kernel, N, ids, normalize/denormalize, sanitizeKernel, writeGRM
// Differences are deliberate and listed in INTEGRATION.md: the genotype/missings fp64 matrices are never
// materialised on the host (they live as 2-bit rows + tensor-core operands on the GPU).
#ifndef DGRM_HOST_H
#define DGRM_HOST_H

#include <iostream>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dgrm.h"

namespace dgrm_host {

struct DissectError : std::runtime_error {
  explicit DissectError(const std::string& m) : std::runtime_error(m) {}
};

enum GroupBy { ungrouped, byPosition, byGene, byGroup, byOrderedFixedSize, byAll };  // genotype.h:42-50

struct Options {
  bool makeGRM;
  std::string genotypeFile;            // --bfile
  std::string genotypeListFile;        // --bfile-list F: one genotype file prefix per line (options.cpp:414-417)
  std::string bgenGenotypeFile;        // --bgen F: F.bgen (+ F.sample when the file holds no sample identifiers), options.cpp:418-423
  std::string outFile;                 // --out            (default "dissect", options.cpp:60)
  std::string fileSNPsToKeep;          // --extract
  std::string fileIndividualsToKeep;   // --keep
  bool flatGRMNormalization;           // --grm-flat-norm  (options.cpp:533-537)
  double minimumNumberOverlapingSNPs;  // --min-overlap-snps (default 10, options.cpp:80)
  double maximumProportionOfElimitaedIndividuals;  // --min-prop-grm-inds-kept (default 0.9, options.cpp:81)
  int defaultBlockSize;                // --default-block-size (16..4096, default 256, options.cpp:244)
  // regional analysis (options.cpp:975-1012): one GRM per SNP group, files <out>.<group>.grm.*
  bool regionalAnalysis;
  GroupBy regionBy;
  std::string regionsFile;             // --groups F: lines `snp group [group ...]`
  int regionSize, regionOverlapping;   // --region-size / --region-overlap (kb on the command line, bp here)
  int minSNPsInRegions;                // --min-snps-region (default 1)
  bool writeLOCO;                      // --gpu-loco: also write <out>.loco.<group> = all SNPs minus the group
                                       // (the (S, N) algebra of singlereml.cpp:293-298, done on the device)
  // engine knobs (no counterpart in the reference)
  int device, operand, countMethod, chunkSNPs, ctaGroup;
  int engines;                         // --gpu-engines G: the lower triangle is split over G engines (GPU device +
                                       // r mod the number of devices for engine r), all driven by this one host rank
  bool lazyHostMatrices;               // --gpu-lazy-host: kernel / N are only brought to the host when asked for
                                       // (ensureHost); --make-grm itself needs files only, so a 50k or 100k GRM never
                                       // has to fit the INT_MAX local-matrix limit of a 1 x 1 grid (matrix.cpp:1773-1779)
  int GRMJoinMethod;                   // --grm-join-method 0 | 1 (options.cpp:76, 497-500; auxiliar.cpp:615-690)
  bool allowBGENLayout1;               // --bgen-l1 (options.cpp:1117-1121)
  bool bgenHostDecode;                 // --gpu-bgen-host-decode: unpack the probabilities on the host and push dosage rows
                                       // (default: the inflated blocks go to the device, dgrm_push_bgen_blocks_host)
  Options();
  void parseOptions(int argc, char** argv);
};

struct Communicator {
  int nProcRows, nProcCols, myRow, myCol;  // BLACS grid coordinates (communicator.cpp:66-103)
  int nDefaultBlockRows, nDefaultBlockCols;
  bool mpiRoot;
  Communicator() : nProcRows(1), nProcCols(1), myRow(0), myCol(0), nDefaultBlockRows(256), nDefaultBlockCols(256), mpiRoot(true) {}
};

struct Misc {
  std::ostream* message;
  Misc() : message(&std::cout) {}
  // misc.cpp:151-170: prints and aborts the job; here it throws so callers (tests, the CLI) can decide
  [[noreturn]] void error(const std::string& e, int) const;
};

extern Communicator* communicator;
extern Options options;
extern Misc misc;

enum DistributionType { global, cyclicDistribution, diagonalDistribution };

class Matrix {
 public:
  DistributionType distribution;
  double* m;  // local column-major block, new double[nRows*nCols] (matrix.cpp:1646); owned, delete[]
  int nGlobRows, nGlobCols, nRows, nCols, nBlockRows, nBlockCols;
  bool symmetric;
  char uplo;  // 'L', 'U' or 'B'
  explicit Matrix(DistributionType dist = cyclicDistribution);
  ~Matrix();
  void initParameters(int ngr, int ngc, int nbr, int nbc);  // matrix.cpp:1700-1782 (numroc, allocation)
  int local2globalRow(int r) const;                          // matrix.cpp:435-475
  int local2globalCol(int c) const;
  void fillWithConstant(double v);                           // matrix.cpp:125-137
  void elementWiseDivision(Matrix* m1, double scale = 1.);   // matrix.cpp:4074-4099
  void elementWiseMultiplication(Matrix* m1, double scale = 1.);
 private:
  Matrix(const Matrix&);
  Matrix& operator=(const Matrix&);
};

struct SNP {  // genotype.h:56-72
  std::string chr, name, allele1, allele2;
  int position;
  int nNonMissing;
  int frequencies[4];
  double p1, p2, standardDev;
};

struct Individual {  // genotype.h:85-95
  std::string familyID, individualID, paternalID, maternalID, sex;
  int phenotype;
};

struct GenotypeAttributes { enum Type { calls, probabilities }; };  // genotype.h: how the genotypes of a file are coded

class Genotype {
 public:
  int nSNPs, nIndividuals;
  std::vector<SNP> SNPs;
  std::vector<std::string> SNPIds;
  std::vector<Individual> individuals;
  std::vector<std::string> individualIds;
  bool normalized;
  // .bed selection exactly as readBIMFile/readFAMFile leave it (genotype.cpp:392-546)
  std::string bedFile;
  int nBEDSNPs, nBEDIndividuals;
  std::vector<int> SNPsBEDIdxsToKeep;            // empty = all (genotype.cpp:452-456)
  std::vector<int> individualBEDIdxsMaskToKeep;  // one flag per .fam line (genotype.cpp:750)
  explicit Genotype(const std::string& fname);   // genotype.cpp:110 -> load :276 (metadata; rows stream later)
  // genotype.cpp:110-130 with GenotypeAttributes::probabilities -> readBGENFile (genotypebgen.cpp:561-721): header,
  // sample identifiers (or F.sample), --keep / --extract, and the identifying data of every kept variant; the
  // probabilities are read when a Kernel is computed (Genotype::streamBGEN).
  Genotype(const std::string& fname, GenotypeAttributes::Type type);
  GenotypeAttributes::Type type;
  std::string bgenFile;
  int bgenLayout, bgenCompression;
  long long bgenFirstVariant;                    // file position of the first variant data block (offset + 4)
  std::vector<unsigned char> bgenKeepVariant;    // one flag per variant of the file (--extract)
  // Calls sink(dosage, missing, rows) for blocks of kept variants: dosage / missing are [rows][nIndividuals] arrays of
  // the kept samples, dosage = P(AB) + 2 P(BB), missing != 0 where the sample has no probabilities
  // (ProbabilitiesSetter::set_value, genotypebgen.cpp:183-236).
  typedef void (*BGENSink)(void* ctx, const double* dosage, const unsigned char* missing, long long rows);
  void streamBGEN(BGENSink sink, void* ctx) const;
  typedef void (*BGENBlockSink)(void* ctx, const unsigned char* data, const long long* offsets, long long rows);
  void streamBGENBlocks(BGENBlockSink sink, void* ctx) const;   // the inflated probability blocks, for the device to unpack
  void removeSNPsWithZeroVariance();             // genotypebgen.cpp:739-779 (on the SNP records; the device drops the rows)
  Genotype();                                    // genotype.cpp:82-91, filled by genotypeOfSNPsGroup
  ~Genotype();
  // SNP groups (genotype.cpp:1293-1470) and the genotype of one group (genotype.cpp:1596-1621)
  GroupBy groupedBy;
  std::map<std::string, std::set<std::string> > groupedSNPs;
  std::map<std::string, int> SNPIdsIdx;
  void groupSNPs(GroupBy groupBy);
  std::vector<std::string> getGroup(const std::string& group, bool writeGroup = true);
  void genotypeOfSNPsGroup(const std::string& group, Genotype* newGenotype, bool writeGroup = true);
  // Resident cohort: the packed .bed is uploaded once (dgrm_store_bed_file) and every Kernel built from this
  // Genotype or from one of its groups re-reads it from HBM.  storeRows[k] = stored row of SNP k.
  void loadResident();
  dgrm_engine* engine() const { return eng_; }
  std::vector<int> storeRows;
 private:
  dgrm_engine* eng_;
  bool ownsEngine_;
  void groupSNPsByPosition();
  void groupSNPsByGene();
  void readBIMFile(const std::string& f, const std::set<std::string>& keepSNPs);
  void readFAMFile(const std::string& f, const std::set<std::string>& keepIndividualIds);
};

enum KernelType { kernelGRM };

class Kernel {
 public:
  KernelType type;
  std::string name;
  Matrix* kernel;  // A (normalized) or S (not), cyclicDistribution, symmetric, uplo 'L'
  Matrix* N;       // pair counts (or m under --grm-flat-norm), same structure
  bool normalized, diagonalized, asymmetric;
  int nIndividuals;
  std::vector<Individual> individuals;
  std::vector<std::string> individualIds;
  std::map<std::string, int> individualIdsIdx;
  std::vector<std::string> randomVarNames;  // = genotype->SNPIds (kernel.cpp:85)
  Kernel(Genotype* genotype, bool normalizeGRM = true);  // kernel.cpp:61-125, computed on the GPU
  Kernel(const std::vector<Genotype*>& genotypes, bool normalizeGRM);  // one GRM over the SNPs of several files (--bfile-list)
  explicit Kernel(const std::string& f);                 // kernel.cpp:127-... -> readGRM :1023-1196 (host only)
  Kernel();                                              // empty, for addKernels results
  ~Kernel();
  // kernel.cpp:1705-1859 (addGRMs): this = s1*grm1 + s2*grm2 on the unnormalised (S, N) pair, s = +1 or -1.
  // Regional / LOCO algebra: global - regional is exact in N and exact to rounding in S (singlereml.cpp:293-298).
  void addKernels(double scalingFactor1, Kernel* grm1, double scalingFactor2, Kernel* grm2);
  void normalize();      // kernel.cpp:382-414
  void denormalize();    // kernel.cpp:416-448
  bool sanitizeKernel(); // kernel.cpp:1993-2037
  void writeGRM(const std::string& f);  // kernel.cpp:893-1009 (.grm.ids / .grm.snps / .grm.dat)
  // Device-side addKernels(1., saved, -1., this) / (1., saved, 1., this): a Kernel that reads the engine's
  // (S, N) view `saved -/+ current` (dgrm_set_view).  `saved` must be the Kernel whose GRM was kept with
  // dgrm_save on the same engine; the view ends when the returned Kernel is deleted.
  Kernel* combinedWithSaved(Kernel* saved, double scalingFactorThis);
  dgrm_engine* engine() { return eng_; }
  void ensureHost();  // materialise kernel / N on the host (no-op when they are there already)
 private:
  dgrm_engine* eng_;                  // engine 0 (owner of the per-SNP statistics; every engine computes the same)
  std::vector<dgrm_engine*> engs_;    // all engines of the partition (engs_[0] == eng_)
  bool ownsEngine_, isView_;
  std::vector<long long> subset_;  // individuals (positions) still in the kernel after sanitizeKernel
  void fetch();
  void compute(const std::vector<Genotype*>& genotypes, bool normalizeGRM);
  std::vector<std::string> minimumIndividualsRemovePairs(const std::vector<int>& i1, const std::vector<int>& i2);
  void filterIndividuals(const std::vector<std::string>& keepIds);
  Kernel(const Kernel&);
  Kernel& operator=(const Kernel&);
};

// auxiliar.cpp:581-700 for the genotype sources of --make-grm (--bfile, or --bfile-list with --grm-join-method)
Kernel* loadGRMUsingOptions();

}  // namespace dgrm_host
#endif
