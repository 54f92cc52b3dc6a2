// Implementation of the host-side mirror (see host.h).  All GRM arithmetic happens in libdgrm.so.
#include "host.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>

#include <zlib.h>

namespace dgrm_host {

Communicator* communicator = new Communicator();
Options options;
Misc misc;

void Misc::error(const std::string& e, int) const {
  std::cerr << "\n**********************************\n" << e << "\n**********************************\n" << std::endl;
  throw DissectError(e);
}

static void engine_check(dgrm_engine* e, int rc) {
  if (rc != DGRM_OK) misc.error(dgrm_last_error(e), 0);
}

static dgrm_engine* create_engine(int rank = 0, int world = 1) {
  dgrm_config cfg;
  dgrm_default_config(&cfg);
  const int ndev = std::max(1, dgrm_device_count());
  cfg.device = (options.device + rank) % ndev;
  cfg.rank = rank;
  cfg.world = world;
  cfg.operand = options.operand;
  cfg.count_method = options.countMethod;
  cfg.flat_norm = options.flatGRMNormalization ? 1 : 0;
  cfg.chunk_snps = options.chunkSNPs;
  cfg.cta_group = options.ctaGroup;
  dgrm_engine* e = NULL;
  if (dgrm_create(&cfg, &e) != DGRM_OK) misc.error(dgrm_last_error(NULL), 0);
  return e;
}

// ------------------------------------------------------------------------------------------ Options
Options::Options()
    : makeGRM(false), outFile("dissect"), flatGRMNormalization(false), minimumNumberOverlapingSNPs(10.),
      maximumProportionOfElimitaedIndividuals(0.9), defaultBlockSize(256), regionalAnalysis(false), regionBy(ungrouped),
      regionSize(-1), regionOverlapping(0), minSNPsInRegions(1), writeLOCO(false), device(0), operand(DGRM_OPERAND_FP16),
      countMethod(DGRM_COUNT_FP4), chunkSNPs(16384), ctaGroup(2), engines(1), lazyHostMatrices(false), GRMJoinMethod(0),
      allowBGENLayout1(false), bgenHostDecode(false) {}

void Options::parseOptions(int argc, char** argv) {
  auto need = [&](int i) -> std::string {
    if (i + 1 >= argc) misc.error(std::string("Error: The option ") + argv[i] + " needs an argument.", 0);
    return argv[i + 1];
  };
  bool seenRegion = false, seenGroups = false, seenJoin = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--make-grm") makeGRM = true;
    else if (a == "--bfile") genotypeFile = need(i++);
    else if (a == "--bfile-list") genotypeListFile = need(i++);                 // options.cpp:414-417
    else if (a == "--bgen") bgenGenotypeFile = need(i++);                       // options.cpp:418-423
    else if (a == "--bgen-l1") allowBGENLayout1 = true;                         // options.cpp:1117-1121
    else if (a == "--gpu-bgen-host-decode") bgenHostDecode = true;
    else if (a == "--grm-join-method") {                                        // options.cpp:497-500
      GRMJoinMethod = atoi(need(i++).c_str());
      seenJoin = true;
      if (GRMJoinMethod < 0 || GRMJoinMethod > 1) misc.error("Error: --grm-join-method is out of the range [0, 1].", 0);
    }
    else if (a == "--out") outFile = need(i++);
    else if (a == "--extract") fileSNPsToKeep = need(i++);
    else if (a == "--keep") fileIndividualsToKeep = need(i++);
    else if (a == "--grm-flat-norm") flatGRMNormalization = true;
    else if (a == "--min-overlap-snps") minimumNumberOverlapingSNPs = atof(need(i++).c_str());
    else if (a == "--min-prop-grm-inds-kept") maximumProportionOfElimitaedIndividuals = atof(need(i++).c_str());
    else if (a == "--default-block-size") {
      defaultBlockSize = atoi(need(i++).c_str());
      if (defaultBlockSize < 16 || defaultBlockSize > 4096)
        misc.error("Error: --default-block-size is out of the range [16, 4096].", 0);
    } else if (a == "--gpu-device") device = atoi(need(i++).c_str());
    else if (a == "--gpu-operand") {
      const std::string v = need(i++);
      operand = v == "bf16" ? DGRM_OPERAND_BF16 : v == "fp16x2" ? DGRM_OPERAND_FP16X2 : DGRM_OPERAND_FP16;
    } else if (a == "--gpu-count") {
      const std::string v = need(i++);
      countMethod = v == "popc" ? DGRM_COUNT_POPC : v == "fp8" ? DGRM_COUNT_MMA : DGRM_COUNT_FP4;
    }
    // options.cpp:975-1012
    else if (a == "--region-size") { regionSize = atoi(need(i++).c_str()) * 1000; regionBy = byPosition; regionalAnalysis = true; seenRegion = true; }
    else if (a == "--region-overlap") { regionOverlapping = atoi(need(i++).c_str()) * 1000; regionBy = byPosition; regionalAnalysis = true; seenRegion = true; }
    else if (a == "--groups") { regionsFile = need(i++); regionBy = byGene; regionalAnalysis = true; seenGroups = true; }
    else if (a == "--group-all") { regionBy = byAll; regionalAnalysis = true; }
    else if (a == "--min-snps-region") minSNPsInRegions = atoi(need(i++).c_str());
    else if (a == "--gpu-loco") writeLOCO = true;
    else if (a == "--gpu-chunk-snps") chunkSNPs = atoi(need(i++).c_str());
    else if (a == "--gpu-engines") {
      engines = atoi(need(i++).c_str());
      if (engines < 1 || engines > 64) misc.error("Error: --gpu-engines is out of the range [1, 64].", 0);
    } else if (a == "--gpu-lazy-host") lazyHostMatrices = true;
    else if (a == "--gpu-eager-host") lazyHostMatrices = false;
    else if (a == "--gpu-cta-group") ctaGroup = atoi(need(i++).c_str());
    else misc.error("Error: Invalid option: " + a, 0);
  }
  // options.cpp:1582-1593
  if (seenRegion && seenGroups)
    misc.error("Error: Only one type of regional analysis is allowed. The options --region-overlap and --region-size or --groups can not be used together.", 0);
  if (regionSize < 0 && regionOverlapping > 0)
    misc.error("Error: The option --region-overlap can only be used together with --region-size.", 0);
  if (!genotypeListFile.empty() && !genotypeFile.empty())   // options.cpp:1193-1196
    misc.error("Error: A genotype file and a genotype list of files can not be specified at the same time.", 0);
  if (seenJoin && genotypeListFile.empty())                   // options.cpp:1288-1291
    misc.error("Error: The option --grm-join-method can only be used together with --bfile-list.", 0);
  if (!bgenGenotypeFile.empty() && (!genotypeFile.empty() || !genotypeListFile.empty()))
    misc.error("Error: A genotype file and a genotype list of files can not be specified at the same time.", 0);
  if (regionalAnalysis && genotypeFile.empty())
    misc.error("Error: To perform a regional analysis, a genotype file must be specified.", 0);
  if (writeLOCO && !regionalAnalysis) misc.error("Error: --gpu-loco needs a regional option (e.g. --groups).", 0);
  if (engines > 1 && regionalAnalysis)
    misc.error("Error: --gpu-engines > 1 is implemented for a single GRM (no regional option) in this adapter.", 0);
}

// ------------------------------------------------------------------------------------------- Matrix
static int numroc(int n, int nb, int iproc, int nprocs) {
  const int nblocks = n / nb;
  int r = (nblocks / nprocs) * nb;
  const int extra = nblocks % nprocs;
  if (iproc < extra) r += nb;
  else if (iproc == extra) r += n % nb;
  return r;
}

Matrix::Matrix(DistributionType dist)
    : distribution(dist), m(NULL), nGlobRows(0), nGlobCols(0), nRows(0), nCols(0), nBlockRows(0), nBlockCols(0),
      symmetric(false), uplo('B') {}

Matrix::~Matrix() { delete[] m; }

void Matrix::initParameters(int ngr, int ngc, int nbr, int nbc) {
  delete[] m;
  m = NULL;
  nGlobRows = ngr; nGlobCols = ngc; nBlockRows = nbr; nBlockCols = nbc;
  nRows = numroc(ngr, nbr, communicator->myRow, communicator->nProcRows);
  nCols = numroc(ngc, nbc, communicator->myCol, communicator->nProcCols);
  // the reference refuses local matrices beyond INT_MAX elements (matrix.cpp:1773-1779)
  if ((double)nRows * (double)nCols > 2147483647.)
    misc.error("Error: The local matrix is too large for this number of processes.", 0);
  m = new (std::nothrow) double[(size_t)std::max(nRows, 1) * (size_t)std::max(nCols, 1)];
  if (!m) misc.error("Error: Sorry, unable to allocate enough memory.", 0);
}

int Matrix::local2globalRow(int r) const {
  return ((r / nBlockRows) * communicator->nProcRows + communicator->myRow) * nBlockRows + r % nBlockRows;
}
int Matrix::local2globalCol(int c) const {
  return ((c / nBlockCols) * communicator->nProcCols + communicator->myCol) * nBlockCols + c % nBlockCols;
}

void Matrix::fillWithConstant(double v) {
  for (size_t i = 0; i < (size_t)nRows * nCols; ++i) m[i] = v;
}

static void check_same_structure(const Matrix* a, const Matrix* b, const char* what) {
  if (a->nGlobRows != b->nGlobRows || a->nGlobCols != b->nGlobCols || a->nRows != b->nRows || a->nCols != b->nCols ||
      a->distribution != b->distribution || a->nBlockRows != b->nBlockRows || a->nBlockCols != b->nBlockCols)
    misc.error(std::string("Error: An internal error was happened. Element wise ") + what +
               " can not be performed on matrices with different structural properties.", 0);
  if (a->symmetric != b->symmetric || a->uplo != b->uplo)
    if ((a->symmetric && a->uplo != 'B') || (b->symmetric && b->uplo != 'B'))
      misc.error(std::string("Error: An internal error was happened. Element wise ") + what +
                 " can not be performed on matrices with different symmetry properties.", 0);
}

void Matrix::elementWiseDivision(Matrix* m1, double scale) {
  check_same_structure(this, m1, "division");
  // only the defined triangle is touched here (the reference sweeps the uninitialised half too)
  for (int c = 0; c < nCols; ++c)
    for (int r = 0; r < nRows; ++r) {
      if (symmetric && uplo == 'L' && local2globalCol(c) > local2globalRow(r)) continue;
      m[(size_t)c * nRows + r] *= scale / m1->m[(size_t)c * nRows + r];
    }
}

void Matrix::elementWiseMultiplication(Matrix* m1, double scale) {
  check_same_structure(this, m1, "multiplication");
  for (int c = 0; c < nCols; ++c)
    for (int r = 0; r < nRows; ++r) {
      if (symmetric && uplo == 'L' && local2globalCol(c) > local2globalRow(r)) continue;
      m[(size_t)c * nRows + r] *= scale * m1->m[(size_t)c * nRows + r];
    }
}

// ----------------------------------------------------------------------------------------- Genotype
static std::vector<std::vector<std::string> > read_table(const std::string& path, size_t min_cols) {
  std::ifstream f(path.c_str());
  if (!f) misc.error("Error: The file [ " + path + " ] does not exist or cannot be opened.", 0);
  std::vector<std::vector<std::string> > rows;
  std::string line;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::vector<std::string> t;
    std::string w;
    while (ss >> w) t.push_back(w);
    if (t.empty()) continue;
    if (t.size() < min_cols) misc.error("Error: The file [ " + path + " ] has a line with too few columns.", 0);
    rows.push_back(t);
  }
  return rows;
}

Genotype::Genotype()
    : nSNPs(0), nIndividuals(0), normalized(false), nBEDSNPs(0), nBEDIndividuals(0), type(GenotypeAttributes::calls),
      bgenLayout(0), bgenCompression(0), bgenFirstVariant(0), groupedBy(ungrouped), eng_(NULL), ownsEngine_(false) {}

Genotype::~Genotype() {
  if (eng_ && ownsEngine_) dgrm_destroy(eng_);
}

Genotype::Genotype(const std::string& fname)
    : nSNPs(0), nIndividuals(0), normalized(false), bedFile(fname + ".bed"), nBEDSNPs(0), nBEDIndividuals(0),
      type(GenotypeAttributes::calls), bgenLayout(0), bgenCompression(0), bgenFirstVariant(0), groupedBy(ungrouped), eng_(NULL),
      ownsEngine_(false) {
  std::set<std::string> keepSNPs, keepInds;
  if (!options.fileSNPsToKeep.empty()) {
    *misc.message << "Only SNPs specified in file [ " << options.fileSNPsToKeep << " ] will be kept." << std::endl;
    for (auto& r : read_table(options.fileSNPsToKeep, 1)) keepSNPs.insert(r[0]);
  }
  if (!options.fileIndividualsToKeep.empty()) {
    *misc.message << "Only individuals specified in file [ " << options.fileIndividualsToKeep << " ] will be kept." << std::endl;
    for (auto& r : read_table(options.fileIndividualsToKeep, 2)) keepInds.insert(r[0] + "@" + r[1]);
  }
  readBIMFile(fname, keepSNPs);
  readFAMFile(fname, keepInds);
  if (nSNPs == 0 || nIndividuals == 0)
    misc.error("Error: There are 0 SNPs and/or 0 individuals kept for the analysis. Aborting...", 0);
  for (int k = 0; k < nSNPs; ++k) SNPIdsIdx[SNPIds[k]] = k;
}


// ------------------------------------------------------------------------------------------- BGEN
// The reference reads BGEN through genfile::bgen (genotypebgen.cpp:284-559), which its tree does not vendor.  What
// follows restates the published v1.1 / v1.2 container for the subset the reference accepts (biallelic, diploid,
// unphased; ProbabilitiesSetter::initialise / set_min_max_ploidy / set_number_of_entries, genotypebgen.cpp:99-172).
namespace {

struct BgenReader {
  FILE* f;
  std::string name;
  long long size;
  explicit BgenReader(const std::string& path) : f(fopen(path.c_str(), "rb")), name(path), size(0) {
    if (!f) misc.error("Error: Unable to read the file [ " + path + " ].", 0);   // genotypebgen.cpp:320-322
    if (fseeko(f, 0, SEEK_END) == 0) size = (long long)ftello(f);
    fseeko(f, 0, SEEK_SET);
  }
  ~BgenReader() { if (f) fclose(f); }
  [[noreturn]] void truncated() { misc.error("Error: Unexpected end of the BGEN file [ " + name + " ].", 0); }
  // every length field of the container is checked against what is left of the file before it is believed
  void need(long long n) { if (n < 0 || (long long)ftello(f) + n > size) truncated(); }
  void bytes(void* dst, size_t n) {
    if (n && fread(dst, 1, n, f) != n) truncated();
  }
  uint32_t u32() { unsigned char b[4]; bytes(b, 4); return b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24); }
  uint16_t u16() { unsigned char b[2]; bytes(b, 2); return (uint16_t)(b[0] | (b[1] << 8)); }
  std::string str(size_t n) { need((long long)n); std::string s(n, '\0'); bytes(&s[0], n); return s; }
  void skip(long long n) { need(n); if (fseeko(f, n, SEEK_CUR) != 0) truncated(); }
};

struct BgenVariant { std::string chr, rsid; uint32_t position; std::vector<std::string> alleles; };

// identifying data of the next variant (genfile::bgen::read_snp_identifying_data)
BgenVariant read_variant_id(BgenReader& r, int layout, uint32_t nSamples) {
  BgenVariant v;
  if (layout == 1 && r.u32() != nSamples) misc.error("Error: Unexpected number of samples in a BGEN variant block.", 0);
  r.skip(r.u16());                 // variant identifier (the reference passes a dummy, genotypebgen.cpp:467-472)
  v.rsid = r.str(r.u16());
  v.chr = r.str(r.u16());
  v.position = r.u32();
  const int k = layout == 2 ? r.u16() : 2;
  for (int a = 0; a < k; ++a) v.alleles.push_back(r.str(r.u32()));
  if (k != 2)   // genotypebgen.cpp:107-110
    misc.error("Error: The bgen file contains non-biallelic genotypes. DISSECT currently can only work with biallelic genotypes.", 0);
  return v;
}

// length in bytes of the genotype data block that follows the identifying data
long long genotype_block_bytes(BgenReader& r, int layout, int compression, uint32_t nSamples) {
  if (layout == 1 && !compression) return 6ll * nSamples;
  return r.u32();
}

// the probability data of one variant, inflated and appended to `out` (genfile::bgen::read_genotype_data_block +
// uncompress_probability_data): layout 2 from the N field on, layout 1 the 6N bytes
void read_probability_block(BgenReader& r, int layout, int compression, uint32_t nSamples, long long len,
                            std::vector<unsigned char>& raw, std::vector<unsigned char>& out) {
  size_t dataLen;
  r.need(len);
  if (layout == 2 && compression) {
    if (len < 4) r.truncated();
    dataLen = r.u32();
    raw.resize((size_t)len - 4);
  } else if (layout == 1 && compression) {
    dataLen = 6u * (size_t)nSamples;
    raw.resize((size_t)len);
  } else {
    dataLen = (size_t)len;
    raw.resize(0);
  }
  const size_t at = out.size();
  out.resize(at + dataLen);
  if (compression) {
    r.bytes(raw.data(), raw.size());
    uLongf got = (uLongf)dataLen;
    if (uncompress(out.data() + at, &got, raw.data(), (uLong)raw.size()) != Z_OK || got != dataLen)
      misc.error("Error: A genotype block of the BGEN file [ " + r.name + " ] could not be decompressed.", 0);
  } else {
    r.bytes(out.data() + at, dataLen);
  }
}

}  // namespace

Genotype::Genotype(const std::string& fname, GenotypeAttributes::Type t)
    : nSNPs(0), nIndividuals(0), normalized(false), nBEDSNPs(0), nBEDIndividuals(0), type(t), bgenFile(fname + ".bgen"),
      bgenLayout(0), bgenCompression(0), bgenFirstVariant(0), groupedBy(ungrouped), eng_(NULL), ownsEngine_(false) {
  if (t != GenotypeAttributes::probabilities) misc.error("Error: An internal error was happened. Unexpected genotype type.", 0);
  std::set<std::string> keepSNPs, keepInds;
  if (!options.fileSNPsToKeep.empty()) {
    *misc.message << "Only SNPs specified in file [ " << options.fileSNPsToKeep << " ] will be kept." << std::endl;
    for (auto& r : read_table(options.fileSNPsToKeep, 1)) keepSNPs.insert(r[0]);
  }
  if (!options.fileIndividualsToKeep.empty()) {
    *misc.message << "Only individuals specified in file [ " << options.fileIndividualsToKeep << " ] will be kept." << std::endl;
    for (auto& r : read_table(options.fileIndividualsToKeep, 2)) keepInds.insert(r[0] + "@" + r[1]);
  }
  *misc.message << "Reading genotype data from file [ " << fname << " ] ..." << std::endl;
  BgenReader r(bgenFile);
  const uint32_t offset = r.u32();
  const uint32_t headerLength = r.u32();
  const uint32_t nVariants = r.u32();
  const uint32_t nSamples = r.u32();
  const std::string magic = r.str(4);
  if (magic != "bgen" && magic != std::string(4, '\0')) misc.error("Error: The file [ " + bgenFile + " ] is not a valid BGEN file.", 0);
  if (headerLength < 20) misc.error("Error: The file [ " + bgenFile + " ] is not a valid BGEN file.", 0);
  r.skip((long long)headerLength - 20);
  const uint32_t flags = r.u32();
  bgenCompression = (int)(flags & 3u);
  bgenLayout = (int)((flags >> 2) & 15u);
  // genotypebgen.cpp:330-351
  if (bgenLayout == 0) misc.error("Error: The BGEN file is in Layout 0. Currently DISSECT only supports Layout 1 and Layout 2.", 0);
  if (bgenLayout == 1 && !options.allowBGENLayout1)
    misc.error("Error: The BGEN file is using Layout 1. Currently DISSECT only supports Layout 1 for testing purposes. You can use the option --bgen-l1 to enable reading BGEN layout 1 files. But we recommend to update the file to Layout 2.", 0);
  if (bgenLayout > 2) misc.error("Error: Unrecognized BGEN layout type. Supported types are Layout 1 and Layout 2.", 0);
  if (bgenCompression == 2) misc.error("Error: zstd-compressed BGEN files are not supported by this adapter (no zstd library in the build).", 0);
  std::vector<std::pair<std::string, std::string> > ids;
  if (flags >> 31) {
    r.u32();
    const uint32_t n2 = r.u32();
    for (uint32_t i = 0; i < n2; ++i) { const std::string id = r.str(r.u16()); ids.push_back(std::make_pair(id, id)); }   // genotypebgen.cpp:353-359
  } else {
    // genotypebgen.cpp:360-365, 414-441: F.sample, header lines `ID_1 ID_2 ...` and `0 0 ...`
    *misc.message << "There are no individual identifiers in the [ " << bgenFile << " ] file. Trying to read them from [ " << fname << ".sample file ]." << std::endl;
    std::vector<std::vector<std::string> > tab = read_table(fname + ".sample", 2);
    if (tab.size() <= 2) misc.error("Error: The file [ " + fname + ".sample ] is empty.", 0);
    if (tab[0][0] != "ID_1" || tab[0][1] != "ID_2" || tab[1][0] != "0" || tab[1][1] != "0")
      misc.error("Error: Unexpected header in the file [" + fname + ".sample].", 0);
    for (size_t i = 2; i < tab.size(); ++i) ids.push_back(std::make_pair(tab[i][0], tab[i][1]));
  }
  if (ids.size() != nSamples) misc.error("Error: An error has happened whe reading bgen file. Unexpected number of individual Ids.", 0);
  nBEDIndividuals = (int)nSamples;
  for (size_t i = 0; i < ids.size(); ++i) {
    const std::string key = ids[i].first + "@" + ids[i].second;
    const bool keep = keepInds.empty() || keepInds.count(key);
    individualBEDIdxsMaskToKeep.push_back(keep ? 1 : 0);
    if (!keep) continue;
    Individual ind;   // genotypebgen.cpp:578-598
    ind.familyID = ids[i].first; ind.individualID = ids[i].second;
    ind.paternalID = ind.maternalID = ind.sex = "NA";
    ind.phenotype = -1;
    if (std::find(individualIds.begin(), individualIds.end(), key) != individualIds.end())
      misc.error("Error: The individual with id: " + ind.individualID + " appears more than one time in the genotypes BGEN file.", 0);
    individuals.push_back(ind);
    individualIds.push_back(key);
  }
  nIndividuals = (int)individuals.size();
  *misc.message << "Reading bgen file [ " << fname << " ] (" << (bgenLayout == 2 ? "v1.2 layout" : "v1.1 layout") << ", " << bgenLayout << ", "
                << (bgenCompression ? "compressed" : "uncompressed") << ") with " << nSamples << " individuals and " << nVariants
                << " variants." << std::endl;
  bgenFirstVariant = (long long)offset + 4;
  if (fseeko(r.f, bgenFirstVariant, SEEK_SET) != 0) misc.error("Error: Unexpected end of the BGEN file [ " + bgenFile + " ].", 0);
  nBEDSNPs = (int)nVariants;
  bgenKeepVariant.assign(nVariants, 0);
  for (uint32_t v = 0; v < nVariants; ++v) {
    const BgenVariant bv = read_variant_id(r, bgenLayout, nSamples);
    r.skip(genotype_block_bytes(r, bgenLayout, bgenCompression, nSamples));
    if (!keepSNPs.empty() && !keepSNPs.count(bv.rsid)) continue;   // genotypebgen.cpp:655-659
    bgenKeepVariant[v] = 1;
    SNP s;
    s.chr = bv.chr; s.name = bv.rsid; s.position = (int)bv.position; s.allele1 = bv.alleles[0]; s.allele2 = bv.alleles[1];
    s.nNonMissing = 0; s.p1 = s.p2 = s.standardDev = 0.;
    for (int c = 0; c < 4; ++c) s.frequencies[c] = 0;
    if (SNPIdsIdx.count(s.name))
      misc.error("Error: The SNP with name: " + s.name + " appears more than one time in the genotypes BGEN file.", 0);
    SNPIdsIdx[s.name] = (int)SNPs.size();
    SNPs.push_back(s);
    SNPIds.push_back(s.name);
  }
  nSNPs = (int)SNPs.size();
  if (nSNPs == 0 || nIndividuals == 0)
    misc.error("Error: There are 0 SNPs and/or 0 individuals kept for the analysis. Aborting...", 0);
}

void Genotype::streamBGEN(BGENSink sink, void* ctx) const {
  BgenReader r(bgenFile);
  if (fseeko(r.f, bgenFirstVariant, SEEK_SET) != 0) misc.error("Error: Unexpected end of the BGEN file [ " + bgenFile + " ].", 0);
  const uint32_t nSamples = (uint32_t)nBEDIndividuals;
  const size_t nk = (size_t)nIndividuals;
  std::vector<int> cols;
  for (uint32_t i = 0; i < nSamples; ++i) if (individualBEDIdxsMaskToKeep[i]) cols.push_back((int)i);
  const size_t blockRows = std::max<size_t>(1, std::min<size_t>(4096, (64u << 20) / (nk * sizeof(double))));
  std::vector<double> dosage(blockRows * nk);
  std::vector<unsigned char> missing(blockRows * nk);
  std::vector<unsigned char> raw, data;
  size_t rows = 0;
  for (size_t v = 0; v < bgenKeepVariant.size(); ++v) {
    read_variant_id(r, bgenLayout, nSamples);
    const long long len = genotype_block_bytes(r, bgenLayout, bgenCompression, nSamples);
    if (!bgenKeepVariant[v]) { r.skip(len); continue; }   // ignoreProbabilities
    // probability data of this variant, decompressed (8 bytes of slack for the bit reader)
    data.clear();
    read_probability_block(r, bgenLayout, bgenCompression, nSamples, len, raw, data);
    data.resize(data.size() + 8, 0);
    double* drow = &dosage[rows * nk];
    unsigned char* mrow = &missing[rows * nk];
    if (bgenLayout == 1) {
      for (size_t c = 0; c < nk; ++c) {
        const unsigned char* p = &data[(size_t)cols[c] * 6];
        const double aa = (double)(p[0] | (p[1] << 8)) / 32768.0, ab = (double)(p[2] | (p[3] << 8)) / 32768.0,
                     bb = (double)(p[4] | (p[5] << 8)) / 32768.0;
        mrow[c] = (aa + ab + bb) == 0. ? 1 : 0;     // genotypebgen.cpp:208-214
        drow[c] = ab + 2. * bb;                      // genotypebgen.cpp:199-205
      }
    } else {
      const unsigned char* p = data.data();
      const uint32_t n2 = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
      const int k2 = p[4] | (p[5] << 8), pmin = p[6], pmax = p[7];
      if (n2 != nSamples || k2 != 2) misc.error("Error: An internal error has happened. Unexpected small number of samples.", 0);
      if (pmin != 2 || pmax != 2)   // genotypebgen.cpp:120-123
        misc.error("Error: The bgen file contains non-diploid genotypes. DISSECT currently can only work with diploid genotypes.", 0);
      const unsigned char* ploidy = p + 8;
      const int phased = p[8 + nSamples], bits = p[9 + nSamples];
      if (phased)                   // genotypebgen.cpp:168-171
        misc.error("Error: Currently can only load unphased genotypes. If you are interested on loading phased genotypes, please, contact us.", 0);
      if (bits < 1 || bits > 32) misc.error("Error: Unexpected number of bits per probability in the BGEN file.", 0);
      const unsigned char* pr = p + 10 + nSamples;
      const uint64_t mask = bits == 32 ? 0xFFFFFFFFull : ((1ull << bits) - 1ull);
      const double scale = (double)mask;
      for (size_t c = 0; c < nk; ++c) {
        const size_t i = (size_t)cols[c];
        if ((ploidy[i] & 0x3F) != 2)
          misc.error("Error: The bgen file contains non-diploid genotypes. DISSECT currently can only work with diploid genotypes.", 0);
        uint64_t val[2];
        for (int h = 0; h < 2; ++h) {
          const uint64_t bit = (uint64_t)(2 * i + h) * (uint64_t)bits;
          uint64_t w;
          memcpy(&w, pr + (bit >> 3), 8);   // little-endian host
          val[h] = (w >> (bit & 7)) & mask;
        }
        const double aa = (double)val[0] / scale, ab = (double)val[1] / scale;
        const double bb = 1.0 - (aa + ab);           // the implied last probability of the format
        mrow[c] = (ploidy[i] & 0x80) ? 1 : 0;       // set_value(MissingValue), genotypebgen.cpp:229-236
        drow[c] = ab + 2. * bb;
      }
    }
    if (++rows == blockRows) { sink(ctx, dosage.data(), missing.data(), (long long)rows); rows = 0; }
  }
  if (rows) sink(ctx, dosage.data(), missing.data(), (long long)rows);
}

// The same walk over the file, handing on the inflated probability blocks themselves (for the device to unpack,
// dgrm_push_bgen_blocks_host): pieces of up to 32 MB, block k of a piece at data[offsets[k] .. offsets[k+1]).
// The file is read sequentially; the zlib inflation of a piece's variants -- the only per-byte host work left on this
// path -- runs on all host threads (each variant is its own zlib stream with a known inflated size).
void Genotype::streamBGENBlocks(BGENBlockSink sink, void* ctx) const {
  BgenReader r(bgenFile);
  if (fseeko(r.f, bgenFirstVariant, SEEK_SET) != 0) misc.error("Error: Unexpected end of the BGEN file [ " + bgenFile + " ].", 0);
  const uint32_t nSamples = (uint32_t)nBEDIndividuals;
  std::vector<unsigned char> comp, data;
  std::vector<long long> offsets(1, 0), cpos(1, 0);
  const bool trace = getenv("DGRM_TRACE_BGEN") != NULL;   // seconds spent reading / inflating / in the sink, to stderr
  double tRead = 0., tInflate = 0., tSink = 0.;
  auto now = []() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; };
  double tMark = now();
  auto flush = [&]() {
    const long long rows = (long long)offsets.size() - 1;
    if (rows == 0) return;
    tRead += now() - tMark;
    tMark = now();
    // one allocation for the whole walk where possible: fresh pages are the expensive part of a large buffer
    if (data.capacity() < (size_t)offsets.back()) data.reserve((size_t)offsets.back() + (size_t)(offsets.back() / rows) * 64);
    data.resize((size_t)offsets.back());
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : failed)
    for (long long k = 0; k < rows; ++k) {
      const size_t dlen = (size_t)(offsets[(size_t)k + 1] - offsets[(size_t)k]);
      const size_t clen = (size_t)(cpos[(size_t)k + 1] - cpos[(size_t)k]);
      unsigned char* dst = data.data() + offsets[(size_t)k];
      const unsigned char* src = comp.data() + cpos[(size_t)k];
      if (bgenCompression) {
        uLongf got = (uLongf)dlen;
        if (uncompress(dst, &got, src, (uLong)clen) != Z_OK || got != dlen) failed += 1;
      } else {
        memcpy(dst, src, dlen);
      }
    }
    if (failed) misc.error("Error: A genotype block of the BGEN file [ " + bgenFile + " ] could not be decompressed.", 0);
    tInflate += now() - tMark;
    tMark = now();
    sink(ctx, data.data(), offsets.data(), rows);
    tSink += now() - tMark;
    tMark = now();
    comp.clear();
    offsets.assign(1, 0);
    cpos.assign(1, 0);
  };
  for (size_t v = 0; v < bgenKeepVariant.size(); ++v) {
    read_variant_id(r, bgenLayout, nSamples);
    const long long len = genotype_block_bytes(r, bgenLayout, bgenCompression, nSamples);
    if (!bgenKeepVariant[v]) { r.skip(len); continue; }
    // stored and inflated sizes of this variant's probability data (genfile::bgen::read_genotype_data_block)
    long long dlen = len, clen = len;
    if (bgenLayout == 2 && bgenCompression) { dlen = r.u32(); clen = len - 4; }
    else if (bgenLayout == 1 && bgenCompression) dlen = 6ll * nSamples;
    r.need(clen);
    if (dlen > (1ll << 32)) misc.error("Error: A genotype block of the BGEN file [ " + bgenFile + " ] could not be decompressed.", 0);
    const size_t at = comp.size();
    comp.resize(at + (size_t)clen);
    r.bytes(comp.data() + at, (size_t)clen);
    cpos.push_back((long long)comp.size());
    offsets.push_back(offsets.back() + dlen);
    if (offsets.back() >= (32ll << 20) || offsets.size() > 4096) flush();
  }
  flush();
  if (trace) fprintf(stderr, "[bgen] read %.3f s, inflate %.3f s, sink %.3f s\n", tRead, tInflate, tSink);
}

// genotypebgen.cpp:739-779: SNPs whose standard deviation is 0 are listed in <out>.snps.novariance and dropped.  The
// device never let their rows into S or N (dgrm_push_dosage_host), so only the records go here.
void Genotype::removeSNPsWithZeroVariance() {
  std::vector<std::string> removeSNPs;
  std::vector<SNP> keptSNPs;
  std::vector<std::string> keptIds;
  for (int k = 0; k < nSNPs; ++k) {
    if (SNPs[(size_t)k].standardDev != 0.) { keptSNPs.push_back(SNPs[(size_t)k]); keptIds.push_back(SNPIds[(size_t)k]); }
    else removeSNPs.push_back(SNPIds[(size_t)k]);
  }
  if (removeSNPs.empty()) return;
  std::ofstream out((options.outFile + ".snps.novariance").c_str());
  for (size_t i = 0; i < removeSNPs.size(); ++i) out << removeSNPs[i] << std::endl;
  *misc.message << "Warning: There are " << removeSNPs.size() << " SNPs which have zero variance and will be removed. They are stored in file [ "
                << (options.outFile + ".snps.novariance") << " ]." << std::endl;
  SNPs.swap(keptSNPs);
  SNPIds.swap(keptIds);
  nSNPs = (int)SNPs.size();
  SNPIdsIdx.clear();
  for (int k = 0; k < nSNPs; ++k) SNPIdsIdx[SNPIds[(size_t)k]] = k;
}

// readBEDFile's role (genotype.cpp:548-724) for a Genotype that will feed several kernels: validate the header
// and keep all rows of the file packed in HBM; nothing is decoded to fp64.
void Genotype::loadResident() {
  if (eng_) return;
  eng_ = create_engine();
  ownsEngine_ = true;
  engine_check(eng_, dgrm_set_individuals(eng_, nBEDIndividuals, individualBEDIdxsMaskToKeep.data()));
  *misc.message << "Reading genotype data from file [ " << bedFile << " ] ..." << std::endl;
  engine_check(eng_, dgrm_store_bed_file(eng_, bedFile.c_str(), nBEDSNPs));
  storeRows.resize((size_t)nSNPs);
  for (int k = 0; k < nSNPs; ++k) storeRows[(size_t)k] = SNPsBEDIdxsToKeep.empty() ? k : SNPsBEDIdxsToKeep[(size_t)k];
}

// genotype.cpp:1293-1345
void Genotype::groupSNPs(GroupBy groupBy) {
  groupedSNPs.clear();
  if (groupBy == byPosition) {
    *misc.message << "Grouping SNPs in regions by position..." << std::endl;
    groupSNPsByPosition();
  } else if (groupBy == byGene || groupBy == byGroup) {
    *misc.message << "Grouping SNPs in regions by group..." << std::endl;
    groupSNPsByGene();
  } else if (groupBy == byAll) {
    groupedBy = byAll;
    groupedSNPs["all"] = std::set<std::string>(SNPIds.begin(), SNPIds.end());
  } else {
    groupedBy = ungrouped;
    return;
  }
  *misc.message << "Created " << groupedSNPs.size() << " SNP group(s)" << std::endl;
  // groups with too few SNPs are dropped and listed in <out>.filteredregions (genotype.cpp:1328-1345)
  std::vector<std::string> small;
  for (auto& g : groupedSNPs)
    if ((int)g.second.size() < options.minSNPsInRegions) small.push_back(g.first);
  if (!small.empty()) {
    std::ofstream f((options.outFile + ".filteredregions").c_str());
    for (auto& g : small) {
      f << g << " " << groupedSNPs[g].size();
      groupedSNPs.erase(g);
    }
    *misc.message << small.size() << " regions with less than " << options.minSNPsInRegions
                  << " SNPs have been filtered. Regions saved on file [ " << options.outFile << ".filteredregions ]." << std::endl;
  }
}

// genotype.cpp:1347-1401.  Region j of a chromosome covers [j*step, j*step + regionSize) with
// step = regionSize - regionOverlapping; a SNP belongs to every region that contains its position.
void Genotype::groupSNPsByPosition() {
  groupedBy = byPosition;
  const int size = options.regionSize;
  const int step = options.regionSize - options.regionOverlapping;
  if (size <= 0 || step <= 0) misc.error("Error: Invalid region size / overlap.", 0);
  for (int i = 0; i < nSNPs; ++i) {
    const int pos = SNPs[(size_t)i].position;
    const int last = pos / step;                               // last region starting at or before pos
    const int first = pos < size ? 0 : (pos - size) / step + 1;  // first region ending after pos
    for (int j = first; j <= last; ++j) {
      std::ostringstream name;
      name << SNPs[(size_t)i].chr << "_" << j * step << "-" << j * step + size;
      groupedSNPs[name.str()].insert(SNPs[(size_t)i].name);
    }
  }
}

// genotype.cpp:1403-1446: lines `snp group [group ...]`; SNPs that are not in the genotype are skipped
void Genotype::groupSNPsByGene() {
  groupedBy = byGene;
  *misc.message << "Reading SNP grouping data from file [ " << options.regionsFile << " ] ..." << std::endl;
  std::ifstream file(options.regionsFile.c_str());
  if (!file) misc.error("Error: The file [ " + options.regionsFile + " ] does not exist or cannot be opened.", 0);
  std::string line;
  while (std::getline(file, line)) {
    std::istringstream ss(line);
    std::string snp, group;
    ss >> snp;
    if (!SNPIdsIdx.count(snp)) continue;
    while (ss >> group) {
      if (groupedSNPs[group].count(snp))
        *misc.message << "Warning: The group " << group << " is defined more than once for SNP " << snp << "." << std::endl;
      groupedSNPs[group].insert(snp);
    }
  }
}

// genotype.cpp:1567-1594
std::vector<std::string> Genotype::getGroup(const std::string& group, bool writeGroup) {
  if (groupedBy == ungrouped)
    misc.error("Error: An internal error was happened. The SNPs within a group can not be retrieved. Groups are not defined.", 0);
  if (!groupedSNPs.count(group))
    misc.error("Error: An internal error was happened. The group " + group + " is not a valid group.", 0);
  std::vector<std::string> list(groupedSNPs[group].begin(), groupedSNPs[group].end());
  if (writeGroup) {
    std::ofstream f((options.outFile + ".region." + group).c_str());
    for (auto& s : list) f << s << "\n";
  }
  return list;
}

// genotype.cpp:1596-1621 -> filterSNPsAndIndividuals(:972): same individuals, the group's SNPs in genotype order,
// per-SNP records copied from the parent (:1111).  The fp64 matrices the reference filters here do not exist:
// the group shares the parent's resident rows.
void Genotype::genotypeOfSNPsGroup(const std::string& group, Genotype* g, bool writeGroup) {
  if (g == this)
    misc.error("Error: An internal error was happened. When getting a genotpe group, the resultat genotype must be different than the source genotype.", 0);
  const std::vector<std::string> list = getGroup(group, writeGroup);
  if (list.empty()) misc.error("Error: The group " + group + " is not a valid group. There are not SNPs inside this group.", 0);
  loadResident();
  std::vector<int> idx;
  for (auto& sname : list) idx.push_back(SNPIdsIdx[sname]);
  std::sort(idx.begin(), idx.end());
  g->nIndividuals = nIndividuals;
  g->individuals = individuals;
  g->individualIds = individualIds;
  g->normalized = normalized;
  g->bedFile = bedFile;
  g->nBEDSNPs = nBEDSNPs;
  g->nBEDIndividuals = nBEDIndividuals;
  g->individualBEDIdxsMaskToKeep = individualBEDIdxsMaskToKeep;
  g->groupedBy = ungrouped;
  g->groupedSNPs.clear();
  g->SNPs.clear(); g->SNPIds.clear(); g->SNPIdsIdx.clear(); g->SNPsBEDIdxsToKeep.clear(); g->storeRows.clear();
  for (int k : idx) {
    g->SNPIdsIdx[SNPIds[(size_t)k]] = (int)g->SNPs.size();
    g->SNPs.push_back(SNPs[(size_t)k]);
    g->SNPIds.push_back(SNPIds[(size_t)k]);
    g->SNPsBEDIdxsToKeep.push_back(storeRows[(size_t)k]);
    g->storeRows.push_back(storeRows[(size_t)k]);
  }
  g->nSNPs = (int)g->SNPs.size();
  if (g->eng_ && g->ownsEngine_) dgrm_destroy(g->eng_);
  g->eng_ = eng_;
  g->ownsEngine_ = false;
}

void Genotype::readBIMFile(const std::string& f, const std::set<std::string>& keepSNPs) {
  *misc.message << "Reading SNPs data from file [ " << f << ".bim ] ..." << std::endl;
  std::ifstream file((f + ".bim").c_str());
  if (!file) misc.error("Error: The file [ " + f + ".bim ] does not exist or cannot be opened.", 0);
  std::set<std::string> seen;
  std::string line;
  int ibed = 0;
  while (std::getline(file, line)) {
    std::istringstream ss(line);
    std::string t[6];
    ss >> t[0] >> t[1] >> t[2] >> t[3] >> t[4] >> t[5];
    if (t[1].empty()) continue;
    if (!keepSNPs.empty() && keepSNPs.find(t[1]) == keepSNPs.end()) { ++ibed; continue; }
    SNPsBEDIdxsToKeep.push_back(ibed++);
    SNP s;
    s.chr = t[0]; s.name = t[1]; s.position = atoi(t[3].c_str()); s.allele1 = t[4]; s.allele2 = t[5];
    s.nNonMissing = 0; s.frequencies[0] = s.frequencies[1] = s.frequencies[2] = s.frequencies[3] = 0;
    s.p1 = s.p2 = s.standardDev = 0.;
    if (!seen.insert(s.name).second)
      misc.error("Error: The SNP with name: " + s.name + " appears more than one time in the genotypes BIM file.", 0);
    SNPs.push_back(s);
    SNPIds.push_back(s.name);
  }
  nBEDSNPs = ibed;
  nSNPs = (int)SNPs.size();
  if (keepSNPs.empty()) {
    *misc.message << nSNPs << " SNPs found." << std::endl;
    SNPsBEDIdxsToKeep.clear();
  } else {
    *misc.message << ibed << " SNPs found. " << nSNPs << " kept for the analysis." << std::endl;
  }
}

void Genotype::readFAMFile(const std::string& f, const std::set<std::string>& keepInds) {
  *misc.message << "Reading Individuals data from file [ " << f << ".fam ] ..." << std::endl;
  std::ifstream file((f + ".fam").c_str());
  if (!file) misc.error("Error: The file [ " + f + ".fam ] does not exist or cannot be opened.", 0);
  std::set<std::string> seen;
  std::string line;
  while (std::getline(file, line)) {
    std::istringstream ss(line);
    Individual ind;
    ss >> ind.familyID >> ind.individualID >> ind.paternalID >> ind.maternalID >> ind.sex;
    if (ind.familyID.empty()) continue;
    if ((ss >> ind.phenotype).fail())
      misc.error("Error. The phenotype of the individual " + ind.familyID + " " + ind.individualID + " is not an integer.", 0);
    const std::string key = ind.familyID + "@" + ind.individualID;
    if (!keepInds.empty() && keepInds.find(key) == keepInds.end()) { individualBEDIdxsMaskToKeep.push_back(0); continue; }
    individualBEDIdxsMaskToKeep.push_back(1);
    if (!seen.insert(key).second)
      misc.error("Error: The individual with family Id: " + ind.familyID + " and individual Id: " + ind.individualID +
                 " appears more than one time in the genotypes FAM file.", 0);
    individuals.push_back(ind);
    individualIds.push_back(key);
  }
  nBEDIndividuals = (int)individualBEDIdxsMaskToKeep.size();
  nIndividuals = (int)individuals.size();
  if (keepInds.empty()) *misc.message << nIndividuals << " individuals found." << std::endl;
  else *misc.message << nBEDIndividuals << " individuals found. " << nIndividuals << " kept for the analysis." << std::endl;
}

// ------------------------------------------------------------------------------------------- Kernel
Kernel::Kernel(Genotype* genotype, bool normalizeGRM)
    : type(kernelGRM), name("default"), kernel(NULL), N(NULL), normalized(false), diagonalized(false), asymmetric(false),
      nIndividuals(genotype->nIndividuals), eng_(NULL), ownsEngine_(true), isView_(false) {
  std::vector<Genotype*> one(1, genotype);
  compute(one, normalizeGRM);
}

// One GRM from several genotype files with the same individuals and disjoint SNPs (--bfile-list): what
// loadGRMUsingOptions builds by adding per-file kernels (method 0, auxiliar.cpp:620-672) or by joining the genotypes
// first (method 1, :673-690).  On the device both are the same thing: S and N simply keep accumulating over the files'
// SNP chunks before the one division.
Kernel::Kernel(const std::vector<Genotype*>& genotypes, bool normalizeGRM)
    : type(kernelGRM), name("default"), kernel(NULL), N(NULL), normalized(false), diagonalized(false), asymmetric(false),
      nIndividuals(genotypes.empty() ? 0 : genotypes[0]->nIndividuals), eng_(NULL), ownsEngine_(true), isView_(false) {
  if (genotypes.empty()) misc.error("Error: No genotype file(s) specified.", 0);
  compute(genotypes, normalizeGRM);
}

void Kernel::compute(const std::vector<Genotype*>& genotypes, bool normalizeGRM) {
  Genotype* genotype = genotypes[0];
  for (int i = 0; i < genotype->nIndividuals; ++i) {
    individuals.push_back(genotype->individuals[i]);
    const std::string key = genotype->individualIds[i];
    individualIds.push_back(key);
    if (individualIdsIdx.count(key))
      misc.error("Error: The individual with family Id: " + genotype->individuals[i].familyID + " and individual Id: " +
                 genotype->individuals[i].individualID + " appears more than one time in the genotypes.", 0);
    individualIdsIdx[key] = i;
  }
  int totalSNPs = 0;
  for (size_t g = 0; g < genotypes.size(); ++g) {
    if (g > 0) {
      if (genotypes[g]->individualIds != individualIds || genotypes[g]->nBEDIndividuals != genotype->nBEDIndividuals ||
          genotypes[g]->individualBEDIdxsMaskToKeep != genotype->individualBEDIdxsMaskToKeep)
        misc.error("Error: An internal error was happened. GRMs with different individuals can not be added.", 0);   // kernel.cpp:1740
      std::set<std::string> seen(randomVarNames.begin(), randomVarNames.end());
      for (size_t k = 0; k < genotypes[g]->SNPIds.size(); ++k)
        if (seen.count(genotypes[g]->SNPIds[k]))   // auxiliar.cpp:651-655
          misc.error("Error: When computing a GRM from multiple genotype files. There are SNPs repeated in different files.", 0);
    }
    randomVarNames.insert(randomVarNames.end(), genotypes[g]->SNPIds.begin(), genotypes[g]->SNPIds.end());
    totalSNPs += genotypes[g]->nSNPs;
  }

  *misc.message << "Starting GRM computation..." << std::endl;
  if (genotype->engine()) {
    if (genotypes.size() != 1) misc.error("Error: An internal error was happened. A resident genotype cannot be joined with others.", 0);
    // the Genotype keeps its packed rows in HBM (loadResident): this kernel's SNPs are re-read from there
    eng_ = genotype->engine();
    engs_.push_back(eng_);
    ownsEngine_ = false;
    engine_check(eng_, dgrm_begin(eng_, genotype->nSNPs));
    std::vector<uint8_t> keep((size_t)genotype->nBEDSNPs, 0);
    for (int k = 0; k < genotype->nSNPs; ++k) keep[(size_t)genotype->storeRows[(size_t)k]] = 1;
    engine_check(eng_, dgrm_push_stored(eng_, 0, genotype->nBEDSNPs, keep.data()));
  } else {
    // One engine, or the lower triangle split over several engines (GPUs) that are all fed by this host rank: each
    // contracts its own tile set; the per-SNP statistics are the same integers everywhere.  The reference distributes
    // fp64 genotype blocks over BLACS instead (genotype.cpp:639-707).
    const int G = std::max(1, options.engines);
    for (int r = 0; r < G; ++r) {
      dgrm_engine* e = create_engine(r, G);
      engs_.push_back(e);
      engine_check(e, dgrm_set_individuals(e, genotype->nBEDIndividuals, genotype->individualBEDIdxsMaskToKeep.data()));
      engine_check(e, dgrm_begin(e, totalSNPs));
    }
    eng_ = engs_[0];
    for (size_t g = 0; g < genotypes.size(); ++g) {
      Genotype* gt = genotypes[g];
      if (gt->type == GenotypeAttributes::probabilities) {
        // readBGENFile's probability loop (genotypebgen.cpp:639-689): every block of dosage rows goes to every engine,
        // which standardises it (ProbabilitiesSetter::finalise) and contracts its own tiles
        struct Ctx { std::vector<dgrm_engine*>* engs; const Genotype* gt; } ctx = {&engs_, gt};
        if (!options.bgenHostDecode) {
          // the inflated blocks go to the device as they are; bgen_unpack_kernel forms the dosage rows there
          gt->streamBGENBlocks(
              [](void* c, const unsigned char* data, const long long* offsets, long long rows) {
                Ctx* cx = static_cast<Ctx*>(c);
                std::vector<dgrm_engine*>& es = *cx->engs;
                std::vector<int64_t> off(offsets, offsets + rows + 1);
                for (size_t r = 0; r < es.size(); ++r)
                  engine_check(es[r], dgrm_push_bgen_blocks_host(es[r], data, off.data(), rows, cx->gt->bgenLayout,
                                                                 cx->gt->nBEDIndividuals, cx->gt->individualBEDIdxsMaskToKeep.data()));
              },
              &ctx);
          continue;
        }
        gt->streamBGEN(
            [](void* c, const double* dosage, const unsigned char* missing, long long rows) {
              std::vector<dgrm_engine*>& es = *static_cast<Ctx*>(c)->engs;
              const long long n = dgrm_num_individuals(es[0]);
              for (size_t r = 0; r < es.size(); ++r)
                engine_check(es[r], dgrm_push_dosage_host(es[r], dosage, n, missing, n, rows));
            },
            &ctx);
        continue;
      }
      *misc.message << "Reading genotype data from file [ " << gt->bedFile << " ] ..." << std::endl;
      if (G == 1) {
        std::vector<int64_t> keepIdx(gt->SNPsBEDIdxsToKeep.begin(), gt->SNPsBEDIdxsToKeep.end());
        engine_check(eng_, dgrm_load_bed_file(eng_, gt->bedFile.c_str(), gt->nBEDSNPs, keepIdx.empty() ? NULL : keepIdx.data(),
                                              (int64_t)keepIdx.size()));
        continue;
      }
      // several engines: the .bed is read once and every slab of rows is handed to every engine
      engine_check(eng_, dgrm_validate_bed_header(eng_, gt->bedFile.c_str()));   // genotype.cpp:578-605
      std::vector<uint8_t> keep;
      if (!gt->SNPsBEDIdxsToKeep.empty()) {
        keep.assign((size_t)gt->nBEDSNPs, 0);
        for (size_t k = 0; k < gt->SNPsBEDIdxsToKeep.size(); ++k) keep[(size_t)gt->SNPsBEDIdxsToKeep[k]] = 1;
      }
      FILE* f = fopen(gt->bedFile.c_str(), "rb");
      if (!f || fseek(f, 3, SEEK_SET) != 0) misc.error("Error: The file [ " + gt->bedFile + " ] does not exist or cannot be opened.", 0);
      const int64_t bpr = ((int64_t)gt->nBEDIndividuals + 3) / 4;
      const int64_t slab = std::max<int64_t>(1, (64ll << 20) / bpr);
      std::vector<uint8_t> buf((size_t)(slab * bpr));   // pageable: the copy is staged before the push returns
      for (int64_t r0 = 0; r0 < gt->nBEDSNPs; r0 += slab) {
        const int64_t cnt = std::min<int64_t>(slab, gt->nBEDSNPs - r0);
        if ((int64_t)fread(buf.data(), (size_t)bpr, (size_t)cnt, f) != cnt) {
          fclose(f);
          misc.error("Error: An internal error has happened when reading the genotypes file.", 0);   // genotype.cpp:693-696
        }
        for (int r = 0; r < G; ++r)
          engine_check(engs_[r], dgrm_push_bed_host(engs_[r], buf.data(), bpr, cnt, keep.empty() ? NULL : keep.data() + r0));
      }
      fclose(f);
    }
  }
  int rc = DGRM_OK;
  for (size_t r = 0; r < engs_.size(); ++r) {
    const int rcr = dgrm_finish(engs_[r]);
    if (rc == DGRM_OK) rc = rcr;
  }
  // per-SNP records, as gatherSNPData leaves them on the root (genotype.cpp:859-870)
  {
    const int m = totalSNPs;
    std::vector<int32_t> freq((size_t)m * 4), nn(m);
    std::vector<double> p1(m), p2(m), sd(m);
    engine_check(eng_, dgrm_snp_stats(eng_, freq.data(), nn.data(), p1.data(), p2.data(), sd.data()));
    int k = 0;
    for (size_t g = 0; g < genotypes.size(); ++g)
      for (int q = 0; q < genotypes[g]->nSNPs; ++q, ++k) {
        SNP& s = genotypes[g]->SNPs[q];
        for (int c = 0; c < 4; ++c) s.frequencies[c] = freq[(size_t)k * 4 + c];
        s.nNonMissing = nn[k]; s.p1 = p1[k]; s.p2 = p2[k]; s.standardDev = sd[k];
      }
  }
  if (rc == DGRM_E_BADSNP) {
    // genotype.cpp:924-939: list the offenders in <out>.badsnps, then fail
    int64_t cnt = 0;
    std::vector<int64_t> idx(65536);
    dgrm_bad_snps(eng_, idx.data(), (int64_t)idx.size(), &cnt);
    std::ofstream bad((options.outFile + ".badsnps").c_str());
    for (int64_t i = 0; i < std::min<int64_t>(cnt, (int64_t)idx.size()); ++i) bad << randomVarNames[(size_t)idx[i]] << std::endl;
    *misc.message << "The SNPs are stored in the file [ " << options.outFile + ".badsnps" << " ]." << std::endl;
  }
  engine_check(eng_, rc);
  for (size_t g = 0; g < genotypes.size(); ++g) genotypes[g]->normalized = true;  // the standardised values exist (on the GPU) from here on
  {
    // BGEN sources: readBGENFile ends with removeSNPsWithZeroVariance (genotypebgen.cpp:717), before the Kernel
    // constructor copies SNPIds into randomVarNames (kernel.cpp:85)
    bool any = false;
    for (size_t g = 0; g < genotypes.size(); ++g)
      if (genotypes[g]->type == GenotypeAttributes::probabilities) { genotypes[g]->removeSNPsWithZeroVariance(); any = true; }
    if (any) {
      randomVarNames.clear();
      totalSNPs = 0;
      for (size_t g = 0; g < genotypes.size(); ++g) {
        randomVarNames.insert(randomVarNames.end(), genotypes[g]->SNPIds.begin(), genotypes[g]->SNPIds.end());
        totalSNPs += genotypes[g]->nSNPs;
      }
    }
  }
  if (!options.flatGRMNormalization)
    *misc.message << "GRM will be normalized using the individuals shared non-missing SNPs." << std::endl;
  else
    *misc.message << "GRM will be normalized for all individuals by the total number of SNPs: " << double(totalSNPs) << "." << std::endl;
  normalized = normalizeGRM;
  if (!options.lazyHostMatrices) fetch();
  *misc.message << "GRM computation finished." << std::endl;
}

// loadGRMUsingOptions (auxiliar.cpp:581-700) for the genotype sources of --make-grm: one file (--bfile) or a list
// (--bfile-list, with --grm-join-method 0 = add per-file GRMs, 1 = join the genotypes first).
Kernel* loadGRMUsingOptions() {
  if (options.genotypeListFile.empty()) {
    if (options.genotypeFile.empty() && options.bgenGenotypeFile.empty()) misc.error("Error: No genotype file(s) specified.", 0);
    // analysis.cpp:515-522: --bfile -> calls, --bgen -> probabilities
    Genotype* genotype = options.bgenGenotypeFile.empty() ? new Genotype(options.genotypeFile)
                                                          : new Genotype(options.bgenGenotypeFile, GenotypeAttributes::probabilities);
    Kernel* grm = new Kernel(genotype);
    delete genotype;
    return grm;
  }
  if (options.GRMJoinMethod == 0)
    *misc.message << "Computing GRM using a list of genotype files. Method 0: First compute GRMs for each genotype, then add GRMs." << std::endl;
  else
    *misc.message << "Computing GRM using a list of genotype files. Method 1: First join genotypes, then compute GRM." << std::endl;
  std::ifstream file(options.genotypeListFile.c_str());
  if (!file) misc.error("Error: The file [ " + options.genotypeListFile + " ] does not exist or cannot be opened.", 0);
  std::vector<Genotype*> genotypes;
  std::set<std::string> seenFiles;
  std::string line;
  while (std::getline(file, line)) {
    std::istringstream ss(line);
    std::string partialFileName;
    if (!(ss >> partialFileName)) continue;
    if (!seenFiles.insert(partialFileName).second)   // genotype.cpp:397-400
      misc.error("Error: The file [ " + partialFileName + ".bim/bed/fam ] appears defined twice.", 0);
    if (!genotypes.empty()) *misc.message << (options.GRMJoinMethod == 0 ? "Adding GRM..." : "Adding genotype...") << std::endl;
    genotypes.push_back(new Genotype(partialFileName));
  }
  Kernel* grm = new Kernel(genotypes, true);
  for (size_t g = 0; g < genotypes.size(); ++g) delete genotypes[g];
  *misc.message << "Genotype loading and GRM computation finished." << std::endl;
  return grm;
}

Kernel::Kernel()
    : type(kernelGRM), name("default"), kernel(NULL), N(NULL), normalized(false), diagonalized(false), asymmetric(false),
      nIndividuals(0), eng_(NULL), ownsEngine_(true), isView_(false) {}

// readGRM (kernel.cpp:1023-1196): <f>.grm.ids / .grm.snps / .grm.dat -> kernel (lower) and N (lower; the file holds
// N's upper triangle, matrix.cpp:2324-2334).  Single host rank (1 x 1 grid) only.
Kernel::Kernel(const std::string& f)
    : type(kernelGRM), name(f), kernel(NULL), N(NULL), normalized(true), diagonalized(false), asymmetric(false),
      nIndividuals(0), eng_(NULL), ownsEngine_(true), isView_(false) {
  if (communicator->nProcRows != 1 || communicator->nProcCols != 1)
    misc.error("Error: reading a GRM file is implemented for a 1x1 process grid in this adapter.", 0);
  *misc.message << "Reading GRM from [ " << f << ".grm.(ids/snps/dat) ]..." << std::endl;
  for (auto& r : read_table(f + ".grm.ids", 2)) {
    Individual ind;
    ind.familyID = r[0]; ind.individualID = r[1]; ind.phenotype = 0;
    const std::string key = r[0] + "@" + r[1];
    if (individualIdsIdx.count(key))
      misc.error("Error: The individual with family Id: " + r[0] + " and individual Id: " + r[1] + " appears more than one time in file [ " + f + ".grm.ids ].", 0);
    individualIdsIdx[key] = (int)individuals.size();
    individuals.push_back(ind);
    individualIds.push_back(key);
  }
  nIndividuals = (int)individuals.size();
  for (auto& r : read_table(f + ".grm.snps", 1)) randomVarNames.push_back(r[0]);
  std::ifstream dat((f + ".grm.dat").c_str(), std::ios::binary);
  if (!dat) misc.error("Error: The file [ " + f + ".grm.dat ] does not exist or cannot be opened.", 0);
  unsigned char header[14];
  dat.read((char*)header, 14);
  // kernel.cpp:1066-1138: magic 'G','R','M',0,0x5A,0x99, version 2, doubles, sizeof(double), 1 = normalized
  if (!dat || header[0] != 'G' || header[1] != 'R' || header[2] != 'M' || header[4] != 0x5A || header[5] != 0x99 ||
      header[7] != 0x1 || header[8] != sizeof(double))
    misc.error("Error: The file [ " + f + ".grm.dat ] is not a valid GRM file.", 0);
  if (header[9] != 0x1) misc.error("Error: Only normalized, non-diagonalized GRM files are supported by this adapter.", 0);
  const int n = nIndividuals;
  std::vector<double> col((size_t)n + 1);
  kernel = new Matrix(cyclicDistribution);
  N = new Matrix(cyclicDistribution);
  kernel->initParameters(n, n, communicator->nDefaultBlockRows, communicator->nDefaultBlockCols);
  N->initParameters(n, n, communicator->nDefaultBlockRows, communicator->nDefaultBlockCols);
  kernel->symmetric = N->symmetric = true;
  kernel->uplo = N->uplo = 'L';
  for (int c = 0; c < n; ++c) {
    dat.read((char*)col.data(), (std::streamsize)((size_t)(n + 1) * sizeof(double)));
    if (!dat) misc.error("Error: An error was happened when reading a matrix. Could the file be corrupt?", 0);
    for (int r = 0; r <= c; ++r) N->m[(size_t)r * n + c] = col[(size_t)r];          // N[r][c] -> lower (c, r)
    for (int r = c; r < n; ++r) kernel->m[(size_t)c * n + r] = col[(size_t)r + 1];  // A[r][c]
  }
  *misc.message << "GRM loaded." << std::endl;
}

static void add_lower(Matrix* a, const Matrix* b, double sa, double sb) {
  check_same_structure(a, b, "addition");
  for (int c = 0; c < a->nCols; ++c)
    for (int r = 0; r < a->nRows; ++r) {
      if (a->local2globalCol(c) > a->local2globalRow(r)) continue;
      double& x = a->m[(size_t)c * a->nRows + r];
      x = sa * x + sb * b->m[(size_t)c * a->nRows + r];
    }
}

static Matrix* clone_matrix(const Matrix* src) {
  Matrix* m = new Matrix(src->distribution);
  m->initParameters(src->nGlobRows, src->nGlobCols, src->nBlockRows, src->nBlockCols);
  m->symmetric = src->symmetric;
  m->uplo = src->uplo;
  memcpy(m->m, src->m, (size_t)src->nRows * src->nCols * sizeof(double));
  return m;
}

void Kernel::addKernels(double s1, Kernel* grm1, double s2, Kernel* grm2) {
  if (grm1 == this || grm2 == this)
    misc.error("Error: An internal error was happened when adding two GRMs. GRMs to add must be different instances than the resulting GRM.", 0);
  if ((s1 != 1. && s1 != -1.) || (s2 != 1. && s2 != -1.))
    misc.error("Error: An internal error was happened. Scaling factor when adding or substracting GRMs can only be 1 or -1.", 0);
  if (!grm1->N || !grm2->N) misc.error("Error: An internal error was happened. GRMs can not be added without the N matrix.", 0);
  if (grm1->individualIds != grm2->individualIds || grm1->nIndividuals != grm2->nIndividuals)
    misc.error("Error: An internal error was happened. GRMs with different individuals can not be added.", 0);
  // SNP bookkeeping (kernel.cpp:1819-1856)
  std::vector<std::string> result;
  std::set<std::string> set1(grm1->randomVarNames.begin(), grm1->randomVarNames.end());
  std::set<std::string> set2(grm2->randomVarNames.begin(), grm2->randomVarNames.end());
  if (s1 == s2) {
    for (auto& x : grm1->randomVarNames)
      if (set2.count(x)) misc.error("Error: Invalid operation. Two GRMs cannot be added together if the SNPs used for computing them intersect.", 0);
    result = grm2->randomVarNames;
    result.insert(result.end(), grm1->randomVarNames.begin(), grm1->randomVarNames.end());
  } else {
    const std::vector<std::string>& big = (s1 == 1.) ? grm1->randomVarNames : grm2->randomVarNames;
    const std::set<std::string>& bigset = (s1 == 1.) ? set1 : set2;
    const std::vector<std::string>& small = (s1 == 1.) ? grm2->randomVarNames : grm1->randomVarNames;
    const std::set<std::string>& smallset = (s1 == 1.) ? set2 : set1;
    for (auto& x : small)
      if (!bigset.count(x)) misc.error("Error: Invalid operation. Two GRMs cannot be substracted if all the SNPs of one are not contained into the other.", 0);
    for (auto& x : big)
      if (!smallset.count(x)) result.push_back(x);
  }
  // arithmetic on the unnormalised pair: S = A o N (denormalize), S = s1 S1 + s2 S2, N = s1 N1 + s2 N2, re-divide
  const bool n1 = grm1->normalized, n2 = grm2->normalized;
  grm1->denormalize();
  grm2->denormalize();
  delete kernel; delete N;
  kernel = clone_matrix(grm1->kernel);
  N = clone_matrix(grm1->N);
  add_lower(kernel, grm2->kernel, s1, s2);
  add_lower(N, grm2->N, s1, s2);
  if (n1) grm1->normalize();
  if (n2) grm2->normalize();
  type = kernelGRM;
  individuals = grm1->individuals;
  individualIds = grm1->individualIds;
  individualIdsIdx = grm1->individualIdsIdx;
  nIndividuals = grm1->nIndividuals;
  normalized = false;
  if (n1) normalize();
  randomVarNames = result;
  name = grm1->name + " " + grm2->name;
}

Kernel::~Kernel() {
  delete kernel;
  delete N;
  if (eng_ && isView_) {
    dgrm_set_output_subset(eng_, NULL, 0);
    dgrm_set_view(eng_, DGRM_VIEW_CURRENT);
  }
  if (ownsEngine_) {
    if (engs_.empty() && eng_) dgrm_destroy(eng_);
    for (size_t r = 0; r < engs_.size(); ++r) dgrm_destroy(engs_[r]);
  }
}

void Kernel::ensureHost() {
  if (!kernel && eng_) fetch();
}

// addKernels(1., saved, +-1., this) on the device (kernel.cpp:1705-1859): the engine combines the unnormalised S and
// the integer N of the saved and the current GRM before the one division; SNP bookkeeping as kernel.cpp:1819-1853.
Kernel* Kernel::combinedWithSaved(Kernel* saved, double scalingFactorThis) {
  if (!eng_ || saved->eng_ != eng_)
    misc.error("Error: An internal error was happened. GRMs from different engines cannot be combined on the device.", 0);
  if (scalingFactorThis != 1. && scalingFactorThis != -1.)
    misc.error("Error: An internal error was happened. Scaling factor when adding or substracting GRMs can only be 1 or -1.", 0);
  if (saved->individualIds != individualIds)
    misc.error("Error: An internal error was happened. GRMs with different individuals can not be added.", 0);
  Kernel* r = new Kernel();
  r->eng_ = eng_;
  r->engs_ = engs_;
  r->ownsEngine_ = false;
  r->isView_ = true;
  r->individuals = individuals;
  r->individualIds = individualIds;
  r->individualIdsIdx = individualIdsIdx;
  r->nIndividuals = nIndividuals;
  std::set<std::string> mine(randomVarNames.begin(), randomVarNames.end());
  if (scalingFactorThis < 0.) {
    size_t found = 0;
    for (auto& sname : saved->randomVarNames) {
      if (mine.count(sname)) ++found; else r->randomVarNames.push_back(sname);
    }
    if (found != mine.size())
      misc.error("Error: Invalid operation. Two GRMs cannot be substracted if all the SNPs of one are not contained into the other.", 0);
  } else {
    for (auto& sname : saved->randomVarNames)
      if (mine.count(sname))
        misc.error("Error: Invalid operation. Two GRMs cannot be added together if the SNPs used for computing them intersect.", 0);
    r->randomVarNames = randomVarNames;
    r->randomVarNames.insert(r->randomVarNames.end(), saved->randomVarNames.begin(), saved->randomVarNames.end());
  }
  r->name = saved->name + " " + name;
  engine_check(eng_, dgrm_set_output_subset(eng_, NULL, 0));
  engine_check(eng_, dgrm_set_view(eng_, scalingFactorThis < 0. ? DGRM_VIEW_SAVED_MINUS_CURRENT : DGRM_VIEW_SAVED_PLUS_CURRENT));
  r->normalized = true;
  if (!options.lazyHostMatrices) r->fetch();
  return r;
}

// (Re)fill kernel and N from the device tile store in this rank's block-cyclic layout.
void Kernel::fetch() {
  delete kernel; delete N;
  kernel = new Matrix(cyclicDistribution);
  N = new Matrix(cyclicDistribution);
  const int bs = communicator->nDefaultBlockRows;
  kernel->initParameters(nIndividuals, nIndividuals, bs, communicator->nDefaultBlockCols);
  N->initParameters(nIndividuals, nIndividuals, bs, communicator->nDefaultBlockCols);
  kernel->symmetric = N->symmetric = true;
  kernel->uplo = N->uplo = 'L';  // matrix.cpp:2679-2680; kernel.cpp:105-108 for the flat case
  // every engine of the partition writes the entries of the tiles it owns (dgrm.h) into the same local block
  for (size_t r = 0; r < engs_.size(); ++r) {
    dgrm_engine* e = engs_[r];
    engine_check(e, dgrm_fetch_block_cyclic(e, normalized ? DGRM_FETCH_A : DGRM_FETCH_S, kernel->m, std::max(kernel->nRows, 1),
                                            kernel->nBlockRows, kernel->nBlockCols, communicator->myRow, communicator->myCol,
                                            communicator->nProcRows, communicator->nProcCols, 'L'));
    engine_check(e, dgrm_fetch_block_cyclic(e, DGRM_FETCH_N, N->m, std::max(N->nRows, 1), N->nBlockRows, N->nBlockCols,
                                            communicator->myRow, communicator->myCol, communicator->nProcRows,
                                            communicator->nProcCols, 'L'));
  }
}

void Kernel::normalize() {
  if (normalized) return;
  ensureHost();
  if (diagonalized) misc.error("Error: Diagonalized Kernel cannot be normalized.", 0);
  if (!N) misc.error("Error: An internal error was happened. GRM cannot be normalized because normalization matrix is not present.", 0);
  if (N->uplo != kernel->uplo)
    misc.error("Error: An internal error was happened when normalizing a GRM. GRM and normalization matrix should have same uplo properties.", 0);
  kernel->elementWiseDivision(N);
  normalized = true;
}

void Kernel::denormalize() {
  if (!normalized) return;
  ensureHost();
  if (diagonalized) misc.error("Error: Diagonalized Kernel cannot be unnormalized.", 0);
  kernel->elementWiseMultiplication(N);
  normalized = false;
}

// kernel.cpp:1861-1926, order-dependent greedy; pairs arrive as (row, col) then (col, row) per hit, hits in
// column-major scan order of the lower triangle (matrix.cpp:4348-4377 on one rank).
std::vector<std::string> Kernel::minimumIndividualsRemovePairs(const std::vector<int>& g1, const std::vector<int>& g2) {
  std::vector<int> freq(nIndividuals, 0);
  for (size_t i = 0; i < g1.size(); ++i)
    if (g1[i] != g2[i]) { freq[g1[i]]++; freq[g2[i]]++; }
  std::set<int> del;
  for (size_t i = 0; i < g1.size(); ++i) {
    if (g1[i] == g2[i]) continue;
    if (freq[g1[i]] < freq[g2[i]]) {
      if (del.find(g1[i]) == del.end()) del.insert(g2[i]);
    } else {
      if (del.find(g2[i]) == del.end()) del.insert(g1[i]);
    }
  }
  std::vector<std::string> keep;
  for (int i = 0; i < nIndividuals; ++i)
    if (del.find(i) == del.end()) keep.push_back(individualIds[i]);
  return keep;
}

bool Kernel::sanitizeKernel() {
  if (!eng_) misc.error("Error: sanitizeKernel on a host-only GRM is not implemented in this adapter.", 0);
  std::vector<std::pair<int, int> > hits;  // (col, row) so that sorting gives the column-major scan order
  for (size_t r = 0; r < engs_.size(); ++r) {   // every engine scans the tiles it owns
    dgrm_engine* e = engs_[r];
    int64_t cnt = 0;
    engine_check(e, dgrm_low_overlap_pairs(e, options.minimumNumberOverlapingSNPs, NULL, 0, &cnt));
    std::vector<int64_t> pairs((size_t)std::max<int64_t>(cnt, 1) * 2);
    if (cnt > 0) engine_check(e, dgrm_low_overlap_pairs(e, options.minimumNumberOverlapingSNPs, pairs.data(), cnt, &cnt));
    for (int64_t i = 0; i < cnt; ++i) hits.push_back(std::make_pair((int)pairs[2 * i + 1], (int)pairs[2 * i]));
  }
  std::sort(hits.begin(), hits.end());
  std::vector<int> i1, i2;
  for (size_t i = 0; i < hits.size(); ++i) {
    i1.push_back(hits[i].second); i2.push_back(hits[i].first);
    i1.push_back(hits[i].first);  i2.push_back(hits[i].second);
  }
  const std::vector<std::string> keep = minimumIndividualsRemovePairs(i1, i2);
  const double proportion = double(keep.size()) / double(nIndividuals);
  if (proportion < options.maximumProportionOfElimitaedIndividuals) {
    *misc.message << "More than 10% of the individuals have to be removed for not having an overlaping of at least "
                  << int(options.minimumNumberOverlapingSNPs) << " SNPs on all pairs. Skipping this GRM." << std::endl;
    return false;
  }
  if (keep.size() == 1) {
    *misc.message << "Sorry, not enough individuals have an overlaping of at least " << int(options.minimumNumberOverlapingSNPs)
                  << " SNPs. Skipping this GRM." << std::endl;
    return false;
  }
  if ((int)keep.size() != nIndividuals) {
    *misc.message << nIndividuals - keep.size() << " individuals will be removed since they do not have at least "
                  << options.minimumNumberOverlapingSNPs << " overlapping non-missing SNPs." << std::endl;
    filterIndividuals(keep);
  }
  return true;
}

// kernel.cpp:1378-1460: rows/columns of the already computed matrices are dropped; nothing is recomputed.
void Kernel::filterIndividuals(const std::vector<std::string>& keepIds) {
  std::set<std::string> keep(keepIds.begin(), keepIds.end());
  std::vector<long long> base = subset_;
  if (base.empty()) for (int i = 0; i < nIndividuals; ++i) base.push_back(i);
  std::vector<long long> next;
  std::vector<Individual> inds;
  std::vector<std::string> ids;
  for (int i = 0; i < nIndividuals; ++i)
    if (keep.count(individualIds[i])) { next.push_back(base[i]); inds.push_back(individuals[i]); ids.push_back(individualIds[i]); }
  subset_ = next;
  individuals = inds;
  individualIds = ids;
  individualIdsIdx.clear();
  for (size_t i = 0; i < ids.size(); ++i) individualIdsIdx[ids[i]] = (int)i;
  nIndividuals = (int)ids.size();
  std::vector<int64_t> idx(subset_.begin(), subset_.end());
  for (size_t r = 0; r < engs_.size(); ++r) engine_check(engs_[r], dgrm_set_output_subset(engs_[r], idx.data(), (int64_t)idx.size()));
  if (kernel || !options.lazyHostMatrices) fetch();
}

void Kernel::writeGRM(const std::string& f) {
  if (asymmetric) misc.error("Error: An internal error was happened. This operation can not be performed on an asymmetric GRM", 0);
  if (!normalized) misc.error("Error: An internal error was happened. GRM can only be stored in normalized form.", 0);
  if (!N && !eng_) misc.error("Error: An internal error was happened. GRM can only be stored when N information is stored.", 0);
  *misc.message << "Writing GRM on [ " << f << ".grm.(ids/snps/dat) ]..." << std::endl;
  if (communicator->mpiRoot) {
    std::ofstream ids((f + ".grm.ids").c_str());
    for (int i = 0; i < nIndividuals; ++i) ids << individuals[i].familyID << " " << individuals[i].individualID << std::endl;
    std::ofstream snps((f + ".grm.snps").c_str());
    for (size_t i = 0; i < randomVarNames.size(); ++i) snps << randomVarNames[i] << std::endl;
  }
  if (eng_) {
    // packMatrices + writeMatrixMPI funnel every block through rank 0 in the reference (matrix.cpp:2308-2341);
    // here the device packs the (n+1) x n array directly
    // (and with several engines every engine writes the runs of the file that its tiles hold)
    for (size_t r = 0; r < engs_.size(); ++r)
      engine_check(engs_[r], dgrm_write_grm_dat_part(engs_[r], (f + ".grm.dat").c_str(), r == 0 ? 1 : 0));
  } else {
    // host-only kernel (read from file / addKernels result): kernel.cpp:944-974 + matrix.cpp:2324-2334 on one rank
    if (communicator->nProcRows != 1 || communicator->nProcCols != 1)
      misc.error("Error: writing a host-side GRM is implemented for a 1x1 process grid in this adapter.", 0);
    std::ofstream dat((f + ".grm.dat").c_str(), std::ios::binary);
    const unsigned char header[14] = {'G', 'R', 'M', 0, 0x5A, 0x99, 0x2, 0x1, (unsigned char)sizeof(double), 0x1, 0, 0, 0, 0};
    dat.write((const char*)header, 14);
    const int n = nIndividuals;
    std::vector<double> col((size_t)n + 1);
    for (int c = 0; c < n; ++c) {
      for (int r = 0; r <= c; ++r) col[(size_t)r] = N->m[(size_t)r * n + c];
      for (int r = c; r < n; ++r) col[(size_t)r + 1] = kernel->m[(size_t)c * n + r];
      dat.write((const char*)col.data(), (std::streamsize)((size_t)(n + 1) * sizeof(double)));
    }
    if (!dat) misc.error("Error: An error was happened when writing a matrix. Is there enough space in the hard disk?", 0);
  }
  *misc.message << "GRM stored." << std::endl;
}

}  // namespace dgrm_host
