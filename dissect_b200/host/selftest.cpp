// CPU self-test of the host mirror's file and algebra paths (no GPU): the invariants the reference checks in its
// manual tests test20 (test.cpp:1677-1745: write -> read round trip exactly equal; GRM of merged SNPs == sum of the
// per-part GRMs, N exactly) and test57 (test.cpp:4146-4417: add/subtract algebra), on files written by the
// reference binary itself (tests/golden/parts_*).
//   selftest roundtrip <in> <out>         read <in>.grm.*, write <out>.grm.*
//   selftest add <a> <b> <out>            out = a + b
//   selftest sub <all> <part> <out>       out = all - part
//   selftest layout n bsR bsC R C r c     Matrix::initParameters / local2global for process (r, c) of an R x C grid
//   selftest bgen <options...>            parse the options (--bgen F [--bgen-l1] [--keep] [--extract]), read the BGEN
//                                         container on the host and print `IND fam ind`, `SNP chr rsid pos a1 a2` and,
//                                         per kept variant, `ROW` followed by `dosage(hex float):missing` per kept
//                                         sample -- the rows Kernel::compute hands to dgrm_push_dosage_host
//   selftest bgenblocks <options...>      the same file as Genotype::streamBGENBlocks hands it to the engines: per kept
//                                         variant `BLOCK <bytes> <FNV-1a 64 of the inflated probability block>`, then
//                                         `TIME <seconds>` for the whole walk (parallel inflate)
//   selftest groups <options...>          parse the options, load the .bim/.fam metadata, group the SNPs
//                                         (Genotype::groupSNPs, genotype.cpp:1293-1470) and print `group: snps...`
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <utility>
#include <vector>

#include "host.h"

using namespace dgrm_host;

int main(int argc, char** argv) {
  try {
    if (argc >= 4 && !strcmp(argv[1], "roundtrip")) {
      Kernel k(argv[2]);
      k.writeGRM(argv[3]);
      return 0;
    }
    if (argc >= 5 && (!strcmp(argv[1], "add") || !strcmp(argv[1], "sub"))) {
      Kernel a(argv[2]), b(argv[3]), r;
      r.addKernels(1., &a, !strcmp(argv[1], "add") ? 1. : -1., &b);
      r.writeGRM(argv[4]);
      return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "bgen")) {
      options.parseOptions(argc - 1, argv + 1);
      Genotype g(options.bgenGenotypeFile, GenotypeAttributes::probabilities);
      for (int i = 0; i < g.nIndividuals; ++i) printf("IND %s %s\n", g.individuals[i].familyID.c_str(), g.individuals[i].individualID.c_str());
      for (int k = 0; k < g.nSNPs; ++k)
        printf("SNP %s %s %d %s %s\n", g.SNPs[k].chr.c_str(), g.SNPs[k].name.c_str(), g.SNPs[k].position, g.SNPs[k].allele1.c_str(),
               g.SNPs[k].allele2.c_str());
      int n = g.nIndividuals;
      g.streamBGEN(
          [](void* ctx, const double* dosage, const unsigned char* missing, long long rows) {
            const int nn = *static_cast<int*>(ctx);
            for (long long r = 0; r < rows; ++r) {
              printf("ROW");
              for (int i = 0; i < nn; ++i) printf(" %a:%d", dosage[r * nn + i], (int)missing[r * nn + i]);
              printf("\n");
            }
          },
          &n);
      return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "bgenblocks")) {
      options.parseOptions(argc - 1, argv + 1);
      Genotype g(options.bgenGenotypeFile, GenotypeAttributes::probabilities);
      struct Out { std::vector<unsigned char> bytes; std::vector<long long> lens; } out;
      struct timespec t0, t1;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      g.streamBGENBlocks(
          [](void* ctx, const unsigned char* data, const long long* offsets, long long rows) {
            Out* o = static_cast<Out*>(ctx);
            o->bytes.insert(o->bytes.end(), data, data + offsets[rows]);
            for (long long r = 0; r < rows; ++r) o->lens.push_back(offsets[r + 1] - offsets[r]);
          },
          &out);
      clock_gettime(CLOCK_MONOTONIC, &t1);
      size_t at = 0;
      for (size_t k = 0; k < out.lens.size(); ++k) {   // hashed outside the timed walk
        unsigned long long h = 1469598103934665603ull;
        for (long long b = 0; b < out.lens[k]; ++b) { h ^= out.bytes[at + (size_t)b]; h *= 1099511628211ull; }
        at += (size_t)out.lens[k];
        printf("BLOCK %lld %016llx\n", out.lens[k], h);
      }
      printf("TIME %.6f\n", (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
      return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "groups")) {
      options.parseOptions(argc - 1, argv + 1);
      Genotype g(options.genotypeFile);
      g.groupSNPs(options.regionBy);
      for (auto& grp : g.groupedSNPs) {
        printf("GROUP %s:", grp.first.c_str());
        for (auto& snp : g.getGroup(grp.first, false)) printf(" %s", snp.c_str());
        printf("\n");
      }
      return 0;
    }
    if (argc >= 9 && !strcmp(argv[1], "layout")) {
      // selftest layout n bsR bsC R C myRow myCol: the mirror's Matrix geometry for one process of an R x C grid,
      // printed as `N rows cols`, `R l g`, `C l g` (reference: matrix.cpp:435-475, numroc: matrix.cpp:1748-1749)
      const int n = atoi(argv[2]), bsR = atoi(argv[3]), bsC = atoi(argv[4]);
      communicator->nProcRows = atoi(argv[5]);
      communicator->nProcCols = atoi(argv[6]);
      communicator->myRow = atoi(argv[7]);
      communicator->myCol = atoi(argv[8]);
      Matrix m(cyclicDistribution);
      m.initParameters(n, n, bsR, bsC);
      printf("N %d %d\n", m.nRows, m.nCols);
      for (int l = 0; l < m.nRows; ++l) printf("R %d %d\n", l, m.local2globalRow(l));
      for (int l = 0; l < m.nCols; ++l) printf("C %d %d\n", l, m.local2globalCol(l));
      return 0;
    }
    fprintf(stderr, "usage: selftest roundtrip|add|sub|groups|bgen|bgenblocks|layout ...\n");
    return 2;
  } catch (const DissectError&) {
    return 1;
  }
}
