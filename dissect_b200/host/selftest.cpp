// CPU self-test of the host mirror's file and algebra paths (no GPU): the invariants the reference checks in its
// manual tests test20 (test.cpp:1677-1745: write -> read round trip exactly equal; GRM of merged SNPs == sum of the
// per-part GRMs, N exactly) and test57 (test.cpp:4146-4417: add/subtract algebra), on files written by the
// reference binary itself (tests/golden/parts_*).
//   selftest roundtrip <in> <out>         read <in>.grm.*, write <out>.grm.*
//   selftest add <a> <b> <out>            out = a + b
//   selftest sub <all> <part> <out>       out = all - part
//   selftest groups <options...>          parse the options, load the .bim/.fam metadata, group the SNPs
//                                         (Genotype::groupSNPs, genotype.cpp:1293-1470) and print `group: snps...`
#include <cstdio>
#include <cstring>

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
    fprintf(stderr, "usage: selftest roundtrip|add|sub|groups ...\n");
    return 2;
  } catch (const DissectError&) {
    return 1;
  }
}
