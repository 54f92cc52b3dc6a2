// Oracle-side layout driver (TEST INFRASTRUCTURE): the reference's own block-cyclic index arithmetic,
// Matrix::local2global (matrix.cpp:435-475) and Matrix::global2local (matrix.cpp:477-488), evaluated for an
// emulated process (myRow, myCol) of an R x C grid.  The 1 x 1 shim cannot run more than one rank, but these two
// functions only read communicator->{nProcRows, nProcCols, myRow, myCol} and the matrix's block sizes, so the
// driver sets those fields and calls the reference's unmodified code.
//     layoutdump <n> <bsR> <bsC> <R> <C> <myRow> <myCol> <nLocalRows> <nLocalCols>
// prints `R <l> <g>` / `C <l> <g>` for every local row / column index, and `G <g> <procRow> <posRow> <procCol> <posCol>`
// for every global index (global2local with the row and the column parameters).
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "communicator.h"
#include "global.h"
#include "matrix.h"
#include "misc.h"
#include "options.h"

Communicator* communicator;
Options options;
Misc misc;

int main(int argc, char** argv) {
  if (argc < 10) {
    fprintf(stderr, "usage: layoutdump n bsR bsC R C myRow myCol nLocalRows nLocalCols\n");
    return 2;
  }
  const int n = atoi(argv[1]), bsR = atoi(argv[2]), bsC = atoi(argv[3]), R = atoi(argv[4]), C = atoi(argv[5]);
  const int myRow = atoi(argv[6]), myCol = atoi(argv[7]), nlr = atoi(argv[8]), nlc = atoi(argv[9]);
  char* only[1] = {argv[0]};
  communicator = new Communicator(1, only);
  communicator->nProcRows = R;
  communicator->nProcCols = C;
  communicator->myRow = myRow;
  communicator->myCol = myCol;
  Matrix m(cyclicDistribution);
  m.nGlobRows = m.nGlobCols = n;
  m.nBlockRows = bsR;
  m.nBlockCols = bsC;
  m.nRows = nlr;
  m.nCols = nlc;
  for (int l = 0; l < nlr; ++l) printf("R %d %d\n", l, m.local2global(row, l));
  for (int l = 0; l < nlc; ++l) printf("C %d %d\n", l, m.local2global(column, l));
  for (int g = 0; g < n; ++g) {
    const LocalPosition pr = m.global2local(g, bsR, R), pc = m.global2local(g, bsC, C);
    printf("G %d %d %d %d %d\n", g, pr.proc, pr.position, pc.proc, pc.position);
  }
  return 0;
}
