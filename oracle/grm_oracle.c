/*
 * ORACLE — plain-C restatement of DISSECT's `--make-grm` arithmetic.  TEST INFRASTRUCTURE ONLY:
 * nothing in the product (dissect_b200/, libdgrm.so) links or calls this file; tests/, smoke() and
 * bench.py's cpu_baseline leg use it as the checker / reported CPU baseline.
 *
 * Pinned against the reference's own binary (oracle/_ref/dissect + snpdump, built from the unmodified
 * sources by oracle/Makefile) and the golden fixtures it wrote (tests/golden/, oracle/make_golden.py):
 * counts, p1/p2/standardDev and the standardised genotypes agree bit for bit, N exactly, A to the
 * summation order of the inner dsyrk (MKL ScaLAPACK in the original — un-vendored, un-pinned,
 * Makefile:27,32,99-113 — a blocked triple loop here).  Compile with -ffp-contract=off: the reference is
 * built for plain x86-64 (no FMA contraction).
 *
 * Citations are file:line in the reference tree.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* genotype.cpp:741-787 parseSNPbyte: 2-bit pattern -> stored code (00->1, 01->0 missing, 10->2, 11->3) */
static const unsigned char kPatternToCode[4] = {1, 0, 2, 3};

/*
 * genotype.cpp:726-739 parseSNP.  rows: m x pitch packed bytes; keep[n_fam] mask or NULL.
 * Outputs codes (m x n_kept, row-major), freq (m x 4), nn, p1, p2, sd.  Returns n_kept.
 */
int oracle_decode_stats(const unsigned char* rows, long pitch, int m, int n_fam, const int* keep,
                        unsigned char* codes, int* freq, int* nn, double* p1, double* p2, double* sd) {
  int n = 0;
  for (int f = 0; f < n_fam; ++f) n += (!keep || keep[f]) ? 1 : 0;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < m; ++k) {
    const unsigned char* r = rows + (long)k * pitch;
    int fr[4] = {0, 0, 0, 0};
    int o = 0;
    for (int f = 0; f < n_fam; ++f) {
      if (keep && !keep[f]) continue;                                  /* :750 */
      const unsigned char c = kPatternToCode[(r[f >> 2] >> (2 * (f & 3))) & 3]; /* :752, :784 */
      codes[(long)k * n + o++] = c;
      fr[c]++;                                                         /* :757-775 */
    }
    for (int c = 0; c < 4; ++c) freq[k * 4 + c] = fr[c];
    nn[k] = fr[1] + fr[2] + fr[3];                                     /* :735 */
    p1[k] = (2. * (double)fr[1] + (double)fr[2]) / (2. * (double)nn[k]); /* :736 */
    p2[k] = (2. * (double)fr[3] + (double)fr[2]) / (2. * (double)nn[k]); /* :737 */
    sd[k] = sqrt(2. * p1[k] * (1. - p1[k]));                           /* :738 */
  }
  return n;
}

/* genotype.cpp:918-961 normalizeGenotypes.  z: m x n row-major.  Returns the number of sd == 0 SNPs
 * (fatal in the reference, :924-939); their rows are left as zeros. */
int oracle_standardise(const unsigned char* codes, int m, int n, const double* p2, const double* sd, double* z) {
  int bad = 0;
  for (int k = 0; k < m; ++k) bad += (sd[k] == 0.);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < m; ++k) {
    const double v = 2. * p2[k];                                       /* :922 */
    for (int i = 0; i < n; ++i) {
      double g = (double)codes[(long)k * n + i];
      const int t = (int)(g != 0.);                                    /* :948 */
      g -= (double)t * (1. + v);                                       /* :949 */
      g = (sd[k] == 0.) ? 0. : g / sd[k];                              /* :959 */
      z[(long)k * n + i] = g;
    }
  }
  return bad;
}

/*
 * kernel.cpp:94,100 -> matrix.cpp:2677-2683 pdsyrk_('L','T'): S = Z^T Z and N = M^T M (lower triangle,
 * column-major n x n, upper left untouched), then matrix.cpp:4091-4098: A = S * (1.0 / N).
 * flat_norm (kernel.cpp:102-109): N = m.
 */
void oracle_grm(const unsigned char* codes, const double* z, int m, int n, int flat_norm, double* S, double* N,
                double* A) {
  const int TB = 64;
#pragma omp parallel for schedule(dynamic)
  for (int jb = 0; jb < n; jb += TB) {
    for (int ib = jb; ib < n; ib += TB) {
      const int je = jb + TB < n ? jb + TB : n, ie = ib + TB < n ? ib + TB : n;
      double accS[64][64];
      double accN[64][64];
      memset(accS, 0, sizeof accS);
      memset(accN, 0, sizeof accN);
      for (int k = 0; k < m; ++k) {
        const double* zr = z + (long)k * n;
        const unsigned char* cr = codes + (long)k * n;
        for (int j = jb; j < je; ++j) {
          const double zj = zr[j];
          const double qj = cr[j] != 0 ? 1. : 0.;
          for (int i = ib; i < ie; ++i) {
            accS[j - jb][i - ib] += zr[i] * zj;
            accN[j - jb][i - ib] += (cr[i] != 0 ? 1. : 0.) * qj;
          }
        }
      }
      for (int j = jb; j < je; ++j)
        for (int i = (ib > j ? ib : j); i < ie; ++i) {
          const double s = accS[j - jb][i - ib];
          const double nv = flat_norm ? (double)m : accN[j - jb][i - ib];
          S[(long)j * n + i] = s;
          N[(long)j * n + i] = nv;
          A[(long)j * n + i] = s * (1.0 / nv);                         /* matrix.cpp:4096 */
        }
    }
  }
}
