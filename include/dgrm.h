/*
 * dgrm.h — C ABI of the B200-native GRM engine (libdgrm.so).
 *
 * This is the drop-in boundary for DISSECT's `--make-grm` hot path.  The reference has no
 * FFI seam of its own; every entry point below names the reference code it
 * replaces (file:line are relative to the reference tree).  Signatures carry only plain
 * pointers and sizes.  All functions return DGRM_OK (0) or a negative DGRM_E_* code and
 * never abort: the host adapter converts a non-zero status into `misc.error(...)`
 * (misc.cpp:151-170) using dgrm_last_error().
 *
 * Threading: call from one host thread per engine (the reference makes every MPI/BLACS call
 * from the rank's main thread).  The library never calls MPI.
 * Ownership: the library only writes into caller-owned buffers (Matrix::m is allocated with
 * new double[] and freed with delete[], matrix.cpp:1646, 99-106).
 *
 * There is no CPU fallback: every compute entry point fails with DGRM_E_CUDA when no
 * sm_100 device is usable.
 */
#ifndef DGRM_H
#define DGRM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGRM_ABI_VERSION 1

/* status codes */
#define DGRM_OK            0
#define DGRM_E_INVALID    -1  /* bad argument / call order                                   */
#define DGRM_E_CUDA       -2  /* CUDA runtime/driver error, or no sm_100 device              */
#define DGRM_E_NOMEM      -3  /* device or host allocation failed   (matrix.cpp:1662-1670)   */
#define DGRM_E_BADBED     -4  /* bad magic / not SNP-major / short read (genotype.cpp:578-605, 693-696) */
#define DGRM_E_BADSNP     -5  /* SNP with standardDev == 0          (genotype.cpp:924-939)   */
#define DGRM_E_NANSNP     -6  /* SNP with no non-missing genotype (p1 = 0/0; the reference lets the NaN through, genotype.cpp:736) */
#define DGRM_E_IO         -7  /* file could not be opened / written                          */

/* operand format of the standardised genotypes fed to the tensor cores */
#define DGRM_OPERAND_FP16    0  /* one pass, fp16 operands (default; |z| <= ~45 fits easily) */
#define DGRM_OPERAND_BF16    1  /* one pass, bf16 operands                                   */
#define DGRM_OPERAND_FP16X2  2  /* hi+lo fp16 split, three passes (hi.hi + lo.hi + hi.lo)    */

/* how the per-pair non-missing count N (kernel.cpp:96-101) is formed */
#define DGRM_COUNT_MMA   0  /* 0/1 e4m3 operands through tcgen05 kind::f8f6f4, exact in fp32  */
#define DGRM_COUNT_POPC  1  /* AND + popcount over bit-packed non-missing masks               */
#define DGRM_COUNT_FP4   2  /* 0/1 e2m1 operands through block-scaled tcgen05 kind::mxf4 (all scales 1.0), exact in fp32 */

/* transport of streamed A values: fp64 as they are (default), or rounded to fp32 on the device and widened back by
 * the library's host threads; the diagonal always arrives as the exact fp64 value.  fp32 adds a relative error
 * <= 2^-24 = 6e-8 to A_ij (three orders below the operand tolerance) and halves the bytes of the largest stream. */
#define DGRM_TRANSPORT_FP64 0
#define DGRM_TRANSPORT_FP32 1
#define DGRM_TRANSPORT_NARROW 2   /* pair counts only */

/* which matrix dgrm_fetch_* returns */
#define DGRM_FETCH_A  0  /* normalised GRM  A = S * (1/N)     (matrix.cpp:4096)               */
#define DGRM_FETCH_S  1  /* unnormalised    S = Z^T Z         (kernel.cpp:94)                 */
#define DGRM_FETCH_N  2  /* pair counts     N = M^T M         (kernel.cpp:100) or m if flat   */

/* kernel classes reported by dgrm_kernel_times() */
#define DGRM_K_COUNT    0  /* bed_count_stats                                                 */
#define DGRM_K_EXPAND   1  /* bed_expand (+ diag partial reduce)                              */
#define DGRM_K_SYRK     2  /* grm_syrk tcgen05, 16-bit operands                               */
#define DGRM_K_PAIRCNT  3  /* pair count (fp8 tcgen05 or popcount)                            */
#define DGRM_K_FINAL    4  /* diagonal fix-up                                                 */
#define DGRM_K_FETCH    5  /* tile -> block-cyclic gather                                     */
#define DGRM_K_GATHER   6  /* row / individual gather into the aligned chunk buffer           */
#define DGRM_K_D2H      7  /* copy-engine transfers of streamed (bound) outputs, not a kernel  */
#define DGRM_K_NUM      8

typedef struct dgrm_engine dgrm_engine;

typedef struct dgrm_config {
  int32_t device;        /* CUDA device ordinal                                              */
  int32_t rank;          /* this engine's index in the tile-set partition                    */
  int32_t world;         /* number of engines sharing the lower triangle (1 = all tiles)     */
  int32_t operand;       /* DGRM_OPERAND_*                                                   */
  int32_t count_method;  /* DGRM_COUNT_*                                                     */
  int32_t flat_norm;     /* --grm-flat-norm: N := m, no pair count (kernel.cpp:102-109)      */
  int32_t chunk_snps;    /* SNPs contracted per launch (fp32 in-chunk, fp64 across); mult. of 256 */
  int32_t cta_group;     /* 1 = one-SM MMA 128x256 tiles, 2 = CTA-pair MMA 256x256 tiles     */
  int32_t profile;       /* 1 = record CUDA events around every kernel launch                */
  int32_t resident_chunks; /* operand slots: SNP chunks whose tensor-core operands stay in HBM at once (0 = auto) */
  int32_t resident_mb;   /* memory budget (MiB) for resident operands when resident_chunks == 0     */
  int32_t reserve_sms;   /* SMs the persistent contraction leaves free for the caller's own kernels (e.g. NCCL
                            all-gather CTAs, which cannot co-reside with it); -1 or 0 = none.  Measured at 2 GPUs:
                            no net gain (the contraction loses what the exchange wins), so the default is 0 */
  int32_t a_transport;   /* DGRM_TRANSPORT_*: how STREAMED values of A (dgrm_bind_output / dgrm_bind_tiles) cross
                            PCIe; the caller's buffer always receives doubles                                 */
  int32_t n_transport;   /* how STREAMED pair counts cross PCIe: DGRM_TRANSPORT_NARROW (default: 16-bit deficiencies
                            m - N or int32, widened to double by the library's host threads) or DGRM_TRANSPORT_FP64
                            (doubles straight into a pinned caller buffer by the copy engine: 4x the PCIe bytes, no
                            host CPU work; the better choice when an engine has only a few host cores to itself)     */
  int32_t reserved[2];
} dgrm_config;

/* Fill *cfg with defaults (device 0, rank 0/1, fp16 operands, fp4 pair count, chunk 16384, cta_group 2). */
int dgrm_default_config(dgrm_config *cfg);

/* Engine life cycle.  Replaces nothing in the reference; owns device buffers and streams. */
int dgrm_create(const dgrm_config *cfg, dgrm_engine **out);
void dgrm_destroy(dgrm_engine *e);

/* Text of the last error on this engine (or of the last failed dgrm_create when e == NULL).
 * Stands in for the message DISSECT passes to misc.error (misc.cpp:151). */
const char *dgrm_last_error(const dgrm_engine *e);

int dgrm_abi_version(void);

/* Number of CUDA devices visible to the process (0 when there is none or the runtime fails). */
int dgrm_device_count(void);

/* Cohort geometry.  n_fam = individuals in the .fam; keep_mask[n_fam] != 0 keeps individual f
 * (individualBEDIdxsMaskToKeep, genotype.cpp:750); NULL keeps everyone. */
int dgrm_set_individuals(dgrm_engine *e, int64_t n_fam, const int32_t *keep_mask);

/* Start a GRM: clears S, N and the per-SNP statistics.  m_capacity >= number of SNPs that
 * will be kept over all following pushes. */
int dgrm_begin(dgrm_engine *e, int64_t m_capacity);

/* Feed SNP-major .bed rows (header stripped): row r starts at rows + r*pitch and holds
 * ceil(n_fam/4) bytes, individual f in byte f/4 bits 2(f%4)..2(f%4)+1 (genotype.cpp:662, 752, 784).
 * row_keep[nrows] != 0 keeps a row (SNPsBEDIdxsToKeep, genotype.cpp:660-690); NULL keeps all.
 * Replaces readBEDFile's decode loop + parseSNP/parseSNPbyte (genotype.cpp:639-787),
 * normalizeGenotypes (genotype.cpp:888-961) and both pdsyrk_ calls (kernel.cpp:94,100) for
 * these rows.  The host variant copies host->device itself (asynchronously when `rows` is
 * pinned: a pinned `rows` buffer must then stay unchanged until dgrm_finish / dgrm_sync returns, or until a
 * stream that dgrm_release_inputs has ordered after the engine's reads gets there); the device variant reads
 * rows already in HBM once `producer_stream` (a cudaStream_t; NULL = the legacy default stream) reaches the
 * point of the call. */
int dgrm_push_bed_host(dgrm_engine *e, const uint8_t *rows, int64_t pitch, int64_t nrows,
                       const uint8_t *row_keep);
int dgrm_push_bed_device(dgrm_engine *e, const uint8_t *d_rows, int64_t pitch, int64_t nrows,
                         const uint8_t *row_keep, void *producer_stream);

/* Feed real-valued genotype rows (expected allele counts, "dosages"): the rows the reference's BGEN reader forms per
 * variant, d_i = P_i(AB) + 2 P_i(BB), with missing[r][i] != 0 where sample i has no probabilities
 * (ProbabilitiesSetter::set_value, genotypebgen.cpp:183-236).  Row r starts at dosage + r*pitch (elements) and holds
 * the KEPT individuals only (dgrm_num_individuals of them; the reader already skips the others,
 * genotypebgen.cpp:139-150); missing may be NULL (nothing missing), its rows start at missing + r*missing_pitch.
 * Replaces ProbabilitiesSetter::finalise (mean over the present samples, std with ALL samples in the divisor,
 * z = (d - mean)/std, 0 where missing; genotypebgen.cpp:239-282), parseBGENSNP's SNP record (p2 = mean/2,
 * p1 = 1 - mean/2, standardDev = std, nNonMissing = -1, zero frequencies; genotypebgen.cpp:723-736; read them with
 * dgrm_snp_stats) and both pdsyrk_ calls for these rows.  Rows whose std is 0 contribute nothing to S or N, which is
 * what removeSNPsWithZeroVariance (genotypebgen.cpp:739-779) achieves by dropping them; the adapter lists them
 * (std_dev == 0) in <out>.snps.novariance.  The rows are copied before the call returns.  Packed and real-valued
 * rows may alternate within one GRM (--bfile-list style joins); each switch ends the open chunk.
 * Parsing the BGEN container itself (the un-vendored `genfile` library in the reference) stays with the caller. */
int dgrm_push_dosage_host(dgrm_engine *e, const double *dosage, int64_t pitch, const uint8_t *missing,
                          int64_t missing_pitch, int64_t nrows);

/* The same rows one step earlier: the probability data blocks of BGEN variants as they come out of the container
 * (after zlib inflation), unpacked on the device -- 2B/8 + 1 bytes per genotype cross the bus instead of 9 and the host
 * does no per-sample work.  Block r is data[offsets[r] .. offsets[r+1]):
 *   layout 2: N u32 | K u16 | Pmin u8 | Pmax u8 | N ploidy/missingness bytes | phased u8 | B u8 | 2N B-bit probabilities
 *   layout 1: N x 3 u16 probabilities (all zero = missing)
 * i.e. what genfile::bgen::read_and_parse_genotype_data_block is handed (genotypebgen.cpp:499-518).  n_samples is the
 * sample count of the file, sample_keep[n_samples] != 0 the samples to keep (individualBGENIdxsMaskToKeep,
 * genotypebgen.cpp:367-379; NULL = all); as many must be kept as dgrm_num_individuals.  The conditions the reference's
 * setter refuses (non-biallelic, non-diploid, phased; genotypebgen.cpp:99-172) fail here with the same messages (the
 * per-sample ploidy check is made on the device and reported by dgrm_finish).  Dosages are formed exactly as
 * ProbabilitiesSetter::set_value does from genfile's doubles: d = P(AB) + 2 P(BB), P = numerator / (2^B - 1)
 * (layout 1: / 32768), P(BB) = 1 - (P(AA) + P(AB)) in layout 2; everything after that is dgrm_push_dosage_host.
 * The blocks are copied before the call returns. */
int dgrm_push_bgen_blocks_host(dgrm_engine *e, const uint8_t *data, const int64_t *offsets, int64_t nrows,
                               int layout, int64_t n_samples, const int32_t *sample_keep);

/* Make `consumer_stream` (a cudaStream_t; NULL = legacy default stream) wait until the engine has read every
 * buffer handed to dgrm_push_bed_device / dgrm_push_bed_host so far, so that the caller may overwrite those
 * buffers (e.g. with the next NCCL all-gather) without a host synchronisation. */
int dgrm_release_inputs(dgrm_engine *e, void *consumer_stream);

/* Whole-file convenience: validates the 3-byte header exactly as genotype.cpp:578-605 and
 * streams the kept rows.  snp_keep_idx[m_keep] are ascending .bim positions, NULL = all
 * m_file rows (SNPsBEDIdxsToKeep). */
int dgrm_load_bed_file(dgrm_engine *e, const char *bed_path, int64_t m_file,
                       const int64_t *snp_keep_idx, int64_t m_keep);

/* The 3-byte header checks of readBEDFile alone (genotype.cpp:578-605), with the reference's messages: for hosts that
 * read the file themselves (one reader feeding several engines). */
int dgrm_validate_bed_header(dgrm_engine *e, const char *bed_path);

/* Flush the last partial chunk, finish S and N.  Fails with DGRM_E_BADSNP / DGRM_E_NANSNP
 * when a kept SNP is monomorphic / entirely missing (genotype.cpp:924-939); the offending
 * SNP indices are then available from dgrm_bad_snps so the adapter can write <out>.badsnps. */
int dgrm_finish(dgrm_engine *e);
int dgrm_sync(dgrm_engine *e);

int64_t dgrm_num_snps(const dgrm_engine *e);         /* SNPs pushed and kept so far          */
int64_t dgrm_num_individuals(const dgrm_engine *e);  /* individuals kept = dimension of the output matrices (LD mode: SNPs of the group) */

/* What `struct SNP` holds on the root rank after readBEDFile/gatherSNPData
 * (genotype.h:65-69, genotype.cpp:735-738, 859-870): bit-exact counts, 0-ulp doubles.
 * freq is [m][4] with freq[k][0] = missing, [1] = hom A1, [2] = het, [3] = hom A2.
 * Any pointer may be NULL. */
int dgrm_snp_stats(dgrm_engine *e, int32_t *freq, int32_t *n_non_missing, double *p1,
                   double *p2, double *std_dev);
int dgrm_bad_snps(dgrm_engine *e, int64_t *idx, int64_t cap, int64_t *count);

/* Fill the caller's local piece of a 2-D block-cyclic n x n Matrix (matrix.cpp:435-487,
 * 1748-1758; grid R x C row-major, source process (0,0)): global (gr,gc) owned by
 * ((gr/bs_r)%R, (gc/bs_c)%C) lands at local[lc*lld + lr].  uplo 'L' writes only gc <= gr
 * (what pdsyrk_('L',..) leaves defined, matrix.cpp:2679-2682), 'U' only gc >= gr, 'B' both.
 * With world > 1 only entries of tiles this engine owns are written (everything else in `local` is left as it
 * was), so the engines of a partition can fill one buffer in turn, or each fill a zeroed panel that the caller
 * then sums over the engines (every entry has exactly one owner: x + 0 + ... + 0 is exact).  `local` may be
 * pageable or pinned host memory, or DEVICE memory of this engine's GPU: the multi-GPU egress reduces such
 * device panels over NVLink to the engine whose host rank owns the block. */
int dgrm_fetch_block_cyclic(dgrm_engine *e, int which, double *local, int64_t lld, int bs_r,
                            int bs_c, int my_row, int my_col, int grid_rows, int grid_cols,
                            char uplo);
/* Same, restricted to local columns [lc_begin, lc_end) of that block; `local` points at column lc_begin.
 * Lets a caller move a large block through a bounded panel (the reference moves one bs x bs block at a time,
 * matrix.cpp:2297-2341). */
int dgrm_fetch_block_cyclic_cols(dgrm_engine *e, int which, double *local, int64_t lld, int bs_r,
                                 int bs_c, int my_row, int my_col, int grid_rows, int grid_cols,
                                 char uplo, int64_t lc_begin, int64_t lc_end);

/* Bind a caller-owned local block-cyclic buffer (same arguments as dgrm_fetch_block_cyclic) that
 * dgrm_finish fills with matrix `which`.  When the buffer is pinned (dgrm_host_alloc) the tiles are
 * stored into it group by group while later tile groups are still being contracted, so the
 * device->host traffic (copy engine) overlaps the tensor-core work; this streaming applies to pinned
 * uplo 'L' bindings and may additionally deposit the symmetric value above the diagonal inside the
 * 256-row diagonal blocks (the reference leaves that half uninitialised, matrix.cpp:1646, 2682).
 * Pageable buffers and uplo 'U'/'B' are filled at the end of dgrm_finish.  One binding per `which`; bindings survive dgrm_begin and are dropped by
 * dgrm_set_individuals / dgrm_unbind_outputs.  Replaces the result hand-over of Kernel::Kernel
 * (kernel.cpp:94-116: kernel->m and N->m are the outputs). */
int dgrm_bind_output(dgrm_engine *e, int which, double *local, int64_t lld, int bs_r, int bs_c,
                     int my_row, int my_col, int grid_rows, int grid_cols, char uplo);
int dgrm_unbind_outputs(dgrm_engine *e);

/* Multi-engine counterpart of dgrm_bind_output: bind a caller buffer [dgrm_num_tiles][256][128] that dgrm_finish
 * fills with this engine's tiles of matrix `which` (same layout as dgrm_fetch_tiles), tile group by tile group
 * while later groups are still being contracted.  One binding per `which`; dropped by dgrm_unbind_outputs. */
int dgrm_bind_tiles(dgrm_engine *e, int which, double *tiles);

/* Pairs (i > j) whose shared non-missing count satisfies -0.1 < N_ij < min_overlap - 0.1,
 * the scan sanitizeKernel performs (kernel.cpp:1993-2037, matrix.cpp:4334-4401).
 * pairs is [cap][2] (row, col), *count receives the total found (may exceed cap). */
int dgrm_low_overlap_pairs(dgrm_engine *e, double min_overlap, int64_t *pairs, int64_t cap,
                           int64_t *count);

/* Restrict every following fetch / write to the individuals idx[0..count) (ascending positions among
 * the kept individuals), i.e. what Kernel::filterIndividuals leaves after sanitizeKernel removed
 * individuals (kernel.cpp:2029-2033, 1378-1460): the matrices themselves are not recomputed.
 * count == 0 or idx == NULL restores the full set. */
int dgrm_set_output_subset(dgrm_engine *e, const int64_t *idx, int64_t count);

/* Write <path> = the .grm.dat of Kernel::writeGRM (kernel.cpp:944-974): 14-byte header then
 * the packed (n+1) x n column-major array of matrix.cpp:2324-2334.  world must be 1. */
int dgrm_write_grm_dat(dgrm_engine *e, const char *path);

/* Multi-engine form of the writer (the reference funnels every block through rank 0: packMatrices
 * matrix.cpp:2262-2349 + writeMatrixMPI :2168-2219).  Every engine of a partition writes, at their offsets in
 * the same file, exactly the entries whose tiles it owns (per file column a few contiguous runs), so an
 * 80 GB result needs no gather.  Exactly one engine calls it with create != 0 FIRST (creates / truncates the
 * file, writes the header, sizes the file); the others call it with create == 0 after that call has returned
 * (the caller's barrier).  With world == 1 and create != 0 it is dgrm_write_grm_dat. */
int dgrm_write_grm_dat_part(dgrm_engine *e, const char *path, int create);

/* Raw tile export for multi-engine runs: this engine's 128x256 tiles, column-major inside a
 * tile.  tile_rc[ntiles][2] = (row block of 128, column panel of 256). */
int64_t dgrm_num_tiles(const dgrm_engine *e);
int dgrm_tile_coords(const dgrm_engine *e, int32_t *tile_rc);
int dgrm_fetch_tiles(dgrm_engine *e, int which, double *tiles /* [ntiles][128*256] */);
/* tiles [tile0, tile0 + count) of the same export (a 100k-individual tile set is tens of GB: move it in pieces) */
int dgrm_fetch_tiles_range(dgrm_engine *e, int which, double *tiles, int64_t tile0, int64_t count);

/* ---- resident cohort: the packed .bed stays in HBM (2 bits per genotype: 12.5 GB at 100k x 500k) ----------
 * The reference loads a Genotype once and then builds many kernels from it: one per SNP group
 * (Analysis::makeGRM regional loop, analysis.cpp:67-110; SingleREML::computeRegional, singlereml.cpp:293-298),
 * each time re-filtering the fp64 block-cyclic genotype matrix (genotypeOfSNPsGroup -> filterSNPsAndIndividuals,
 * genotype.cpp:1596-1621, 972).  Here the stored rows are simply pushed again with another keep mask.
 * dgrm_store_bed_file validates the header like dgrm_load_bed_file and appends all m_file rows;
 * dgrm_store_bed_host appends rows from memory; dgrm_store_clear empties the store (dgrm_set_individuals keeps it
 * while n_fam is unchanged).  dgrm_push_stored feeds stored rows [first, first + nrows) to the GRM begun by
 * dgrm_begin, row_keep[nrows] as in dgrm_push_bed_host. */
int dgrm_store_bed_file(dgrm_engine *e, const char *bed_path, int64_t m_file);
int dgrm_store_bed_host(dgrm_engine *e, const uint8_t *rows, int64_t pitch, int64_t nrows);
/* same, from rows already in this GPU's memory, readable once `producer_stream` reaches the point of the call */
int dgrm_store_bed_device(dgrm_engine *e, const uint8_t *d_rows, int64_t pitch, int64_t nrows, void *producer_stream);
int dgrm_store_clear(dgrm_engine *e);
int64_t dgrm_stored_snps(const dgrm_engine *e);
int dgrm_push_stored(dgrm_engine *e, int64_t first, int64_t nrows, const uint8_t *row_keep);

/* The standardised genotypes of the GRM's SNPs, what Genotype::normalizeGenotypes leaves in
 * genotype->genotypes (genotype.cpp:888-961: g -= (1 + 2 p2) for non-missing, then g /= standardDev; missing
 * stays 0) for callers that keep the Genotype (returnGenotype, auxiliar.cpp:606-613): an m x n matrix
 * (row = SNP, column = kept individual) in the caller's block-cyclic layout, 0-ulp.  Needs every SNP of the
 * current GRM to have come from the resident store (dgrm_push_stored). */
int dgrm_fetch_standardised(dgrm_engine *e, double *local, int64_t lld, int bs_r, int bs_c, int my_row,
                            int my_col, int grid_rows, int grid_cols);

/* ---- (S, N) algebra between two GRMs of the same individuals (Kernel::addKernels, kernel.cpp:1705-1859) ----
 * The reference denormalises both kernels (S = A o N), adds or subtracts S and N (pdgeadd_) and divides again;
 * LOCO is `globalGRM->addKernels(1., grm, -1., regionalGRM)` (singlereml.cpp:298).  The engine keeps a second
 * tile store: dgrm_save copies (DGRM_SAVE_COPY) or accumulates (DGRM_SAVE_ADD, sum of disjoint SNP groups) the
 * finished GRM's unnormalised S, integer N and SNP count into it, and dgrm_set_view selects what every fetch /
 * bind / write / scan sees from then on: the current GRM, the saved one, saved - current (LOCO) or
 * saved + current.  S and N are combined exactly (fp64 / int32) before the one division. */
#define DGRM_SAVE_COPY 0
#define DGRM_SAVE_ADD 1
#define DGRM_SAVE_DROP 2
#define DGRM_VIEW_CURRENT 0
#define DGRM_VIEW_SAVED 1
#define DGRM_VIEW_SAVED_MINUS_CURRENT 2
#define DGRM_VIEW_SAVED_PLUS_CURRENT 3
int dgrm_save(dgrm_engine *e, int mode);
int dgrm_set_view(dgrm_engine *e, int view);

/* ---- LD mode: SNP x SNP correlations of a group of SNPs ---------------------------------
 * Replaces the contraction of GWAS::getLessSignificantCorrelatedSNPs (gwas.cpp:1156-1172; igwas.cpp:979-995):
 *     correlations->multiply(genotypes, 'N', genotypes, 'T')      G G^T     (pdsyrk_('L','N'), matrix.cpp:2677-2683)
 *     normalization->multiply(missings, 'N', missings, 'T')       M M^T
 *     correlations->elementWiseDivision(normalization)            matrix.cpp:4091-4098
 *     correlations->getGlobalIndexOutsideRange(-t, t, ...)        matrix.cpp:4265-4330
 * on the standardised genotypes (gwas.cpp:149) of the m_group SNPs pushed after dgrm_ld_begin with the ordinary
 * dgrm_push_bed_host / dgrm_push_bed_device calls (SNP rows in group order; row_keep as usual).  Same kernels as
 * the GRM with the roles of SNPs and individuals exchanged: the contraction runs over the kept individuals.
 * After dgrm_finish every reader works on the m_group x m_group matrices: dgrm_fetch_block_cyclic(DGRM_FETCH_A) is
 * `correlations`, DGRM_FETCH_N the shared non-missing counts, dgrm_snp_stats the group's struct SNP values, and
 * dgrm_ld_pairs lists the strictly-lower (row, col) pairs with value < -threshold or > threshold, the pairs
 * gwas.cpp:1191 acts on (pairs is [cap][2]; *count receives the total found, which may exceed cap).
 * dgrm_set_individuals returns the engine to GRM mode. */
int dgrm_ld_begin(dgrm_engine *e, int64_t n_fam, const int32_t *keep_mask, int64_t m_group);
int dgrm_ld_pairs(dgrm_engine *e, double threshold, int64_t *pairs, int64_t cap, int64_t *count);

/* Bytes the copy engine moved device -> host for the streamed (bound) outputs of the last dgrm_finish. */
int64_t dgrm_last_d2h_bytes(const dgrm_engine *e);

/* Pinned host memory helpers (Matrix::m replacement allocation for zero-staging fetches). */
void *dgrm_host_alloc(int64_t bytes);
void dgrm_host_free(void *p);

/* Per-kernel-class device time (ms, CUDA events on the launching stream; needs cfg.profile = 1) and
 * launch counts accumulated since the previous call (the call resets them). */
int dgrm_kernel_times(dgrm_engine *e, double ms[DGRM_K_NUM], int64_t launches[DGRM_K_NUM]);

#ifdef __cplusplus
}
#endif
#endif /* DGRM_H */
