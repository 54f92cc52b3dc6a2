// libdgrm: C-ABI engine around the sm_100a kernels (see include/dgrm.h for the contract and the
// reference file:line each entry point replaces).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/dgrm.h"
#include "bed_kernels.cuh"
#include "dosage_kernels.cuh"
#include "finalize_kernels.cuh"
#include "paircount_popc.cuh"
#include "syrk_tcgen05.cuh"

using namespace dgrm;

namespace {

thread_local std::string g_create_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct EventPair {
  cudaEvent_t a, b;
};

}  // namespace

struct dgrm_engine {
  dgrm_config cfg;
  std::string err;
  int sm_count = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  EncodeTiledFn encode = nullptr;

  // cohort geometry
  int64_t n_fam = 0, n = 0, n_pad = 0;   // n = dimension of the output matrices (individuals; SNPs of the group in LD mode)
  int64_t n_row = 0;        // individuals per compacted packed row (== n except in LD mode)
  bool ld_mode = false;     // dgrm_ld_begin: SNP x SNP correlations, the contraction runs over the individuals
  std::vector<int32_t> ld_keep;  // keep mask of the current LD geometry (to recognise an unchanged dgrm_ld_begin)
  int64_t b_fam = 0;        // bytes per raw .bed row
  int64_t b_row = 0;        // bytes per compacted row
  int64_t chunk_pitch = 0;  // device chunk row pitch (multiple of 128)
  int32_t* d_src_ind = nullptr;

  // tile store
  int nrb = 0, ncp = 0;
  std::vector<int4> work;       // launch work list (CG form)
  std::vector<int4> work_cg1;   // 128x256 form (popcount kernel)
  std::vector<int32_t> tile_rc; // [ntiles][2]
  std::vector<int32_t> h_index; // host copy of d_index: [nrb][ncp] -> storage tile or -1 (not owned)
  std::vector<int32_t> h_subset; // host copy of d_subset
  std::vector<int32_t> need_blocks;  // 128-individual blocks that this engine's tiles touch (ascending)
  int32_t* d_need_blocks = nullptr;
  int4* d_work = nullptr;
  int4* d_work_cg1 = nullptr;
  int32_t* d_index = nullptr;
  int2* d_tile_rc = nullptr;   // per storage tile: first global row / column
  int64_t ntiles = 0;
  double* d_S = nullptr;
  int32_t* d_N = nullptr;
  double* d_diag = nullptr;
  int32_t* d_subset = nullptr;  // output subset (sanitizeKernel's filterIndividuals)
  int64_t n_out = 0;

  // output bindings (filled while dgrm_finish runs, tile group by tile group)
  struct Binding {
    int which = -1;
    double* host = nullptr;     // caller pointer
    double* dev = nullptr;      // device-visible alias when the buffer is pinned (zero-copy), else NULL
    CyclicGeom g{};
    int uplo = 0;
  };
  Binding bind[3];
  int nbind = 0;
  struct TileBinding {
    int which = -1;
    double* host = nullptr;
    bool pinned = false;
  };
  TileBinding tbind[3];
  int ntbind = 0;
  cudaStream_t out_stream = nullptr;   // packs finished row groups into staging
  cudaStream_t d2h_stream = nullptr;   // copy engine: staging -> caller buffer
  std::vector<cudaEvent_t> group_ev;
  static constexpr int kOutSlots = 3;
  void* d_out_stage[kOutSlots] = {nullptr, nullptr, nullptr};
  void* h_out_stage[kOutSlots] = {nullptr, nullptr, nullptr};  // pinned; staged jobs are finished by host threads
  size_t out_stage_bytes = 0;
  cudaEvent_t stage_packed[kOutSlots] = {nullptr, nullptr, nullptr}, stage_copied_out[kOutSlots] = {nullptr, nullptr, nullptr};
  struct OutJob {
    int b;                  // binding
    int row0, row1;         // global rows of the raster row group
    int lr0, lr1, lc1;      // local rectangle rows [lr0, lr1) x cols [0, lc1)
    int ev;                 // tile-group event after which the rectangle is final
    bool direct;            // copy engine writes straight into the caller's (pinned) buffer
    int ntype;              // transport type: 0 double; 1 int32 / 2 uint16 deficiency (pair counts only), 3 float (A, cfg.a_transport); 1-3 are widened by the host
    bool tiles;             // tile-range job (dgrm_bind_tiles): b indexes tbind, elements [t0, t1) of the tile store
    int64_t t0, t1;
  };
  std::vector<OutJob> jobs;
  size_t jobs_issued = 0, jobs_done = 0;
  int64_t d2h_bytes = 0;      // bytes the copy engine moved for the streamed outputs of the last dgrm_finish
  double* d_diag_a = nullptr; // fp32 transport: exact fp64 A_ii of the visible individuals, delivered after the last job
  double* h_diag_a = nullptr;
  std::vector<int> rg_end;    // per raster row group: end index in `work` of its last owned supertile
  std::vector<int> rg_row1;   // per raster row group: one past its last global row

  // resident operand slots: `slots` chunks of kc_max SNPs live side by side in Z / Q / MB
  int slots = 0;
  int chunks_in_super = 0;
  int launched_s = 0, launched_n = 0;  // resident chunks of this super already contracted into S / N
  // streaming outputs: the first tile groups (work items [0, pre_end)) are contracted chunk by chunk while the rows
  // are still arriving, so that the tensor cores are not idle during the host-to-device phase of an e2e build
  int pre_end = -1;                    // -1 = not decided yet for this GRM, 0 = no prefix
  int pre_s = 0, pre_n = 0;            // resident chunks already contracted over the prefix
  int prefix_pct = 30;                 // DGRM_PREFIX_PCT
  int64_t supers_done = 0;
  int slot_kpad[kMaxChunks] = {0};
  int l2_policy = 3;  // see SyrkParams::l2_policy (DGRM_L2_POLICY overrides, for experiments)

  // chunk state
  int kc_max = 0;
  uint8_t* d_chunk = nullptr;
  int fill = 0;
  uint16_t* d_Z = nullptr;
  int64_t z_pitch = 0;
  uint8_t* d_Q = nullptr;
  int64_t q_pitch = 0;
  uint8_t* d_Q4 = nullptr;
  int64_t q4_pitch = 0;
  uint32_t* d_MB = nullptr;
  int64_t mb_pitch = 0;
  double* d_zt = nullptr;
  double* d_diag_part = nullptr;
  CUtensorMap tmap_z, tmap_q, tmap_q4;
  int64_t chunks_done = 0;

  // per-SNP statistics
  int64_t m_cap = 0, m = 0;
  int32_t* d_freq = nullptr;
  int32_t* d_nn = nullptr;
  double *d_p1 = nullptr, *d_p2 = nullptr, *d_sd = nullptr;
  int32_t* d_bad_count = nullptr;
  int64_t* d_bad_list = nullptr;
  static constexpr int kBadCap = 65536;

  // host -> device staging (double buffered)
  uint8_t* d_stage[2] = {nullptr, nullptr};
  int64_t stage_pitch = 0;
  int stage_rows = 0;
  cudaEvent_t stage_copied[2] = {nullptr, nullptr}, stage_free[2] = {nullptr, nullptr};
  int stage_i = 0;
  int32_t* h_rowidx[2] = {nullptr, nullptr};
  int32_t* d_rowidx[2] = {nullptr, nullptr};
  cudaEvent_t rowidx_free[2] = {nullptr, nullptr};
  int rowidx_i = 0;
  cudaEvent_t producer_ev = nullptr;
  // `prep` decodes (gather, count, expand) while `stream` contracts the chunks that are already resident
  cudaStream_t prep = nullptr;
  cudaEvent_t prep_ev = nullptr, mma_ev = nullptr;
  cudaEvent_t slot_ev[kMaxChunks] = {nullptr};  // recorded on `stream` once every contraction that reads the slot is queued
  int slots_evented = 0;
  double out_wait_ms = 0, out_host_ms = 0;  // DGRM_TRACE_OUT: host time waiting for / finishing output jobs
  int reserve_sms = 0;
  int32_t* d_miss = nullptr;   // per-individual missing genotypes of the current GRM (only while streaming outputs are bound)
  bool miss_valid = false;
  bool n_u16 = false;          // streamed pair counts travel as 16-bit deficiencies m - N
  int overlap = 1;  // DGRM_OVERLAP=0: contract only in dgrm_finish (experiments)
  int fused_decode = 1;  // DGRM_FUSED_DECODE=0: separate gather and count kernels (A/B measurements)
  int expand_k_fast = 0, expand_stream_stores = 0;   // DGRM_EXPAND_KFAST / DGRM_EXPAND_CS (A/B measurements)
  int k1_transposed = 1; // DGRM_K1_TRANSPOSED=0: row-major chunk + bed_expand_kernel (round-1 decode path, A/B measurements)
  uint32_t* d_chunkT = nullptr;   // transposed chunk: [kc_max / 16][t_pitch] words (GRM mode)
  int64_t t_pitch = 0;
  int32_t* d_counts = nullptr;    // [kc_max][4] per-row pattern counts being accumulated
  int32_t* d_miss_part = nullptr; // [kc_max / 128][n_pad]

  // real-valued genotype rows (dgrm_push_dosage_host): pinned host staging (transfer pieces) -> device staging (up to a
  // whole chunk of rows, so that one statistics launch carries many rows) -> operands
  double* h_dos[2] = {nullptr, nullptr};
  uint8_t* h_dmiss[2] = {nullptr, nullptr};
  int h_cap = 0, h_fill = 0, dos_i = 0;   // rows per pinned buffer, rows staged in the current one
  bool h_any_missing = false;             // a row staged in the current pinned buffer came with a missing mask
  double* d_dos = nullptr;
  uint8_t* d_dmiss = nullptr;
  int dev_cap = 0, dev_fill = 0;          // rows of the device staging buffer, rows copied (or being copied) into it
  bool dev_any_missing = false;
  double *d_dos_mean = nullptr, *d_dos_sd = nullptr;
  cudaEvent_t dos_copied[2] = {nullptr, nullptr}, dos_ready = nullptr, dos_read = nullptr;
  int dosage_stage_mb = 2048;             // DGRM_DOSAGE_STAGE_MB: size of the device staging buffer
  // BGEN probability blocks (dgrm_push_bgen_blocks_host): pinned raw pieces -> device raw buffer -> bgen_unpack_kernel -> d_dos
  uint8_t* h_raw[2] = {nullptr, nullptr};
  int64_t* h_boff[2] = {nullptr, nullptr};
  uint8_t* d_raw = nullptr;
  int64_t* d_boff = nullptr;
  int64_t raw_cap = 0, raw_fill = 0;      // bytes per raw buffer, bytes staged in the current pinned one
  int raw_rows_cap = 0, raw_rows = 0, raw_i = 0;
  int raw_layout = 0;
  int64_t raw_samples = 0;                // samples per block of the file being pushed
  cudaEvent_t raw_copied[2] = {nullptr, nullptr}, raw_read = nullptr;
  std::vector<int32_t> bgen_keep;         // the sample mask the device column list was built from
  int32_t* d_bgen_cols = nullptr;
  bool bgen_all = true;                   // every sample of the file is kept (no column list)
  bool fill_dosage = false;               // the `fill` rows of the open chunk are already expanded (dosage rows)

  bool have_ind = false, begun = false, finished = false;

  // resident packed cohort (dgrm_store_*): SNP-major rows of ceil(n_fam/4) bytes at pitch bed_pitch
  uint8_t* d_bed = nullptr;
  int64_t bed_pitch = 0, bed_rows = 0, bed_cap = 0, bed_nfam = 0;
  cudaEvent_t bed_ev = nullptr;           // last upload into the store
  std::vector<int32_t> grm_rows;          // stored row of every SNP of the current GRM, in GRM order
  bool grm_from_store = true;             // false once a SNP of the current GRM came from somewhere else
  int32_t* d_grm_rows = nullptr;
  int64_t d_grm_rows_cap = 0;

  // saved (S, N) store and the view every reader sees (dgrm_save / dgrm_set_view)
  double* d_S2 = nullptr;
  int32_t* d_N2 = nullptr;
  int64_t m2 = 0;
  bool saved = false, saved_flat = false;
  int view = DGRM_VIEW_CURRENT;

  // profiling
  std::vector<EventPair> ev[DGRM_K_NUM];
  std::vector<EventPair> ev_pool;
  double ms_acc[DGRM_K_NUM] = {0};
  int64_t launches[DGRM_K_NUM] = {0};
};

namespace {

int fail(dgrm_engine* e, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t _st = (call);                                                                            \
    if (_st != cudaSuccess)                                                                              \
      return fail(e, _st == cudaErrorMemoryAllocation ? DGRM_E_NOMEM : DGRM_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                  cudaGetErrorString(_st), __FILE__, __LINE__);                                          \
  } while (0)

struct Prof {
  dgrm_engine* e;
  int k;
  cudaStream_t st;
  EventPair p{nullptr, nullptr};
  Prof(dgrm_engine* e_, int k_, cudaStream_t st_ = nullptr) : e(e_), k(k_), st(st_ ? st_ : e_->stream) {
    e->launches[k]++;
    if (!e->cfg.profile) return;
    if (!e->ev_pool.empty()) {
      p = e->ev_pool.back();
      e->ev_pool.pop_back();
    } else {
      cudaEventCreate(&p.a);
      cudaEventCreate(&p.b);
    }
    cudaEventRecord(p.a, st);
  }
  ~Prof() {
    if (!e->cfg.profile) return;
    cudaEventRecord(p.b, st);
    e->ev[k].push_back(p);
  }
};

void free_dev(void* p) {
  if (p) cudaFree(p);
}

void release_dosage_staging(dgrm_engine* e) {
  for (int i = 0; i < 2; ++i) {
    if (e->h_dos[i]) cudaFreeHost(e->h_dos[i]);
    if (e->h_dmiss[i]) cudaFreeHost(e->h_dmiss[i]);
    e->h_dos[i] = nullptr; e->h_dmiss[i] = nullptr;
  }
  free_dev(e->d_dos); e->d_dos = nullptr;
  free_dev(e->d_dmiss); e->d_dmiss = nullptr;
  free_dev(e->d_dos_mean); e->d_dos_mean = nullptr;
  free_dev(e->d_dos_sd); e->d_dos_sd = nullptr;
  e->h_cap = e->h_fill = e->dev_cap = e->dev_fill = 0;
  for (int i = 0; i < 2; ++i) {
    if (e->h_raw[i]) cudaFreeHost(e->h_raw[i]);
    if (e->h_boff[i]) cudaFreeHost(e->h_boff[i]);
    e->h_raw[i] = nullptr; e->h_boff[i] = nullptr;
  }
  free_dev(e->d_raw); e->d_raw = nullptr;
  free_dev(e->d_boff); e->d_boff = nullptr;
  free_dev(e->d_bgen_cols); e->d_bgen_cols = nullptr;
  e->bgen_keep.clear();
  e->bgen_all = true;
  e->raw_cap = e->raw_fill = 0;
  e->raw_rows_cap = e->raw_rows = 0;
}

void release_geometry(dgrm_engine* e) {
  free_dev(e->d_src_ind); e->d_src_ind = nullptr;
  free_dev(e->d_work); e->d_work = nullptr;
  free_dev(e->d_work_cg1); e->d_work_cg1 = nullptr;
  free_dev(e->d_index); e->d_index = nullptr;
  free_dev(e->d_need_blocks); e->d_need_blocks = nullptr;
  free_dev(e->d_tile_rc); e->d_tile_rc = nullptr;
  free_dev(e->d_S); e->d_S = nullptr;
  free_dev(e->d_N); e->d_N = nullptr;
  free_dev(e->d_diag); e->d_diag = nullptr;
  free_dev(e->d_miss); e->d_miss = nullptr;
  free_dev(e->d_diag_a); e->d_diag_a = nullptr;
  if (e->h_diag_a) cudaFreeHost(e->h_diag_a);
  e->h_diag_a = nullptr;
  free_dev(e->d_subset); e->d_subset = nullptr;
  e->h_subset.clear();
  e->n_out = 0;
  free_dev(e->d_S2); e->d_S2 = nullptr;
  free_dev(e->d_N2); e->d_N2 = nullptr;
  e->saved = false;
  e->view = DGRM_VIEW_CURRENT;
  free_dev(e->d_chunk); e->d_chunk = nullptr;
  release_dosage_staging(e);
  free_dev(e->d_chunkT); e->d_chunkT = nullptr;
  free_dev(e->d_counts); e->d_counts = nullptr;
  free_dev(e->d_miss_part); e->d_miss_part = nullptr;
  free_dev(e->d_Z); e->d_Z = nullptr;
  free_dev(e->d_Q); e->d_Q = nullptr;
  free_dev(e->d_Q4); e->d_Q4 = nullptr;
  free_dev(e->d_MB); e->d_MB = nullptr;
  free_dev(e->d_zt); e->d_zt = nullptr;
  free_dev(e->d_diag_part); e->d_diag_part = nullptr;
  for (int i = 0; i < 2; ++i) {
    free_dev(e->d_stage[i]); e->d_stage[i] = nullptr;
    free_dev(e->d_rowidx[i]); e->d_rowidx[i] = nullptr;
    if (e->h_rowidx[i]) cudaFreeHost(e->h_rowidx[i]);
    e->h_rowidx[i] = nullptr;
  }
  e->have_ind = false;
}

void release_stats(dgrm_engine* e) {
  free_dev(e->d_freq); e->d_freq = nullptr;
  free_dev(e->d_nn); e->d_nn = nullptr;
  free_dev(e->d_p1); e->d_p1 = nullptr;
  free_dev(e->d_p2); e->d_p2 = nullptr;
  free_dev(e->d_sd); e->d_sd = nullptr;
  e->m_cap = 0;
}

int make_tmap(dgrm_engine* e, CUtensorMap* map, void* base, int elem_bytes, int64_t cols, int64_t rows,
              int64_t pitch_bytes) {
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), 128u};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = e->encode(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                               base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(e, DGRM_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DGRM_OK;
}

// Lower-triangle supertiles (256x256) in an L2-friendly raster: groups of 8 supertile rows, swept column by column.
// Engine r of a partition of G owns the contiguous raster segment [r*K/G, (r+1)*K/G) of the K supertiles: equal work
// (every tile costs the same), and its concurrent CTAs share A/B panels exactly as on one GPU (dealing tiles
// round-robin gave each engine one row of every column: 1 + 74 panels per k-block instead of 8 + 9, and the
// contraction ran 20 % slower at 8 GPUs).
// With four or more engines the raster is additionally cut into a 4 x 4 grid of macro blocks that are visited one
// after the other (6 full squares + 4 diagonal halves = 8 units of area, ordered so that the diagonal halves come in
// pairs: every unit-sized piece of the order is one square or two triangles), so that a contiguous segment is a
// COMPACT region of the triangle instead of a band of rows across all columns: with 8 engines an engine's tiles touch
// only about half of the individuals (a quarter as rows, a quarter as columns), and bed_expand writes tensor-core
// operands for those only.  The decode is the one part of the path that is replicated on every
// engine (it costs n*m per engine while the contraction costs n^2 m / G), so this halves what limits weak scaling.
int build_tiles(dgrm_engine* e) {
  const int T = (int)(e->n_pad / 256);
  const int GR = 8;
  e->nrb = 2 * T;
  e->ncp = T;
  e->work.clear();
  e->work_cg1.clear();
  e->tile_rc.clear();
  std::vector<int32_t>& index = e->h_index;   // host copy: ownership tests of the multi-engine readers
  index.assign((size_t)e->nrb * e->ncp, -1);
  e->rg_end.clear();
  e->rg_row1.clear();
  int64_t s = 0, ls = 0;
  const int64_t total = (int64_t)T * (T + 1) / 2;
  const int64_t own0 = total * e->cfg.rank / e->cfg.world, own1 = total * (e->cfg.rank + 1) / e->cfg.world;
  const int MB = e->cfg.world >= 4 ? 4 : 1;                 // macro blocks per side
  const int mb_size = (T + MB - 1) / MB;                    // supertile rows per macro block
  static const int kOrder4[10][2] = {{0, 0}, {1, 1}, {1, 0}, {2, 0}, {2, 1}, {2, 2}, {3, 3}, {3, 2}, {3, 0}, {3, 1}};
  static const int kOrder1[1][2] = {{0, 0}};
  const int (*order)[2] = MB == 4 ? kOrder4 : kOrder1;
  const int norder = MB == 4 ? 10 : 1;
  std::vector<uint8_t> need((size_t)2 * T, 0);              // 128-row blocks of individuals this engine's tiles touch
  for (int ob = 0; ob < norder; ++ob) {
    {
      const int bI = order[ob][0], bJ = order[ob][1];
      if (bI * mb_size >= T) continue;
      const int i_lo = bI * mb_size, i_hi = std::min(T, i_lo + mb_size);
      const int j_lo = bJ * mb_size, j_hi = std::min(T, j_lo + mb_size);
      for (int g = i_lo; g < i_hi; g += GR) {
        const int imax = std::min(g + GR, i_hi) - 1;
        if (!(ob == 0 && g == 0)) e->rg_end.push_back((int)(e->cfg.cta_group == 2 ? e->work.size() : e->work_cg1.size()));
        for (int J = j_lo; J < j_hi && J <= imax; ++J) {
          for (int I = std::max(g, J); I <= imax; ++I, ++s) {
            if (s < own0 || s >= own1) continue;
            const int s0 = (int)(2 * ls), s1 = (int)(2 * ls + 1);
            index[(size_t)(2 * I) * e->ncp + J] = s0;
            index[(size_t)(2 * I + 1) * e->ncp + J] = s1;
            e->tile_rc.push_back(2 * I); e->tile_rc.push_back(J);
            e->tile_rc.push_back(2 * I + 1); e->tile_rc.push_back(J);
            e->work_cg1.push_back(make_int4(I * 256, J * 256, s0, s0));
            e->work_cg1.push_back(make_int4(I * 256 + 128, J * 256, s1, s1));
            if (e->cfg.cta_group == 2) e->work.push_back(make_int4(I * 256, J * 256, s0, s1));
            need[(size_t)2 * I] = need[(size_t)2 * I + 1] = need[(size_t)2 * J] = need[(size_t)2 * J + 1] = 1;
            ++ls;
          }
        }
      }
    }
  }
  if (e->cfg.cta_group == 1) e->work = e->work_cg1;
  e->rg_end.push_back((int)e->work.size());
  for (int g = 0; g < T; g += GR) e->rg_row1.push_back(std::min(g + GR, T) * 256);   // (rectangle jobs: world == 1, MB == 1)
  e->ntiles = 2 * ls;
  e->need_blocks.clear();
  for (int b2 = 0; b2 < 2 * T; ++b2)
    if (need[(size_t)b2]) e->need_blocks.push_back(b2);
  // upload
  CK(cudaMalloc(&e->d_index, index.size() * sizeof(int32_t)));
  CK(cudaMemcpy(e->d_index, index.data(), index.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (!e->need_blocks.empty()) {
    CK(cudaMalloc(&e->d_need_blocks, e->need_blocks.size() * sizeof(int32_t)));
    CK(cudaMemcpy(e->d_need_blocks, e->need_blocks.data(), e->need_blocks.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  if (!e->work_cg1.empty()) {
    std::vector<int2> rc(e->work_cg1.size());
    for (size_t t = 0; t < e->work_cg1.size(); ++t) rc[(size_t)e->work_cg1[t].z] = make_int2(e->work_cg1[t].x, e->work_cg1[t].y);
    CK(cudaMalloc(&e->d_tile_rc, rc.size() * sizeof(int2)));
    CK(cudaMemcpy(e->d_tile_rc, rc.data(), rc.size() * sizeof(int2), cudaMemcpyHostToDevice));
  }
  if (!e->work.empty()) {
    CK(cudaMalloc(&e->d_work, e->work.size() * sizeof(int4)));
    CK(cudaMemcpy(e->d_work, e->work.data(), e->work.size() * sizeof(int4), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&e->d_work_cg1, e->work_cg1.size() * sizeof(int4)));
    CK(cudaMemcpy(e->d_work_cg1, e->work_cg1.data(), e->work_cg1.size() * sizeof(int4), cudaMemcpyHostToDevice));
  }
  return DGRM_OK;
}

TileStore tile_store(const dgrm_engine* e) {
  TileStore ts;
  ts.S = e->d_S;
  ts.N = e->cfg.flat_norm ? nullptr : e->d_N;
  ts.index = e->d_index;
  ts.ncp = e->ncp;
  ts.n = (int)(e->d_subset ? e->n_out : e->n);
  ts.flat_n = (double)e->m;
  ts.subset = e->d_subset;
  ts.S2 = e->d_S2;
  ts.N2 = e->saved_flat ? nullptr : e->d_N2;
  ts.flat_n2 = (double)e->m2;
  ts.view = e->saved ? e->view : DGRM_VIEW_CURRENT;
  return ts;
}

template <int CG, int KIND, typename OutT>
int launch_syrk(dgrm_engine* e, const CUtensorMap& tmap, const SyrkParams& p) {
  using Cfg = SyrkCfg<CG>;
  auto kern = grm_syrk_kernel<CG, KIND, OutT>;   // its dynamic shared memory limit is raised per device in dgrm_create
  const int units = std::min<int>(p.work_end - p.work_begin, (e->sm_count - e->reserve_sms) / CG);
  if (units <= 0) return DGRM_OK;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(units * CG);
  lc.blockDim = dim3(Cfg::kThreads);
  lc.dynamicSmemBytes = Cfg::kSmemBytes;
  lc.stream = e->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  CK(cudaLaunchKernelEx(&lc, kern, tmap, p));
  return DGRM_OK;
}

constexpr int kCountMergeSnps = 32768;  // SNPs per pair-count accumulation (exact in fp32: << 2^24)

int flush_super(dgrm_engine* e, bool final, bool wrap = false);
int flush_dosage(dgrm_engine* e);
int process_chunk(dgrm_engine* e);
int flush_prefix(dgrm_engine* e);
bool streaming_bound(const dgrm_engine* e);

// The next resident operand slot (`chunks_in_super`) becomes writable on the decode stream.
int acquire_slot(dgrm_engine* e) {
  if (e->chunks_in_super == e->slots) {  // every resident slot is taken: contract them before reusing slot 0
    const int rc = flush_super(e, false, true);
    if (rc) return rc;
  }
  if (e->supers_done > 0) CK(cudaStreamWaitEvent(e->prep, e->slot_ev[e->chunks_in_super], 0));  // the slot's previous chunk is contracted
  return DGRM_OK;
}

// The open slot holds a finished chunk of kc SNPs (kc_pad operand columns).
int close_chunk(dgrm_engine* e, int kc, int kc_pad) {
  e->slot_kpad[e->chunks_in_super] = kc_pad;
  e->m += kc;
  e->chunks_done++;
  e->fill = 0;
  ++e->chunks_in_super;
  // Without streaming output bindings the chunk is contracted right away on the MMA stream, so that the decode of
  // the next chunk (prep stream) overlaps it; with them the contraction waits for dgrm_finish, which walks the
  // tiles group by group so that finished groups leave the device early.
  if (e->overlap && !streaming_bound(e)) return flush_super(e, false);
  if (e->overlap) return flush_prefix(e);
  return DGRM_OK;
}

// One chunk of packed rows: per-SNP statistics, then expansion into resident operand slot `chunks_in_super`.
int process_chunk(dgrm_engine* e) {
  const int kc = e->fill;
  if (kc == 0) return DGRM_OK;
  if (e->ld_mode) {
    // LD mode: the chunk's SNP rows become operand ROWS [m, m + kc), all individuals along K
    Prof pr(e, DGRM_K_EXPAND, e->prep);
    LdExpandArgs a;
    a.chunk = e->d_chunk; a.pitch = e->chunk_pitch; a.kc = kc; a.n = (int)e->n_row; a.zt = e->d_zt;
    a.freq = e->d_freq + e->m * 4;
    a.Z = e->d_Z + e->m * e->z_pitch; a.z_pitch = e->z_pitch; a.lo_off = (int64_t)e->slots * e->kc_max;
    a.Q = e->d_Q ? e->d_Q + e->m * e->q_pitch : nullptr; a.q_pitch = e->q_pitch;
    a.Q4 = e->d_Q4 ? e->d_Q4 + e->m * e->q4_pitch : nullptr; a.q4_pitch = e->q4_pitch;
    a.MB = e->d_MB ? e->d_MB + e->m * e->mb_pitch : nullptr; a.mb_pitch = e->mb_pitch;
    a.diag = e->d_diag + e->m;
    const dim3 grid((unsigned)(((e->n_row + 7) / 8 + 255) / 256), (unsigned)kc);
    const int op = e->cfg.operand;
    if (op == DGRM_OPERAND_BF16) ld_expand_kernel<true, false><<<grid, 256, 0, e->prep>>>(a);
    else if (op == DGRM_OPERAND_FP16X2) ld_expand_kernel<false, true><<<grid, 256, 0, e->prep>>>(a);
    else ld_expand_kernel<false, false><<<grid, 256, 0, e->prep>>>(a);
    CK(cudaGetLastError());
    e->m += kc;
    e->chunks_done++;
    e->fill = 0;
    return DGRM_OK;
  }
  const int kc_pad = (int)round_up(kc, 256);  // whole 128-byte k-blocks for every operand width (fp4: 256 SNPs)
  const bool split = e->cfg.operand == DGRM_OPERAND_FP16X2;
  const bool bf16 = e->cfg.operand == DGRM_OPERAND_BF16;
  const bool count_mma = !e->cfg.flat_norm && e->cfg.count_method == DGRM_COUNT_MMA;
  const bool count_fp4 = !e->cfg.flat_norm && e->cfg.count_method == DGRM_COUNT_FP4;
  const bool count_popc = !e->cfg.flat_norm && e->cfg.count_method == DGRM_COUNT_POPC;
  if (e->fill_dosage) {
    // dosage rows were expanded into the open slot as they arrived (expand_dosage_rows): only the diagonal is left
    diag_reduce_kernel<<<(unsigned)((e->n_pad + 31) / 32), 256, 0, e->prep>>>(e->d_diag_part, kc_pad / kExpK, (int)e->n_pad,
                                                                             e->d_diag);
    CK(cudaGetLastError());
    e->fill_dosage = false;
    return close_chunk(e, kc, kc_pad);
  }
  int rc0 = acquire_slot(e);
  if (rc0) return rc0;
  const int slot = e->chunks_in_super;
  const int64_t col0 = (int64_t)slot * e->kc_max;

  // (per-SNP counts, statistics and the zt table of these rows were produced while the rows were gathered into the
  // chunk buffer: bed_decode_kernel in push_device_rows)
  const bool transposed = e->d_chunkT != nullptr;
  if (!transposed && !e->cfg.flat_norm && e->miss_valid) {
    if (streaming_bound(e)) {
      const int nbytes = (int)e->b_row;
      const dim3 grid((unsigned)((nbytes + 127) / 128), (unsigned)((kc + 255) / 256));
      bed_missing_per_individual_kernel<<<grid, 128, 0, e->prep>>>(e->d_chunk, e->chunk_pitch, kc, nbytes, (int)e->n_row, e->d_miss);
      CK(cudaGetLastError());
    } else {
      e->miss_valid = false;   // a chunk went by uncounted: streamed pair counts fall back to int32
    }
  }
  {
    Prof pr(e, DGRM_K_EXPAND, e->prep);
    dim3 grid((unsigned)(e->n_pad / kExpI), (unsigned)(kc_pad / kExpK));
    const int qk = count_mma ? 1 : count_fp4 ? 2 : count_popc ? 3 : 0;
    if (transposed) {
      ExpandTArgs a;
      // operands only for the individuals this engine's tiles touch (all of them unless the partition is compact)
      a.blocks = (e->need_blocks.size() < (size_t)(e->n_pad / kExpI)) ? e->d_need_blocks : nullptr;
      if (a.blocks) grid.x = (unsigned)e->need_blocks.size();
      if (grid.x == 0) grid.x = 1, a.blocks = nullptr;
      a.k_fast = e->expand_k_fast;
      a.stream_stores = e->expand_stream_stores;
      if (a.k_fast) std::swap(grid.x, grid.y);
      a.T = e->d_chunkT; a.t_pitch = e->t_pitch; a.kc = kc; a.n = (int)e->n_row; a.zt = e->d_zt;
      a.Z = e->d_Z + col0; a.z_pitch = e->z_pitch; a.lo_off = (int64_t)e->slots * e->kc_max;
      a.Q = count_mma ? e->d_Q + col0 : nullptr; a.q_pitch = e->q_pitch;
      a.Q4 = count_fp4 ? e->d_Q4 + col0 / 2 : nullptr; a.q4_pitch = e->q4_pitch;
      a.MB = count_popc ? e->d_MB + col0 / 32 : nullptr; a.mb_pitch = e->mb_pitch;
      a.diag_part = e->d_diag_part; a.miss_part = e->cfg.flat_norm ? nullptr : e->d_miss_part; a.n_pad = (int)e->n_pad;
#define DGRM_EXPAND_T(B, S, K) bed_expand_t_kernel<B, S, K><<<grid, kExpTThreads, sizeof(ExpandTSmem<K, S>), e->prep>>>(a)
#define DGRM_EXPAND_TQ(B, S)                                                   \
  do {                                                                         \
    if (qk == 0) DGRM_EXPAND_T(B, S, 0); else if (qk == 1) DGRM_EXPAND_T(B, S, 1); \
    else if (qk == 2) DGRM_EXPAND_T(B, S, 2); else DGRM_EXPAND_T(B, S, 3);     \
  } while (0)
      if (bf16) DGRM_EXPAND_TQ(true, false);
      else if (split) DGRM_EXPAND_TQ(false, true);
      else DGRM_EXPAND_TQ(false, false);
#undef DGRM_EXPAND_TQ
#undef DGRM_EXPAND_T
      CK(cudaGetLastError());
      diag_reduce_kernel<<<(unsigned)((e->n_pad + 31) / 32), 256, 0, e->prep>>>(
          e->d_diag_part, kc_pad / kExpK, (int)e->n_pad, e->d_diag, a.miss_part, a.miss_part ? e->d_miss : nullptr);
      CK(cudaGetLastError());
    } else {
      ExpandArgs a;
      a.chunk = e->d_chunk; a.pitch = e->chunk_pitch; a.kc = kc; a.n = (int)e->n_row; a.zt = e->d_zt;
      a.Z = e->d_Z + col0; a.z_pitch = e->z_pitch; a.lo_off = (int64_t)e->slots * e->kc_max;
      a.Q = count_mma ? e->d_Q + col0 : nullptr; a.q_pitch = e->q_pitch;
      a.Q4 = count_fp4 ? e->d_Q4 + col0 / 2 : nullptr; a.q4_pitch = e->q4_pitch;
      a.MB = count_popc ? e->d_MB + col0 / 32 : nullptr; a.mb_pitch = e->mb_pitch;
      a.diag_part = e->d_diag_part; a.n_pad = (int)e->n_pad;
      const size_t smem = sizeof(ExpandSmem);
#define DGRM_EXPAND(B, S, K) bed_expand_kernel<B, S, K><<<grid, 256, smem, e->prep>>>(a)
#define DGRM_EXPAND_Q(B, S)                                              \
  do {                                                                   \
    if (qk == 0) DGRM_EXPAND(B, S, 0); else if (qk == 1) DGRM_EXPAND(B, S, 1); \
    else if (qk == 2) DGRM_EXPAND(B, S, 2); else DGRM_EXPAND(B, S, 3);   \
  } while (0)
      if (bf16) DGRM_EXPAND_Q(true, false);
      else if (split) DGRM_EXPAND_Q(false, true);
      else DGRM_EXPAND_Q(false, false);
#undef DGRM_EXPAND_Q
#undef DGRM_EXPAND
      CK(cudaGetLastError());
      diag_reduce_kernel<<<(unsigned)((e->n_pad + 31) / 32), 256, 0, e->prep>>>(e->d_diag_part, kc_pad / kExpK,
                                                                               (int)e->n_pad, e->d_diag);
      CK(cudaGetLastError());
    }
  }
  return close_chunk(e, kc, kc_pad);
}

// Rectangles of a block-cyclic binding stream out row group by row group only when this engine owns the whole
// triangle: a partial owner would have to leave the other engines' entries of the caller's buffer alone, which the
// strided copy-engine transfers cannot do.  With world > 1 such bindings are filled at the end of dgrm_finish
// (owned entries only); dgrm_bind_tiles is the streaming egress of a multi-engine run.
bool binding_streams(const dgrm_engine* e, const dgrm_engine::Binding& bd) { return bd.uplo == 0 && e->cfg.world == 1; }

bool streaming_bound(const dgrm_engine* e) {
  bool s = e->ntbind > 0;
  for (int b = 0; b < e->nbind; ++b) s |= binding_streams(e, e->bind[b]);
  return s;
}

// Does this engine own the tile that holds lower-triangle element (r >= c) of the visible matrix?
inline bool owns_element(const dgrm_engine* e, int r, int c) {
  if (!e->h_subset.empty()) { r = e->h_subset[(size_t)r]; c = e->h_subset[(size_t)c]; }
  return e->h_index[(size_t)(r >> 7) * e->ncp + (c >> 8)] >= 0;
}

// first local index whose global index is >= target (block-cyclic, monotone)
int first_local_ge(int target, int nloc, int bs, int me, int P) {
  int a = 0, b = nloc;
  while (a < b) {
    const int mid = (a + b) / 2;
    const int gidx = ((mid / bs) * P + me) * bs + mid % bs;
    if (gidx >= target) b = mid; else a = mid + 1;
  }
  return a;
}

inline size_t transport_size(int ntype) {
  return ntype == 2 ? sizeof(uint16_t) : (ntype == 1 || ntype == 3) ? sizeof(int32_t) : sizeof(double);
}

// Issue output job k: pack kernel (out_stream) -> device staging -> copy engine (d2h_stream) -> either the
// caller's pinned buffer (direct) or the library's pinned host staging (finished by finish_job).
int issue_job(dgrm_engine* e, size_t k) {
  const dgrm_engine::OutJob& j = e->jobs[k];
  if (j.tiles) {
    const dgrm_engine::TileBinding& tb = e->tbind[j.b];
    const int sl = (int)(k % dgrm_engine::kOutSlots);
    const int64_t cnt = j.t1 - j.t0;
    TileStore ts = tile_store(e);
    ts.S = ts.S + j.t0;
    ts.N = ts.N ? ts.N + j.t0 : nullptr;
    ts.S2 = ts.S2 ? ts.S2 + j.t0 : nullptr;
    ts.N2 = ts.N2 ? ts.N2 + j.t0 : nullptr;
    CK(cudaStreamWaitEvent(e->out_stream, e->group_ev[(size_t)j.ev], 0));
    {
      Prof pr(e, DGRM_K_FETCH, e->out_stream);
      const unsigned blocks = (unsigned)((cnt + 255) / 256);
      if (j.ntype == 2)
        export_tiles_kernel<uint16_t><<<blocks, 256, 0, e->out_stream>>>(ts, tb.which, cnt, (uint16_t*)e->d_out_stage[sl],
                                                                        e->d_tile_rc, j.t0 / kTileElems, e->d_diag);
      else if (j.ntype == 1)
        export_tiles_kernel<int><<<blocks, 256, 0, e->out_stream>>>(ts, tb.which, cnt, (int*)e->d_out_stage[sl],
                                                                   e->d_tile_rc, j.t0 / kTileElems, e->d_diag);
      else if (j.ntype == 3)
        export_tiles_kernel<float><<<blocks, 256, 0, e->out_stream>>>(ts, tb.which, cnt, (float*)e->d_out_stage[sl],
                                                                     e->d_tile_rc, j.t0 / kTileElems, e->d_diag);
      else
        export_tiles_kernel<double><<<blocks, 256, 0, e->out_stream>>>(ts, tb.which, cnt, (double*)e->d_out_stage[sl],
                                                                      e->d_tile_rc, j.t0 / kTileElems, e->d_diag);
      CK(cudaGetLastError());
    }
    CK(cudaEventRecord(e->stage_packed[sl], e->out_stream));
    CK(cudaStreamWaitEvent(e->d2h_stream, e->stage_packed[sl], 0));
    {
      Prof pr(e, DGRM_K_D2H, e->d2h_stream);
      const size_t es = transport_size(j.ntype);
      void* dst = j.direct ? (void*)(tb.host + j.t0) : e->h_out_stage[sl];
      CK(cudaMemcpyAsync(dst, e->d_out_stage[sl], (size_t)cnt * es, cudaMemcpyDeviceToHost, e->d2h_stream));
      e->d2h_bytes += (int64_t)cnt * (int64_t)es;
    }
    CK(cudaEventRecord(e->stage_copied_out[sl], e->d2h_stream));
    return DGRM_OK;
  }
  const dgrm_engine::Binding& bd = e->bind[j.b];
  const CyclicGeom& g = bd.g;
  const int sl = (int)(k % dgrm_engine::kOutSlots);
  const TileStore ts = tile_store(e);
  CK(cudaStreamWaitEvent(e->out_stream, e->group_ev[(size_t)j.ev], 0));
  {
    Prof pr(e, DGRM_K_FETCH, e->out_stream);
    const dim3 blk(128, 2);
    const dim3 grid((unsigned)((j.lr1 - j.lr0 + 127) / 128), (unsigned)((j.lc1 + 1) / 2));
    if (j.ntype == 2)
      pack_rows_kernel<uint16_t><<<grid, blk, 0, e->out_stream>>>(ts, bd.which, e->d_diag, (uint16_t*)e->d_out_stage[sl], g,
                                                                 j.lr0, j.lr1, j.lc1);
    else if (j.ntype == 1)
      pack_rows_kernel<int><<<grid, blk, 0, e->out_stream>>>(ts, bd.which, e->d_diag, (int*)e->d_out_stage[sl], g, j.lr0,
                                                            j.lr1, j.lc1);
    else if (j.ntype == 3)
      pack_rows_kernel<float><<<grid, blk, 0, e->out_stream>>>(ts, bd.which, e->d_diag, (float*)e->d_out_stage[sl], g, j.lr0,
                                                              j.lr1, j.lc1);
    else
      pack_rows_kernel<double><<<grid, blk, 0, e->out_stream>>>(ts, bd.which, e->d_diag, (double*)e->d_out_stage[sl], g,
                                                               j.lr0, j.lr1, j.lc1);
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(e->stage_packed[sl], e->out_stream));
  CK(cudaStreamWaitEvent(e->d2h_stream, e->stage_packed[sl], 0));
  {
    // copy engine: everything left of the diagonal block in one strided copy, the diagonal block itself as
    // 256-column strips that start at their own diagonal (so almost nothing above the diagonal is moved)
    Prof pr(e, DGRM_K_D2H, e->d2h_stream);
    const int h = j.lr1 - j.lr0;
    const size_t es = transport_size(j.ntype);
    auto copy_rect = [&](int lra, int lrb, int lca, int lcb) -> cudaError_t {
      if (lrb <= lra || lcb <= lca) return cudaSuccess;
      e->d2h_bytes += (int64_t)(lrb - lra) * (lcb - lca) * (int64_t)es;
      const char* src = (const char*)e->d_out_stage[sl] + ((int64_t)lca * h + (lra - j.lr0)) * es;
      if (j.direct)
        return cudaMemcpy2DAsync(bd.host + (int64_t)lca * g.lld + lra, (size_t)g.lld * sizeof(double), src, (size_t)h * es,
                                 (size_t)(lrb - lra) * es, (size_t)(lcb - lca), cudaMemcpyDeviceToHost, e->d2h_stream);
      char* dst = (char*)e->h_out_stage[sl] + ((int64_t)lca * h + (lra - j.lr0)) * es;
      return cudaMemcpy2DAsync(dst, (size_t)h * es, src, (size_t)h * es, (size_t)(lrb - lra) * es, (size_t)(lcb - lca),
                               cudaMemcpyDeviceToHost, e->d2h_stream);
    };
    if (!j.direct && es < sizeof(double)) {
      // narrow staged elements: the rectangle is contiguous in both staging buffers, and one flat copy beats a
      // strided one whose segments would be only h * 2 (or 4) bytes long; the few entries above the diagonal of the
      // diagonal block ride along (the host ignores them)
      CK(cudaMemcpyAsync(e->h_out_stage[sl], e->d_out_stage[sl], (size_t)h * (size_t)j.lc1 * es, cudaMemcpyDeviceToHost,
                         e->d2h_stream));
      e->d2h_bytes += (int64_t)h * j.lc1 * (int64_t)es;
      CK(cudaEventRecord(e->stage_copied_out[sl], e->d2h_stream));
      return DGRM_OK;
    }
    const int lc0d = first_local_ge(j.row0, g.nloc_c, g.bs_c, g.my_col, g.C);
    CK(copy_rect(j.lr0, j.lr1, 0, lc0d));
    for (int g0 = j.row0; g0 < j.row1; g0 += 256) {
      const int lca = first_local_ge(g0, g.nloc_c, g.bs_c, g.my_col, g.C);
      const int lcb = first_local_ge(std::min(g0 + 256, j.row1), g.nloc_c, g.bs_c, g.my_col, g.C);
      const int lra = first_local_ge(g0, g.nloc_r, g.bs_r, g.my_row, g.R);
      CK(copy_rect(lra, j.lr1, lca, lcb));
    }
  }
  CK(cudaEventRecord(e->stage_copied_out[sl], e->d2h_stream));
  return DGRM_OK;
}

// int32 pair counts -> the caller's fp64 matrix.  The destination is written once and not read again by this
// library: non-temporal stores skip the read-for-ownership of every line (a third of the host memory traffic).
#if defined(__x86_64__)
__attribute__((target("avx2"))) void widen_i32_stream(double* dst, const int32_t* src, int cnt) {
  int i = 0;
  for (; i < cnt && ((uintptr_t)(dst + i) & 31); ++i) dst[i] = (double)src[i];
  for (; i + 4 <= cnt; i += 4)
    _mm256_stream_pd(dst + i, _mm256_cvtepi32_pd(_mm_loadu_si128((const __m128i*)(src + i))));
  for (; i < cnt; ++i) dst[i] = (double)src[i];
}
#endif

void widen_i32(double* dst, const int32_t* src, int cnt, bool stream) {
#if defined(__x86_64__)
  if (stream) return widen_i32_stream(dst, src, cnt);
#endif
  for (int i = 0; i < cnt; ++i) dst[i] = (double)src[i];
}

// fp32 transport of A -> the caller's fp64 matrix
#if defined(__x86_64__)
__attribute__((target("avx2"))) void widen_f32_stream(double* dst, const float* src, int cnt) {
  int i = 0;
  for (; i < cnt && ((uintptr_t)(dst + i) & 31); ++i) dst[i] = (double)src[i];
  for (; i + 4 <= cnt; i += 4) _mm256_stream_pd(dst + i, _mm256_cvtps_pd(_mm_loadu_ps(src + i)));
  for (; i < cnt; ++i) dst[i] = (double)src[i];
}
#endif

void widen_f32(double* dst, const float* src, int cnt, bool stream) {
#if defined(__x86_64__)
  if (stream) return widen_f32_stream(dst, src, cnt);
#endif
  for (int i = 0; i < cnt; ++i) dst[i] = (double)src[i];
}

// 16-bit deficiencies m - N -> fp64 pair counts (0xFFFF = no value = 0)
#if defined(__x86_64__)
__attribute__((target("avx2"))) void widen_u16_stream(double* dst, const uint16_t* src, int cnt, double m_total) {
  int i = 0;
  for (; i < cnt && ((uintptr_t)(dst + i) & 31); ++i) dst[i] = src[i] == 0xFFFF ? 0. : m_total - (double)src[i];
  const __m256d mt = _mm256_set1_pd(m_total);
  const __m128i sentinel = _mm_set1_epi32(0xFFFF);
  for (; i + 4 <= cnt; i += 4) {
    const __m128i v = _mm_cvtepu16_epi32(_mm_loadl_epi64((const __m128i*)(src + i)));
    const __m256d none = _mm256_castsi256_pd(_mm256_cvtepi32_epi64(_mm_cmpeq_epi32(v, sentinel)));
    _mm256_stream_pd(dst + i, _mm256_andnot_pd(none, _mm256_sub_pd(mt, _mm256_cvtepi32_pd(v))));
  }
  for (; i < cnt; ++i) dst[i] = src[i] == 0xFFFF ? 0. : m_total - (double)src[i];
}
#endif

void widen_u16(double* dst, const uint16_t* src, int cnt, double m_total, bool stream) {
#if defined(__x86_64__)
  if (stream) return widen_u16_stream(dst, src, cnt, m_total);
#endif
  for (int i = 0; i < cnt; ++i) dst[i] = src[i] == 0xFFFF ? 0. : m_total - (double)src[i];
}

// Wait for job k's transfer; staged jobs are then moved (and pair counts widened to double) into the caller's
// buffer by host threads, exactly the rows on or below the diagonal of every column.
int finish_job(dgrm_engine* e, size_t k) {
  const dgrm_engine::OutJob& j = e->jobs[k];
  const int sl = (int)(k % dgrm_engine::kOutSlots);
  const auto tw0 = std::chrono::steady_clock::now();
  CK(cudaEventSynchronize(e->stage_copied_out[sl]));
  const auto tw1 = std::chrono::steady_clock::now();
  e->out_wait_ms += std::chrono::duration<double, std::milli>(tw1 - tw0).count();
  struct Acc {
    dgrm_engine* e;
    std::chrono::steady_clock::time_point t;
    ~Acc() { e->out_host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); }
  } acc{e, tw1};
  if (j.direct) return DGRM_OK;
#if defined(__x86_64__)
  static const bool nt = __builtin_cpu_supports("avx2") && !getenv("DGRM_NO_NT");
#else
  static const bool nt = false;
#endif
  const double m_total = (double)e->m;
  if (j.tiles) {
    const void* src = e->h_out_stage[sl];
    double* dst = e->tbind[j.b].host + j.t0;
    const int64_t cnt = j.t1 - j.t0;
    const int ntype = j.ntype;
#pragma omp parallel for schedule(static)
    for (int64_t o = 0; o < cnt; o += 1 << 16) {
      const int c = (int)std::min<int64_t>(1 << 16, cnt - o);
      if (ntype == 2) widen_u16(dst + o, (const uint16_t*)src + o, c, m_total, nt);
      else if (ntype == 1) widen_i32(dst + o, (const int32_t*)src + o, c, nt);
      else if (ntype == 3) widen_f32(dst + o, (const float*)src + o, c, nt);
      else memcpy(dst + o, (const double*)src + o, (size_t)c * sizeof(double));
#if defined(__x86_64__)
      _mm_sfence();   // non-temporal stores are weakly ordered
#endif
    }
    return DGRM_OK;
  }
  const dgrm_engine::Binding& bd = e->bind[j.b];
  const CyclicGeom g = bd.g;
  const int h = j.lr1 - j.lr0;
  const void* hs = e->h_out_stage[sl];
  double* host = bd.host;
  const int ntype = j.ntype;
  const int lr0 = j.lr0, lr1 = j.lr1;
#pragma omp parallel
  {
#pragma omp for schedule(dynamic, 16) nowait
  for (int lc = 0; lc < j.lc1; ++lc) {
    const int gc = ((lc / g.bs_c) * g.C + g.my_col) * g.bs_c + lc % g.bs_c;
    const int lra = std::max(lr0, first_local_ge(gc, g.nloc_r, g.bs_r, g.my_row, g.R));
    if (lra >= lr1) continue;
    double* dst = host + (int64_t)lc * g.lld + lra;
    const int cnt = lr1 - lra;
    if (ntype == 2) {
      widen_u16(dst, (const uint16_t*)hs + (int64_t)lc * h + (lra - lr0), cnt, m_total, nt);
    } else if (ntype == 1) {
      widen_i32(dst, (const int32_t*)hs + (int64_t)lc * h + (lra - lr0), cnt, nt);
    } else if (ntype == 3) {
      widen_f32(dst, (const float*)hs + (int64_t)lc * h + (lra - lr0), cnt, nt);
    } else {
      memcpy(dst, (const double*)hs + (int64_t)lc * h + (lra - lr0), (size_t)cnt * sizeof(double));
    }
  }
#if defined(__x86_64__)
  _mm_sfence();
#endif
  }
  return DGRM_OK;
}

// fp32 transport: the streamed A values were rounded to fp32 on their way out; the diagonal (which the engine
// holds as the exact fp64 sum of squares) is delivered separately as doubles and patched into the bound buffers.
int deliver_exact_diagonal(dgrm_engine* e) {
  if (e->cfg.a_transport != DGRM_TRANSPORT_FP32) return DGRM_OK;
  bool any_rect = false, any_tiles = false;
  for (int b = 0; b < e->nbind; ++b) any_rect |= binding_streams(e, e->bind[b]) && e->bind[b].which == DGRM_FETCH_A;
  for (int b = 0; b < e->ntbind; ++b) any_tiles |= e->tbind[b].which == DGRM_FETCH_A;
  if (!any_rect && !any_tiles) return DGRM_OK;
  if (!e->d_diag_a) {
    CK(cudaMalloc(&e->d_diag_a, (size_t)e->n_pad * sizeof(double)));
    CK(cudaHostAlloc(&e->h_diag_a, (size_t)e->n_pad * sizeof(double), cudaHostAllocDefault));
  }
  // stream order puts this after the last tile group; cnt values of A_ii as doubles
  auto fetch_diag = [&](const TileStore& ts, int cnt) -> int {
    diag_values_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, e->stream>>>(ts, DGRM_FETCH_A, e->d_diag, e->d_diag_a, cnt);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->h_diag_a, e->d_diag_a, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->d2h_bytes += (int64_t)cnt * 8;
    return DGRM_OK;
  };
  if (any_rect) {   // block-cyclic bindings see the visible individuals (output subset)
    const int n_vis = (int)(e->d_subset ? e->n_out : e->n);
    const int rc = fetch_diag(tile_store(e), n_vis);
    if (rc) return rc;
    for (int b = 0; b < e->nbind; ++b) {
      const dgrm_engine::Binding& bd = e->bind[b];
      if (!binding_streams(e, bd) || bd.which != DGRM_FETCH_A) continue;
      const CyclicGeom& g = bd.g;
      for (int i = 0; i < n_vis; ++i) {
        if ((i / g.bs_r) % g.R != g.my_row || (i / g.bs_c) % g.C != g.my_col || !owns_element(e, i, i)) continue;
        const int lr = (i / g.bs_r / g.R) * g.bs_r + i % g.bs_r, lc = (i / g.bs_c / g.C) * g.bs_c + i % g.bs_c;
        bd.host[(int64_t)lc * g.lld + lr] = e->h_diag_a[i];
      }
    }
  }
  if (any_tiles) {  // tile bindings hold the engine's own tiles: every individual, no subset
    TileStore ts = tile_store(e);
    ts.subset = nullptr;
    ts.n = (int)e->n;
    const int rc = fetch_diag(ts, (int)e->n);
    if (rc) return rc;
    for (int b = 0; b < e->ntbind; ++b) {
      if (e->tbind[b].which != DGRM_FETCH_A) continue;
      for (int i = 0; i < (int)e->n; ++i) {
        const int sidx = e->h_index[(size_t)(i >> 7) * e->ncp + (i >> 8)];
        if (sidx >= 0) e->tbind[b].host[(int64_t)sidx * kTileElems + (int64_t)(i & 255) * kTileM + (i & 127)] = e->h_diag_a[i];
      }
    }
  }
  return DGRM_OK;
}

// issue every job that is final after tile group `gi` for which a staging slot is free
int issue_ready_jobs(dgrm_engine* e, int gi) {
  while (e->jobs_issued < e->jobs.size() && e->jobs[e->jobs_issued].ev <= gi &&
         e->jobs_issued - e->jobs_done < (size_t)dgrm_engine::kOutSlots) {
    const int rc = issue_job(e, e->jobs_issued);
    if (rc) return rc;
    e->jobs_issued++;
  }
  return DGRM_OK;
}

// Tile groups of a streamed build: one per raster row group, ends rounded up to whole waves of the persistent grid
// (no ragged tail, and a row group is final as soon as its launches are done).
void stream_groups(const dgrm_engine* e, std::vector<int>* group_end, std::vector<int>* rg_group) {
  const int nwork = (int)e->work.size();
  const int units = std::max(1, (e->sm_count - e->reserve_sms) / e->cfg.cta_group);
  group_end->clear();
  rg_group->assign(e->rg_end.size(), 0);
  for (size_t rg = 0; rg < e->rg_end.size(); ++rg) {
    const int end = std::min(nwork, (e->rg_end[rg] + units - 1) / units * units);
    if (end > (group_end->empty() ? 0 : group_end->back())) group_end->push_back(end);
    (*rg_group)[rg] = std::max(0, (int)group_end->size() - 1);
  }
  if (group_end->empty() || group_end->back() != nwork) group_end->push_back(nwork);
}

// S += Z^T Z of resident chunk j0 over work items [w0, w1)
int contract_s(dgrm_engine* e, int w0, int w1, int j0) {
  const bool split = e->cfg.operand == DGRM_OPERAND_FP16X2;
  const bool bf16 = e->cfg.operand == DGRM_OPERAND_BF16;
  const int lo = e->slots * e->kc_max;
  Prof pr(e, DGRM_K_SYRK);
  SyrkParams p{};
  p.work = e->d_work;
  p.l2_policy = e->l2_policy;
  p.work_begin = w0;
  p.work_end = w1;
  p.nchunks = 1;
  p.chunk_col[0] = j0 * e->kc_max;
  p.chunk_kpad[0] = e->slot_kpad[j0];
  p.accumulate = e->supers_done > 0 || j0 > 0;
  p.npass = split ? 3 : 1;
  p.a_off[0] = 0; p.a_off[1] = lo; p.a_off[2] = 0;
  p.b_off[0] = 0; p.b_off[1] = 0; p.b_off[2] = lo;
  p.out = e->d_S;
  p.idesc = make_idesc(e->cfg.cta_group, bf16 ? 1 : 0);
  return e->cfg.cta_group == 2 ? launch_syrk<2, 0, double>(e, e->tmap_z, p) : launch_syrk<1, 0, double>(e, e->tmap_z, p);
}

// N += M^T M of the resident chunks [j0, j1) (one accumulation: at most 32768 SNPs) over work items [w0, w1)
int contract_n(dgrm_engine* e, int w0, int w1, int j0, int j1) {
  const bool count_fp4 = e->cfg.count_method == DGRM_COUNT_FP4;
  Prof pr(e, DGRM_K_PAIRCNT);
  if (e->cfg.count_method == DGRM_COUNT_POPC) {
    const int nw1 = w1 * e->cfg.cta_group, nw0 = w0 * e->cfg.cta_group;  // 128-row tiles of this group
    pair_count_popc_kernel<<<std::min(nw1 - nw0, e->sm_count * 2), 256, 0, e->stream>>>(
        e->d_MB + (int64_t)j0 * e->kc_max / 32, e->mb_pitch, e->slot_kpad[j0] / 32, e->d_work_cg1 + nw0, nw1 - nw0,
        e->d_N, e->supers_done > 0 || j0 > 0);
    CK(cudaGetLastError());
    return DGRM_OK;
  }
  int kpad = 0;
  for (int j = j0; j < j1; ++j) kpad += e->slot_kpad[j];   // slots are adjacent; only the last can be short
  SyrkParams p{};
  p.work = e->d_work;
  p.l2_policy = e->l2_policy;
  p.work_begin = w0;
  p.work_end = w1;
  p.nchunks = 1;
  p.npass = 1;
  p.a_off[0] = p.b_off[0] = 0;
  p.out = e->d_N;
  p.accumulate = e->supers_done > 0 || j0 > 0;
  if (count_fp4) {  // operand addressed in bytes: two SNPs per byte
    p.chunk_col[0] = j0 * e->kc_max / 2;
    p.chunk_kpad[0] = kpad / 2;
    p.idesc = make_idesc_mxf4(e->cfg.cta_group);
    return e->cfg.cta_group == 2 ? launch_syrk<2, 2, int>(e, e->tmap_q4, p) : launch_syrk<1, 2, int>(e, e->tmap_q4, p);
  }
  p.chunk_col[0] = j0 * e->kc_max;
  p.chunk_kpad[0] = kpad;
  p.idesc = make_idesc(e->cfg.cta_group, 0 /* e4m3 */);
  return e->cfg.cta_group == 2 ? launch_syrk<2, 1, int>(e, e->tmap_q, p) : launch_syrk<1, 1, int>(e, e->tmap_q, p);
}

inline int count_merge(const dgrm_engine* e) {
  // pair count: 0/1 products accumulate exactly in fp32 up to 2^24, so resident chunks are merged up to
  // 32768 SNPs per TMEM accumulation (one int32 read-modify-write of the tile per merged group)
  const bool mma = !e->cfg.flat_norm && (e->cfg.count_method == DGRM_COUNT_MMA || e->cfg.count_method == DGRM_COUNT_FP4);
  return mma ? std::max(1, kCountMergeSnps / e->kc_max) : 1;
}

// Resident chunks [from_s, nch) into S and merged groups [from_n, n_to) into N, over work items [w0, w1).
int contract_range(dgrm_engine* e, int w0, int w1, int from_s, int from_n, int nch, int n_to) {
  if (w1 <= w0) return DGRM_OK;
  const int merge = count_merge(e);
  for (int j0 = std::min(from_s, from_n); j0 < nch; ++j0) {
    if (j0 >= from_s) {
      const int rc = contract_s(e, w0, w1, j0);
      if (rc) return rc;
    }
    if (e->cfg.flat_norm || j0 < from_n || j0 >= n_to || j0 % merge != 0) continue;
    const int rc = contract_n(e, w0, w1, j0, std::min(n_to, j0 + merge));
    if (rc) return rc;
  }
  return DGRM_OK;
}

// Streamed build, while rows are still being pushed: contract the chunks that are resident so far over the prefix
// of the work list.  dgrm_finish then has less left to do for the first tile groups, whose outputs leave first.
int flush_prefix(dgrm_engine* e) {
  const int nwork = (int)e->work.size();
  if (e->pre_end < 0) {
    e->pre_end = 0;
    if (e->prefix_pct > 0 && nwork > 0) {
      std::vector<int> group_end, rg_group;
      stream_groups(e, &group_end, &rg_group);
      for (int end : group_end)
        if ((int64_t)end * 100 >= (int64_t)nwork * e->prefix_pct) {
          if (end < nwork) e->pre_end = end;
          break;
        }
    }
  }
  if (e->pre_end <= 0) return DGRM_OK;
  const int nch = e->chunks_in_super, merge = count_merge(e);
  CK(cudaEventRecord(e->prep_ev, e->prep));
  CK(cudaStreamWaitEvent(e->stream, e->prep_ev, 0));
  const int n_to = nch / merge * merge;
  const int rc = contract_range(e, 0, e->pre_end, e->pre_s, e->pre_n, nch, n_to);
  if (rc) return rc;
  e->pre_s = nch;
  e->pre_n = std::max(e->pre_n, n_to);
  return DGRM_OK;
}

// Contract resident chunks that have not been contracted yet (fp32 in TMEM per chunk, fp64 / int32
// read-modify-write of the tile across chunks).  S: chunks [launched_s, nch); N: merged groups from launched_n up to
// the last complete group (every chunk when `final` or `wrap`).
// `wrap`: every slot is taken and slot 0 is about to be overwritten.
// `final`: the last contraction of this GRM; with streaming bindings tiles then leave the device group by
// group while later groups are still being computed.
int flush_super(dgrm_engine* e, bool final, bool wrap) {
  const int nch = e->chunks_in_super;
  if (nch == 0) return DGRM_OK;
  const int merge = count_merge(e);
  const int ls = e->launched_s, ln = e->launched_n;
  const int n_to = (final || wrap) ? nch : nch / merge * merge;
  const int nwork = (int)e->work.size();
  const int pre_end = std::max(0, e->pre_end);   // work items [0, pre_end) are ahead by (pre_s, pre_n): flush_prefix
  // the operands of every chunk expanded so far must be complete before the MMA stream reads them
  CK(cudaEventRecord(e->prep_ev, e->prep));
  CK(cudaStreamWaitEvent(e->stream, e->prep_ev, 0));
  const bool stream_out = streaming_bound(e);
  if (nwork > 0) {
    // Tile groups.  Streaming outputs: one group per raster row group (stream_groups); otherwise one launch range
    // (two when a prefix of the work list is ahead of the rest).
    std::vector<int> group_end;
    std::vector<int> rg_group(e->rg_end.size(), 0);
    if (final && stream_out) {
      stream_groups(e, &group_end, &rg_group);
    } else {
      if (pre_end > 0 && pre_end < nwork) group_end.push_back(pre_end);
      group_end.push_back(nwork);
    }
    if (final && stream_out) {
      // Pair counts leave the device as 16-bit deficiencies m - N_ij when the per-individual missing counts prove
      // that every deficiency fits (m - N_ij <= missing_i + missing_j), else as int32; A as fp64 or fp32 (cfg).
      e->n_u16 = false;
      if (e->miss_valid && !e->cfg.flat_norm && !(e->saved && e->view != DGRM_VIEW_CURRENT) && e->d_miss) {
        std::vector<int32_t> miss((size_t)e->n);
        CK(cudaMemcpyAsync(miss.data(), e->d_miss, (size_t)e->n * sizeof(int32_t), cudaMemcpyDeviceToHost, e->prep));
        CK(cudaStreamSynchronize(e->prep));
        int64_t mx = 0;
        for (int32_t v : miss) mx = std::max<int64_t>(mx, v);
        e->n_u16 = 2 * mx <= 65534 && !getenv("DGRM_NO_U16");
      }
      // DGRM_TRANSPORT_FP64 (a zero-initialised config reads as that too): counts leave as doubles, no host widening
      const bool narrow_n = e->cfg.n_transport == DGRM_TRANSPORT_NARROW;
      // output jobs: one per (raster row group, streaming binding), in the order the row groups finish
      e->jobs.clear();
      e->jobs_issued = e->jobs_done = 0;
      e->d2h_bytes = 0;
      size_t need = 0;
      for (size_t rg = 0; rg < e->rg_row1.size(); ++rg) {
        const int row0 = rg == 0 ? 0 : e->rg_row1[rg - 1], row1 = std::min<int>(e->rg_row1[rg], (int)e->n);
        if (row0 >= row1) continue;
        for (int b = 0; b < e->nbind; ++b) {
          const dgrm_engine::Binding& bd = e->bind[b];
          if (!binding_streams(e, bd)) continue;
          dgrm_engine::OutJob j;
          j.b = b; j.row0 = row0; j.row1 = row1; j.ev = std::max(0, rg_group[rg]);
          j.lr0 = first_local_ge(row0, bd.g.nloc_r, bd.g.bs_r, bd.g.my_row, bd.g.R);
          j.lr1 = first_local_ge(row1, bd.g.nloc_r, bd.g.bs_r, bd.g.my_row, bd.g.R);
          j.lc1 = first_local_ge(row1, bd.g.nloc_c, bd.g.bs_c, bd.g.my_col, bd.g.C);
          if (j.lr1 <= j.lr0 || j.lc1 <= 0) continue;
          j.ntype = (bd.which == DGRM_FETCH_N && !e->cfg.flat_norm && narrow_n) ? (e->n_u16 ? 2 : 1)
                    : (bd.which == DGRM_FETCH_A && e->cfg.a_transport == DGRM_TRANSPORT_FP32) ? 3 : 0;
          j.direct = bd.dev != nullptr && j.ntype == 0;
          j.tiles = false;
          j.t0 = j.t1 = 0;
          need = std::max(need, (size_t)(j.lr1 - j.lr0) * (size_t)j.lc1 * sizeof(double));
          e->jobs.push_back(j);
        }
      }
      // tile bindings: one job per tile group (the group's tiles are contiguous in the store)
      if (e->ntbind > 0) {
        const int64_t per_work = (e->cfg.cta_group == 2 ? 2 : 1) * (int64_t)kTileElems;
        std::vector<dgrm_engine::OutJob> tj;
        int w0g = 0;
        for (size_t gi = 0; gi < group_end.size(); w0g = group_end[gi], ++gi) {
          for (int b = 0; b < e->ntbind; ++b) {
            dgrm_engine::OutJob j{};
            j.b = b; j.ev = (int)gi; j.tiles = true;
            j.ntype = (e->tbind[b].which == DGRM_FETCH_N && !e->cfg.flat_norm && narrow_n) ? (e->n_u16 ? 2 : 1)
                      : (e->tbind[b].which == DGRM_FETCH_A && e->cfg.a_transport == DGRM_TRANSPORT_FP32) ? 3 : 0;
            j.direct = e->tbind[b].pinned && j.ntype == 0;
            j.t0 = (int64_t)w0g * per_work;
            j.t1 = (int64_t)group_end[gi] * per_work;
            need = std::max(need, (size_t)(j.t1 - j.t0) * sizeof(double));
            tj.push_back(j);
          }
        }
        // merge with the rectangle jobs, ordered by the group that completes them
        std::vector<dgrm_engine::OutJob> all;
        size_t a = 0, b2 = 0;
        while (a < e->jobs.size() || b2 < tj.size()) {
          if (b2 >= tj.size() || (a < e->jobs.size() && e->jobs[a].ev <= tj[b2].ev)) all.push_back(e->jobs[a++]);
          else all.push_back(tj[b2++]);
        }
        e->jobs.swap(all);
      }
      if (need > e->out_stage_bytes) {
        CK(cudaStreamSynchronize(e->d2h_stream));
        CK(cudaStreamSynchronize(e->out_stream));
        for (int i = 0; i < dgrm_engine::kOutSlots; ++i) {
          free_dev(e->d_out_stage[i]);
          e->d_out_stage[i] = nullptr;
          if (e->h_out_stage[i]) cudaFreeHost(e->h_out_stage[i]);
          e->h_out_stage[i] = nullptr;
          CK(cudaMalloc(&e->d_out_stage[i], need));
          CK(cudaHostAlloc(&e->h_out_stage[i], need, cudaHostAllocDefault));
        }
        e->out_stage_bytes = need;
      }
    }
    // Within a tile group the chunks are contracted by one launch each (chunk-major inside the group): the
    // launch boundary re-synchronises the persistent CTAs every few hundred microseconds, which is what keeps
    // the A/B panels they share hot in L2.  (One tile-major launch over all chunks lets the CTAs drift apart
    // and re-read every panel from HBM: measured 233 GB per launch instead of ~40 GB.)
    int w0 = 0;
    for (size_t gi = 0; gi < group_end.size(); w0 = group_end[gi], ++gi) {
      const int w1 = group_end[gi];
      const bool ahead = w1 <= pre_end;   // this group belongs to the prefix that flush_prefix already advanced
      const int rc0 = contract_range(e, w0, w1, ahead ? e->pre_s : ls, ahead ? e->pre_n : ln, nch, n_to);
      if (rc0) return rc0;
      if (final && stream_out) {
        while (e->group_ev.size() <= gi) {
          cudaEvent_t ev;
          CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
          e->group_ev.push_back(ev);
        }
        CK(cudaEventRecord(e->group_ev[gi], e->stream));
        const int rc = issue_ready_jobs(e, (int)gi);
        if (rc) return rc;
      }
    }
    if (final && stream_out) {
      // everything is queued on the compute stream: now feed the copy engine and finish the staged jobs
      e->out_wait_ms = e->out_host_ms = 0;
      while (e->jobs_done < e->jobs.size()) {
        int rc = issue_ready_jobs(e, (int)group_end.size());
        if (rc) return rc;
        rc = finish_job(e, e->jobs_done);
        if (rc) return rc;
        e->jobs_done++;
      }
      {
        const int rc = deliver_exact_diagonal(e);
        if (rc) return rc;
      }
      if (getenv("DGRM_TRACE_OUT"))
        fprintf(stderr, "[dgrm] %zu output jobs: waited %.2f ms for the copy engine, %.2f ms in host conversion\n",
                e->jobs.size(), e->out_wait_ms, e->out_host_ms);
    }
  }
  e->launched_s = nch;
  e->launched_n = n_to;
  e->pre_s = nch;
  e->pre_n = n_to;
  // slots whose contractions are all queued can be overwritten once the MMA stream gets here: process_chunk makes
  // the decode stream wait for the slot's own event only, so the decode of the next super overlaps the rest of this one
  const int done = e->cfg.flat_norm ? nch : std::min(nch, n_to);
  for (; e->slots_evented < done; ++e->slots_evented) {
    cudaEvent_t& ev = e->slot_ev[e->slots_evented];
    if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ev, e->stream));
  }
  if (final || wrap) {
    e->chunks_in_super = 0;
    e->launched_s = e->launched_n = 0;
    e->pre_s = e->pre_n = 0;
    e->slots_evented = 0;
    e->supers_done++;
  }
  return DGRM_OK;
}

// append rows of a device-resident block to the chunk buffer, contracting whenever it fills up
int push_device_rows(dgrm_engine* e, const uint8_t* d_rows, int64_t pitch, int64_t nrows, const uint8_t* row_keep) {
  std::vector<int32_t> idx;
  int64_t kept = nrows;
  if (row_keep) {
    idx.reserve(nrows);
    for (int64_t r = 0; r < nrows; ++r)
      if (row_keep[r]) idx.push_back((int32_t)r);
    kept = (int64_t)idx.size();
  }
  if (e->h_fill > 0 || e->raw_rows > 0 || e->dev_fill > 0 || e->fill_dosage) {  // real-valued rows came before: they end their chunk
    int rc = flush_dosage(e);
    if (rc) return rc;
    if (e->fill_dosage) rc = process_chunk(e);
    if (rc) return rc;
  }
  if (e->m + e->fill + kept > e->m_cap)
    return fail(e, DGRM_E_INVALID, "more SNPs pushed (%lld) than dgrm_begin's capacity (%lld)",
                (long long)(e->m + e->fill + kept), (long long)e->m_cap);
  int64_t pos = 0;
  while (pos < kept) {
    const int take = (int)std::min<int64_t>(kept - pos, e->kc_max - e->fill);
    const int32_t* d_idx = nullptr;
    if (row_keep) {
      const int s = e->rowidx_i;
      e->rowidx_i ^= 1;
      CK(cudaEventSynchronize(e->rowidx_free[s]));
      memcpy(e->h_rowidx[s], idx.data() + pos, (size_t)take * sizeof(int32_t));
      CK(cudaMemcpyAsync(e->d_rowidx[s], e->h_rowidx[s], (size_t)take * sizeof(int32_t), cudaMemcpyHostToDevice,
                         e->prep));
      d_idx = e->d_rowidx[s];
    }
    {
      const int nbytes = (int)(e->d_src_ind ? e->b_row : e->b_fam);
      uint8_t* dst = e->d_chunk + (int64_t)e->fill * e->chunk_pitch;
      const uint64_t align = (uint64_t)(uintptr_t)d_rows | (uint64_t)pitch;
      const int vec = e->d_src_ind ? 1 : (align % 16 == 0 ? 16 : align % 8 == 0 ? 8 : align % 4 == 0 ? 4 : 1);
      SnpStatsOut so{e->d_freq, e->d_nn, e->d_p1, e->d_p2, e->d_sd, e->d_bad_count, e->d_bad_list, dgrm_engine::kBadCap};
      const int64_t snp0 = e->m + e->fill;
      if (e->d_chunkT) {
        // GRM mode: one pass produces the transposed chunk (16-SNP words per individual) and the per-row counts;
        // sources that cannot be read with aligned 32-bit loads (or need the --keep compaction of individuals) are
        // first gathered into the row-major chunk buffer
        Prof pr(e, DGRM_K_GATHER, e->prep);
        if (e->fill == 0) CK(cudaMemsetAsync(e->d_counts, 0, (size_t)e->kc_max * 4 * sizeof(int32_t), e->prep));
        const uint8_t* tsrc = d_rows;
        int64_t tpitch = pitch, trow0 = pos;
        const int32_t* tidx = d_idx;
        if (vec < 4) {
          const int per_row = (nbytes / vec + 255) / 256;
          const dim3 grid((unsigned)std::max(1, std::min(per_row, 64)), (unsigned)take);
          bed_gather_kernel<uint8_t><<<grid, 256, 0, e->prep>>>(d_rows, pitch, d_idx, pos, take, dst, e->chunk_pitch, nbytes,
                                                                e->d_src_ind, (int)e->n_row);
          CK(cudaGetLastError());
          tsrc = dst; tpitch = e->chunk_pitch; trow0 = 0; tidx = nullptr;
        }
        const int tbytes = (int)(e->d_src_ind ? e->b_row : nbytes);
        const int nwords = (int)((e->n_row + 15) / 16);
        const int groups = ((e->fill + take - 1) >> 4) - (e->fill >> 4) + 1;
        bed_decode_t_kernel<<<dim3((unsigned)((nwords + 255) / 256), (unsigned)groups), 256, 0, e->prep>>>(
            tsrc, tpitch, tidx, trow0, e->fill, take, e->d_chunkT, e->t_pitch, tbytes, (int)e->n_row, e->d_counts);
        CK(cudaGetLastError());
        bed_stats_finish_kernel<<<(take + 255) / 256, 256, 0, e->prep>>>(e->d_counts, e->fill, take, (int)e->n_row, snp0, so,
                                                                        e->d_zt);
        CK(cudaGetLastError());
      } else if (vec > 1 && e->fused_decode) {
        // one pass: row selection / alignment copy + per-SNP counts and statistics, one warp per row
        Prof pr(e, DGRM_K_GATHER, e->prep);
        const int blocks = (take * 32 + 255) / 256;
#define DGRM_DECODE(T)                                                                                          \
  bed_decode_kernel<T><<<blocks, 256, 0, e->prep>>>(d_rows, pitch, d_idx, pos, take, dst, e->chunk_pitch, nbytes, \
                                                    (int)e->n_row, snp0, so, e->d_zt, e->fill)
        if (vec == 16) DGRM_DECODE(uint4);
        else if (vec == 8) DGRM_DECODE(uint2);
        else DGRM_DECODE(uint32_t);
#undef DGRM_DECODE
        CK(cudaGetLastError());
      } else {
        // --keep compaction of individuals, or rows that are not even 4-byte aligned: gather, then count
        {
          Prof pr(e, DGRM_K_GATHER, e->prep);
          const int per_row = (nbytes / vec + 255) / 256;
          const dim3 grid((unsigned)std::max(1, std::min(per_row, 64)), (unsigned)take);
#define DGRM_GATHER(T)                                                                                            \
  bed_gather_kernel<T><<<grid, 256, 0, e->prep>>>(d_rows, pitch, d_idx, pos, take, dst, e->chunk_pitch, nbytes, \
                                                  e->d_src_ind, (int)e->n_row)
          if (vec == 16) DGRM_GATHER(uint4);
          else if (vec == 8) DGRM_GATHER(uint2);
          else if (vec == 4) DGRM_GATHER(uint32_t);
          else DGRM_GATHER(uint8_t);
#undef DGRM_GATHER
          CK(cudaGetLastError());
        }
        Prof pr(e, DGRM_K_COUNT, e->prep);
        bed_count_stats_kernel<<<(take * 32 + 255) / 256, 256, 0, e->prep>>>(dst, e->chunk_pitch, take, (int)e->n_row, snp0, so,
                                                                            e->d_zt + (int64_t)e->fill * 4);
        CK(cudaGetLastError());
      }
      if (row_keep) CK(cudaEventRecord(e->rowidx_free[e->rowidx_i ^ 1], e->prep));
    }
    e->fill += take;
    pos += take;
    if (e->fill == e->kc_max) {
      const int rc = process_chunk(e);
      if (rc) return rc;
    }
  }
  return DGRM_OK;
}

// Staging for real-valued rows: two pinned host buffers of 64 MB of dosages (the caller's rows are copied here, so the
// row-at-a-time pushes of a variant parser still cross the bus in large pieces) and a device buffer of dev_cap rows, up
// to one chunk: the statistics kernel runs one sequential chain per row, so a launch should carry many rows.
int ensure_dosage_staging(dgrm_engine* e) {
  if (e->dev_cap > 0) return DGRM_OK;
  const int64_t row_bytes = std::max<int64_t>(1, e->n * (int64_t)sizeof(double));
  const int64_t hrows = std::max<int64_t>(1, std::min<int64_t>((64ll << 20) / row_bytes, e->kc_max));
  int64_t drows = ((int64_t)e->dosage_stage_mb << 20) / (row_bytes + e->n);
  drows = std::max<int64_t>(256, std::min<int64_t>(drows / 256 * 256, e->kc_max));
  for (int i = 0; i < 2; ++i) {
    if (cudaHostAlloc(&e->h_dos[i], (size_t)hrows * e->n * sizeof(double), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc(&e->h_dmiss[i], (size_t)hrows * e->n, cudaHostAllocDefault) != cudaSuccess) {
      release_dosage_staging(e);
      return fail(e, DGRM_E_NOMEM, "dgrm_push_dosage_host: no pinned memory for %lld staged rows", (long long)hrows);
    }
    if (!e->dos_copied[i]) CK(cudaEventCreateWithFlags(&e->dos_copied[i], cudaEventDisableTiming));
  }
  if (!e->dos_read) CK(cudaEventCreateWithFlags(&e->dos_read, cudaEventDisableTiming));
  if (!e->dos_ready) CK(cudaEventCreateWithFlags(&e->dos_ready, cudaEventDisableTiming));
  if (cudaMalloc(&e->d_dos, (size_t)drows * e->n * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&e->d_dmiss, (size_t)drows * e->n) != cudaSuccess ||
      cudaMalloc(&e->d_dos_mean, (size_t)drows * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&e->d_dos_sd, (size_t)drows * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    release_dosage_staging(e);   // nothing half-allocated stays behind
    return fail(e, DGRM_E_NOMEM, "dgrm_push_dosage_host: no device memory for %lld staged rows (DGRM_DOSAGE_STAGE_MB = %d)",
                (long long)drows, e->dosage_stage_mb);
  }
  e->h_cap = (int)hrows;
  e->dev_cap = (int)drows;
  e->h_fill = e->dev_fill = 0;
  e->h_any_missing = e->dev_any_missing = false;
  return DGRM_OK;
}

// The dev_fill rows of the device staging buffer: missing mask folded in, statistics, then expansion into the open
// operand slot at column `fill`.  Full buffers are multiples of 256 rows, so the open chunk only ever ends on a
// partial k-block when this is the last batch before something else (packed rows, dgrm_finish): that batch closes
// the chunk.
int process_dosage_device(dgrm_engine* e) {
  const int rows = e->dev_fill;
  if (rows == 0) return DGRM_OK;
  CK(cudaEventRecord(e->dos_ready, e->copy_stream));
  CK(cudaStreamWaitEvent(e->prep, e->dos_ready, 0));
  {
    Prof pr(e, DGRM_K_COUNT, e->prep);
    if (e->dev_any_missing) {
      const int64_t count = (int64_t)rows * e->n;
      dosage_fold_mask_kernel<<<(unsigned)std::min<int64_t>((count + 255) / 256, 148 * 16), 256, 0, e->prep>>>(e->d_dos, e->d_dmiss,
                                                                                                             count);
      CK(cudaGetLastError());
    }
    DosageStatsOut so{e->d_freq, e->d_nn, e->d_p1, e->d_p2, e->d_sd, e->d_bad_count + 3};
    dosage_stats_kernel<<<(rows + kStatRows - 1) / kStatRows, kStatThreads, sizeof(DosageStatSmem), e->prep>>>(
        e->d_dos, e->n, rows, (int)e->n, e->m + e->fill, so, e->d_dos_mean, e->d_dos_sd);
    CK(cudaGetLastError());
  }
  e->dev_fill = 0;
  e->dev_any_missing = false;
  const bool split = e->cfg.operand == DGRM_OPERAND_FP16X2;
  const bool bf16 = e->cfg.operand == DGRM_OPERAND_BF16;
  const int qk = e->cfg.flat_norm ? 0
                 : e->cfg.count_method == DGRM_COUNT_MMA ? 1
                 : e->cfg.count_method == DGRM_COUNT_FP4 ? 2
                 : e->cfg.count_method == DGRM_COUNT_POPC ? 3 : 0;
  int pos = 0;
  while (pos < rows) {
    if (e->fill == 0) {
      const int rc = acquire_slot(e);
      if (rc) return rc;
    }
    const int take = std::min(rows - pos, e->kc_max - e->fill);
    const int64_t col = (int64_t)e->chunks_in_super * e->kc_max + e->fill;
    DosageExpandArgs a;
    a.dos = e->d_dos + (int64_t)pos * e->n; a.pitch = e->n;
    a.mean = e->d_dos_mean + pos; a.sd = e->d_dos_sd + pos;
    a.kc = take; a.n = (int)e->n;
    a.Z = e->d_Z + col; a.z_pitch = e->z_pitch; a.lo_off = (int64_t)e->slots * e->kc_max;
    a.Q = qk == 1 ? e->d_Q + col : nullptr; a.q_pitch = e->q_pitch;
    a.Q4 = qk == 2 ? e->d_Q4 + col / 2 : nullptr; a.q4_pitch = e->q4_pitch;
    a.MB = qk == 3 ? e->d_MB + col / 32 : nullptr; a.mb_pitch = e->mb_pitch;
    a.diag_part = e->d_diag_part + (int64_t)(e->fill / kExpK) * e->n_pad; a.n_pad = (int)e->n_pad;
    const dim3 grid((unsigned)(e->n_pad / kExpI), (unsigned)(round_up(take, 256) / kExpK));
    {
      Prof pr(e, DGRM_K_EXPAND, e->prep);
#define DGRM_DOSAGE(B, S, K) dosage_expand_kernel<B, S, K><<<grid, 256, sizeof(DosageSmem), e->prep>>>(a)
#define DGRM_DOSAGE_Q(B, S)                                                     \
  do {                                                                          \
    if (qk == 0) DGRM_DOSAGE(B, S, 0); else if (qk == 1) DGRM_DOSAGE(B, S, 1);  \
    else if (qk == 2) DGRM_DOSAGE(B, S, 2); else DGRM_DOSAGE(B, S, 3);          \
  } while (0)
      if (bf16) DGRM_DOSAGE_Q(true, false);
      else if (split) DGRM_DOSAGE_Q(false, true);
      else DGRM_DOSAGE_Q(false, false);
#undef DGRM_DOSAGE_Q
#undef DGRM_DOSAGE
      CK(cudaGetLastError());
    }
    e->fill += take;
    e->fill_dosage = true;
    pos += take;
    if (e->fill == e->kc_max || e->fill % 256 != 0) {
      const int rc = process_chunk(e);  // diagonal + bookkeeping of the finished chunk
      if (rc) return rc;
    }
  }
  CK(cudaEventRecord(e->dos_read, e->prep));
  // per-individual missing counts are not kept for real-valued rows: streamed pair counts travel as int32
  if (!e->cfg.flat_norm) e->miss_valid = false;
  return DGRM_OK;
}

// The h_fill rows of the current pinned buffer go to the device staging buffer (asynchronously, copy stream); every
// time that buffer is full it is processed.
int send_dosage(dgrm_engine* e) {
  const int rows = e->h_fill;
  if (rows == 0) return DGRM_OK;
  const int s = e->dos_i;
  const bool with_miss = e->h_any_missing;
  const int64_t n = e->n;
  int off = 0;
  while (off < rows) {
    // the kernels of the previous batch have read the device buffer (also when this batch was opened by unpacked BGEN
    // rows, which are written on the decode stream and say nothing to the copy stream)
    CK(cudaStreamWaitEvent(e->copy_stream, e->dos_read, 0));
    const int take = std::min(rows - off, e->dev_cap - e->dev_fill);
    CK(cudaMemcpyAsync(e->d_dos + (int64_t)e->dev_fill * n, e->h_dos[s] + (int64_t)off * n, (size_t)take * n * sizeof(double),
                       cudaMemcpyHostToDevice, e->copy_stream));
    if (with_miss) {
      if (!e->dev_any_missing && e->dev_fill > 0) CK(cudaMemsetAsync(e->d_dmiss, 0, (size_t)e->dev_fill * n, e->copy_stream));
      CK(cudaMemcpyAsync(e->d_dmiss + (int64_t)e->dev_fill * n, e->h_dmiss[s] + (int64_t)off * n, (size_t)take * n,
                         cudaMemcpyHostToDevice, e->copy_stream));
      e->dev_any_missing = true;
    } else if (e->dev_any_missing) {
      CK(cudaMemsetAsync(e->d_dmiss + (int64_t)e->dev_fill * n, 0, (size_t)take * n, e->copy_stream));
    }
    e->dev_fill += take;
    off += take;
    if (e->dev_fill == e->dev_cap) {
      const int rc = process_dosage_device(e);
      if (rc) return rc;
    }
  }
  CK(cudaEventRecord(e->dos_copied[s], e->copy_stream));
  e->dos_i ^= 1;
  e->h_fill = 0;
  e->h_any_missing = false;
  return DGRM_OK;
}

// Raw staging for BGEN probability blocks: two pinned pieces and one device piece of raw_cap bytes (at least one block).
int ensure_bgen_staging(dgrm_engine* e, int64_t block_bytes) {
  const int64_t want = std::max<int64_t>(64ll << 20, block_bytes);
  if (e->raw_cap >= want) return DGRM_OK;
  CK(cudaStreamSynchronize(e->copy_stream));
  CK(cudaStreamSynchronize(e->prep));
  for (int i = 0; i < 2; ++i) {
    if (e->h_raw[i]) cudaFreeHost(e->h_raw[i]);
    if (e->h_boff[i]) cudaFreeHost(e->h_boff[i]);
    e->h_raw[i] = nullptr; e->h_boff[i] = nullptr;
  }
  free_dev(e->d_raw); e->d_raw = nullptr;
  free_dev(e->d_boff); e->d_boff = nullptr;
  e->raw_cap = 0;
  const int rows_cap = 16384;
  for (int i = 0; i < 2; ++i) {
    if (!e->raw_copied[i]) CK(cudaEventCreateWithFlags(&e->raw_copied[i], cudaEventDisableTiming));
  }
  if (!e->raw_read) CK(cudaEventCreateWithFlags(&e->raw_read, cudaEventDisableTiming));
  bool ok = true;
  for (int i = 0; i < 2 && ok; ++i)
    ok = cudaHostAlloc(&e->h_raw[i], (size_t)want, cudaHostAllocDefault) == cudaSuccess &&
         cudaHostAlloc(&e->h_boff[i], (size_t)(rows_cap + 1) * sizeof(int64_t), cudaHostAllocDefault) == cudaSuccess;
  ok = ok && cudaMalloc(&e->d_raw, (size_t)want) == cudaSuccess &&
       cudaMalloc(&e->d_boff, (size_t)(rows_cap + 1) * sizeof(int64_t)) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    for (int i = 0; i < 2; ++i) {   // nothing half-allocated stays behind
      if (e->h_raw[i]) cudaFreeHost(e->h_raw[i]);
      if (e->h_boff[i]) cudaFreeHost(e->h_boff[i]);
      e->h_raw[i] = nullptr; e->h_boff[i] = nullptr;
    }
    free_dev(e->d_raw); e->d_raw = nullptr;
    free_dev(e->d_boff); e->d_boff = nullptr;
    return fail(e, DGRM_E_NOMEM, "dgrm_push_bgen_blocks_host: no memory for %lld bytes of staged blocks", (long long)want);
  }
  e->raw_cap = want;
  e->raw_rows_cap = rows_cap;
  return DGRM_OK;
}

// The raw_rows blocks of the current pinned piece go to the device and are unpacked into the device staging buffer,
// which is processed every time it is full.
int send_bgen(dgrm_engine* e) {
  const int rows = e->raw_rows;
  if (rows == 0) return DGRM_OK;
  const int s = e->raw_i;
  CK(cudaStreamWaitEvent(e->copy_stream, e->raw_read, 0));   // the unpack kernels of the previous piece have read d_raw
  CK(cudaMemcpyAsync(e->d_raw, e->h_raw[s], (size_t)e->raw_fill, cudaMemcpyHostToDevice, e->copy_stream));
  CK(cudaMemcpyAsync(e->d_boff, e->h_boff[s], (size_t)(rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, e->copy_stream));
  CK(cudaEventRecord(e->raw_copied[s], e->copy_stream));
  CK(cudaStreamWaitEvent(e->prep, e->raw_copied[s], 0));
  e->raw_i ^= 1;
  e->raw_rows = 0;
  e->raw_fill = 0;
  const int64_t n = e->n;
  int off = 0;
  while (off < rows) {
    const int take = std::min(rows - off, e->dev_cap - e->dev_fill);
    {
      Prof pr(e, DGRM_K_GATHER, e->prep);
      BgenUnpackArgs a;
      a.raw = e->d_raw; a.offsets = e->d_boff; a.row0 = off; a.rows = take; a.layout = e->raw_layout; a.n_samples = (int)e->raw_samples;
      a.cols = e->bgen_all ? nullptr : e->d_bgen_cols; a.n = (int)n;
      a.dos = e->d_dos + (int64_t)e->dev_fill * n; a.errors = e->d_bad_count + 4;
      bgen_unpack_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)((take + kUnpackRows - 1) / kUnpackRows)), 256, 0, e->prep>>>(a);
      CK(cudaGetLastError());
      if (e->dev_any_missing) CK(cudaMemsetAsync(e->d_dmiss + (int64_t)e->dev_fill * n, 0, (size_t)take * n, e->prep));
    }
    e->dev_fill += take;
    off += take;
    if (e->dev_fill == e->dev_cap) {
      const int rc = process_dosage_device(e);
      if (rc) return rc;
    }
  }
  CK(cudaEventRecord(e->raw_read, e->prep));
  return DGRM_OK;
}

// everything staged so far reaches the operands
int flush_dosage(dgrm_engine* e) {
  int rc = send_dosage(e);
  if (rc) return rc;
  rc = send_bgen(e);
  if (rc) return rc;
  return process_dosage_device(e);
}

// (Re)allocate the resident operand buffers for `slots` chunks of kc_max SNPs.
int ensure_operands(dgrm_engine* e, int slots) {
  if (slots == e->slots && e->d_Z) return DGRM_OK;
  CK(cudaStreamSynchronize(e->prep));
  CK(cudaStreamSynchronize(e->stream));
  free_dev(e->d_Z); e->d_Z = nullptr;
  free_dev(e->d_Q); e->d_Q = nullptr;
  free_dev(e->d_Q4); e->d_Q4 = nullptr;
  free_dev(e->d_MB); e->d_MB = nullptr;
  const bool split = e->cfg.operand == DGRM_OPERAND_FP16X2;
  const int64_t cols = (int64_t)slots * e->kc_max;
  e->slots = slots;
  e->z_pitch = cols * (split ? 2 : 1);
  e->q_pitch = cols;
  e->mb_pitch = cols / 32;
  CK(cudaMalloc(&e->d_Z, (size_t)e->n_pad * e->z_pitch * sizeof(uint16_t)));
  int rc = make_tmap(e, &e->tmap_z, e->d_Z, 2, e->z_pitch, e->n_pad, e->z_pitch * 2);
  if (rc) return rc;
  if (!e->cfg.flat_norm && e->cfg.count_method == DGRM_COUNT_MMA) {
    CK(cudaMalloc(&e->d_Q, (size_t)e->n_pad * e->q_pitch));
    rc = make_tmap(e, &e->tmap_q, e->d_Q, 1, e->q_pitch, e->n_pad, e->q_pitch);
    if (rc) return rc;
  }
  if (!e->cfg.flat_norm && e->cfg.count_method == DGRM_COUNT_FP4) {
    e->q4_pitch = cols / 2;
    CK(cudaMalloc(&e->d_Q4, (size_t)e->n_pad * e->q4_pitch));
    rc = make_tmap(e, &e->tmap_q4, e->d_Q4, 1, e->q4_pitch, e->n_pad, e->q4_pitch);
    if (rc) return rc;
  }
  if (!e->cfg.flat_norm && e->cfg.count_method == DGRM_COUNT_POPC)
    CK(cudaMalloc(&e->d_MB, (size_t)e->n_pad * e->mb_pitch * sizeof(uint32_t)));
  return DGRM_OK;
}

// Open a PLINK .bed and validate its 3-byte header with the reference's checks and messages
// (genotype.cpp:578-605).
int open_bed(dgrm_engine* e, const char* bed_path, FILE** out) {
  FILE* f = fopen(bed_path, "rb");
  if (!f) return fail(e, DGRM_E_IO, "Error: The file [ %s ] does not exist or cannot be opened.", bed_path);
  unsigned char header[3];
  if (fread(header, 1, 3, f) != 3) {
    fclose(f);
    return fail(e, DGRM_E_BADBED, "Error: The data file is not a valid binary BED > v1.0");
  }
  // genotype.cpp:581-605
  if (header[0] != 0x6C || header[1] != 0x1B) {
    fclose(f);
    return fail(e, DGRM_E_BADBED, "Error: The data file is not a valid binary BED > v1.0");
  }
  if (header[2] == 0x0) {
    fclose(f);
    return fail(e, DGRM_E_BADBED,
                "Error: The BED file is in Individual-major mode. This mode is not currently supported. Please, read "
                "PLINK manual.");
  }
  if (header[2] != 0x1) {
    fclose(f);
    return fail(e, DGRM_E_BADBED,
                "Error: The data file is not a valid binary BED.\nIt is not specified whether the file is in SNP-major "
                "o Individual-major mode.");
  }
  *out = f;
  return DGRM_OK;
}

int check_ready(dgrm_engine* e, bool need_finished) {
  if (!e) return DGRM_E_INVALID;
  if (!e->have_ind) return fail(e, DGRM_E_INVALID, "dgrm_set_individuals has not been called");
  if (!e->begun) return fail(e, DGRM_E_INVALID, "dgrm_begin has not been called");
  if (need_finished && !e->finished) return fail(e, DGRM_E_INVALID, "dgrm_finish has not been called");
  return DGRM_OK;
}

int numroc(int n, int nb, int iproc, int nprocs) {
  // ScaLAPACK numroc_ with isrcproc = 0 (matrix.cpp:1748-1749)
  const int nblocks = n / nb;
  int r = (nblocks / nprocs) * nb;
  const int extra = nblocks % nprocs;
  if (iproc < extra) r += nb;
  else if (iproc == extra) r += n % nb;
  return r;
}

}  // namespace

// ================================================================================================
extern "C" {

int dgrm_abi_version(void) { return DGRM_ABI_VERSION; }

int dgrm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int dgrm_validate_bed_header(dgrm_engine* e, const char* bed_path) {
  if (!e || !bed_path) return DGRM_E_INVALID;
  FILE* f = nullptr;
  const int rc = open_bed(e, bed_path, &f);
  if (f) fclose(f);
  return rc;
}

int dgrm_default_config(dgrm_config* cfg) {
  if (!cfg) return DGRM_E_INVALID;
  memset(cfg, 0, sizeof *cfg);
  cfg->device = 0;
  cfg->rank = 0;
  cfg->world = 1;
  cfg->operand = DGRM_OPERAND_FP16;
  cfg->count_method = DGRM_COUNT_FP4;
  cfg->flat_norm = 0;
  cfg->chunk_snps = 16384;
  cfg->cta_group = 2;
  cfg->profile = 0;
  cfg->resident_chunks = 0;  // auto
  cfg->resident_mb = 16384;
  cfg->reserve_sms = -1;
  cfg->a_transport = DGRM_TRANSPORT_FP64;
  cfg->n_transport = DGRM_TRANSPORT_NARROW;
  return DGRM_OK;
}

const char* dgrm_last_error(const dgrm_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int dgrm_create(const dgrm_config* cfg, dgrm_engine** out) {
  dgrm_engine* e = nullptr;
  if (!cfg || !out) return fail(nullptr, DGRM_E_INVALID, "dgrm_create: null argument");
  *out = nullptr;
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world)
    return fail(nullptr, DGRM_E_INVALID, "dgrm_create: bad rank/world %d/%d", cfg->rank, cfg->world);
  if (cfg->chunk_snps < 256 || cfg->chunk_snps % 256 != 0 || cfg->chunk_snps > 32768)
    return fail(nullptr, DGRM_E_INVALID, "dgrm_create: chunk_snps must be a multiple of 256 in [256, 32768]");
  if (cfg->cta_group != 1 && cfg->cta_group != 2)
    return fail(nullptr, DGRM_E_INVALID, "dgrm_create: cta_group must be 1 or 2");
  if (cfg->resident_chunks < 0 || cfg->resident_chunks > kMaxChunks || cfg->resident_mb < 0)
    return fail(nullptr, DGRM_E_INVALID, "dgrm_create: resident_chunks must be in [0, %d]", kMaxChunks);
  if (cfg->operand < 0 || cfg->operand > DGRM_OPERAND_FP16X2 || cfg->count_method < 0 ||
      cfg->count_method > DGRM_COUNT_FP4)
    return fail(nullptr, DGRM_E_INVALID, "dgrm_create: bad operand/count_method");
  if (cfg->a_transport != DGRM_TRANSPORT_FP64 && cfg->a_transport != DGRM_TRANSPORT_FP32)
    return fail(nullptr, DGRM_E_INVALID, "dgrm_create: bad a_transport");
  if (cfg->n_transport != DGRM_TRANSPORT_FP64 && cfg->n_transport != DGRM_TRANSPORT_NARROW && cfg->n_transport != 0)
    return fail(nullptr, DGRM_E_INVALID, "dgrm_create: bad n_transport");
  int ndev = 0;
  cudaError_t st = cudaGetDeviceCount(&ndev);
  if (st != cudaSuccess || ndev == 0)
    return fail(nullptr, DGRM_E_CUDA, "no CUDA device: %s (this library has no CPU fallback)",
                cudaGetErrorString(st));
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(nullptr, DGRM_E_INVALID, "dgrm_create: device %d out of range (%d devices)", cfg->device, ndev);
  cudaDeviceProp prop;
  st = cudaGetDeviceProperties(&prop, cfg->device);
  if (st != cudaSuccess) return fail(nullptr, DGRM_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(st));
  if (prop.major != 10)
    return fail(nullptr, DGRM_E_CUDA, "device %d is sm_%d%d; libdgrm is built for sm_100a only (no fallback path)",
                cfg->device, prop.major, prop.minor);
  e = new dgrm_engine();
  e->cfg = *cfg;
  e->sm_count = prop.multiProcessorCount;
  e->kc_max = cfg->chunk_snps;
  if (const char* pol = getenv("DGRM_L2_POLICY")) e->l2_policy = atoi(pol);
  if (const char* ov = getenv("DGRM_OVERLAP")) e->overlap = atoi(ov);
  if (const char* fd = getenv("DGRM_FUSED_DECODE")) e->fused_decode = atoi(fd);
  if (const char* kt = getenv("DGRM_K1_TRANSPOSED")) e->k1_transposed = atoi(kt);
  if (const char* kf = getenv("DGRM_EXPAND_KFAST")) e->expand_k_fast = atoi(kf);
  if (const char* cs = getenv("DGRM_EXPAND_CS")) e->expand_stream_stores = atoi(cs);
  if (const char* pp = getenv("DGRM_PREFIX_PCT")) e->prefix_pct = std::max(0, std::min(90, atoi(pp)));
  if (const char* ds = getenv("DGRM_DOSAGE_STAGE_MB")) e->dosage_stage_mb = std::max(1, std::min(65536, atoi(ds)));
  e->reserve_sms = cfg->reserve_sms > 0 ? cfg->reserve_sms : 0;
  if (const char* rs = getenv("DGRM_RESERVE_SMS")) e->reserve_sms = atoi(rs);
  e->reserve_sms = std::max(0, std::min(e->reserve_sms, e->sm_count - 2));
#define CKC(call)                                                                          \
  do {                                                                                     \
    cudaError_t _s = (call);                                                               \
    if (_s != cudaSuccess) {                                                               \
      fail(nullptr, DGRM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(_s));          \
      delete e;                                                                            \
      return DGRM_E_CUDA;                                                                  \
    }                                                                                      \
  } while (0)
  CKC(cudaSetDevice(cfg->device));
  CKC(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&e->prep, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&e->out_stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
  for (int i = 0; i < dgrm_engine::kOutSlots; ++i) {
    CKC(cudaEventCreateWithFlags(&e->stage_packed[i], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&e->stage_copied_out[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < 2; ++i) {
    CKC(cudaEventCreateWithFlags(&e->stage_copied[i], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&e->stage_free[i], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&e->rowidx_free[i], cudaEventDisableTiming));
  }
  CKC(cudaEventCreateWithFlags(&e->producer_ev, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&e->bed_ev, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&e->prep_ev, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&e->mma_ev, cudaEventDisableTiming));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) {
    fail(nullptr, DGRM_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    delete e;
    return DGRM_E_CUDA;
  }
  e->encode = (EncodeTiledFn)fn;
  // The output kernels run on their own stream while grm_syrk CTAs (129 KB of shared memory, 168 registers x 256
  // threads each) occupy every SM: they can only be co-scheduled onto those SMs when they ask for the same L1/shared-memory split.
  CKC(cudaFuncSetAttribute(pack_rows_kernel<double>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           (int)cudaSharedmemCarveoutMaxShared));
  CKC(cudaFuncSetAttribute(pack_rows_kernel<int>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           (int)cudaSharedmemCarveoutMaxShared));
  CKC(cudaFuncSetAttribute(pack_rows_kernel<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           (int)cudaSharedmemCarveoutMaxShared));
  CKC(cudaFuncSetAttribute(pack_rows_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           (int)cudaSharedmemCarveoutMaxShared));
  // kernel attributes are per device: every engine sets them for its own (a process may drive several GPUs)
#define DGRM_SYRK_ATTR(CG, KIND, T)                                                                          \
  CKC(cudaFuncSetAttribute(grm_syrk_kernel<CG, KIND, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SyrkCfg<CG>::kSmemBytes))
  DGRM_SYRK_ATTR(1, 0, double); DGRM_SYRK_ATTR(2, 0, double);
  DGRM_SYRK_ATTR(1, 1, int); DGRM_SYRK_ATTR(2, 1, int);
  DGRM_SYRK_ATTR(1, 2, int); DGRM_SYRK_ATTR(2, 2, int);
#undef DGRM_SYRK_ATTR
#define DGRM_EXPAND_ATTR(B, S, K)                                                                          \
  CKC(cudaFuncSetAttribute(bed_expand_kernel<B, S, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExpandSmem)))
  DGRM_EXPAND_ATTR(false, false, 0); DGRM_EXPAND_ATTR(false, false, 1); DGRM_EXPAND_ATTR(false, false, 2); DGRM_EXPAND_ATTR(false, false, 3);
  DGRM_EXPAND_ATTR(true, false, 0); DGRM_EXPAND_ATTR(true, false, 1); DGRM_EXPAND_ATTR(true, false, 2); DGRM_EXPAND_ATTR(true, false, 3);
  DGRM_EXPAND_ATTR(false, true, 0); DGRM_EXPAND_ATTR(false, true, 1); DGRM_EXPAND_ATTR(false, true, 2); DGRM_EXPAND_ATTR(false, true, 3);
#undef DGRM_EXPAND_ATTR
#define DGRM_EXPAND_T_ATTR(B, S, K)                                                                          \
  CKC(cudaFuncSetAttribute(bed_expand_t_kernel<B, S, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExpandTSmem<K, S>)))
  DGRM_EXPAND_T_ATTR(false, false, 0); DGRM_EXPAND_T_ATTR(false, false, 1); DGRM_EXPAND_T_ATTR(false, false, 2); DGRM_EXPAND_T_ATTR(false, false, 3);
  DGRM_EXPAND_T_ATTR(true, false, 0); DGRM_EXPAND_T_ATTR(true, false, 1); DGRM_EXPAND_T_ATTR(true, false, 2); DGRM_EXPAND_T_ATTR(true, false, 3);
  DGRM_EXPAND_T_ATTR(false, true, 0); DGRM_EXPAND_T_ATTR(false, true, 1); DGRM_EXPAND_T_ATTR(false, true, 2); DGRM_EXPAND_T_ATTR(false, true, 3);
#undef DGRM_EXPAND_T_ATTR
#define DGRM_DOSAGE_ATTR(B, S, K)                                                                          \
  CKC(cudaFuncSetAttribute(dosage_expand_kernel<B, S, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DosageSmem)))
  DGRM_DOSAGE_ATTR(false, false, 0); DGRM_DOSAGE_ATTR(false, false, 1); DGRM_DOSAGE_ATTR(false, false, 2); DGRM_DOSAGE_ATTR(false, false, 3);
  DGRM_DOSAGE_ATTR(true, false, 0); DGRM_DOSAGE_ATTR(true, false, 1); DGRM_DOSAGE_ATTR(true, false, 2); DGRM_DOSAGE_ATTR(true, false, 3);
  DGRM_DOSAGE_ATTR(false, true, 0); DGRM_DOSAGE_ATTR(false, true, 1); DGRM_DOSAGE_ATTR(false, true, 2); DGRM_DOSAGE_ATTR(false, true, 3);
#undef DGRM_DOSAGE_ATTR
  CKC(cudaFuncSetAttribute(dosage_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DosageStatSmem)));
#undef CKC
  *out = e;
  return DGRM_OK;
}

void dgrm_destroy(dgrm_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  release_geometry(e);
  release_stats(e);
  free_dev(e->d_bad_count);
  free_dev(e->d_bad_list);
  for (int k = 0; k < DGRM_K_NUM; ++k)
    for (auto& p : e->ev[k]) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto& p : e->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (int i = 0; i < 2; ++i) {
    if (e->stage_copied[i]) cudaEventDestroy(e->stage_copied[i]);
    if (e->stage_free[i]) cudaEventDestroy(e->stage_free[i]);
    if (e->rowidx_free[i]) cudaEventDestroy(e->rowidx_free[i]);
  }
  if (e->producer_ev) cudaEventDestroy(e->producer_ev);
  for (int i = 0; i < 2; ++i)
    if (e->dos_copied[i]) cudaEventDestroy(e->dos_copied[i]);
  if (e->dos_read) cudaEventDestroy(e->dos_read);
  if (e->dos_ready) cudaEventDestroy(e->dos_ready);
  for (int i = 0; i < 2; ++i)
    if (e->raw_copied[i]) cudaEventDestroy(e->raw_copied[i]);
  if (e->raw_read) cudaEventDestroy(e->raw_read);
  if (e->bed_ev) cudaEventDestroy(e->bed_ev);
  free_dev(e->d_bed);
  free_dev(e->d_grm_rows);
  if (e->prep_ev) cudaEventDestroy(e->prep_ev);
  if (e->mma_ev) cudaEventDestroy(e->mma_ev);
  for (auto ev : e->slot_ev) if (ev) cudaEventDestroy(ev);
  if (e->prep) cudaStreamDestroy(e->prep);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->out_stream) cudaStreamDestroy(e->out_stream);
  if (e->d2h_stream) cudaStreamDestroy(e->d2h_stream);
  for (auto ev : e->group_ev) cudaEventDestroy(ev);
  for (int i = 0; i < dgrm_engine::kOutSlots; ++i) {
    free_dev(e->d_out_stage[i]);
    if (e->h_out_stage[i]) cudaFreeHost(e->h_out_stage[i]);
    if (e->stage_packed[i]) cudaEventDestroy(e->stage_packed[i]);
    if (e->stage_copied_out[i]) cudaEventDestroy(e->stage_copied_out[i]);
  }
  delete e;
}

// Cohort + output geometry.  out_dim == 0: GRM mode, the output matrices are individuals x individuals;
// out_dim > 0: LD mode, they are out_dim x out_dim (SNPs of a group) and the individuals are the contraction axis.
static int setup_geometry(dgrm_engine* e, int64_t n_fam, const int32_t* keep_mask, int64_t out_dim) {
  if (!e) return DGRM_E_INVALID;
  if (n_fam <= 0 || n_fam > (1 << 30)) return fail(e, DGRM_E_INVALID, "dgrm_set_individuals: bad n_fam");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->prep));
  CK(cudaStreamSynchronize(e->stream));
  release_geometry(e);
  if (e->d_bed && e->bed_nfam != n_fam) {  // the resident rows belong to another .fam
    free_dev(e->d_bed); e->d_bed = nullptr;
    e->bed_rows = e->bed_cap = 0;
  }
  e->nbind = 0;
  e->ntbind = 0;
  e->begun = e->finished = false;
  std::vector<int32_t> src;
  bool identity = true;
  if (keep_mask) {
    for (int64_t f = 0; f < n_fam; ++f)
      if (keep_mask[f]) src.push_back((int32_t)f); else identity = false;
  }
  e->n_fam = n_fam;
  e->n_row = (keep_mask && !identity) ? (int64_t)src.size() : n_fam;
  if (e->n_row <= 0) return fail(e, DGRM_E_INVALID, "dgrm_set_individuals: no individual kept");
  e->ld_mode = out_dim > 0;
  e->n = e->ld_mode ? out_dim : e->n_row;
  e->n_pad = round_up(e->n, 256);
  e->b_fam = (n_fam + 3) / 4;
  e->b_row = (e->n_row + 3) / 4;
  e->chunk_pitch = round_up(std::max(e->b_fam, e->b_row), 128);
  if (keep_mask && !identity) {
    CK(cudaMalloc(&e->d_src_ind, src.size() * sizeof(int32_t)));
    CK(cudaMemcpy(e->d_src_ind, src.data(), src.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  {
    const int rc = build_tiles(e);
    if (rc) return rc;
  }
  const int kc = e->kc_max;
  if (e->ntiles > 0) {
    CK(cudaMalloc(&e->d_S, (size_t)e->ntiles * kTileElems * sizeof(double)));
    if (!e->cfg.flat_norm) CK(cudaMalloc(&e->d_N, (size_t)e->ntiles * kTileElems * sizeof(int32_t)));
  }
  CK(cudaMalloc(&e->d_diag, (size_t)e->n_pad * sizeof(double)));
  CK(cudaMalloc(&e->d_miss, (size_t)e->n_pad * sizeof(int32_t)));
  CK(cudaMalloc(&e->d_chunk, (size_t)kc * e->chunk_pitch));
  CK(cudaMemset(e->d_chunk, 0x55, (size_t)kc * e->chunk_pitch));
  CK(cudaMalloc(&e->d_zt, (size_t)kc * 4 * sizeof(double)));
  if (!e->ld_mode && e->k1_transposed) {
    e->t_pitch = round_up(e->n_row, 32);
    CK(cudaMalloc(&e->d_chunkT, (size_t)(kc / 16) * e->t_pitch * sizeof(uint32_t)));
    CK(cudaMalloc(&e->d_counts, (size_t)kc * 4 * sizeof(int32_t)));
    CK(cudaMalloc(&e->d_miss_part, (size_t)(kc / kExpK) * e->n_pad * sizeof(int32_t)));
    CK(cudaMemset(e->d_miss_part, 0, (size_t)(kc / kExpK) * e->n_pad * sizeof(int32_t)));
  }
  CK(cudaMalloc(&e->d_diag_part, (size_t)(kc / kExpK) * e->n_pad * sizeof(double)));
  CK(cudaMemset(e->d_diag_part, 0, (size_t)(kc / kExpK) * e->n_pad * sizeof(double)));   // blocks nobody expands stay 0
  e->slots = 0;  // operand slots are sized by dgrm_begin, once the SNP count is known
  // host staging: slabs of up to kc rows
  e->stage_rows = kc;
  e->stage_pitch = round_up(e->b_fam, 16);
  for (int i = 0; i < 2; ++i) {
    CK(cudaMalloc(&e->d_stage[i], (size_t)e->stage_rows * e->stage_pitch));
    CK(cudaMalloc(&e->d_rowidx[i], (size_t)kc * sizeof(int32_t)));
    CK(cudaHostAlloc(&e->h_rowidx[i], (size_t)kc * sizeof(int32_t), cudaHostAllocDefault));
  }
  if (!e->d_bad_count) {
    CK(cudaMalloc(&e->d_bad_count, 8 * sizeof(int32_t)));
    CK(cudaMalloc(&e->d_bad_list, (size_t)dgrm_engine::kBadCap * sizeof(int64_t)));
  }
  e->have_ind = true;
  return DGRM_OK;
}

int dgrm_set_individuals(dgrm_engine* e, int64_t n_fam, const int32_t* keep_mask) {
  return setup_geometry(e, n_fam, keep_mask, 0);
}

static int begin_common(dgrm_engine* e, int64_t m_capacity);

int dgrm_begin(dgrm_engine* e, int64_t m_capacity) {
  if (!e) return DGRM_E_INVALID;
  if (!e->have_ind) return fail(e, DGRM_E_INVALID, "dgrm_begin: call dgrm_set_individuals first");
  if (e->ld_mode) return fail(e, DGRM_E_INVALID, "dgrm_begin: the engine is in LD mode (dgrm_ld_begin); call dgrm_set_individuals for a GRM");
  return begin_common(e, m_capacity);
}

// LD mode (GWAS::getLessSignificantCorrelatedSNPs, gwas.cpp:1156-1172): correlations between the m_group SNPs that
// the following pushes deliver.  The operands are SNP-major (row = SNP, K = individuals, every individual chunk
// resident at once), the tile store is m_group x m_group, everything downstream of the contraction is shared.
int dgrm_ld_begin(dgrm_engine* e, int64_t n_fam, const int32_t* keep_mask, int64_t m_group) {
  if (!e) return DGRM_E_INVALID;
  if (m_group <= 0 || m_group > (1 << 30)) return fail(e, DGRM_E_INVALID, "dgrm_ld_begin: bad m_group");
  if (e->cfg.flat_norm) return fail(e, DGRM_E_INVALID, "dgrm_ld_begin: flat normalisation does not exist for SNP correlations");
  // GWAS calls this once per SNP group: keep the geometry (tile store, operand buffers, staging) when nothing changed
  std::vector<int32_t> km;
  if (keep_mask) km.assign(keep_mask, keep_mask + n_fam);
  const bool same = e->have_ind && e->ld_mode && e->n_fam == n_fam && e->n == m_group && e->ld_keep == km &&
                    e->nbind == 0 && e->ntbind == 0;
  if (!same) {
    int rc = setup_geometry(e, n_fam, keep_mask, m_group);
    if (rc) return rc;
    e->ld_keep.swap(km);
  }
  return begin_common(e, m_group);
}

static int begin_common(dgrm_engine* e, int64_t m_capacity) {
  if (m_capacity <= 0) return fail(e, DGRM_E_INVALID, "dgrm_begin: m_capacity must be positive");
  CK(cudaSetDevice(e->cfg.device));
  if (m_capacity > e->m_cap) {
    CK(cudaStreamSynchronize(e->prep));
    CK(cudaStreamSynchronize(e->stream));
    release_stats(e);
    CK(cudaMalloc(&e->d_freq, (size_t)m_capacity * 4 * sizeof(int32_t)));
    CK(cudaMalloc(&e->d_nn, (size_t)m_capacity * sizeof(int32_t)));
    CK(cudaMalloc(&e->d_p1, (size_t)m_capacity * sizeof(double)));
    CK(cudaMalloc(&e->d_p2, (size_t)m_capacity * sizeof(double)));
    CK(cudaMalloc(&e->d_sd, (size_t)m_capacity * sizeof(double)));
    e->m_cap = m_capacity;
  }
  if (e->ld_mode) {
    // every individual chunk is resident: slots cover the padded individual count
    const int64_t kpad = round_up(e->n_row, 256);
    const int64_t need = (kpad + e->kc_max - 1) / e->kc_max;
    if (need > kMaxChunks)
      return fail(e, DGRM_E_INVALID, "dgrm_ld_begin: %lld individuals need more than %d chunks of %d", (long long)e->n_row,
                  kMaxChunks, e->kc_max);
    const int rc = ensure_operands(e, (int)need);
    if (rc) return rc;
    // rows past the group's SNPs and columns past the individuals must contribute nothing
    const bool split = e->cfg.operand == DGRM_OPERAND_FP16X2;
    CK(cudaMemsetAsync(e->d_Z, 0, (size_t)e->n_pad * e->z_pitch * sizeof(uint16_t), e->prep));
    if (e->d_Q) CK(cudaMemsetAsync(e->d_Q, 0, (size_t)e->n_pad * e->q_pitch, e->prep));
    if (e->d_Q4) CK(cudaMemsetAsync(e->d_Q4, 0, (size_t)e->n_pad * e->q4_pitch, e->prep));
    if (e->d_MB) CK(cudaMemsetAsync(e->d_MB, 0, (size_t)e->n_pad * e->mb_pitch * sizeof(uint32_t), e->prep));
    (void)split;
  } else {
    // resident operand slots: as many chunks as the SNP count needs, bounded by the memory budget
    const int64_t need = (m_capacity + e->kc_max - 1) / e->kc_max;
    const bool split = e->cfg.operand == DGRM_OPERAND_FP16X2;
    const double per_chunk = (double)e->n_pad * e->kc_max * ((split ? 4 : 2) + 1.0 + 1.0 / 8);
    int64_t cap = e->cfg.resident_chunks > 0
                      ? e->cfg.resident_chunks
                      : (int64_t)std::max(1.0, std::floor((double)(e->cfg.resident_mb > 0 ? e->cfg.resident_mb : 16384) * 1048576.0 / per_chunk));
    cap = std::min<int64_t>(cap, kMaxChunks);
    const int slots = (int)std::max<int64_t>(1, std::min(need, cap));
    const int rc = ensure_operands(e, slots);
    if (rc) return rc;
  }
  e->m = 0;
  e->fill = 0;
  e->h_fill = e->dev_fill = 0;
  e->raw_rows = 0;
  e->raw_fill = 0;
  e->h_any_missing = e->dev_any_missing = false;
  e->fill_dosage = false;
  e->chunks_done = 0;
  e->chunks_in_super = 0;
  e->launched_s = e->launched_n = 0;
  e->pre_s = e->pre_n = 0;
  e->pre_end = -1;
  e->slots_evented = 0;
  e->supers_done = 0;
  e->finished = false;
  e->grm_rows.clear();
  e->grm_from_store = true;
  if (e->d_subset) {
    CK(cudaStreamSynchronize(e->stream));
    free_dev(e->d_subset);
    e->d_subset = nullptr;
    e->h_subset.clear();
    e->n_out = 0;
  }
  // a GRM that was abandoned before dgrm_finish may still have contractions in flight
  CK(cudaEventRecord(e->mma_ev, e->stream));
  CK(cudaStreamWaitEvent(e->prep, e->mma_ev, 0));
  CK(cudaMemsetAsync(e->d_miss, 0, (size_t)e->n_pad * sizeof(int32_t), e->prep));
  e->miss_valid = !e->ld_mode;
  e->n_u16 = false;
  CK(cudaMemsetAsync(e->d_bad_count, 0, 8 * sizeof(int32_t), e->prep));
  CK(cudaMemsetAsync(e->d_diag, 0, (size_t)e->n_pad * sizeof(double), e->prep));
  e->begun = true;
  return DGRM_OK;
}

int dgrm_push_bed_device(dgrm_engine* e, const uint8_t* d_rows, int64_t pitch, int64_t nrows, const uint8_t* row_keep,
                         void* producer_stream) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  if (e->finished) return fail(e, DGRM_E_INVALID, "dgrm_push_bed_device after dgrm_finish");
  if (!d_rows || nrows < 0 || pitch < e->b_fam) return fail(e, DGRM_E_INVALID, "dgrm_push_bed_device: bad rows/pitch");
  CK(cudaSetDevice(e->cfg.device));
  e->grm_from_store = false;
  // order after the producer; NULL is the legacy default stream (what torch uses unless told otherwise), which the
  // engine's non-blocking streams do not synchronise with implicitly
  CK(cudaEventRecord(e->producer_ev, (cudaStream_t)producer_stream));
  CK(cudaStreamWaitEvent(e->prep, e->producer_ev, 0));
  // bound the host-side index list
  for (int64_t r0 = 0; r0 < nrows; r0 += (1 << 20)) {
    const int64_t cnt = std::min<int64_t>(nrows - r0, 1 << 20);
    rc = push_device_rows(e, d_rows + r0 * pitch, pitch, cnt, row_keep ? row_keep + r0 : nullptr);
    if (rc) return rc;
  }
  return DGRM_OK;
}

int dgrm_release_inputs(dgrm_engine* e, void* consumer_stream) {
  if (!e) return DGRM_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  // readers of caller memory: the row gather into the chunk buffer (decode stream) and the host-to-device copies
  // of dgrm_push_bed_host (copy stream)
  CK(cudaEventRecord(e->producer_ev, e->prep));
  CK(cudaStreamWaitEvent((cudaStream_t)consumer_stream, e->producer_ev, 0));
  CK(cudaEventRecord(e->producer_ev, e->copy_stream));
  CK(cudaStreamWaitEvent((cudaStream_t)consumer_stream, e->producer_ev, 0));
  return DGRM_OK;
}

int dgrm_push_bed_host(dgrm_engine* e, const uint8_t* rows, int64_t pitch, int64_t nrows, const uint8_t* row_keep) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  if (e->finished) return fail(e, DGRM_E_INVALID, "dgrm_push_bed_host after dgrm_finish");
  if (!rows || nrows < 0 || pitch < e->b_fam) return fail(e, DGRM_E_INVALID, "dgrm_push_bed_host: bad rows/pitch");
  CK(cudaSetDevice(e->cfg.device));
  e->grm_from_store = false;
  for (int64_t r0 = 0; r0 < nrows; r0 += e->stage_rows) {
    const int64_t cnt = std::min<int64_t>(nrows - r0, e->stage_rows);
    const int s = e->stage_i;
    e->stage_i ^= 1;
    CK(cudaEventSynchronize(e->stage_free[s]));  // the gather that last read this slot is done
    CK(cudaMemcpy2DAsync(e->d_stage[s], (size_t)e->stage_pitch, rows + r0 * pitch, (size_t)pitch, (size_t)e->b_fam,
                         (size_t)cnt, cudaMemcpyHostToDevice, e->copy_stream));
    CK(cudaEventRecord(e->stage_copied[s], e->copy_stream));
    CK(cudaStreamWaitEvent(e->prep, e->stage_copied[s], 0));
    rc = push_device_rows(e, e->d_stage[s], e->stage_pitch, cnt, row_keep ? row_keep + r0 : nullptr);
    if (rc) return rc;
    CK(cudaEventRecord(e->stage_free[s], e->prep));
  }
  return DGRM_OK;
}

int dgrm_push_dosage_host(dgrm_engine* e, const double* dosage, int64_t pitch, const uint8_t* missing, int64_t missing_pitch,
                          int64_t nrows) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  if (e->finished) return fail(e, DGRM_E_INVALID, "dgrm_push_dosage_host after dgrm_finish");
  if (e->ld_mode) return fail(e, DGRM_E_INVALID, "dgrm_push_dosage_host: not available in LD mode");
  if (!dosage || nrows < 0 || pitch < e->n || (missing && missing_pitch < e->n))
    return fail(e, DGRM_E_INVALID, "dgrm_push_dosage_host: bad rows/pitch (rows hold the %lld kept individuals)", (long long)e->n);
  const int64_t pending = e->m + e->fill + e->dev_fill + e->h_fill + e->raw_rows;
  if (pending + nrows > e->m_cap)
    return fail(e, DGRM_E_INVALID, "more SNPs pushed (%lld) than dgrm_begin's capacity (%lld)", (long long)(pending + nrows),
                (long long)e->m_cap);
  CK(cudaSetDevice(e->cfg.device));
  rc = ensure_dosage_staging(e);
  if (rc) return rc;
  e->grm_from_store = false;
  if (e->fill > 0 && !e->fill_dosage) {  // packed rows came before: they end their chunk
    rc = process_chunk(e);
    if (rc) return rc;
  }
  rc = send_bgen(e);   // probability blocks pushed before these rows keep their place in the SNP order
  if (rc) return rc;
  const int64_t n = e->n;
  for (int64_t r = 0; r < nrows;) {
    const int s = e->dos_i;
    if (e->h_fill == 0) CK(cudaEventSynchronize(e->dos_copied[s]));  // the copies that last read this pinned buffer are done
    const int64_t take = std::min<int64_t>(nrows - r, e->h_cap - e->h_fill);
    double* hd = e->h_dos[s] + (size_t)e->h_fill * n;
    uint8_t* hm = e->h_dmiss[s] + (size_t)e->h_fill * n;
    if (missing && !e->h_any_missing && e->h_fill > 0) memset(e->h_dmiss[s], 0, (size_t)e->h_fill * n);
    const bool clear_mask = !missing && e->h_any_missing;
#pragma omp parallel for schedule(static) if (take * n >= (1 << 20))
    for (int64_t k = 0; k < take; ++k) {
      memcpy(hd + k * n, dosage + (r + k) * pitch, (size_t)n * sizeof(double));
      if (missing) memcpy(hm + k * n, missing + (r + k) * missing_pitch, (size_t)n);
      else if (clear_mask) memset(hm + k * n, 0, (size_t)n);
    }
    if (missing) e->h_any_missing = true;
    e->h_fill += (int)take;
    r += take;
    if (e->h_fill == e->h_cap) {
      rc = send_dosage(e);
      if (rc) return rc;
    }
  }
  return DGRM_OK;
}

int dgrm_push_bgen_blocks_host(dgrm_engine* e, const uint8_t* data, const int64_t* offsets, int64_t nrows, int layout,
                               int64_t n_samples, const int32_t* sample_keep) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  if (e->finished) return fail(e, DGRM_E_INVALID, "dgrm_push_bgen_blocks_host after dgrm_finish");
  if (e->ld_mode) return fail(e, DGRM_E_INVALID, "dgrm_push_bgen_blocks_host: not available in LD mode");
  if (!data || !offsets || nrows < 0 || (layout != 1 && layout != 2) || n_samples <= 0 || n_samples > INT32_MAX)
    return fail(e, DGRM_E_INVALID, "dgrm_push_bgen_blocks_host: bad arguments");
  const int64_t pending = e->m + e->fill + e->dev_fill + e->h_fill + e->raw_rows;
  if (pending + nrows > e->m_cap)
    return fail(e, DGRM_E_INVALID, "more SNPs pushed (%lld) than dgrm_begin's capacity (%lld)", (long long)(pending + nrows),
                (long long)e->m_cap);
  // the format conditions the reference's setter checks per variant (genotypebgen.cpp:99-172), on the block headers
  int64_t max_block = 0;
  for (int64_t r = 0; r < nrows; ++r) {
    const int64_t len = offsets[r + 1] - offsets[r];
    const uint8_t* b = data + offsets[r];
    max_block = std::max(max_block, len);
    if (layout == 1) {
      if (len != 6 * n_samples) return fail(e, DGRM_E_INVALID, "Error: An internal error has happened. Unexpected small number of samples.");
      continue;
    }
    if (len < 10 + n_samples) return fail(e, DGRM_E_INVALID, "Error: An internal error has happened. Unexpected small number of samples.");
    const uint32_t nb = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
    const int kb = b[4] | (b[5] << 8), pmin = b[6], pmax = b[7], phased = b[8 + n_samples], bits = b[9 + n_samples];
    if ((int64_t)nb != n_samples) return fail(e, DGRM_E_INVALID, "Error: An internal error has happened. Unexpected small number of samples.");
    if (kb != 2)
      return fail(e, DGRM_E_INVALID, "Error: The bgen file contains non-biallelic genotypes. DISSECT currently can only work with biallelic genotypes.");
    if (pmin != 2 || pmax != 2)
      return fail(e, DGRM_E_INVALID, "Error: The bgen file contains non-diploid genotypes. DISSECT currently can only work with diploid genotypes.");
    if (phased)
      return fail(e, DGRM_E_INVALID, "Error: Currently can only load unphased genotypes. If you are interested on loading phased genotypes, please, contact us.");
    if (bits < 1 || bits > 32 || len < 10 + n_samples + (2 * n_samples * bits + 7) / 8)
      return fail(e, DGRM_E_INVALID, "Error: Unexpected number of probabilities when reading bgen file.");
  }
  CK(cudaSetDevice(e->cfg.device));
  rc = ensure_dosage_staging(e);
  if (rc) return rc;
  e->grm_from_store = false;
  if (e->fill > 0 && !e->fill_dosage) {  // packed rows came before: they end their chunk
    rc = process_chunk(e);
    if (rc) return rc;
  }
  rc = send_dosage(e);   // dosage rows pushed before these blocks keep their place in the SNP order
  if (rc) return rc;
  // the kept samples: a device list of file indices (individualBGENIdxsMaskToKeep, genotypebgen.cpp:367-379)
  {
    std::vector<int32_t> km;
    if (sample_keep) km.assign(sample_keep, sample_keep + n_samples);
    if (e->raw_rows > 0 && (km != e->bgen_keep || n_samples != e->raw_samples || layout != e->raw_layout)) {
      rc = send_bgen(e);   // blocks of another file / selection are still staged
      if (rc) return rc;
    }
    if (km != e->bgen_keep || n_samples != e->raw_samples || !e->d_bgen_cols) {
      std::vector<int32_t> cols;
      for (int64_t i = 0; i < n_samples; ++i)
        if (!sample_keep || sample_keep[i]) cols.push_back((int32_t)i);
      if ((int64_t)cols.size() != e->n)
        return fail(e, DGRM_E_INVALID, "dgrm_push_bgen_blocks_host: %lld samples are kept, the engine was set up for %lld individuals",
                    (long long)cols.size(), (long long)e->n);
      CK(cudaStreamSynchronize(e->prep));
      free_dev(e->d_bgen_cols);
      e->d_bgen_cols = nullptr;
      CK(cudaMalloc(&e->d_bgen_cols, cols.size() * sizeof(int32_t)));
      CK(cudaMemcpy(e->d_bgen_cols, cols.data(), cols.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      e->bgen_all = (int64_t)cols.size() == n_samples;
      e->bgen_keep.swap(km);
    }
    e->raw_samples = n_samples;
    e->raw_layout = layout;
  }
  if (max_block > e->raw_cap) {
    rc = send_bgen(e);
    if (rc) return rc;
  }
  rc = ensure_bgen_staging(e, max_block);
  if (rc) return rc;
  for (int64_t r = 0; r < nrows;) {
    const int64_t len0 = offsets[r + 1] - offsets[r];
    if (e->raw_rows == e->raw_rows_cap || e->raw_fill + len0 > e->raw_cap) {
      rc = send_bgen(e);
      if (rc) return rc;
    }
    const int s = e->raw_i;
    if (e->raw_rows == 0) {
      CK(cudaEventSynchronize(e->raw_copied[s]));   // the copy that last read this pinned piece is done
      e->h_boff[s][0] = 0;
    }
    // the run of blocks [r, r2) that fits into the open piece; the caller's blocks need not be adjacent in `data`
    int64_t r2 = r, fill = e->raw_fill;
    while (r2 < nrows && e->raw_rows + (r2 - r) < e->raw_rows_cap && fill + (offsets[r2 + 1] - offsets[r2]) <= e->raw_cap) {
      fill += offsets[r2 + 1] - offsets[r2];
      e->h_boff[s][e->raw_rows + (r2 - r) + 1] = fill;
      ++r2;
    }
    const int64_t* pos = e->h_boff[s] + e->raw_rows;
    uint8_t* dst = e->h_raw[s];
#pragma omp parallel for schedule(static) if (fill - e->raw_fill >= (1 << 20))
    for (int64_t k = r; k < r2; ++k) memcpy(dst + pos[k - r], data + offsets[k], (size_t)(offsets[k + 1] - offsets[k]));
    e->raw_rows += (int)(r2 - r);
    e->raw_fill = fill;
    r = r2;
  }
  return DGRM_OK;
}

int dgrm_load_bed_file(dgrm_engine* e, const char* bed_path, int64_t m_file, const int64_t* snp_keep_idx,
                       int64_t m_keep) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  if (!bed_path || m_file <= 0) return fail(e, DGRM_E_INVALID, "dgrm_load_bed_file: bad arguments");
  FILE* f = nullptr;
  rc = open_bed(e, bed_path, &f);
  if (rc) return rc;
  std::vector<uint8_t> keep;
  if (snp_keep_idx) {
    keep.assign((size_t)m_file, 0);
    int64_t prev = -1;
    for (int64_t i = 0; i < m_keep; ++i) {
      const int64_t s = snp_keep_idx[i];
      if (s < 0 || s >= m_file || s <= prev) {
        fclose(f);
        return fail(e, DGRM_E_INVALID, "dgrm_load_bed_file: snp_keep_idx must be ascending positions in [0, m_file)");
      }
      keep[(size_t)s] = 1;
      prev = s;
    }
  }
  const int64_t slab = e->stage_rows;
  uint8_t* buf[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i)
    if (cudaHostAlloc(&buf[i], (size_t)slab * e->b_fam, cudaHostAllocDefault) != cudaSuccess) {
      if (buf[0]) cudaFreeHost(buf[0]);
      fclose(f);
      return fail(e, DGRM_E_NOMEM, "Error: There are not enough memory for reading the binary file.");
    }
  int which = 0;
  rc = DGRM_OK;
  for (int64_t r0 = 0; r0 < m_file && rc == DGRM_OK; r0 += slab, which ^= 1) {
    const int64_t cnt = std::min<int64_t>(slab, m_file - r0);
    // the previous push from this pinned buffer went through the same double-buffered stage: it is
    // complete once the stage slot it used has been recycled, which push_bed_host waits for.
    if (r0 >= 2 * slab) cudaStreamSynchronize(e->copy_stream);
    const size_t got = fread(buf[which], (size_t)e->b_fam, (size_t)cnt, f);
    if ((int64_t)got != cnt) {
      rc = fail(e, DGRM_E_BADBED, "Error: An internal error has happened when reading the genotypes file.");
      break;
    }
    rc = dgrm_push_bed_host(e, buf[which], e->b_fam, cnt, snp_keep_idx ? keep.data() + r0 : nullptr);
  }
  cudaStreamSynchronize(e->copy_stream);
  cudaFreeHost(buf[0]);
  cudaFreeHost(buf[1]);
  fclose(f);
  return rc;
}

int dgrm_finish(dgrm_engine* e) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  CK(cudaSetDevice(e->cfg.device));
  if (!e->finished) {
    rc = flush_dosage(e);
    if (rc) return rc;
    rc = process_chunk(e);
    if (rc) return rc;
    if (e->ld_mode) {   // every individual chunk is resident: contract them all now
      const int64_t kpad = round_up(e->n_row, 256);
      e->chunks_in_super = e->chunks_done > 0 ? e->slots : 0;
      for (int j = 0; j < e->slots; ++j) e->slot_kpad[j] = (int)std::min<int64_t>(e->kc_max, kpad - (int64_t)j * e->kc_max);
    }
    rc = flush_super(e, true);
    if (rc) return rc;
    CK(cudaEventRecord(e->prep_ev, e->prep));  // statistics, bad-SNP list and the exact diagonal live on the decode stream
    CK(cudaStreamWaitEvent(e->stream, e->prep_ev, 0));
    if (e->ntiles > 0 && e->chunks_done == 0) {  // no SNP at all: a well-defined empty GRM
      CK(cudaMemsetAsync(e->d_S, 0, (size_t)e->ntiles * kTileElems * sizeof(double), e->stream));
      if (e->d_N) CK(cudaMemsetAsync(e->d_N, 0, (size_t)e->ntiles * kTileElems * sizeof(int32_t), e->stream));
    }
    if (e->ntiles > 0 && e->m > 0) {
      Prof pr(e, DGRM_K_FINAL);
      fix_diag_kernel<<<(unsigned)((e->n + 255) / 256), 256, 0, e->stream>>>(e->d_S, e->d_index, e->ncp, (int)e->n,
                                                                          e->d_diag);
      CK(cudaGetLastError());
    }
    e->finished = true;
    CK(cudaStreamSynchronize(e->out_stream));
    CK(cudaStreamSynchronize(e->d2h_stream));
    if (e->chunks_done == 0 || e->work.empty())
      for (int b = 0; b < e->ntbind; ++b) {
        rc = dgrm_fetch_tiles(e, e->tbind[b].which, e->tbind[b].host);
        if (rc) return rc;
      }
    // bindings that could not stream (pageable memory, uplo U/B, or nothing was contracted): fill them now
    for (int b = 0; b < e->nbind; ++b) {
      const dgrm_engine::Binding& bd = e->bind[b];
      if (binding_streams(e, bd) && e->chunks_done > 0 && !e->work.empty()) continue;
      const char up = bd.uplo == 0 ? 'L' : bd.uplo == 1 ? 'U' : 'B';
      rc = dgrm_fetch_block_cyclic(e, bd.which, bd.host, bd.g.lld, bd.g.bs_r, bd.g.bs_c, bd.g.my_row, bd.g.my_col, bd.g.R,
                                   bd.g.C, up);
      if (rc) return rc;
    }
  }
  int32_t bad[8];   // {sd == 0, no non-missing, listed, zero-variance dosage rows, BGEN samples that are not diploid, ...}
  CK(cudaMemcpyAsync(bad, e->d_bad_count, sizeof bad, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaGetLastError());
  if (bad[0] > 0)
    return fail(e, DGRM_E_BADSNP,
                "Error: There are SNPs with same genotype for all samples or insignificant variability. Please, filter "
                "them before performing this analysis. (%d SNPs)",
                bad[0]);
  if (bad[1] > 0)
    return fail(e, DGRM_E_NANSNP, "Error: %d SNPs have no non-missing genotype (allele frequency is 0/0).", bad[1]);
  if (bad[4] > 0)   // genotypebgen.cpp:120-123
    return fail(e, DGRM_E_INVALID,
                "Error: The bgen file contains non-diploid genotypes. DISSECT currently can only work with diploid genotypes.");
  if (bad[3] > 0 && e->cfg.flat_norm)
    return fail(e, DGRM_E_BADSNP,
                "Error: %d dosage SNPs have zero variance; under --grm-flat-norm they change the SNP count the GRM is "
                "divided by (removeSNPsWithZeroVariance): filter them (dgrm_snp_stats: std_dev == 0) and rebuild.",
                bad[3]);
  return DGRM_OK;
}

int dgrm_sync(dgrm_engine* e) {
  if (!e) return DGRM_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->copy_stream));
  CK(cudaStreamSynchronize(e->prep));
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaStreamSynchronize(e->out_stream));
  CK(cudaStreamSynchronize(e->d2h_stream));
  return DGRM_OK;
}

int dgrm_bind_output(dgrm_engine* e, int which, double* local, int64_t lld, int bs_r, int bs_c, int my_row, int my_col,
                     int R, int C, char uplo) {
  if (!e) return DGRM_E_INVALID;
  if (!e->have_ind) return fail(e, DGRM_E_INVALID, "dgrm_bind_output: call dgrm_set_individuals first");
  if (!local || which < 0 || which > 2 || bs_r < 1 || bs_c < 1 || R < 1 || C < 1 || my_row < 0 || my_row >= R ||
      my_col < 0 || my_col >= C)
    return fail(e, DGRM_E_INVALID, "dgrm_bind_output: bad arguments");
  const int up = (uplo == 'L' || uplo == 'l') ? 0 : (uplo == 'U' || uplo == 'u') ? 1 : (uplo == 'B' || uplo == 'b') ? 2 : -1;
  if (up < 0) return fail(e, DGRM_E_INVALID, "dgrm_bind_output: uplo must be L, U or B");
  CK(cudaSetDevice(e->cfg.device));
  int slot = e->nbind;
  for (int b = 0; b < e->nbind; ++b)
    if (e->bind[b].which == which) slot = b;
  if (slot >= 3) return fail(e, DGRM_E_INVALID, "dgrm_bind_output: at most three bindings");
  dgrm_engine::Binding& bd = e->bind[slot];
  bd.which = which;
  bd.host = local;
  bd.uplo = up;
  bd.g.bs_r = bs_r; bd.g.bs_c = bs_c; bd.g.my_row = my_row; bd.g.my_col = my_col; bd.g.R = R; bd.g.C = C;
  bd.g.nloc_r = numroc((int)e->n, bs_r, my_row, R);
  bd.g.nloc_c = numroc((int)e->n, bs_c, my_col, C);
  bd.g.lld = lld;
  if (lld < std::max(bd.g.nloc_r, 1)) return fail(e, DGRM_E_INVALID, "dgrm_bind_output: lld < local rows");
  cudaPointerAttributes attr;
  bd.dev = nullptr;
  if (cudaPointerGetAttributes(&attr, local) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
    bd.dev = (double*)attr.devicePointer;
  cudaGetLastError();
  if (slot == e->nbind) e->nbind++;
  return DGRM_OK;
}

int dgrm_unbind_outputs(dgrm_engine* e) {
  if (!e) return DGRM_E_INVALID;
  e->nbind = 0;
  e->ntbind = 0;
  return DGRM_OK;
}

int dgrm_bind_tiles(dgrm_engine* e, int which, double* tiles) {
  if (!e) return DGRM_E_INVALID;
  if (!e->have_ind) return fail(e, DGRM_E_INVALID, "dgrm_bind_tiles: call dgrm_set_individuals first");
  if (!tiles || which < 0 || which > 2) return fail(e, DGRM_E_INVALID, "dgrm_bind_tiles: bad arguments");
  CK(cudaSetDevice(e->cfg.device));
  int slot = e->ntbind;
  for (int b = 0; b < e->ntbind; ++b)
    if (e->tbind[b].which == which) slot = b;
  if (slot >= 3) return fail(e, DGRM_E_INVALID, "dgrm_bind_tiles: at most three bindings");
  e->tbind[slot].which = which;
  e->tbind[slot].host = tiles;
  cudaPointerAttributes attr;
  e->tbind[slot].pinned = cudaPointerGetAttributes(&attr, tiles) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (slot == e->ntbind) e->ntbind++;
  return DGRM_OK;
}

int64_t dgrm_num_snps(const dgrm_engine* e) { return e ? e->m + e->fill + e->dev_fill + e->h_fill + e->raw_rows : 0; }
int64_t dgrm_num_individuals(const dgrm_engine* e) { return e ? e->n : 0; }

int dgrm_snp_stats(dgrm_engine* e, int32_t* freq, int32_t* nn, double* p1, double* p2, double* sd) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->prep));
  CK(cudaStreamSynchronize(e->stream));
  const size_t m = (size_t)e->m;
  if (m == 0) return DGRM_OK;
  if (freq) CK(cudaMemcpy(freq, e->d_freq, m * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (nn) CK(cudaMemcpy(nn, e->d_nn, m * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (p1) CK(cudaMemcpy(p1, e->d_p1, m * sizeof(double), cudaMemcpyDeviceToHost));
  if (p2) CK(cudaMemcpy(p2, e->d_p2, m * sizeof(double), cudaMemcpyDeviceToHost));
  if (sd) CK(cudaMemcpy(sd, e->d_sd, m * sizeof(double), cudaMemcpyDeviceToHost));
  return DGRM_OK;
}

int dgrm_bad_snps(dgrm_engine* e, int64_t* idx, int64_t cap, int64_t* count) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->prep));
  CK(cudaStreamSynchronize(e->stream));
  int32_t bad[4];
  CK(cudaMemcpy(bad, e->d_bad_count, sizeof bad, cudaMemcpyDeviceToHost));
  const int64_t have = std::min<int64_t>(bad[2], dgrm_engine::kBadCap);
  if (count) *count = bad[2];
  if (idx && cap > 0 && have > 0) {
    std::vector<int64_t> tmp((size_t)have);
    CK(cudaMemcpy(tmp.data(), e->d_bad_list, (size_t)have * sizeof(int64_t), cudaMemcpyDeviceToHost));
    for (auto& v : tmp)
      if (v < 0) v = -v - 1;  // NaN SNPs are stored negative-coded
    std::sort(tmp.begin(), tmp.end());
    for (int64_t i = 0; i < std::min(cap, have); ++i) idx[i] = tmp[(size_t)i];
  }
  return DGRM_OK;
}

// local columns [lcb, lce) of the caller's block (lce < 0: all of them); `local` points at column lcb
int dgrm_fetch_block_cyclic_cols(dgrm_engine* e, int which, double* local, int64_t lld, int bs_r, int bs_c, int my_row,
                                 int my_col, int R, int C, char uplo, int64_t lcb, int64_t lce) {
  int rc = check_ready(e, true);
  if (rc) return rc;
  if (!local || which < 0 || which > 2 || bs_r < 1 || bs_c < 1 || R < 1 || C < 1 || my_row < 0 || my_row >= R ||
      my_col < 0 || my_col >= C)
    return fail(e, DGRM_E_INVALID, "dgrm_fetch_block_cyclic: bad arguments");
  const int up = (uplo == 'L' || uplo == 'l') ? 0 : (uplo == 'U' || uplo == 'u') ? 1 : (uplo == 'B' || uplo == 'b') ? 2 : -1;
  if (up < 0) return fail(e, DGRM_E_INVALID, "dgrm_fetch_block_cyclic: uplo must be L, U or B");
  CK(cudaSetDevice(e->cfg.device));
  CyclicGeom g;
  g.bs_r = bs_r; g.bs_c = bs_c; g.my_row = my_row; g.my_col = my_col; g.R = R; g.C = C;
  const int n_vis = (int)(e->d_subset ? e->n_out : e->n);
  g.nloc_r = numroc(n_vis, bs_r, my_row, R);
  g.nloc_c = numroc(n_vis, bs_c, my_col, C);
  if (g.nloc_r == 0 || g.nloc_c == 0) return DGRM_OK;
  if (lld < g.nloc_r) return fail(e, DGRM_E_INVALID, "dgrm_fetch_block_cyclic: lld %lld < local rows %d", (long long)lld, g.nloc_r);
  if (lce < 0) { lcb = 0; lce = g.nloc_c; }
  if (lcb < 0 || lce > g.nloc_c || lcb > lce)
    return fail(e, DGRM_E_INVALID, "dgrm_fetch_block_cyclic_cols: bad local column range [%lld, %lld) of %d", (long long)lcb,
                (long long)lce, g.nloc_c);
  if (lcb == lce) return DGRM_OK;
  const int lc_begin = (int)lcb, lc_end = (int)lce;
  const TileStore ts = tile_store(e);
  const dim3 blk(128, 2);

  cudaPointerAttributes attr;
  const bool have_attr = cudaPointerGetAttributes(&attr, local) == cudaSuccess;
  cudaGetLastError();
  const bool on_device = have_attr && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) &&
                         attr.devicePointer != nullptr;
  const bool pinned = have_attr && attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
  if (pinned || on_device) {
    // pinned host memory: zero-copy, the kernel stores straight into the caller's Matrix::m over PCIe.
    // device memory (a multi-engine caller's exchange panel, dgrm.h): plain stores.  Either way only entries of
    // tiles this engine owns are written.
    Prof pr(e, DGRM_K_FETCH);
    g.lld = lld;
    const int PC = 4096;
    for (int lc0 = lc_begin; lc0 < lc_end; lc0 += PC) {
      const int lc1 = std::min(lc0 + PC, lc_end);
      const dim3 grid((g.nloc_r + 127) / 128, (lc1 - lc0 + 1) / 2);
      fetch_block_cyclic_kernel<<<grid, blk, 0, e->stream>>>(ts, which, (double*)attr.devicePointer + (int64_t)(lc0 - lc_begin) * lld,
                                                            g, lc0, lc1, up);
      CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(e->stream));
    return DGRM_OK;
  }
  // pageable destination: device panel -> pinned panel -> memcpy of exactly the requested (and owned) rows
  const int PC = (int)std::max<int64_t>(1, std::min<int64_t>(lc_end - lc_begin, (64 << 20) / ((int64_t)g.nloc_r * 8)));
  struct Panels {
    double* d[2] = {nullptr, nullptr};
    double* h[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    ~Panels() {
      for (int i = 0; i < 2; ++i) {
        if (d[i]) cudaFree(d[i]);
        if (h[i]) cudaFreeHost(h[i]);
        if (done[i]) cudaEventDestroy(done[i]);
      }
    }
  } pn;
  for (int i = 0; i < 2; ++i) {
    CK(cudaMalloc(&pn.d[i], (size_t)PC * g.nloc_r * sizeof(double)));
    CK(cudaHostAlloc(&pn.h[i], (size_t)PC * g.nloc_r * sizeof(double), cudaHostAllocDefault));
    CK(cudaEventCreateWithFlags(&pn.done[i], cudaEventDisableTiming));
  }
  g.lld = g.nloc_r;
  const bool partial = e->cfg.world > 1;   // then only the runs of owned tiles may touch the caller's buffer
  auto first_ge = [&](int target) { return first_local_ge(target, g.nloc_r, bs_r, my_row, R); };
  auto drain = [&](int slot, int lc0, int lc1) -> int {
    CK(cudaEventSynchronize(pn.done[slot]));
    for (int lc = lc0; lc < lc1; ++lc) {
      const int gc = ((lc / bs_c) * C + my_col) * bs_c + lc % bs_c;
      int lo = 0, hi = g.nloc_r;  // local row range to copy
      if (up == 0) lo = first_ge(gc); else if (up == 1) hi = first_ge(gc + 1);
      if (hi <= lo) continue;
      double* dst = local + (int64_t)(lc - lc_begin) * lld;
      const double* src = pn.h[slot] + (int64_t)(lc - lc0) * g.nloc_r;
      if (!partial) {
        memcpy(dst + lo, src + lo, (size_t)(hi - lo) * sizeof(double));
        continue;
      }
      if (!e->h_subset.empty()) {   // ownership follows the individuals behind the subset: element by element
        for (int lr = lo; lr < hi; ++lr) {
          const int gr = ((lr / bs_r) * R + my_row) * bs_r + lr % bs_r;
          if (owns_element(e, std::max(gr, gc), std::min(gr, gc))) dst[lr] = src[lr];
        }
        continue;
      }
      // runs inside one block-cyclic block, one 128-row tile and one side of the diagonal share an owner
      for (int lr = lo; lr < hi;) {
        const int gr = ((lr / bs_r) * R + my_row) * bs_r + lr % bs_r;
        int run = std::min(hi - lr, bs_r - lr % bs_r);
        run = std::min(run, 128 - (gr & 127));
        if (gr < gc) run = std::min(run, gc - gr);
        if (owns_element(e, std::max(gr, gc), std::min(gr, gc)))
          memcpy(dst + lr, src + lr, (size_t)run * sizeof(double));
        lr += run;
      }
    }
    return DGRM_OK;
  };
  int prev_slot = -1, prev0 = 0, prev1 = 0, slot = 0;
  for (int lc0 = lc_begin; lc0 < lc_end; lc0 += PC, slot ^= 1) {
    const int lc1 = std::min(lc0 + PC, lc_end);
    {
      Prof pr(e, DGRM_K_FETCH);
      CK(cudaMemsetAsync(pn.d[slot], 0, (size_t)(lc1 - lc0) * g.nloc_r * sizeof(double), e->stream));
      const dim3 grid((g.nloc_r + 127) / 128, (lc1 - lc0 + 1) / 2);
      fetch_block_cyclic_kernel<<<grid, blk, 0, e->stream>>>(ts, which, pn.d[slot], g, lc0, lc1, up);
      CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(pn.h[slot], pn.d[slot], (size_t)(lc1 - lc0) * g.nloc_r * sizeof(double),
                       cudaMemcpyDeviceToHost, e->stream));
    CK(cudaEventRecord(pn.done[slot], e->stream));
    if (prev_slot >= 0) {
      rc = drain(prev_slot, prev0, prev1);
      if (rc) return rc;
    }
    prev_slot = slot; prev0 = lc0; prev1 = lc1;
  }
  if (prev_slot >= 0) rc = drain(prev_slot, prev0, prev1);
  return rc;
}

int dgrm_fetch_block_cyclic(dgrm_engine* e, int which, double* local, int64_t lld, int bs_r, int bs_c, int my_row,
                            int my_col, int R, int C, char uplo) {
  return dgrm_fetch_block_cyclic_cols(e, which, local, lld, bs_r, bs_c, my_row, my_col, R, C, uplo, 0, -1);
}

int dgrm_low_overlap_pairs(dgrm_engine* e, double min_overlap, int64_t* pairs, int64_t cap, int64_t* count) {
  int rc = check_ready(e, true);
  if (rc) return rc;
  if (cap < 0 || (cap > 0 && !pairs) || !count) return fail(e, DGRM_E_INVALID, "dgrm_low_overlap_pairs: bad arguments");
  CK(cudaSetDevice(e->cfg.device));
  struct Bufs {
    int64_t* pairs = nullptr;
    unsigned long long* count = nullptr;
    ~Bufs() { if (pairs) cudaFree(pairs); if (count) cudaFree(count); }
  } bf;
  CK(cudaMalloc(&bf.count, sizeof(unsigned long long)));
  if (cap > 0) CK(cudaMalloc(&bf.pairs, (size_t)cap * 2 * sizeof(int64_t)));
  CK(cudaMemsetAsync(bf.count, 0, sizeof(unsigned long long), e->stream));
  const int64_t n_vis = e->d_subset ? e->n_out : e->n;
  const dim3 grid((unsigned)((n_vis + 255) / 256), (unsigned)std::min<int64_t>(n_vis, 32768));
  low_overlap_kernel<<<grid, 256, 0, e->stream>>>(tile_store(e), min_overlap, bf.pairs, cap, bf.count);
  CK(cudaGetLastError());
  unsigned long long c = 0;
  CK(cudaMemcpyAsync(&c, bf.count, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *count = (int64_t)c;
  if (cap > 0 && c > 0)
    CK(cudaMemcpy(pairs, bf.pairs, (size_t)std::min<int64_t>(cap, (int64_t)c) * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost));
  return DGRM_OK;
}

int dgrm_ld_pairs(dgrm_engine* e, double threshold, int64_t* pairs, int64_t cap, int64_t* count) {
  int rc = check_ready(e, true);
  if (rc) return rc;
  if (!e->ld_mode) return fail(e, DGRM_E_INVALID, "dgrm_ld_pairs: the engine is not in LD mode (dgrm_ld_begin)");
  if (cap < 0 || (cap > 0 && !pairs) || !count) return fail(e, DGRM_E_INVALID, "dgrm_ld_pairs: bad arguments");
  CK(cudaSetDevice(e->cfg.device));
  struct Bufs {
    int64_t* pairs = nullptr;
    unsigned long long* count = nullptr;
    ~Bufs() { if (pairs) cudaFree(pairs); if (count) cudaFree(count); }
  } bf;
  CK(cudaMalloc(&bf.count, sizeof(unsigned long long)));
  if (cap > 0) CK(cudaMalloc(&bf.pairs, (size_t)cap * 2 * sizeof(int64_t)));
  CK(cudaMemsetAsync(bf.count, 0, sizeof(unsigned long long), e->stream));
  const int64_t n_vis = e->d_subset ? e->n_out : e->n;
  const dim3 grid((unsigned)((n_vis + 255) / 256), (unsigned)std::min<int64_t>(n_vis, 32768));
  outside_range_kernel<<<grid, 256, 0, e->stream>>>(tile_store(e), threshold, bf.pairs, cap, bf.count);
  CK(cudaGetLastError());
  unsigned long long c = 0;
  CK(cudaMemcpyAsync(&c, bf.count, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  *count = (int64_t)c;
  if (cap > 0 && c > 0)
    CK(cudaMemcpy(pairs, bf.pairs, (size_t)std::min<int64_t>(cap, (int64_t)c) * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost));
  return DGRM_OK;
}

// Boundaries, in VISIBLE index space, at which the 128-row block / 256-column panel of the individual behind a visible
// index changes (the output subset is ascending, so each block is one run of visible indices).  Without a subset these
// are the multiples of 128 / 256.
static void visible_block_starts(const dgrm_engine* e, int n, std::vector<int>* rb, std::vector<int>* cp) {
  rb->clear();
  cp->clear();
  int last_rb = -1, last_cp = -1;
  for (int v = 0; v < n; ++v) {
    const int ind = e->h_subset.empty() ? v : e->h_subset[(size_t)v];
    if ((ind >> 7) != last_rb) { rb->push_back(v); last_rb = ind >> 7; }
    if ((ind >> 8) != last_cp) { cp->push_back(v); last_cp = ind >> 8; }
  }
  rb->push_back(n);
  cp->push_back(n);
}

// Owned runs of file column c of the packed (n+1) x n array: rows pr <= c hold N[c][pr] (tile of (c, pr)), rows
// pr > c hold A[pr-1][c] (tile of (pr-1, c)).  Appends [begin, end) row ranges, merged where adjacent.  Ownership is
// constant between two block starts, so a column costs O(n / 128) tests.
static void grm_dat_owned_runs(const dgrm_engine* e, int n, int c, const std::vector<int>& rb, const std::vector<int>& cp,
                               std::vector<std::pair<int, int>>* runs) {
  auto add = [&](int a, int b) {
    if (b <= a) return;
    if (!runs->empty() && runs->back().second == a) runs->back().second = b;
    else runs->push_back({a, b});
  };
  if (e->cfg.world == 1) { add(0, n + 1); return; }
  // N part: visible columns pr in [0, c], cut at the 256-panel starts
  for (size_t k = 0; k + 1 < cp.size() && cp[k] <= c; ++k)
    if (owns_element(e, c, cp[k])) add(cp[k], std::min(cp[k + 1], c + 1));
  // A part: visible rows r in [c, n), cut at the 128-block starts; P row = r + 1
  size_t k = std::upper_bound(rb.begin(), rb.end(), c) - rb.begin() - 1;   // block that contains row c
  for (; k + 1 < rb.size(); ++k) {
    const int r0 = std::max(rb[k], c), r1 = rb[k + 1];
    if (r0 < r1 && owns_element(e, r0, c)) add(r0 + 1, r1 + 1);
  }
}

int dgrm_write_grm_dat_part(dgrm_engine* e, const char* path, int create) {
  int rc = check_ready(e, true);
  if (rc) return rc;
  if (!path) return fail(e, DGRM_E_INVALID, "dgrm_write_grm_dat: null path");
  CK(cudaSetDevice(e->cfg.device));
  const int n = (int)(e->d_subset ? e->n_out : e->n);
  const int64_t col_elems = (int64_t)n + 1;
  struct Res {
    int fd = -1;
    bool created = false, ok = false;
    std::string path;
    double* d_panel = nullptr;
    double* h_panel = nullptr;
    ~Res() {
      if (d_panel) cudaFree(d_panel);
      if (h_panel) cudaFreeHost(h_panel);
      if (fd >= 0) close(fd);
      if (created && !ok) unlink(path.c_str());   // no half-written <out>.grm.dat left behind by its creator
    }
  } rs;
  rs.path = path;
  rs.fd = create ? open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644) : open(path, O_WRONLY);
  if (rs.fd < 0) return fail(e, DGRM_E_IO, "cannot open [ %s ] for writing", path);
  rs.created = create != 0;
  if (create) {
    // kernel.cpp:944-966
    const unsigned char header[14] = {'G', 'R', 'M', 0, 0x5A, 0x99, 0x2, 0x1, (unsigned char)sizeof(double), 0x1, 0, 0, 0, 0};
    if (pwrite(rs.fd, header, 14, 0) != 14 || ftruncate(rs.fd, 14 + col_elems * (int64_t)n * 8) != 0)
      return fail(e, DGRM_E_IO, "short write to [ %s ]", path);
  }
  const int PC = (int)std::max<int64_t>(1, std::min<int64_t>(n, (64 << 20) / (col_elems * 8)));
  CK(cudaMalloc(&rs.d_panel, (size_t)PC * col_elems * sizeof(double)));
  CK(cudaHostAlloc(&rs.h_panel, (size_t)PC * col_elems * sizeof(double), cudaHostAllocDefault));
  const TileStore ts = tile_store(e);
  std::vector<std::vector<std::pair<int, int>>> runs((size_t)PC);
  std::vector<int> rb_starts, cp_starts;
  visible_block_starts(e, n, &rb_starts, &cp_starts);
  for (int c0 = 0; c0 < n; c0 += PC) {
    const int c1 = std::min(c0 + PC, n);
    // the row range of P this engine has anything in, over the panel's columns (the whole column when world == 1)
    int pr0 = n + 1, pr1 = 0;
    for (int c = c0; c < c1; ++c) {
      auto& rv = runs[(size_t)(c - c0)];
      rv.clear();
      grm_dat_owned_runs(e, n, c, rb_starts, cp_starts, &rv);
      if (!rv.empty()) { pr0 = std::min(pr0, rv.front().first); pr1 = std::max(pr1, rv.back().second); }
    }
    if (pr1 <= pr0) continue;
    const int h = pr1 - pr0;
    const dim3 grid((unsigned)((h + 255) / 256), (unsigned)(c1 - c0));
    pack_grm_kernel<<<grid, 256, 0, e->stream>>>(ts, rs.d_panel, c0, c1, pr0, pr1);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rs.h_panel, rs.d_panel, (size_t)(c1 - c0) * h * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (e->cfg.world == 1) {   // one contiguous piece of the file
      const size_t bytes = (size_t)(c1 - c0) * col_elems * sizeof(double);
      const char* src = (const char*)rs.h_panel;
      for (size_t off = 0; off < bytes;) {
        const ssize_t w = pwrite(rs.fd, src + off, bytes - off, 14 + (int64_t)c0 * col_elems * 8 + (int64_t)off);
        if (w <= 0) return fail(e, DGRM_E_IO, "short write to [ %s ]", path);
        off += (size_t)w;
      }
      continue;
    }
    for (int c = c0; c < c1; ++c)
      for (const auto& r : runs[(size_t)(c - c0)]) {
        const char* src = (const char*)(rs.h_panel + (int64_t)(c - c0) * h + (r.first - pr0));
        const size_t bytes = (size_t)(r.second - r.first) * sizeof(double);
        for (size_t off = 0; off < bytes;) {
          const ssize_t w = pwrite(rs.fd, src + off, bytes - off, 14 + ((int64_t)c * col_elems + r.first) * 8 + (int64_t)off);
          if (w <= 0) return fail(e, DGRM_E_IO, "short write to [ %s ]", path);
          off += (size_t)w;
        }
      }
  }
  const int fd = rs.fd;
  rs.fd = -1;
  if (close(fd) != 0) return fail(e, DGRM_E_IO, "close failed on [ %s ]", path);
  rs.ok = true;
  return DGRM_OK;
}

int dgrm_write_grm_dat(dgrm_engine* e, const char* path) {
  if (e && e->cfg.world != 1)
    return fail(e, DGRM_E_INVALID,
                "dgrm_write_grm_dat needs the whole triangle (world == 1); engines of a partition write their own "
                "entries with dgrm_write_grm_dat_part");
  return dgrm_write_grm_dat_part(e, path, 1);
}

int dgrm_set_output_subset(dgrm_engine* e, const int64_t* idx, int64_t count) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->stream));
  free_dev(e->d_subset);
  e->d_subset = nullptr;
  e->h_subset.clear();
  e->n_out = 0;
  if (!idx || count <= 0) return DGRM_OK;
  std::vector<int32_t> v((size_t)count);
  int64_t prev = -1;
  for (int64_t i = 0; i < count; ++i) {
    if (idx[i] <= prev || idx[i] >= e->n)
      return fail(e, DGRM_E_INVALID, "dgrm_set_output_subset: indices must be ascending and < %lld", (long long)e->n);
    prev = idx[i];
    v[(size_t)i] = (int32_t)idx[i];
  }
  CK(cudaMalloc(&e->d_subset, (size_t)count * sizeof(int32_t)));
  CK(cudaMemcpy(e->d_subset, v.data(), (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice));
  e->h_subset = v;
  e->n_out = count;
  return DGRM_OK;
}

int64_t dgrm_num_tiles(const dgrm_engine* e) { return e ? e->ntiles : 0; }

int dgrm_tile_coords(const dgrm_engine* e, int32_t* tile_rc) {
  if (!e || !tile_rc) return DGRM_E_INVALID;
  memcpy(tile_rc, e->tile_rc.data(), e->tile_rc.size() * sizeof(int32_t));
  return DGRM_OK;
}

int dgrm_fetch_tiles_range(dgrm_engine* e, int which, double* tiles, int64_t tile0, int64_t count) {
  int rc = check_ready(e, true);
  if (rc) return rc;
  if (!tiles || which < 0 || which > 2 || tile0 < 0 || count < 0 || tile0 + count > e->ntiles)
    return fail(e, DGRM_E_INVALID, "dgrm_fetch_tiles: bad arguments");
  if (count == 0) return DGRM_OK;
  CK(cudaSetDevice(e->cfg.device));
  const int64_t first = tile0 * kTileElems;
  const int64_t total = count * kTileElems;
  const int64_t slab = 32ll << 20;  // elements per pass (256 MB), two staging buffers
  struct Stage {
    double* d[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    ~Stage() {
      for (int i = 0; i < 2; ++i) {
        if (d[i]) cudaFree(d[i]);
        if (done[i]) cudaEventDestroy(done[i]);
      }
    }
  } st;
  for (int i = 0; i < 2; ++i) {
    CK(cudaMalloc(&st.d[i], (size_t)std::min(total, slab) * sizeof(double)));
    CK(cudaEventCreateWithFlags(&st.done[i], cudaEventDisableTiming));
  }
  TileStore ts = tile_store(e);
  int sl = 0;
  for (int64_t o = 0; o < total; o += slab, sl ^= 1) {
    const int64_t cnt = std::min(slab, total - o);
    TileStore t2 = ts;
    t2.S = ts.S + first + o;
    t2.N = ts.N ? ts.N + first + o : nullptr;
    t2.S2 = ts.S2 ? ts.S2 + first + o : nullptr;
    t2.N2 = ts.N2 ? ts.N2 + first + o : nullptr;
    CK(cudaStreamWaitEvent(e->stream, st.done[sl], 0));
    {
      Prof pr(e, DGRM_K_FETCH);
      export_tiles_kernel<double><<<(unsigned)((cnt + 255) / 256), 256, 0, e->stream>>>(t2, which, cnt, st.d[sl], e->d_tile_rc,
                                                                               (first + o) / kTileElems, e->d_diag);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->stage_packed[sl], e->stream));
    CK(cudaStreamWaitEvent(e->d2h_stream, e->stage_packed[sl], 0));
    CK(cudaMemcpyAsync(tiles + o, st.d[sl], (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, e->d2h_stream));
    CK(cudaEventRecord(st.done[sl], e->d2h_stream));
  }
  CK(cudaStreamSynchronize(e->d2h_stream));
  return DGRM_OK;
}

int dgrm_fetch_tiles(dgrm_engine* e, int which, double* tiles) {
  return dgrm_fetch_tiles_range(e, which, tiles, 0, e ? e->ntiles : 0);
}

// ---- resident cohort -------------------------------------------------------------------------------------

static int store_reserve(dgrm_engine* e, int64_t rows_total) {
  if (rows_total <= e->bed_cap) return DGRM_OK;
  const int64_t pitch = round_up(e->b_fam, 16);
  const int64_t cap = std::max<int64_t>(rows_total, e->bed_cap + e->bed_cap / 2);
  uint8_t* nb = nullptr;
  CK(cudaMalloc(&nb, (size_t)cap * pitch));
  if (e->d_bed && e->bed_rows > 0) {
    cudaError_t st = cudaStreamSynchronize(e->copy_stream);
    if (st == cudaSuccess) st = cudaStreamSynchronize(e->prep);
    if (st == cudaSuccess) st = cudaMemcpy(nb, e->d_bed, (size_t)e->bed_rows * pitch, cudaMemcpyDeviceToDevice);
    if (st != cudaSuccess) {
      cudaFree(nb);
      return fail(e, DGRM_E_CUDA, "growing the resident store failed: %s", cudaGetErrorString(st));
    }
  }
  free_dev(e->d_bed);
  e->d_bed = nb;
  e->bed_cap = cap;
  e->bed_pitch = pitch;
  e->bed_nfam = e->n_fam;
  return DGRM_OK;
}

int dgrm_store_clear(dgrm_engine* e) {
  if (!e) return DGRM_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->copy_stream));
  CK(cudaStreamSynchronize(e->prep));
  free_dev(e->d_bed);
  e->d_bed = nullptr;
  e->bed_rows = e->bed_cap = 0;
  return DGRM_OK;
}

int64_t dgrm_stored_snps(const dgrm_engine* e) { return e ? e->bed_rows : 0; }

int dgrm_store_bed_host(dgrm_engine* e, const uint8_t* rows, int64_t pitch, int64_t nrows) {
  if (!e) return DGRM_E_INVALID;
  if (!e->have_ind) return fail(e, DGRM_E_INVALID, "dgrm_store_bed_host: call dgrm_set_individuals first");
  if (!rows || nrows < 0 || pitch < e->b_fam) return fail(e, DGRM_E_INVALID, "dgrm_store_bed_host: bad rows/pitch");
  if (e->bed_rows + nrows > INT32_MAX) return fail(e, DGRM_E_INVALID, "dgrm_store_bed_host: too many rows");
  CK(cudaSetDevice(e->cfg.device));
  const int rc = store_reserve(e, e->bed_rows + nrows);
  if (rc) return rc;
  if (nrows > 0) {
    CK(cudaMemcpy2DAsync(e->d_bed + e->bed_rows * e->bed_pitch, (size_t)e->bed_pitch, rows, (size_t)pitch,
                         (size_t)e->b_fam, (size_t)nrows, cudaMemcpyHostToDevice, e->copy_stream));
    CK(cudaEventRecord(e->bed_ev, e->copy_stream));
    // pageable sources are staged by the runtime: the call may return before the DMA has read a pinned one
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, rows) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) CK(cudaStreamSynchronize(e->copy_stream));
  }
  e->bed_rows += nrows;
  return DGRM_OK;
}

int dgrm_store_bed_device(dgrm_engine* e, const uint8_t* d_rows, int64_t pitch, int64_t nrows, void* producer_stream) {
  if (!e) return DGRM_E_INVALID;
  if (!e->have_ind) return fail(e, DGRM_E_INVALID, "dgrm_store_bed_device: call dgrm_set_individuals first");
  if (!d_rows || nrows < 0 || pitch < e->b_fam) return fail(e, DGRM_E_INVALID, "dgrm_store_bed_device: bad rows/pitch");
  if (e->bed_rows + nrows > INT32_MAX) return fail(e, DGRM_E_INVALID, "dgrm_store_bed_device: too many rows");
  CK(cudaSetDevice(e->cfg.device));
  const int rc = store_reserve(e, e->bed_rows + nrows);
  if (rc) return rc;
  if (nrows > 0) {
    CK(cudaEventRecord(e->producer_ev, (cudaStream_t)producer_stream));
    CK(cudaStreamWaitEvent(e->copy_stream, e->producer_ev, 0));
    CK(cudaMemcpy2DAsync(e->d_bed + e->bed_rows * e->bed_pitch, (size_t)e->bed_pitch, d_rows, (size_t)pitch,
                         (size_t)e->b_fam, (size_t)nrows, cudaMemcpyDeviceToDevice, e->copy_stream));
    CK(cudaEventRecord(e->bed_ev, e->copy_stream));
    CK(cudaStreamSynchronize(e->copy_stream));
  }
  e->bed_rows += nrows;
  return DGRM_OK;
}

int64_t dgrm_last_d2h_bytes(const dgrm_engine* e) { return e ? e->d2h_bytes : 0; }

int dgrm_store_bed_file(dgrm_engine* e, const char* bed_path, int64_t m_file) {
  if (!e) return DGRM_E_INVALID;
  if (!e->have_ind) return fail(e, DGRM_E_INVALID, "dgrm_store_bed_file: call dgrm_set_individuals first");
  if (!bed_path || m_file <= 0) return fail(e, DGRM_E_INVALID, "dgrm_store_bed_file: bad arguments");
  FILE* f = nullptr;
  int rc = open_bed(e, bed_path, &f);
  if (rc) return rc;
  CK(cudaSetDevice(e->cfg.device));
  rc = store_reserve(e, e->bed_rows + m_file);
  const int64_t slab = std::max<int64_t>(1, (64 << 20) / e->b_fam);
  uint8_t* buf[2] = {nullptr, nullptr};
  for (int i = 0; i < 2 && rc == DGRM_OK; ++i)
    if (cudaHostAlloc(&buf[i], (size_t)slab * e->b_fam, cudaHostAllocDefault) != cudaSuccess)
      rc = fail(e, DGRM_E_NOMEM, "Error: There are not enough memory for reading the binary file.");
  cudaEvent_t done[2] = {nullptr, nullptr};
  for (int i = 0; i < 2 && rc == DGRM_OK; ++i) cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
  int which = 0;
  for (int64_t r0 = 0; r0 < m_file && rc == DGRM_OK; r0 += slab, which ^= 1) {
    const int64_t cnt = std::min<int64_t>(slab, m_file - r0);
    cudaEventSynchronize(done[which]);  // the copy that last read this buffer
    const size_t got = fread(buf[which], (size_t)e->b_fam, (size_t)cnt, f);
    if ((int64_t)got != cnt) {
      rc = fail(e, DGRM_E_BADBED, "Error: An internal error has happened when reading the genotypes file.");
      break;
    }
    if (cudaMemcpy2DAsync(e->d_bed + e->bed_rows * e->bed_pitch, (size_t)e->bed_pitch, buf[which], (size_t)e->b_fam,
                          (size_t)e->b_fam, (size_t)cnt, cudaMemcpyHostToDevice, e->copy_stream) != cudaSuccess) {
      rc = fail(e, DGRM_E_CUDA, "dgrm_store_bed_file: host to device copy failed");
      break;
    }
    cudaEventRecord(done[which], e->copy_stream);
    e->bed_rows += cnt;
  }
  cudaEventRecord(e->bed_ev, e->copy_stream);
  cudaStreamSynchronize(e->copy_stream);
  for (int i = 0; i < 2; ++i) {
    if (buf[i]) cudaFreeHost(buf[i]);
    if (done[i]) cudaEventDestroy(done[i]);
  }
  fclose(f);
  return rc;
}

int dgrm_push_stored(dgrm_engine* e, int64_t first, int64_t nrows, const uint8_t* row_keep) {
  int rc = check_ready(e, false);
  if (rc) return rc;
  if (e->finished) return fail(e, DGRM_E_INVALID, "dgrm_push_stored after dgrm_finish");
  if (!e->d_bed || first < 0 || nrows < 0 || first + nrows > e->bed_rows)
    return fail(e, DGRM_E_INVALID, "dgrm_push_stored: rows [%lld, %lld) are not in the resident store (%lld rows)",
                (long long)first, (long long)(first + nrows), (long long)e->bed_rows);
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamWaitEvent(e->prep, e->bed_ev, 0));
  for (int64_t r = 0; r < nrows; ++r)
    if (!row_keep || row_keep[r]) e->grm_rows.push_back((int32_t)(first + r));
  for (int64_t r0 = 0; r0 < nrows; r0 += (1 << 20)) {
    const int64_t cnt = std::min<int64_t>(nrows - r0, 1 << 20);
    rc = push_device_rows(e, e->d_bed + (first + r0) * e->bed_pitch, e->bed_pitch, cnt, row_keep ? row_keep + r0 : nullptr);
    if (rc) return rc;
  }
  return DGRM_OK;
}

int dgrm_fetch_standardised(dgrm_engine* e, double* local, int64_t lld, int bs_r, int bs_c, int my_row, int my_col,
                            int R, int C) {
  int rc = check_ready(e, true);
  if (rc) return rc;
  if (!local || bs_r < 1 || bs_c < 1 || R < 1 || C < 1 || my_row < 0 || my_row >= R || my_col < 0 || my_col >= C)
    return fail(e, DGRM_E_INVALID, "dgrm_fetch_standardised: bad arguments");
  if (!e->grm_from_store || !e->d_bed || (int64_t)e->grm_rows.size() != e->m)
    return fail(e, DGRM_E_INVALID,
                "dgrm_fetch_standardised: the genotypes of this GRM are not resident (use dgrm_store_bed_* + dgrm_push_stored)");
  CK(cudaSetDevice(e->cfg.device));
  CyclicGeom g;
  g.bs_r = bs_r; g.bs_c = bs_c; g.my_row = my_row; g.my_col = my_col; g.R = R; g.C = C;
  g.nloc_r = numroc((int)e->m, bs_r, my_row, R);
  g.nloc_c = numroc((int)e->n, bs_c, my_col, C);
  if (g.nloc_r == 0 || g.nloc_c == 0) return DGRM_OK;
  if (lld < g.nloc_r) return fail(e, DGRM_E_INVALID, "dgrm_fetch_standardised: lld %lld < local rows %d", (long long)lld, g.nloc_r);
  if (e->m > e->d_grm_rows_cap) {
    free_dev(e->d_grm_rows);
    e->d_grm_rows = nullptr;
    CK(cudaMalloc(&e->d_grm_rows, (size_t)e->m * sizeof(int32_t)));
    e->d_grm_rows_cap = e->m;
  }
  CK(cudaMemcpyAsync(e->d_grm_rows, e->grm_rows.data(), (size_t)e->m * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  StandardisedArgs a;
  a.bed = e->d_bed; a.pitch = e->bed_pitch; a.rows = e->d_grm_rows; a.src_ind = e->d_src_ind;
  a.p2 = e->d_p2; a.sd = e->d_sd; a.m = (int)e->m; a.n = (int)e->n;
  const dim3 blk(128, 2);
  cudaPointerAttributes attr;
  const bool pinned = cudaPointerGetAttributes(&attr, local) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
                      attr.devicePointer != nullptr;
  cudaGetLastError();
  // columns go through a device panel and the copy engine; a pinned destination receives them directly
  const int PC = (int)std::max<int64_t>(1, std::min<int64_t>(g.nloc_c, (128 << 20) / ((int64_t)g.nloc_r * 8)));
  double* d_panel[2] = {nullptr, nullptr};
  double* h_panel[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  rc = DGRM_OK;
  for (int i = 0; i < 2 && rc == DGRM_OK; ++i) {
    if (cudaMalloc(&d_panel[i], (size_t)PC * g.nloc_r * sizeof(double)) != cudaSuccess ||
        (!pinned && cudaHostAlloc(&h_panel[i], (size_t)PC * g.nloc_r * sizeof(double), cudaHostAllocDefault) != cudaSuccess) ||
        cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess)
      rc = fail(e, DGRM_E_NOMEM, "dgrm_fetch_standardised: cannot allocate the staging panels");
  }
  g.lld = g.nloc_r;
  int prev = -1, p0 = 0, p1 = 0, slot = 0;
  auto drain = [&](int sl, int lc0, int lc1) {
    cudaEventSynchronize(done[sl]);
    if (!pinned)
      for (int lc = lc0; lc < lc1; ++lc)
        memcpy(local + (int64_t)lc * lld, h_panel[sl] + (int64_t)(lc - lc0) * g.nloc_r, (size_t)g.nloc_r * sizeof(double));
  };
  for (int lc0 = 0; lc0 < g.nloc_c && rc == DGRM_OK; lc0 += PC, slot ^= 1) {
    const int lc1 = std::min(lc0 + PC, g.nloc_c);
    {
      Prof pr(e, DGRM_K_FETCH);
      const dim3 grid((g.nloc_r + 127) / 128, (lc1 - lc0 + 1) / 2);
      bed_standardised_kernel<<<grid, blk, 0, e->stream>>>(a, d_panel[slot], g, lc0, lc1);
    }
    cudaError_t st = cudaGetLastError();
    if (st == cudaSuccess)
      st = pinned ? cudaMemcpy2DAsync(local + (int64_t)lc0 * lld, (size_t)lld * sizeof(double), d_panel[slot],
                                      (size_t)g.nloc_r * sizeof(double), (size_t)g.nloc_r * sizeof(double),
                                      (size_t)(lc1 - lc0), cudaMemcpyDeviceToHost, e->stream)
                  : cudaMemcpyAsync(h_panel[slot], d_panel[slot], (size_t)(lc1 - lc0) * g.nloc_r * sizeof(double),
                                    cudaMemcpyDeviceToHost, e->stream);
    if (st != cudaSuccess) {
      rc = fail(e, DGRM_E_CUDA, "dgrm_fetch_standardised: %s", cudaGetErrorString(st));
      break;
    }
    cudaEventRecord(done[slot], e->stream);
    if (prev >= 0) drain(prev, p0, p1);
    prev = slot; p0 = lc0; p1 = lc1;
  }
  if (prev >= 0 && rc == DGRM_OK) drain(prev, p0, p1);
  cudaStreamSynchronize(e->stream);
  for (int i = 0; i < 2; ++i) {
    free_dev(d_panel[i]);
    if (h_panel[i]) cudaFreeHost(h_panel[i]);
    if (done[i]) cudaEventDestroy(done[i]);
  }
  return rc;
}

// ---- (S, N) algebra --------------------------------------------------------------------------------------

int dgrm_save(dgrm_engine* e, int mode) {
  if (!e) return DGRM_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  if (mode == DGRM_SAVE_DROP) {
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaStreamSynchronize(e->out_stream));
    free_dev(e->d_S2); e->d_S2 = nullptr;
    free_dev(e->d_N2); e->d_N2 = nullptr;
    e->saved = false;
    e->view = DGRM_VIEW_CURRENT;
    return DGRM_OK;
  }
  int rc = check_ready(e, true);
  if (rc) return rc;
  if (mode != DGRM_SAVE_COPY && mode != DGRM_SAVE_ADD) return fail(e, DGRM_E_INVALID, "dgrm_save: bad mode");
  if (e->d_subset) return fail(e, DGRM_E_INVALID, "dgrm_save: clear the output subset first");
  const bool flat = e->cfg.flat_norm != 0;
  const size_t total = (size_t)e->ntiles * kTileElems;
  if (!e->d_S2 && total > 0) {
    CK(cudaMalloc(&e->d_S2, total * sizeof(double)));
    if (!flat) CK(cudaMalloc(&e->d_N2, total * sizeof(int32_t)));
  }
  if (mode == DGRM_SAVE_ADD && e->saved) {
    if (e->saved_flat != flat) return fail(e, DGRM_E_INVALID, "dgrm_save: flat and per-pair normalisation cannot be added");
    if (total > 0) {
      accumulate_store_kernel<<<e->sm_count * 8, 256, 0, e->stream>>>(e->d_S2, flat ? nullptr : e->d_N2, e->d_S, e->d_N,
                                                                    (int64_t)total);
      CK(cudaGetLastError());
    }
    e->m2 += e->m;
  } else {
    if (total > 0) {
      CK(cudaMemcpyAsync(e->d_S2, e->d_S, total * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
      if (!flat) CK(cudaMemcpyAsync(e->d_N2, e->d_N, total * sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
    }
    e->m2 = e->m;
    e->saved_flat = flat;
  }
  CK(cudaStreamSynchronize(e->stream));
  e->saved = true;
  return DGRM_OK;
}

int dgrm_set_view(dgrm_engine* e, int view) {
  if (!e) return DGRM_E_INVALID;
  if (view < DGRM_VIEW_CURRENT || view > DGRM_VIEW_SAVED_PLUS_CURRENT) return fail(e, DGRM_E_INVALID, "dgrm_set_view: bad view");
  if (view != DGRM_VIEW_CURRENT && !e->saved) return fail(e, DGRM_E_INVALID, "dgrm_set_view: nothing has been saved (dgrm_save)");
  if (view >= DGRM_VIEW_SAVED_MINUS_CURRENT && e->saved_flat != (e->cfg.flat_norm != 0))
    return fail(e, DGRM_E_INVALID, "dgrm_set_view: flat and per-pair normalisation cannot be combined");
  e->view = view;
  return DGRM_OK;
}


void* dgrm_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0) return nullptr;
  if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void dgrm_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int dgrm_kernel_times(dgrm_engine* e, double ms[DGRM_K_NUM], int64_t launches[DGRM_K_NUM]) {
  if (!e) return DGRM_E_INVALID;
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize(e->prep));
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaStreamSynchronize(e->out_stream));
  CK(cudaStreamSynchronize(e->d2h_stream));
  if (getenv("DGRM_TRACE")) {  // launch timeline of everything recorded since the last call (stderr; experiments)
    static const char* names[DGRM_K_NUM] = {"count", "expand", "syrk", "paircnt", "final", "fetch", "gather", "d2h"};
    cudaEvent_t base = nullptr;
    float best = 0.f;
    for (int k = 0; k < DGRM_K_NUM; ++k)
      for (auto& p : e->ev[k]) {
        float t = 0.f;
        if (!base) { base = p.a; continue; }
        if (cudaEventElapsedTime(&t, base, p.a) == cudaSuccess && t < best) { base = p.a; best = 0.f; }
      }
    struct Row { float t0, dt; int k; };
    std::vector<Row> rows;
    for (int k = 0; k < DGRM_K_NUM; ++k)
      for (auto& p : e->ev[k]) {
        Row r{0.f, 0.f, k};
        cudaEventElapsedTime(&r.t0, base, p.a);
        cudaEventElapsedTime(&r.dt, p.a, p.b);
        rows.push_back(r);
      }
    std::sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.t0 < b.t0; });
    for (auto& r : rows) fprintf(stderr, "[dgrm trace r%d] %9.3f ms  +%8.3f  %s\n", e->cfg.rank, r.t0, r.dt, names[r.k]);
  }
  for (int k = 0; k < DGRM_K_NUM; ++k) {
    for (auto& p : e->ev[k]) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) e->ms_acc[k] += t;
      e->ev_pool.push_back(p);
    }
    e->ev[k].clear();
    if (ms) ms[k] = e->ms_acc[k];
    if (launches) launches[k] = e->launches[k];
    e->ms_acc[k] = 0;
    e->launches[k] = 0;
  }
  cudaGetLastError();
  return DGRM_OK;
}

}  // extern "C"
