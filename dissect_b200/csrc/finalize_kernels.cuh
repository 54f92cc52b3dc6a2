// Kernels that read the tile store: exact-diagonal fix-up, A = S * (1/N) on the fly, block-cyclic
// gather (Matrix::m layout, matrix.cpp:435-487), sanitizeKernel's low-overlap scan
// (kernel.cpp:1993-2037) and the packed .grm.dat layout (matrix.cpp:2324-2334).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "syrk_tcgen05.cuh"

namespace dgrm {

struct TileStore {
  const double* S;        // [tiles][256][128]
  const int32_t* N;       // [tiles][256][128], NULL with flat normalisation
  const int32_t* index;   // [nrb][ncp] -> storage index or -1 (not owned / above the diagonal)
  int ncp;                // column panels
  int n;                  // individuals visible to fetch / pack (after an output subset)
  double flat_n;          // value of N under --grm-flat-norm
  const int32_t* subset;  // visible index -> individual (ascending), NULL = identity
  // (S, N) algebra with the saved store (Kernel::addKernels, kernel.cpp:1705-1859); same tile layout
  const double* S2;       // saved S (diagonal already exact)
  const int32_t* N2;      // saved N, NULL when the saved GRM is flat-normalised
  double flat_n2;
  int view;               // 0 current, 1 saved, 2 saved - current, 3 saved + current
};

// Unnormalised S and N at storage element `slot` under the store's view.  `s_cur` is the current S at that
// element (the caller may have replaced it by the exact diagonal).
__device__ __forceinline__ void view_sn(const TileStore& ts, int64_t slot, double s_cur, double* s, double* n) {
  const double n_cur = ts.N ? (double)ts.N[slot] : ts.flat_n;
  if (ts.view == 0) { *s = s_cur; *n = n_cur; return; }
  const double s2 = ts.S2[slot];
  const double n2 = ts.N2 ? (double)ts.N2[slot] : ts.flat_n2;
  if (ts.view == 1) { *s = s2; *n = n2; }
  else if (ts.view == 2) { *s = __dsub_rn(s2, s_cur); *n = n2 - n_cur; }   // pdgeadd_, matrix.cpp:3005
  else { *s = __dadd_rn(s2, s_cur); *n = n2 + n_cur; }
}

__device__ __forceinline__ double view_value(int which, double s, double n) {
  if (which == 2) return n;
  if (which == 1) return s;
  // matrix.cpp:4096: m[i] *= scale / m1->m[i] with scale = 1.  -> two roundings
  return __dmul_rn(s, __ddiv_rn(1., n));
}

__device__ __forceinline__ int64_t tile_slot(const TileStore& ts, int r, int c) {
  if (ts.subset) { r = ts.subset[r]; c = ts.subset[c]; }
  const int sidx = ts.index[(r >> 7) * ts.ncp + (c >> 8)];
  if (sidx < 0) return -1;
  return (int64_t)sidx * kTileElems + (int64_t)(c & 255) * kTileM + (r & 127);
}

// value of matrix `which` at lower-triangle coordinate (r >= c); false when the tile is not owned
__device__ __forceinline__ bool tile_value(const TileStore& ts, int which, int r, int c, double* out) {
  const int64_t slot = tile_slot(ts, r, c);
  if (slot < 0) return false;
  double s, n;
  view_sn(ts, slot, ts.S[slot], &s, &n);
  *out = view_value(which, s, n);
  return true;
}

// Overwrite the tensor-core diagonal of S with the exact fp64 sum of squares from bed_expand.
__global__ void fix_diag_kernel(double* __restrict__ S, const int32_t* __restrict__ index, int ncp, int n,
                                const double* __restrict__ diag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int sidx = index[(i >> 7) * ncp + (i >> 8)];
  if (sidx < 0) return;
  S[(int64_t)sidx * kTileElems + (int64_t)(i & 255) * kTileM + (i & 127)] = diag[i];
}

struct CyclicGeom {
  int bs_r, bs_c, my_row, my_col, R, C;
  int nloc_r, nloc_c;  // numroc
  int64_t lld;
};

__device__ __forceinline__ int local_to_global(int l, int bs, int me, int P) {
  return ((l / bs) * P + me) * bs + (l % bs);
}

// dst[(lc - lc0) * lld + lr] for local columns [lc0, lc1): 2-D launch, x along local rows.
__global__ void fetch_block_cyclic_kernel(const TileStore ts, int which, double* __restrict__ dst, CyclicGeom g,
                                          int lc0, int lc1, int uplo /* 0 L, 1 U, 2 B */) {
  const int lr = blockIdx.x * blockDim.x + threadIdx.x;
  const int lc = lc0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (lr >= g.nloc_r || lc >= lc1) return;
  const int gr = local_to_global(lr, g.bs_r, g.my_row, g.R);
  const int gc = local_to_global(lc, g.bs_c, g.my_col, g.C);
  if (gr >= ts.n || gc >= ts.n) return;
  const bool lower = gc <= gr;
  if (uplo == 0 && !lower) return;
  if (uplo == 1 && gc < gr) return;
  double v;
  if (!tile_value(ts, which, lower ? gr : gc, lower ? gc : gr, &v)) return;
  dst[(int64_t)(lc - lc0) * g.lld + lr] = v;
}

// Streaming output: local rows [lr0, lr1) x local columns [0, lc1) of the caller's block-cyclic matrix, packed
// column-major into a device staging buffer (pitch lr1 - lr0) that the copy engine then moves to the host while
// the tensor cores work on later tile groups.  The rectangle belongs to one finished raster row group; entries
// above the diagonal (inside its diagonal block) get the symmetric value, tiles this engine does not own get 0.
// `diag` supplies the exact fp64 diagonal of S (the tile store itself is patched after the last group).
// Transport encodings of a value leaving the device.  double: as is.  int: exact pair counts as int32.
// uint16_t: pair counts as the deficiency m - N (the number of SNPs at which either individual is missing), used
// when the host has established 2 * max_i(missing_i) <= 65534; 0xFFFF marks "no value" (decoded as N = 0).
template <typename T>
__device__ __forceinline__ T transport_encode(double v, bool valid, double m_total) {
  return (T)v;
}
template <>
__device__ __forceinline__ uint16_t transport_encode<uint16_t>(double v, bool valid, double m_total) {
  return valid ? (uint16_t)(int)(m_total - v) : (uint16_t)0xFFFF;
}

template <>
__device__ __forceinline__ float transport_encode<float>(double v, bool valid, double m_total) {
  return valid ? (float)v : 0.f;
}

// out[i] = matrix `which` at (i, i) of the visible individuals under the store's view, from the exact fp64 diagonal
// of S (0 when this engine does not own the diagonal tile).
__global__ void diag_values_kernel(const TileStore ts, int which, const double* __restrict__ diag, double* __restrict__ out,
                                   int n_vis) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vis) return;
  double v = 0.;
  const int64_t slot = tile_slot(ts, i, i);
  if (slot >= 0) {
    double sv, nv;
    view_sn(ts, slot, diag[ts.subset ? ts.subset[i] : i], &sv, &nv);
    v = view_value(which, sv, nv);
  }
  out[i] = v;
}

template <typename T>
__global__ void pack_rows_kernel(const TileStore ts, int which, const double* __restrict__ diag,
                                 T* __restrict__ stage, const CyclicGeom g, int lr0, int lr1, int lc1) {
  const int lr = lr0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int lc = blockIdx.y * blockDim.y + threadIdx.y;
  if (lr >= lr1 || lc >= lc1) return;
  const int gr = local_to_global(lr, g.bs_r, g.my_row, g.R);
  const int gc = local_to_global(lc, g.bs_c, g.my_col, g.C);
  double v = 0.;
  bool valid = false;
  if (gr < ts.n && gc < ts.n) {
    const int r = gr >= gc ? gr : gc, c = gr >= gc ? gc : gr;
    const int64_t slot = tile_slot(ts, r, c);
    if (slot >= 0) {
      double sv, nv;
      view_sn(ts, slot, (r == c) ? diag[ts.subset ? ts.subset[r] : r] : ts.S[slot], &sv, &nv);
      v = view_value(which, sv, nv);
      valid = true;
    }
  }
  // T = int / uint16_t only for the (exact, integer) pair counts
  stage[(int64_t)lc * (lr1 - lr0) + (lr - lr0)] = transport_encode<T>(v, valid, ts.flat_n);
}

// sanitizeKernel's scan: strictly-lower pairs with -0.1 < N < min_overlap - 0.1
__global__ void low_overlap_kernel(const TileStore ts, double min_overlap, int64_t* __restrict__ pairs, int64_t cap,
                                   unsigned long long* __restrict__ count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ts.n) return;
  for (int c = blockIdx.y; c < r; c += gridDim.y) {
    double nv;
    if (!tile_value(ts, 2, r, c, &nv)) continue;
    if (nv > -0.1 && nv < min_overlap - 0.1) {
      const unsigned long long k = atomicAdd(count, 1ull);
      if ((int64_t)k < cap) {
        pairs[2 * k] = r;
        pairs[2 * k + 1] = c;
      }
    }
  }
}

// LD mode: strictly-lower pairs whose value lies outside [-threshold, threshold] — what
// Matrix::getGlobalIndexOutsideRange(-t, t) reports for a symmetric 'L' matrix (matrix.cpp:4279-4307) and
// GWAS::getLessSignificantCorrelatedSNPs then keeps (idxS1 > idxS2, gwas.cpp:1191).  NaN compares false, as there.
__global__ void outside_range_kernel(const TileStore ts, double threshold, int64_t* __restrict__ pairs, int64_t cap,
                                     unsigned long long* __restrict__ count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ts.n) return;
  for (int c = blockIdx.y; c < r; c += gridDim.y) {
    double v;
    if (!tile_value(ts, 0, r, c, &v)) continue;
    if (v < -threshold || v > threshold) {
      const unsigned long long k = atomicAdd(count, 1ull);
      if ((int64_t)k < cap) {
        pairs[2 * k] = r;
        pairs[2 * k + 1] = c;
      }
    }
  }
}

// Packed (n+1) x n column-major array of packMatrices (matrix.cpp:2324-2334), columns [c0, c1):
//   P[r][c] = N[r][c]   r = 0..c       (upper triangle of N incl. diagonal)
//   P[r+1][c] = A[r][c] r = c..n-1     (lower triangle of A incl. diagonal)
// Only rows [pr0, pr1) of P are produced (pitch pr1 - pr0): an engine of a partition packs just the row range its
// tiles reach.  Entries of tiles the engine does not own are written as 0 and never leave the host staging.
__global__ void pack_grm_kernel(const TileStore ts, double* __restrict__ dst, int c0, int c1, int pr0, int pr1) {
  const int pr = pr0 + blockIdx.x * blockDim.x + threadIdx.x;  // row of P, 0..n
  const int c = c0 + blockIdx.y;
  if (pr >= pr1 || c >= c1) return;
  double v = 0.;
  if (pr <= c) tile_value(ts, 2, c, pr, &v);   // N[pr][c] = N[c][pr]
  else tile_value(ts, 0, pr - 1, c, &v);       // A[pr-1][c]
  dst[(int64_t)(c - c0) * (pr1 - pr0) + (pr - pr0)] = v;
}

// Per-individual count of missing genotypes over a packed chunk (rows = SNPs, 4 individuals per byte, pattern 01 =
// missing).  Bounds every pair's deficiency m - N_ij <= missing_i + missing_j, which is what lets streamed pair
// counts cross PCIe as 16-bit values.  One thread per byte column, 256 rows per block row.
__global__ void bed_missing_per_individual_kernel(const uint8_t* __restrict__ chunk, int64_t pitch, int kc, int nbytes,
                                                  int n, int32_t* __restrict__ miss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbytes) return;
  const int r0 = blockIdx.y * 256, r1 = min(kc, r0 + 256);
  uint32_t acc = 0;  // four 8-bit counters (<= 256 each: the last increment may carry, so flush at 255)
  int c[4] = {0, 0, 0, 0};
  int since = 0;
  for (int r = r0; r < r1; ++r) {
    const uint32_t v = chunk[(int64_t)r * pitch + b];
    const uint32_t m = v & ~(v >> 1) & 0x55u;        // bit 2j set <=> field j == 01
    acc += (m & 1u) | ((m & 4u) << 6) | ((m & 16u) << 12) | ((m & 64u) << 18);
    if (++since == 255) {
      for (int j = 0; j < 4; ++j) c[j] += (acc >> (8 * j)) & 0xFF;
      acc = 0; since = 0;
    }
  }
  for (int j = 0; j < 4; ++j) {
    const int cnt = c[j] + (int)((acc >> (8 * j)) & 0xFF);
    const int i = 4 * b + j;
    if (cnt > 0 && i < n) atomicAdd(&miss[i], cnt);
  }
}

// dgrm_save(DGRM_SAVE_ADD): saved (S, N) += current (S, N), element by element of the tile store.
__global__ void accumulate_store_kernel(double* __restrict__ S2, int32_t* __restrict__ N2, const double* __restrict__ S,
                                        const int32_t* __restrict__ N, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    S2[i] = __dadd_rn(S2[i], S[i]);
    if (N2) N2[i] += N[i];
  }
}

// Standardised genotypes (genotype.cpp:943-961) of the current GRM's SNPs straight from the resident packed rows,
// into a block-cyclic m x n matrix (row = SNP, column = kept individual), local columns [lc0, lc1).
struct StandardisedArgs {
  const uint8_t* bed;      // resident store
  int64_t pitch;
  const int32_t* rows;     // [m] stored row of every SNP of the GRM
  const int32_t* src_ind;  // kept individual -> .fam position, NULL = identity
  const double* p2;
  const double* sd;
  int m, n;
};

__global__ void bed_standardised_kernel(const StandardisedArgs a, double* __restrict__ dst, const CyclicGeom g, int lc0,
                                        int lc1) {
  const int lr = blockIdx.x * blockDim.x + threadIdx.x;
  const int lc = lc0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (lr >= g.nloc_r || lc >= lc1) return;
  const int snp = local_to_global(lr, g.bs_r, g.my_row, g.R);
  const int ind = local_to_global(lc, g.bs_c, g.my_col, g.C);
  if (snp >= a.m || ind >= a.n) return;
  const int f = a.src_ind ? a.src_ind[ind] : ind;
  const int pat = (a.bed[(int64_t)a.rows[snp] * a.pitch + (f >> 2)] >> (2 * (f & 3))) & 3;
  double z = 0.;  // 01 = missing
  if (pat != 1) {
    const double code = pat == 0 ? 1. : (pat == 2 ? 2. : 3.);   // genotype.cpp:752-776
    const double shift = __dadd_rn(1., __dmul_rn(2., a.p2[snp]));
    z = __ddiv_rn(__dsub_rn(code, shift), a.sd[snp]);
  }
  dst[(int64_t)(lc - lc0) * g.lld + lr] = z;
}

// tile export as doubles (multi-engine gather).  ts.S / ts.N already point at the first exported element;
// tile_rc[t] = (first global row, first global column) of storage tile tile0 + t lets the exact fp64 diagonal
// override the tensor-core one (the tile store itself is only patched after the last tile group).
template <typename T>
__global__ void export_tiles_kernel(const TileStore ts, int which, int64_t total, T* __restrict__ dst,
                                    const int2* __restrict__ tile_rc, int64_t tile0, const double* __restrict__ diag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double sc = ts.S[i];
  const int2 rc = tile_rc[tile0 + i / kTileElems];
  const int e = (int)(i % kTileElems);
  const int gr = rc.x + (e & 127), gc = rc.y + (e >> 7);
  if (gr == gc && gr < ts.n) sc = diag[gr];
  double sv, nv;
  view_sn(ts, i, sc, &sv, &nv);
  const double v = (which == 0 && nv == 0.) ? 0. : view_value(which, sv, nv);   // padding rows: N = 0
  dst[i] = transport_encode<T>(v, gr < ts.n && gc < ts.n, ts.flat_n);
}

}  // namespace dgrm
