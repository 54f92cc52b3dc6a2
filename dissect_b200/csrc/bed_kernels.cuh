// PLINK .bed 2-bit decode, per-SNP statistics, standardisation and operand staging.
//
// Reference being replaced (all in genotype.cpp):
//   parseSNPbyte :741-787   2-bit pattern -> code {00:1 hom A1, 01:0 missing, 10:2 het, 11:3 hom A2},
//                           frequencies[code]++, missings = (code != 0)
//   parseSNP     :726-739   nNonMissing, p1, p2, standardDev
//   normalizeGenotypes :888-961   z = (code - 1*(1 + 2*p2)) / sd, missing stays 0
// The fp64 genotype / missing matrices the reference materialises (16 bytes per genotype) never
// exist here: the device keeps the packed 2-bit rows and writes tensor-core operand tiles directly.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgrm {

// ------------------------------------------------------------------------------------------------
// Row / individual gather into the aligned chunk buffer.
// dst row r (pitch dst_pitch, a multiple of 128) <- src row rows[r] (or row0 + r when rows == NULL).
// src_ind == NULL: plain byte copy of nbytes per row.  Otherwise output individual i takes the 2-bit
// field of .fam individual src_ind[i] (individualBEDIdxsMaskToKeep compaction, genotype.cpp:750-781).
// ------------------------------------------------------------------------------------------------
template <typename VecT>
__global__ void bed_gather_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, const int32_t* __restrict__ rows,
                                  int64_t row0, int nrows, uint8_t* __restrict__ dst, int64_t dst_pitch,
                                  int nbytes, const int32_t* __restrict__ src_ind, int n) {
  constexpr int V = (int)sizeof(VecT);
  const int r = blockIdx.y;
  if (r >= nrows) return;
  const int64_t srow = rows ? (int64_t)rows[r] : row0 + r;
  const uint8_t* s = src + srow * src_pitch;
  uint8_t* d = dst + (int64_t)r * dst_pitch;
  if (src_ind == nullptr) {
    // plain row copy, VecT chosen by the host from the alignment of (src, src_pitch); dst rows are 128-aligned
    const int nvec = nbytes / V;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nvec; b += gridDim.x * blockDim.x)
      reinterpret_cast<VecT*>(d)[b] = __ldg(reinterpret_cast<const VecT*>(s) + b);
    if (blockIdx.x == 0)
      for (int b = nvec * V + threadIdx.x; b < nbytes; b += blockDim.x) d[b] = s[b];
    return;
  }
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nbytes; b += gridDim.x * blockDim.x) {
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = 4 * b + j;
      if (i < n) {
        const int f = src_ind[i];
        o |= (uint32_t)((s[f >> 2] >> (2 * (f & 3))) & 3) << (2 * j);
      }
    }
    d[b] = (uint8_t)o;
  }
}

// ------------------------------------------------------------------------------------------------
// K1a: per-SNP allele / missing counts and statistics.  One warp per SNP row, 128-bit loads,
// popcount on the lo/hi bit planes, warp-shuffle reduction, lane 0 does the fp64 arithmetic in the
// reference's operation order (no FMA contraction: explicit round-to-nearest intrinsics).
//   zt[row][pattern] = standardised value for .bed bit pattern {00, 01, 10, 11}
// ------------------------------------------------------------------------------------------------
struct SnpStatsOut {
  int32_t* freq;  // [m][4]
  int32_t* nn;    // [m]
  double* p1;
  double* p2;
  double* sd;
  int32_t* bad_count;  // [2]: {sd == 0, no non-missing}
  int64_t* bad_list;   // first bad_cap offenders (global SNP index, negative-coded for NaN)
  int bad_cap;
};

__global__ void __launch_bounds__(256)
bed_count_stats_kernel(const uint8_t* __restrict__ chunk, int64_t pitch, int nrows, int n, int64_t snp0,
                       SnpStatsOut o, double* __restrict__ zt) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const uint4* p = reinterpret_cast<const uint4*>(chunk + (int64_t)row * pitch);
  const int nvec = (n + 63) >> 6;  // 64 individuals per 16-byte vector
  int c0 = 0, c2 = 0, c3 = 0;
  for (int j = lane; j < nvec; j += 32) {
    const uint4 d = __ldg(p + j);
    const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i0 = (j * 4 + t) * 16;
      int valid = n - i0;
      valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
      const uint32_t km = (valid == 16) ? 0x55555555u : (((1u << (2 * valid)) - 1u) & 0x55555555u);
      const uint32_t lo = w[t] & 0x55555555u;
      const uint32_t hi = (w[t] >> 1) & 0x55555555u;
      c3 += __popc(lo & hi & km);   // 11 hom A2
      c2 += __popc(hi & ~lo & km);  // 10 het
      c0 += __popc(lo & ~hi & km);  // 01 missing
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, s);
    c2 += __shfl_xor_sync(0xffffffffu, c2, s);
    c3 += __shfl_xor_sync(0xffffffffu, c3, s);
  }
  if (lane == 0) {
    const int c1 = n - c0 - c2 - c3;  // 00 hom A1
    const int64_t g = snp0 + row;
    o.freq[g * 4 + 0] = c0;
    o.freq[g * 4 + 1] = c1;
    o.freq[g * 4 + 2] = c2;
    o.freq[g * 4 + 3] = c3;
    const int nn = c1 + c2 + c3;
    o.nn[g] = nn;
    // genotype.cpp:735-738
    const double den = __dmul_rn(2., (double)nn);
    const double p1 = __ddiv_rn(__dadd_rn(__dmul_rn(2., (double)c1), (double)c2), den);
    const double p2 = __ddiv_rn(__dadd_rn(__dmul_rn(2., (double)c3), (double)c2), den);
    const double sd = __dsqrt_rn(__dmul_rn(__dmul_rn(2., p1), __dsub_rn(1., p1)));
    o.p1[g] = p1;
    o.p2[g] = p2;
    o.sd[g] = sd;
    // genotype.cpp:922-923, 948-949, 959
    const double shift = __dadd_rn(1., __dmul_rn(2., p2));
    double z1 = __ddiv_rn(__dsub_rn(1., shift), sd);
    double z2 = __ddiv_rn(__dsub_rn(2., shift), sd);
    double z3 = __ddiv_rn(__dsub_rn(3., shift), sd);
    const bool nan_snp = (nn == 0);
    const bool zero_sd = !nan_snp && (sd == 0.);
    if (nan_snp || zero_sd) {
      z1 = z2 = z3 = 0.;  // keep NaN/Inf out of the accumulators; dgrm_finish reports the error
      const int slot = atomicAdd(&o.bad_count[nan_snp ? 1 : 0], 1);
      const int total = atomicAdd(&o.bad_count[2], 1);
      (void)slot;
      if (total < o.bad_cap) o.bad_list[total] = nan_snp ? -(g + 1) : g;
    }
    double* z = zt + (int64_t)row * 4;
    z[0] = z1;  // 00
    z[1] = 0.;  // 01 missing
    z[2] = z2;  // 10
    z[3] = z3;  // 11
  }
}

// lane 0 of the counting warp: statistics in the reference's operation order (genotype.cpp:735-738, 922-959)
__device__ __forceinline__ void snp_stats_finish(int c0, int c2, int c3, int n, int64_t g, int row, const SnpStatsOut& o,
                                                 double* __restrict__ zt) {
  const int c1 = n - c0 - c2 - c3;  // 00 hom A1
  o.freq[g * 4 + 0] = c0;
  o.freq[g * 4 + 1] = c1;
  o.freq[g * 4 + 2] = c2;
  o.freq[g * 4 + 3] = c3;
  const int nn = c1 + c2 + c3;
  o.nn[g] = nn;
  const double den = __dmul_rn(2., (double)nn);
  const double p1 = __ddiv_rn(__dadd_rn(__dmul_rn(2., (double)c1), (double)c2), den);
  const double p2 = __ddiv_rn(__dadd_rn(__dmul_rn(2., (double)c3), (double)c2), den);
  const double sd = __dsqrt_rn(__dmul_rn(__dmul_rn(2., p1), __dsub_rn(1., p1)));
  o.p1[g] = p1;
  o.p2[g] = p2;
  o.sd[g] = sd;
  const double shift = __dadd_rn(1., __dmul_rn(2., p2));
  double z1 = __ddiv_rn(__dsub_rn(1., shift), sd);
  double z2 = __ddiv_rn(__dsub_rn(2., shift), sd);
  double z3 = __ddiv_rn(__dsub_rn(3., shift), sd);
  const bool nan_snp = (nn == 0);
  const bool zero_sd = !nan_snp && (sd == 0.);
  if (nan_snp || zero_sd) {
    z1 = z2 = z3 = 0.;  // keep NaN/Inf out of the accumulators; dgrm_finish reports the error
    atomicAdd(&o.bad_count[nan_snp ? 1 : 0], 1);
    const int total = atomicAdd(&o.bad_count[2], 1);
    if (total < o.bad_cap) o.bad_list[total] = nan_snp ? -(g + 1) : g;
  }
  double* z = zt + (int64_t)row * 4;
  z[0] = z1;  // 00
  z[1] = 0.;  // 01 missing
  z[2] = z2;  // 10
  z[3] = z3;  // 11
}

// ------------------------------------------------------------------------------------------------
// K1a+b fused: ONE pass over the caller's rows does the row selection / alignment copy into the chunk buffer
// (bed_gather_kernel) and the per-SNP counts and statistics (bed_count_stats_kernel): one warp per kept row,
// VecT loads chosen by the host from the alignment of (src, src_pitch), popcounts on the lo/hi bit planes with the
// padding bits of the last byte masked, warp-shuffle reduction, lane 0 finishes the statistics.
// 1/4 B read + 1/4 B written per genotype instead of 1/4 + 1/4 (gather) + 1/4 (count).
// ------------------------------------------------------------------------------------------------
template <typename VecT>
__global__ void __launch_bounds__(256)
bed_decode_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, const int32_t* __restrict__ rows, int64_t row0,
                  int nrows, uint8_t* __restrict__ dst, int64_t dst_pitch, int nbytes, int n, int64_t snp0, SnpStatsOut o,
                  double* __restrict__ zt, int zt_row0) {
  constexpr int V = (int)sizeof(VecT);
  constexpr int W = V / 4;  // 32-bit words per vector
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const int64_t srow = rows ? (int64_t)rows[row] : row0 + row;
  const uint8_t* s = src + srow * src_pitch;
  uint8_t* d = dst + (int64_t)row * dst_pitch;
  const int nvec = nbytes / V;
  int c0 = 0, c2 = 0, c3 = 0;
  auto count_word = [&](uint32_t w, int i0) {   // 16 individuals starting at i0
    int valid = n - i0;
    valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
    const uint32_t km = (valid == 16) ? 0x55555555u : (((1u << (2 * valid)) - 1u) & 0x55555555u);
    const uint32_t lo = w & 0x55555555u;
    const uint32_t hi = (w >> 1) & 0x55555555u;
    c3 += __popc(lo & hi & km);   // 11 hom A2
    c2 += __popc(hi & ~lo & km);  // 10 het
    c0 += __popc(lo & ~hi & km);  // 01 missing
  };
#pragma unroll 4
  for (int j = lane; j < nvec; j += 32) {
    const VecT v = __ldg(reinterpret_cast<const VecT*>(s) + j);
    reinterpret_cast<VecT*>(d)[j] = v;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
    for (int t = 0; t < W; ++t) count_word(w[t], (j * W + t) * 16);
  }
  for (int b = nvec * V + lane; b < nbytes; b += 32) {   // the row's last nbytes % V bytes, 4 individuals each
    const uint32_t v = s[b];
    d[b] = (uint8_t)v;
    int valid = n - b * 4;
    valid = valid < 0 ? 0 : (valid > 4 ? 4 : valid);
    const uint32_t km = ((1u << (2 * valid)) - 1u) & 0x55u;
    const uint32_t lo = v & 0x55u, hi = (v >> 1) & 0x55u;
    c3 += __popc(lo & hi & km);
    c2 += __popc(hi & ~lo & km);
    c0 += __popc(lo & ~hi & km);
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, sft);
    c2 += __shfl_xor_sync(0xffffffffu, c2, sft);
    c3 += __shfl_xor_sync(0xffffffffu, c3, sft);
  }
  if (lane == 0) snp_stats_finish(c0, c2, c3, n, snp0 + row, zt_row0 + row, o, zt);
}

// ------------------------------------------------------------------------------------------------
// K1a+b, GRM mode: the same single pass (row selection + per-SNP counts), but the chunk leaves it TRANSPOSED in
// 16-SNP groups:   T[kg][i] = one 32-bit word = the 2-bit patterns of SNPs 16 kg .. 16 kg + 15 of individual i
// (SNP t at bits 2t).  bed_expand, whose threads run along individuals, then needs ONE coalesced 32-bit global load
// per 16 genotypes instead of sixteen byte loads from a staged shared-memory tile, and the word's nibbles index the
// SNP-pair tables directly.  Same 1/4 byte per genotype as the row-major chunk.
// Thread = (16-row group, word column j = individuals 16 j .. 16 j + 15): 16 coalesced 32-bit loads (one per row),
// a 16 x 16 transpose of 2-bit fields in registers (4 butterfly stages), 64 contiguous bytes out.
// Rows arrive in arbitrary batches: a launch covers chunk rows [fill, fill + nrows); a group whose first row is in
// this launch starts from "all missing" (01), otherwise the word written by the previous launch is updated.
// Per-row counts are reduced over the warp, then the block, then added to counts[row][3] with atomics;
// bed_stats_finish_kernel turns them into struct SNP values once the rows are complete.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void transpose16x16_2bit(uint32_t (&w)[16]) {
#define DGRM_TSTAGE(S, MASK)                                   \
  _Pragma("unroll") for (int j = 0; j < 16; ++j)               \
    if ((j & S) == 0) {                                        \
      const uint32_t t = ((w[j] >> (2 * S)) ^ w[j + S]) & MASK; \
      w[j + S] ^= t;                                           \
      w[j] ^= t << (2 * S);                                    \
    }
  DGRM_TSTAGE(8, 0x0000FFFFu)
  DGRM_TSTAGE(4, 0x00FF00FFu)
  DGRM_TSTAGE(2, 0x0F0F0F0Fu)
  DGRM_TSTAGE(1, 0x33333333u)
#undef DGRM_TSTAGE
}

__global__ void __launch_bounds__(256)
bed_decode_t_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, const int32_t* __restrict__ rows, int64_t row0,
                    int fill, int nrows, uint32_t* __restrict__ T, int64_t t_pitch, int nbytes, int n,
                    int32_t* __restrict__ counts /* [chunk rows][4]: missing, -, het, hom A2 */) {
  __shared__ uint32_t s_w[8][16];                   // per warp: packed counts of the group's 16 rows
  __shared__ int64_t s_off[16];                     // byte offset of the group's source rows, -1 = not in this launch
  const int kg = (fill >> 4) + blockIdx.y;          // 16-row group of the chunk
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int nwords = (n + 15) >> 4;
  const int g0 = kg << 4;                           // first chunk row of the group
  const int ra = max(g0, fill), rb = min(g0 + 16, fill + nrows);   // chunk rows of this group delivered by this launch
  if (threadIdx.x < 16) {
    const int r = g0 + threadIdx.x;
    s_off[threadIdx.x] = (r >= ra && r < rb) ? (rows ? (int64_t)rows[r - fill] : row0 + (r - fill)) * src_pitch : -1;
  }
  __syncthreads();
  uint32_t w[16];
  uint32_t cnt[16];                                 // per row: missing | het << 10 | hom A2 << 20 (<= 16 each per lane)
  if (j < nwords) {
    const int valid = min(16, n - 16 * j);
    uint32_t act = 0;                               // rows delivered by this launch
    if (4 * j + 4 <= nbytes) {                      // sixteen independent, coalesced 32-bit loads
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int64_t off = s_off[t];
        act |= (off >= 0 ? 1u : 0u) << t;
        w[t] = off >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(src + off + 4 * j)) : 0x55555555u;
      }
    } else {                                        // the row's last, partial word
#pragma unroll 1
      for (int t = 0; t < 16; ++t) {
        const int64_t off = s_off[t];
        uint32_t v = 0x55555555u;
        if (off >= 0) {
          act |= 1u << t;
          v = 0;
          for (int b = 0; 4 * j + b < nbytes; ++b) v |= (uint32_t)src[off + 4 * j + b] << (8 * b);
        }
        w[t] = v;
      }
    }
    if (valid < 16) {                               // individuals past n read as missing (and so are never counted as data)
      const uint32_t keep = (1u << (2 * valid)) - 1u;
#pragma unroll
      for (int t = 0; t < 16; ++t) w[t] = (w[t] & keep) | (0x55555555u & ~keep);
    }
    const int pad = 16 - valid;                     // those padding fields count as "missing" below: taken out again
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const uint32_t v = w[t];
      const uint32_t lo = v & 0x55555555u, hi = (v >> 1) & 0x55555555u;
      const uint32_t c = (uint32_t)(__popc(lo & ~hi) - pad) | ((uint32_t)__popc(hi & ~lo) << 10) | ((uint32_t)__popc(lo & hi) << 20);
      cnt[t] = ((act >> t) & 1u) ? c : 0u;
    }
    transpose16x16_2bit(w);                         // w[c] = individual 16 j + c, SNP t at bits 2t
    uint32_t* out = T + (int64_t)kg * t_pitch + 16 * j;
    if (ra == g0 && rb == g0 + 16) {                // the whole group in one launch: plain 16-byte stores
#pragma unroll
      for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(out)[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    } else {
      // rows [ra, rb) of the group: merge into what earlier launches wrote (or into "all missing" when this launch
      // delivers the group's first row)
      const uint32_t lo_bits = (ra - g0) == 0 ? 0u : ((1u << (2 * (ra - g0))) - 1u);
      const uint32_t hi_bits = (rb - g0) == 16 ? 0xFFFFFFFFu : ((1u << (2 * (rb - g0))) - 1u);
      const uint32_t mine = hi_bits & ~lo_bits;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const uint32_t old = ra == g0 ? 0x55555555u : out[c];
        out[c] = (old & ~mine) | (w[c] & mine);
      }
    }
  } else {
#pragma unroll
    for (int t = 0; t < 16; ++t) cnt[t] = 0;
  }
  // counts: butterfly over the warp that halves the values a lane carries at every step (16 shuffles instead of 80);
  // lane l ends with the warp total of row l >> 1 (three 10-bit counters per register, <= 16 * 32 each)
#define DGRM_BFLY(H, BIT)                                                        \
  _Pragma("unroll") for (int i = 0; i < H; ++i) {                                \
    const bool up = (lane & BIT) != 0;                                           \
    const uint32_t send = up ? cnt[i] : cnt[i + H];                              \
    const uint32_t mine = up ? cnt[i + H] : cnt[i];                              \
    cnt[i] = mine + __shfl_xor_sync(0xffffffffu, send, BIT);                     \
  }
  DGRM_BFLY(8, 16)
  DGRM_BFLY(4, 8)
  DGRM_BFLY(2, 4)
  DGRM_BFLY(1, 2)
#undef DGRM_BFLY
  const uint32_t tot = cnt[0] + __shfl_xor_sync(0xffffffffu, cnt[0], 1);
  if ((lane & 1) == 0) s_w[threadIdx.x >> 5][lane >> 1] = tot;
  __syncthreads();
  if (threadIdx.x < 48) {
    const int t = threadIdx.x / 3, k = threadIdx.x % 3;
    int v = 0;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) v += (int)((s_w[wp][t] >> (10 * k)) & 0x3FFu);
    if (v) atomicAdd(&counts[(int64_t)(g0 + t) * 4 + (k == 0 ? 0 : k + 1)], v);
  }
}

// struct SNP values of chunk rows [fill, fill + nrows) from their complete counts (reference operation order)
__global__ void bed_stats_finish_kernel(const int32_t* __restrict__ counts, int fill, int nrows, int n, int64_t snp0,
                                        SnpStatsOut o, double* __restrict__ zt) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int32_t* c = counts + (int64_t)(fill + r) * 4;
  snp_stats_finish(c[0], c[2], c[3], n, snp0 + r, fill + r, o, zt);
}

// ------------------------------------------------------------------------------------------------
// K1c: expand one chunk into tensor-core operands.
//   Z  [n_pad][z_pitch] 16-bit, K-major: Z[i][k] = round16(zt[k][pattern(k,i)])         (hi part)
//      with SPLIT also Z[i][lo_off + k] = round16(zt - hi)                                 (lo part)
//   Q  [n_pad][q_pitch] e4m3 0/1, K-major: Q[i][k] = (pattern != 01)                      (missings)
//   Q4 [n_pad][q4_pitch] e2m1 0/1, two SNPs per byte (same nibble order for both MMA operands)   (fp4 count)
//   MB [n_pad][mb_pitch] bit k of row i = (pattern != 01)                                  (popcount path)
//   diag_part[ktile][i] = sum over the tile's SNPs of zt^2 in fp64 (exact diagonal of S)
// Block = 128 SNPs x 128 individuals, 256 threads; lanes run along individuals so every smem access
// is conflict free, the transposed tile is staged in smem and written out in full 128/256-byte rows.
// ------------------------------------------------------------------------------------------------
constexpr int kExpK = 128;          // SNPs per block
constexpr int kExpI = 128;          // individuals per block
constexpr int kExpZPitch = 272;     // bytes per staged Z row (256 + 16 pad)
constexpr int kExpQPitch = 136;     // bytes per staged Q row (128 + 8 pad)
constexpr int kExpQ4Pitch = 68;     // bytes per staged Q4 row (64 + 4 pad)
constexpr int kExpMPitch = 20;      // bytes per staged mask row (16 + 4 pad)

struct ExpandSmem {
  uint8_t pk[kExpK][32];            // packed genotypes: row = SNP, 32 bytes = 128 individuals
  uint16_t lut_hi[kExpK][4];
  uint16_t lut_lo[kExpK][4];
  double lut_sq[kExpK][4];
  double dpart[2][kExpI];
  uint8_t outz[kExpI * kExpZPitch];
  uint8_t outq[kExpI * kExpQPitch]; // count operand staging: e4m3 bytes, e2m1 nibbles or mask bits (one kind per launch)
};

template <bool BF16>
__device__ __forceinline__ uint16_t round16(double v) {
  if constexpr (BF16) {
    return __bfloat16_as_ushort(__double2bfloat16(v));
  } else {
    return __half_as_ushort(__double2half(v));
  }
}
template <bool BF16>
__device__ __forceinline__ double widen16(uint16_t b) {
  if constexpr (BF16) {
    return (double)__bfloat162float(__ushort_as_bfloat16(b));
  } else {
    return (double)__half2float(__ushort_as_half(b));
  }
}

struct ExpandArgs {
  const uint8_t* chunk;
  int64_t pitch;
  int kc;            // valid SNP rows in the chunk
  int n;             // individuals
  const double* zt;  // [kc][4]
  uint16_t* Z;
  int64_t z_pitch;   // elements
  int64_t lo_off;    // element offset of the lo part (SPLIT)
  uint8_t* Q;        // e4m3 0/1 operand (QK 1)
  int64_t q_pitch;
  uint8_t* Q4;       // e2m1 0/1 operand, two SNPs per byte (QK 2)
  int64_t q4_pitch;  // bytes
  uint32_t* MB;      // non-missing bit masks (QK 3)
  int64_t mb_pitch;  // 32-bit words
  double* diag_part; // [ktiles][n_pad]
  int n_pad;
};

// QK: which count operand this launch writes (0 none, 1 e4m3 bytes, 2 e2m1 nibbles, 3 bit masks)
template <bool BF16, bool SPLIT, int QK>
__global__ void __launch_bounds__(256) bed_expand_kernel(const ExpandArgs a) {
  extern __shared__ __align__(16) uint8_t expand_smem_raw[];
  ExpandSmem& sm = *reinterpret_cast<ExpandSmem*>(expand_smem_raw);
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * kExpI;
  const int k0 = blockIdx.y * kExpK;

  // packed tile: 128 rows x 32 bytes, one 16-byte piece per thread; rows past the chunk and individuals past n
  // become "missing" (01) here, so the hot loop below needs no bounds logic at all
  {
    const int r = tid >> 1, h = tid & 1;
    uint4 v = make_uint4(0x55555555u, 0x55555555u, 0x55555555u, 0x55555555u);
    if (k0 + r < a.kc) {
      const int64_t off = (int64_t)(k0 + r) * a.pitch + (i0 >> 2) + h * 16;
      if ((i0 >> 2) + h * 16 < a.pitch) v = __ldg(reinterpret_cast<const uint4*>(a.chunk + off));
    }
    if (i0 + kExpI > a.n) {  // block-uniform: only the last block of individuals
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int valid = a.n - (i0 + (h * 4 + q) * 16);  // individuals of this 32-bit word that exist
        valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
        const uint32_t keep = valid == 16 ? 0xFFFFFFFFu : ((1u << (2 * valid)) - 1u);
        w[q] = (w[q] & keep) | (0x55555555u & ~keep);
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(&sm.pk[r][h * 16]) = v;
  }
  if (tid < kExpK) {
    double z[4] = {0., 0., 0., 0.};
    if (k0 + tid < a.kc) {
      const double* zp = a.zt + (int64_t)(k0 + tid) * 4;
      z[0] = zp[0]; z[2] = zp[2]; z[3] = zp[3];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint16_t hi = round16<BF16>(z[c]);
      sm.lut_hi[tid][c] = hi;
      sm.lut_lo[tid][c] = SPLIT ? round16<BF16>(z[c] - widen16<BF16>(hi)) : (uint16_t)0;
      sm.lut_sq[tid][c] = __dmul_rn(z[c], z[c]);
    }
  }
  __syncthreads();

  const int il = tid & (kExpI - 1);
  const int kh = tid >> 7;
  const int bshift = 2 * (il & 3);
  const uint8_t* pkcol = &sm.pk[0][il >> 2];
  double dsum = 0.;
#pragma unroll 1
  for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
    const uint16_t* lut = part == 0 ? &sm.lut_hi[0][0] : &sm.lut_lo[0][0];
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
      const int kg = kh + 2 * j;
      const uint8_t* pkc = pkcol + kg * 8 * 32;
      const uint16_t* lh = lut + kg * 8 * 4;
      const double* lq = &sm.lut_sq[kg * 8][0];
      uint32_t h[4];
      uint32_t c16 = 0;  // the 8 two-bit patterns of this individual, SNP t at bits 2t
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const uint32_t code = ((uint32_t)pkc[t * 32] >> bshift) & 3u;
        const uint32_t v = lh[t * 4 + code];
        if (t & 1) h[t >> 1] = v * 65536u + h[t >> 1]; else h[t >> 1] = v;
        if (part == 0) {
          c16 = code * (1u << (2 * t)) + c16;
          dsum = __dadd_rn(dsum, lq[t * 4 + code]);
        }
      }
      *reinterpret_cast<uint4*>(&sm.outz[il * kExpZPitch + kg * 16]) = make_uint4(h[0], h[1], h[2], h[3]);
      if (part == 0 && QK != 0) {
        // present <=> pattern != 01 <=> hi bit set or lo bit clear; one flag per SNP at bit 2t
        const uint32_t lo = c16 & 0x5555u, hi = (c16 >> 1) & 0x5555u;
        uint32_t x = (hi | (lo ^ 0x5555u)) & 0x5555u;
        if constexpr (QK == 2) {
          // spread flag t from bit 2t to bit 4t + 1 (e2m1 1.0 = 0b0010 in nibble t)
          x = (x & 0xFFu) | ((x & 0xFF00u) << 8);           // SNPs 4..7 to the upper half-word
          x = (x & 0x000F000Fu) | ((x & 0x00F000F0u) << 4);  // pairs to separate bytes
          x = (x & 0x01010101u) | ((x & 0x04040404u) << 2);  // each flag to the base of its nibble
          *reinterpret_cast<uint32_t*>(&sm.outq[il * kExpQ4Pitch + kg * 4]) = x << 1;
        } else {
          // compress the flags to 8 adjacent bits
          x = (x | (x >> 1)) & 0x3333u;
          x = (x | (x >> 2)) & 0x0F0Fu;
          x = (x | (x >> 4)) & 0x00FFu;
          if constexpr (QK == 3) {
            sm.outq[il * kExpMPitch + kg] = (uint8_t)x;
          } else {  // e4m3 1.0 = 0x38, one byte per SNP
            const uint32_t ql = (((x & 0xFu) * 0x00204081u) & 0x01010101u) * 0x38u;
            const uint32_t qh = (((x >> 4) * 0x00204081u) & 0x01010101u) * 0x38u;
            *reinterpret_cast<uint2*>(&sm.outq[il * kExpQPitch + kg * 8]) = make_uint2(ql, qh);
          }
        }
      }
    }
    if (part == 0) sm.dpart[kh][il] = dsum;
    __syncthreads();
    // write-out: full rows, 16 lanes per row
    {
      const int c = tid & 15;
      uint16_t* zbase = a.Z + (part == 0 ? 0 : a.lo_off) + k0;
#pragma unroll
      for (int ps = 0; ps < kExpI / 16; ++ps) {
        const int r = ps * 16 + (tid >> 4);
        const uint4 v = *reinterpret_cast<const uint4*>(&sm.outz[r * kExpZPitch + c * 16]);
        *reinterpret_cast<uint4*>(zbase + (int64_t)(i0 + r) * a.z_pitch + c * 8) = v;
      }
      if (part == 0) {
        if constexpr (QK == 1) {
#pragma unroll
          for (int ps = 0; ps < kExpI / 16; ++ps) {
            const int r = ps * 16 + (tid >> 4);
            const uint2 v = *reinterpret_cast<const uint2*>(&sm.outq[r * kExpQPitch + c * 8]);
            *reinterpret_cast<uint2*>(a.Q + (int64_t)(i0 + r) * a.q_pitch + k0 + c * 8) = v;
          }
        }
        if constexpr (QK == 2) {
#pragma unroll
          for (int ps = 0; ps < kExpI / 16; ++ps) {
            const int r = ps * 16 + (tid >> 4);
            const uint32_t v = *reinterpret_cast<const uint32_t*>(&sm.outq[r * kExpQ4Pitch + c * 4]);
            *reinterpret_cast<uint32_t*>(a.Q4 + (int64_t)(i0 + r) * a.q4_pitch + (k0 >> 1) + c * 4) = v;
          }
        }
        if constexpr (QK == 3) {
          if (tid < kExpI) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&sm.outq[tid * kExpMPitch]);
            *reinterpret_cast<uint4*>(a.MB + (int64_t)(i0 + tid) * a.mb_pitch + (k0 >> 5)) =
                make_uint4(src[0], src[1], src[2], src[3]);
          }
        }
        if (tid < kExpI)
          a.diag_part[(int64_t)blockIdx.y * a.n_pad + i0 + tid] = __dadd_rn(sm.dpart[0][tid], sm.dpart[1][tid]);
      }
    }
    if (SPLIT) __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// K1c, GRM mode, from the transposed chunk T (bed_decode_t_kernel): same outputs as bed_expand_kernel.
// A thread owns one individual and walks 16-SNP words: one coalesced 32-bit global load per 16 genotypes; each nibble
// of the word (the patterns of two adjacent SNPs) indexes a 16-entry table of the SNP PAIR: two 16-bit operand
// values in one 32-bit shared-memory read, and the sum of the two exact squares in one 64-bit read.  That is 1.5
// shared-memory wavefronts per genotype instead of ~4 (byte + 16-bit + 64-bit lookups per genotype), which was what
// bounded the row-major kernel (61 M wavefronts per 16384-SNP chunk, profiles/r1b_ncu_full_summary.csv).
// Also emits, per individual, the number of missing genotypes of the tile (the bound that lets streamed pair counts
// travel as 16-bit values), which used to be a third pass over the chunk.
// ------------------------------------------------------------------------------------------------
struct ExpandTArgs {
  const uint32_t* T;   // [kc / 16][t_pitch]
  int64_t t_pitch;     // words
  int kc;              // valid SNP rows in the chunk
  int n;               // individuals
  const double* zt;    // [kc][4]
  uint16_t* Z;
  int64_t z_pitch;
  int64_t lo_off;
  uint8_t* Q;
  int64_t q_pitch;
  uint8_t* Q4;
  int64_t q4_pitch;
  uint32_t* MB;
  int64_t mb_pitch;
  double* diag_part;   // [ktiles][n_pad]
  int32_t* miss_part;  // [ktiles][n_pad] or NULL
  int n_pad;
  const int32_t* blocks;  // 128-individual blocks to expand (the individual-block index indexes it), NULL = all of them
  int k_fast;             // 1: blockIdx.x runs along the SNP tiles (concurrent blocks write adjacent pieces of the same rows)
  int stream_stores;      // 1: operand stores bypass L2 retention (st.global.cs)
};

constexpr int kExpTThreads = 512;    // 16 warps per block: the 64 KB of staging then carry 48 warps per SM
constexpr int kExpTKh = kExpTThreads / kExpI;   // threads per individual (each takes every kExpTKh-th 16-SNP word)

template <int QK, bool SPLIT>
struct ExpandTSmem {
  uint16_t lut_hi[kExpK][4];
  uint16_t lut_lo[SPLIT ? kExpK : 1][4];
  alignas(16) double lut_sq[kExpK][4];
  uint32_t pair_hi[kExpK / 2][16];   // operand values of SNPs (2p, 2p+1): low half-word = SNP 2p
  uint32_t pair_lo[SPLIT ? kExpK / 2 : 1][16];
  alignas(16) double pair_sq[kExpK / 2][16];
  double dpart[kExpTKh][kExpI];
  int mpart[kExpTKh][kExpI];
  alignas(16) uint8_t outz[kExpI * kExpZPitch];
  alignas(16) uint8_t outq[kExpI * (QK == 1 ? kExpQPitch : QK == 2 ? kExpQ4Pitch : kExpMPitch)];
};

template <bool BF16, bool SPLIT, int QK>
__global__ void __launch_bounds__(kExpTThreads) bed_expand_t_kernel(const ExpandTArgs a) {
  extern __shared__ __align__(16) uint8_t expand_smem_raw[];
  ExpandTSmem<QK, SPLIT>& sm = *reinterpret_cast<ExpandTSmem<QK, SPLIT>*>(expand_smem_raw);
  const int tid = threadIdx.x;
  const int bi = a.k_fast ? blockIdx.y : blockIdx.x, bk = a.k_fast ? blockIdx.x : blockIdx.y;
  const int i0 = (a.blocks ? a.blocks[bi] : bi) * kExpI;
  const int k0 = bk * kExpK;
  if (tid < kExpK) {
    double z[4] = {0., 0., 0., 0.};
    if (k0 + tid < a.kc) {
      const double* zp = a.zt + (int64_t)(k0 + tid) * 4;
      z[0] = zp[0]; z[2] = zp[2]; z[3] = zp[3];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint16_t hi = round16<BF16>(z[c]);
      sm.lut_hi[tid][c] = hi;
      if (SPLIT) sm.lut_lo[tid][c] = round16<BF16>(z[c] - widen16<BF16>(hi));
      sm.lut_sq[tid][c] = __dmul_rn(z[c], z[c]);
    }
  }
  __syncthreads();
  // pair tables: entry idx = pattern(SNP 2p) + 4 * pattern(SNP 2p + 1); 64 pairs x 16 entries, 2 per thread
#pragma unroll
  for (int q = 0; q < 1024 / kExpTThreads; ++q) {
    const int ent = tid * (1024 / kExpTThreads) + q, p = ent >> 4, idx = ent & 15;
    const int ca = idx & 3, cb = idx >> 2;
    sm.pair_hi[p][idx] = (uint32_t)sm.lut_hi[2 * p][ca] | ((uint32_t)sm.lut_hi[2 * p + 1][cb] << 16);
    if (SPLIT) sm.pair_lo[p][idx] = (uint32_t)sm.lut_lo[2 * p][ca] | ((uint32_t)sm.lut_lo[2 * p + 1][cb] << 16);
    sm.pair_sq[p][idx] = __dadd_rn(sm.lut_sq[2 * p][ca], sm.lut_sq[2 * p + 1][cb]);
  }
  __syncthreads();

  const int il = tid & (kExpI - 1);
  const int kh = tid >> 7;
  const bool real = i0 + il < a.n;
  constexpr int kWords = (kExpK / 16) / kExpTKh;   // 16-SNP words per thread
  // the individual's words of this tile (16 SNPs each); groups past the chunk and individuals past n: all missing
  uint32_t words[kWords];
#pragma unroll
  for (int j = 0; j < kWords; ++j) {
    const int kgl = kh + kExpTKh * j;           // 16-SNP group within the tile
    words[j] = 0x55555555u;
    if (real && k0 + 16 * kgl < a.kc) words[j] = __ldg(a.T + (int64_t)((k0 >> 4) + kgl) * a.t_pitch + i0 + il);
  }
  double dsum = 0.;
  int msum = 0;
#pragma unroll 1
  for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
    const uint32_t* ptab = part == 0 ? &sm.pair_hi[0][0] : &sm.pair_lo[0][0];
#pragma unroll
    for (int j = 0; j < kWords; ++j) {
      const int kgl = kh + kExpTKh * j;
      const uint32_t w = words[j];
      uint32_t h[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const uint32_t idx = (w >> (4 * p)) & 15u;
        h[p] = ptab[(kgl * 8 + p) * 16 + idx];
        if (part == 0) dsum = __dadd_rn(dsum, sm.pair_sq[kgl * 8 + p][idx]);
      }
      uint4* oz = reinterpret_cast<uint4*>(&sm.outz[il * kExpZPitch + kgl * 32]);
      oz[0] = make_uint4(h[0], h[1], h[2], h[3]);
      oz[1] = make_uint4(h[4], h[5], h[6], h[7]);
      if (part == 0) {
        const uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
        msum += __popc(lo & ~hi);
        if (QK != 0) {
          // present <=> pattern != 01; one flag per SNP at bit 2t (16 SNPs)
          const uint32_t x = (hi | (lo ^ 0x55555555u)) & 0x55555555u;
          if constexpr (QK == 2) {
            // spread flag t from bit 2t to bit 4t + 1 (e2m1 1.0 = 0b0010 in nibble t), 8 SNPs per 32-bit word
            uint32_t out2[2];
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              uint32_t y = (x >> (16 * hlf)) & 0xFFFFu;
              y = (y & 0xFFu) | ((y & 0xFF00u) << 8);
              y = (y & 0x000F000Fu) | ((y & 0x00F000F0u) << 4);
              y = (y & 0x01010101u) | ((y & 0x04040404u) << 2);
              out2[hlf] = y << 1;
            }
            uint32_t* oq = reinterpret_cast<uint32_t*>(&sm.outq[il * kExpQ4Pitch + kgl * 8]);   // 68-byte pitch: 4-byte aligned
            oq[0] = out2[0];
            oq[1] = out2[1];
          } else {
            uint32_t y = x;                      // compress the 16 flags to adjacent bits
            y = (y | (y >> 1)) & 0x33333333u;
            y = (y | (y >> 2)) & 0x0F0F0F0Fu;
            y = (y | (y >> 4)) & 0x00FF00FFu;
            y = (y | (y >> 8)) & 0x0000FFFFu;
            if constexpr (QK == 3) {
              *reinterpret_cast<uint16_t*>(&sm.outq[il * kExpMPitch + kgl * 2]) = (uint16_t)y;
            } else {                             // e4m3 1.0 = 0x38, one byte per SNP
              uint32_t qb[4];
#pragma unroll
              for (int v = 0; v < 4; ++v) qb[v] = ((((y >> (4 * v)) & 0xFu) * 0x00204081u) & 0x01010101u) * 0x38u;
              uint2* oq = reinterpret_cast<uint2*>(&sm.outq[il * kExpQPitch + kgl * 16]);         // 136-byte pitch: 8-byte aligned
              oq[0] = make_uint2(qb[0], qb[1]);
              oq[1] = make_uint2(qb[2], qb[3]);
            }
          }
        }
      }
    }
    if (part == 0) { sm.dpart[kh][il] = dsum; sm.mpart[kh][il] = msum; }
    __syncthreads();
    // write-out: full rows, 16 lanes per row
    {
      const int c = tid & 15;
      uint16_t* zbase = a.Z + (part == 0 ? 0 : a.lo_off) + k0;
#pragma unroll
      for (int ps = 0; ps < kExpI / (kExpTThreads / 16); ++ps) {
        const int r = ps * (kExpTThreads / 16) + (tid >> 4);
        const uint4 v = *reinterpret_cast<const uint4*>(&sm.outz[r * kExpZPitch + c * 16]);
        uint4* zdst = reinterpret_cast<uint4*>(zbase + (int64_t)(i0 + r) * a.z_pitch + c * 8);
        if (a.stream_stores) __stcs(zdst, v); else *zdst = v;
      }
      if (part == 0) {
        if constexpr (QK == 1) {
#pragma unroll
          for (int ps = 0; ps < kExpI / (kExpTThreads / 16); ++ps) {
            const int r = ps * (kExpTThreads / 16) + (tid >> 4);
            const uint2 v = *reinterpret_cast<const uint2*>(&sm.outq[r * kExpQPitch + c * 8]);
            *reinterpret_cast<uint2*>(a.Q + (int64_t)(i0 + r) * a.q_pitch + k0 + c * 8) = v;
          }
        }
        if constexpr (QK == 2) {
#pragma unroll
          for (int ps = 0; ps < kExpI / (kExpTThreads / 16); ++ps) {
            const int r = ps * (kExpTThreads / 16) + (tid >> 4);
            const uint32_t v = *reinterpret_cast<const uint32_t*>(&sm.outq[r * kExpQ4Pitch + c * 4]);
            *reinterpret_cast<uint32_t*>(a.Q4 + (int64_t)(i0 + r) * a.q4_pitch + (k0 >> 1) + c * 4) = v;
          }
        }
        if constexpr (QK == 3) {
          if (tid < kExpI) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&sm.outq[tid * kExpMPitch]);
            *reinterpret_cast<uint4*>(a.MB + (int64_t)(i0 + tid) * a.mb_pitch + (k0 >> 5)) =
                make_uint4(src[0], src[1], src[2], src[3]);
          }
        }
        if (tid < kExpI) {
          double ds = sm.dpart[0][tid];
          int ms = sm.mpart[0][tid];
#pragma unroll
          for (int h2 = 1; h2 < kExpTKh; ++h2) { ds = __dadd_rn(ds, sm.dpart[h2][tid]); ms += sm.mpart[h2][tid]; }
          a.diag_part[(int64_t)bk * a.n_pad + i0 + tid] = ds;
          if (a.miss_part)   // padding individuals and SNP rows past the chunk were read as "missing": not real ones
            a.miss_part[(int64_t)bk * a.n_pad + i0 + tid] =
                (i0 + tid < a.n) ? ms - max(0, min(kExpK, k0 + kExpK - a.kc)) : 0;
        }
      }
    }
    if (SPLIT) __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// LD mode (GWAS::getLessSignificantCorrelatedSNPs, gwas.cpp:1163-1165: genotypes * genotypes^T, missings * missings^T):
// the operands are SNP-major, which is the .bed's own order — row k of the chunk becomes operand row k with the
// individuals along K, no transposition:
//   Z [row][i] = round16(zt[row][pattern(row, i)]),  Q/Q4/MB [row][i] = (pattern != 01)
//   diag[row]  = f1 z1^2 + f2 z2^2 + f3 z3^2   (sum over the individuals of z^2 in fp64, from the exact counts)
// One thread expands 8 individuals (two packed bytes): a 16-byte Z store, 4 bytes of e2m1 nibbles.
// ------------------------------------------------------------------------------------------------
struct LdExpandArgs {
  const uint8_t* chunk;
  int64_t pitch;
  int kc;              // SNP rows in the chunk
  int n;               // individuals per row
  const double* zt;    // [kc][4]
  const int32_t* freq; // [kc][4] counts of the chunk's SNPs (missing, hom A1, het, hom A2)
  uint16_t* Z;         // operand row of the chunk's first SNP
  int64_t z_pitch;     // elements
  int64_t lo_off;      // element offset of the lo part (SPLIT)
  uint8_t* Q;
  int64_t q_pitch;
  uint8_t* Q4;
  int64_t q4_pitch;
  uint32_t* MB;
  int64_t mb_pitch;    // 32-bit words
  double* diag;        // [kc]
};

template <bool BF16, bool SPLIT>
__global__ void __launch_bounds__(256) ld_expand_kernel(const LdExpandArgs a) {
  const int row = blockIdx.y;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;   // group of 8 individuals
  const int ngroups = (a.n + 7) >> 3;
  const double* zp = a.zt + (int64_t)row * 4;
  const double z0 = zp[0], z2 = zp[2], z3 = zp[3];
  if (g == 0) {
    const int32_t* f = a.freq + (int64_t)row * 4;
    a.diag[row] = __dadd_rn(__dadd_rn(__dmul_rn((double)f[1], __dmul_rn(z0, z0)), __dmul_rn((double)f[2], __dmul_rn(z2, z2))),
                            __dmul_rn((double)f[3], __dmul_rn(z3, z3)));
  }
  if (g >= ngroups) return;
  uint32_t hi[4], lo[4];
  {
    const double zz[4] = {z0, 0., z2, z3};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      hi[c] = round16<BF16>(zz[c]);
      lo[c] = SPLIT ? round16<BF16>(zz[c] - widen16<BF16>((uint16_t)hi[c])) : 0u;
    }
  }
  const uint8_t* src = a.chunk + (int64_t)row * a.pitch + 2 * g;
  uint32_t w = src[0];
  if (8 * g + 4 < a.n) w |= (uint32_t)src[1] << 8;
  const int valid = min(8, a.n - 8 * g);
  if (valid < 8) w = (w & ((1u << (2 * valid)) - 1u)) | (0x5555u & ~((1u << (2 * valid)) - 1u));   // past n: missing
  uint32_t oh[4], ol[4], q4 = 0, q8lo = 0, q8hi = 0, bits = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const uint32_t code = (w >> (2 * t)) & 3u;
    const uint32_t vh = hi[code], vl = lo[code];
    if (t & 1) { oh[t >> 1] |= vh << 16; ol[t >> 1] |= vl << 16; } else { oh[t >> 1] = vh; ol[t >> 1] = vl; }
    const uint32_t present = code != 1u;
    q4 |= present << (4 * t + 1);          // e2m1 1.0 = 0b0010 in nibble t
    if (t < 4) q8lo |= (present * 0x38u) << (8 * t); else q8hi |= (present * 0x38u) << (8 * (t - 4));   // e4m3 1.0
    bits |= present << t;
  }
  // padding individuals (i >= n) inside the last group must not contribute: their pattern was forced to 01
  uint16_t* zrow = a.Z + (int64_t)row * a.z_pitch + 8 * g;
  *reinterpret_cast<uint4*>(zrow) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
  if (SPLIT) *reinterpret_cast<uint4*>(zrow + a.lo_off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  if (a.Q4) *reinterpret_cast<uint32_t*>(a.Q4 + (int64_t)row * a.q4_pitch + 4 * g) = q4;
  if (a.Q) *reinterpret_cast<uint2*>(a.Q + (int64_t)row * a.q_pitch + 8 * g) = make_uint2(q8lo, q8hi);
  if (a.MB) reinterpret_cast<uint8_t*>(a.MB + (int64_t)row * a.mb_pitch)[g] = (uint8_t)bits;
}

// diag[i] += sum_kt diag_part[kt][i], fixed order (deterministic).
// Block = 32 individuals x 8 warps; warp w sums the k-tiles w, w + 8, ... (coalesced 256-byte reads), the eight
// partial sums are then combined in warp order: the summation order is fixed by the launch geometry alone.
__global__ void __launch_bounds__(256)
diag_reduce_kernel(const double* __restrict__ part, int nkt, int n_pad, double* __restrict__ diag,
                   const int32_t* __restrict__ miss_part = nullptr, int32_t* __restrict__ miss = nullptr) {
  __shared__ double s_d[8][32];
  __shared__ int s_m[8][32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  double s = 0.;
  int mc = 0;
  if (i < n_pad) {
    for (int kt = wp; kt < nkt; kt += 8) {
      s = __dadd_rn(s, part[(int64_t)kt * n_pad + i]);
      if (miss_part) mc += miss_part[(int64_t)kt * n_pad + i];
    }
  }
  s_d[wp][lane] = s;
  s_m[wp][lane] = mc;
  __syncthreads();
  if (wp == 0 && i < n_pad) {
    double t = diag[i];
    int m = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { t = __dadd_rn(t, s_d[q][lane]); m += s_m[q][lane]; }
    diag[i] = t;
    if (miss_part) miss[i] += m;   // per-individual missing genotypes of the chunk (exact integers)
  }
}

}  // namespace dgrm
