// Real-valued genotypes (dosages): the ingest below the reference's BGEN reader.
//
// The reference turns every BGEN variant into one row of expected allele counts, d_i = P_i(AB) + 2 P_i(BB), plus a
// 0/1 "present" flag per sample (ProbabilitiesSetter::set_value, genotypebgen.cpp:183-236), and standardises the row
// on the spot (ProbabilitiesSetter::finalise, genotypebgen.cpp:239-282):
//     mean = sum over present d_i / #present
//     std  = sqrt( sum over present (d_i - mean)^2 / nSamples )          (all samples in the divisor, not #present)
//     z_i  = present ? (d_i - mean) / std : 0          (here: (d_i - mean) * (1 / std), see dosage_expand_kernel)
// and records p2 = mean / 2, p1 = 1 - mean / 2, standardDev = std, nNonMissing = -1 in the SNP
// (Genotype::parseBGENSNP, genotypebgen.cpp:723-736).  SNPs whose std is 0 are then dropped from the genotype
// matrices (Genotype::removeSNPsWithZeroVariance, genotypebgen.cpp:739-779): here they contribute nothing to S or N.
// From there the GRM is the same Z'Z / M'M as for a .bed (Kernel::Kernel, kernel.cpp:93-100).
//
// Kernels: the missing mask folded into the rows as NaN, per-row statistics, then the same 128 x 128 transposing
// expansion into tensor-core operands that the
// 2-bit path ends in (bed_expand_kernel), reading 8-byte dosages instead of 2-bit patterns.  HBM-bound: 8 bytes in per
// genotype and pass, 2-5 bytes out.
#pragma once
#include "bed_kernels.cuh"

namespace dgrm {

struct DosageStatsOut {
  int32_t* freq;     // [m][4], zeroed for dosage SNPs (genotype counts do not exist)
  int32_t* nn;       // nNonMissing = -1 (genotypebgen.cpp:733)
  double *p1, *p2, *sd;
  int32_t* removed;  // number of zero-variance dosage SNPs of this GRM
};

// BGEN probability blocks, as they come out of the container after inflation, unpacked on the device into the same
// dosage rows (NaN = missing) that dgrm_push_dosage_host stages: 3 bytes per genotype cross the bus instead of 9, and
// the host does no per-sample work.  One thread per (kept sample, variant).
//   layout 2 block: N u32 | K u16 | Pmin u8 | Pmax u8 | N ploidy bytes (bit 7 = missing) | phased u8 | B u8 |
//                   2N probabilities of B bits, little-endian bit order: P(AA), P(AB) numerators over 2^B - 1
//   layout 1 block: N x (P(AA), P(AB), P(BB)) u16 numerators over 32768; all three zero = missing
// Arithmetic as the host readers do it (dissect_b200/bgenio.py, host.cpp streamBGEN), so the rows are bit-identical:
//   aa = v0 / scale, ab = v1 / scale, bb = 1 - (aa + ab) (layout 2) or v2 / scale (layout 1), d = ab + 2 bb
// which is what genfile hands to ProbabilitiesSetter::set_value and what that forms (genotypebgen.cpp:199-205).
struct BgenUnpackArgs {
  const uint8_t* raw;       // the piece's blocks, back to back
  const int64_t* offsets;   // [rows of the piece + 1] byte offsets into raw
  int row0, rows;           // rows [row0, row0 + rows) of the piece are handled by this launch
  int layout;
  int n_samples;            // samples in the file
  const int32_t* cols;      // [n] file index of every kept sample, NULL = all samples
  int n;                    // kept samples
  double* dos;              // [rows][n]
  int32_t* errors;          // samples whose ploidy is not 2 (genotypebgen.cpp:120-123)
};

constexpr int kUnpackRows = 16;   // variants per block: the per-block set-up (table of quotients) is shared by them

__global__ void __launch_bounds__(256) bgen_unpack_kernel(const BgenUnpackArgs a) {
  __shared__ double lut[1024];   // numerator / (2^B - 1) for B <= 10: two table reads instead of two fp64 divisions
  const int c = blockIdx.x * 256 + threadIdx.x;
  const bool active = c < a.n;
  const int i = active ? (a.cols ? a.cols[c] : c) : 0;
  const int N = a.n_samples;
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  int lut_bits = -1;             // depth the table was built for (block-uniform)
  for (int rr = 0; rr < kUnpackRows; ++rr) {
    const int r = blockIdx.y * kUnpackRows + rr;
    if (r >= a.rows) break;
    const int64_t o0 = a.offsets[a.row0 + r], o1 = a.offsets[a.row0 + r + 1];
    const uint8_t* base = a.raw + o0;
    double d;
    if (a.layout == 1) {
      const uint8_t* p = base + (int64_t)i * 6;
      const double aa = __ddiv_rn((double)(p[0] | (p[1] << 8)), 32768.), ab = __ddiv_rn((double)(p[2] | (p[3] << 8)), 32768.),
                   bb = __ddiv_rn((double)(p[4] | (p[5] << 8)), 32768.);
      const bool missing = __dadd_rn(__dadd_rn(aa, ab), bb) == 0.;            // genotypebgen.cpp:208-214
      d = missing ? nan : __dadd_rn(ab, __dmul_rn(2., bb));
    } else {
      const int bits = base[9 + N];
      const uint64_t mask = bits == 32 ? 0xFFFFFFFFull : ((1ull << bits) - 1ull);
      const double scale = (double)mask;
      const bool use_lut = bits <= 10;
      if (use_lut && bits != lut_bits) {   // block-uniform: bits belongs to the variant
        __syncthreads();
        for (int k = threadIdx.x; k <= (int)mask; k += 256) lut[k] = __ddiv_rn((double)k, scale);
        __syncthreads();
        lut_bits = bits;
      }
      const uint32_t ploidy = base[8 + i];
      const uint8_t* prob = base + 10 + N;
      const int64_t avail = o1 - o0 - 10 - N;
      uint64_t v[2];
      if (bits == 8) {          // the common depths: whole bytes
        v[0] = prob[2 * (int64_t)i];
        v[1] = prob[2 * (int64_t)i + 1];
      } else if (bits == 16) {
        const uint8_t* q = prob + 4 * (int64_t)i;
        v[0] = q[0] | (q[1] << 8);
        v[1] = q[2] | (q[3] << 8);
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t bit = ((int64_t)2 * i + h) * bits;
          const int64_t byte = bit >> 3;
          uint64_t w = 0;
#pragma unroll
          for (int b = 0; b < 5; ++b)   // (bit & 7) + bits <= 39 bits
            if (byte + b < avail) w |= (uint64_t)prob[byte + b] << (8 * b);
          v[h] = (w >> (bit & 7)) & mask;
        }
      }
      const double aa = use_lut ? lut[v[0]] : __ddiv_rn((double)v[0], scale);
      const double ab = use_lut ? lut[v[1]] : __ddiv_rn((double)v[1], scale);
      const double bb = __dsub_rn(1., __dadd_rn(aa, ab));
      if (active && (ploidy & 0x3Fu) != 2u) atomicAdd(a.errors, 1);
      d = (ploidy & 0x80u) ? nan : __dadd_rn(ab, __dmul_rn(2., bb));          // genotypebgen.cpp:229-236
    }
    if (active) a.dos[(int64_t)r * a.n + c] = d;
  }
}

// Missing genotypes travel inside the staged rows as NaN from here on: one pass folds the caller's byte mask into the
// dosages, so that the statistics and the expansion read 8 bytes per genotype and nothing else.  (A NaN that the
// caller put into a dosage row itself therefore also means "missing".)
__global__ void __launch_bounds__(256)
dosage_fold_mask_kernel(double* __restrict__ dos, const uint8_t* __restrict__ miss, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    if (miss[i]) dos[i] = __longlong_as_double(0x7FF8000000000000ll);
}

// Statistics of kStatRows SNP rows per block, one thread per row, in the reference's own summation order: the mean is
// the running sum over the present samples in sample order (set_value :206, finalise :257), the variance the running
// sum of (d - mean)^2 in sample order (finalise :260-268) -- so mean, std, p1, p2 are bit-identical to the reference's
// loops, and so is its `standardDev != 0.` decision about which SNPs are dropped (a constant row whose running sum
// rounds leaves a tiny non-zero std there, and here).  The chains of different rows are independent, which is all the
// parallelism this needs: a launch should therefore carry as many rows as possible (the device staging buffer holds up
// to a whole chunk).  The block streams 64-sample tiles of its rows into shared memory with cp.async (8-byte pieces:
// rows are only 8-byte aligned), one tile ahead of the row threads.
constexpr int kStatRows = 32;
constexpr int kStatThreads = 128;
constexpr int kStatTile = 64;    // samples per tile: 33 KB of staging per block, so the row chains of a whole chunk are resident at once

struct DosageStatSmem {
  double v[2][kStatRows][kStatTile + 1];
};

__global__ void __launch_bounds__(kStatThreads)
dosage_stats_kernel(const double* __restrict__ dos, int64_t pitch, int rows, int n, int64_t snp0, DosageStatsOut o,
                    double* __restrict__ mean_out, double* __restrict__ sd_out) {
  extern __shared__ __align__(16) uint8_t dosage_stat_smem_raw[];
  DosageStatSmem& sm = *reinterpret_cast<DosageStatSmem*>(dosage_stat_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * kStatRows;
  constexpr int kPerWarp = kStatRows / (kStatThreads / 32);   // rows each warp fetches per tile
  const int ntiles = (n + kStatTile - 1) / kStatTile;
  auto fetch_tile = [&](int t, int buf) {
#pragma unroll
    for (int k = 0; k < kPerWarp; ++k) {
      const int rl = warp * kPerWarp + k;
      const int r = min(r0 + rl, rows - 1);
#pragma unroll
      for (int c = 0; c < kStatTile / 32; ++c) {
        const int i = t * kStatTile + c * 32 + lane;
        if (i < n) {
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm.v[buf][rl][c * 32 + lane]);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(dos + (int64_t)r * pitch + i) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double mean = 0.;
  int present_total = 0;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    double acc = 0.;
    int cnt = 0;
    fetch_tile(0, 0);
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      if (t + 1 < ntiles) {
        fetch_tile(t + 1, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      if (tid < kStatRows) {
        const double* v = sm.v[buf][tid];
        const int jn = min(kStatTile, n - t * kStatTile);
        if (pass == 0) {
#pragma unroll 8
          for (int j = 0; j < jn; ++j) {
            const double d = v[j];
            if (d == d) { acc = __dadd_rn(acc, d); ++cnt; }
          }
        } else {
#pragma unroll 8
          for (int j = 0; j < jn; ++j) {
            const double d = v[j];
            if (d == d) { const double x = __dsub_rn(d, mean); acc = __dadd_rn(acc, __dmul_rn(x, x)); }
          }
        }
      }
      __syncthreads();   // the tile may be overwritten by the fetch of the next iteration
    }
    if (pass == 0) {
      present_total = cnt;
      mean = cnt > 0 ? __ddiv_rn(acc, (double)cnt) : 0.;
    } else if (tid < kStatRows && r0 + tid < rows) {
      const double sd = present_total > 0 ? __dsqrt_rn(__ddiv_rn(acc, (double)n)) : 0.;
      const int r = r0 + tid;
      mean_out[r] = mean;
      sd_out[r] = sd;
      const int64_t g = snp0 + r;
      o.freq[g * 4 + 0] = 0; o.freq[g * 4 + 1] = 0; o.freq[g * 4 + 2] = 0; o.freq[g * 4 + 3] = 0;
      o.nn[g] = -1;
      o.p2[g] = __ddiv_rn(mean, 2.);
      o.p1[g] = __dsub_rn(1., __ddiv_rn(mean, 2.));
      o.sd[g] = sd;
      if (!(sd != 0.)) atomicAdd(o.removed, 1);
    }
  }
}

struct DosageExpandArgs {
  const double* dos;    // [kc][pitch], NaN = missing (dosage_fold_mask_kernel)
  int64_t pitch;
  const double* mean;   // [kc]
  const double* sd;     // [kc]
  int kc;               // valid rows
  int n;                // individuals
  uint16_t* Z;          // operand pointers already offset to the first column of these rows
  int64_t z_pitch;
  int64_t lo_off;
  uint8_t* Q;
  int64_t q_pitch;
  uint8_t* Q4;
  int64_t q4_pitch;
  uint32_t* MB;
  int64_t mb_pitch;
  double* diag_part;    // already offset to the first 128-SNP tile of these rows
  int n_pad;
};

struct DosageSmem {
  double mean[kExpK];
  double inv[kExpK];    // 1 / std; 0 = the row contributes nothing (past the end of the rows, or no variance)
  double dpart[2][kExpI];
  alignas(16) uint8_t outz[kExpI * kExpZPitch];
  alignas(16) uint8_t outq[kExpI * kExpQPitch];
};

// Block = 128 SNPs x 128 individuals, 256 threads; lanes run along individuals, so the 8-byte loads of a warp are one
// contiguous 256-byte piece of a dosage row; the transposed tile is staged in shared memory and leaves in full rows.
// The 8 loads of a thread's 8-SNP group are unconditional (indices clamped into the buffer) so that they are all in
// flight before the first use; z = (d - mean) * (1 / std): one correctly rounded
// reciprocal per SNP instead of the reference's per-genotype division, at most 1 ulp of fp64 apart before the value is
// rounded to the 16-bit operand.
template <bool BF16, bool SPLIT, int QK>
__global__ void __launch_bounds__(256, 3) dosage_expand_kernel(const DosageExpandArgs a) {
  extern __shared__ __align__(16) uint8_t dosage_smem_raw[];
  DosageSmem& sm = *reinterpret_cast<DosageSmem*>(dosage_smem_raw);
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * kExpI;
  const int k0 = blockIdx.y * kExpK;
  if (tid < kExpK) {
    const bool live = k0 + tid < a.kc;
    const double sd = live ? a.sd[k0 + tid] : 0.;
    sm.mean[tid] = live ? a.mean[k0 + tid] : 0.;
    sm.inv[tid] = sd > 0. ? __ddiv_rn(1., sd) : 0.;
  }
  __syncthreads();
  const int il = tid & (kExpI - 1);
  const int kh = tid >> 7;
  const bool ind_ok = i0 + il < a.n;
  const int i = ind_ok ? i0 + il : a.n - 1;
  double dsum = 0.;
#pragma unroll 1
  for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      const int kg = kh + 2 * j;
      double d[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int row = min(k0 + kg * 8 + t, a.kc - 1);
        d[t] = __ldg(a.dos + (int64_t)row * a.pitch + i);
      }
      uint32_t h[4];
      uint32_t flags = 0;  // bit t = genotype of SNP kg * 8 + t is present
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int kl = kg * 8 + t;
        const double inv = sm.inv[kl];
        const bool present = ind_ok && inv != 0. && d[t] == d[t];
        const double z = present ? __dmul_rn(__dsub_rn(d[t], sm.mean[kl]), inv) : 0.;
        const uint16_t hi = round16<BF16>(z);
        const uint32_t v = part == 0 ? hi : round16<BF16>(z - widen16<BF16>(hi));
        if (t & 1) h[t >> 1] = v * 65536u + h[t >> 1]; else h[t >> 1] = v;
        if (part == 0) {
          flags |= (present ? 1u : 0u) << t;
          dsum = __dadd_rn(dsum, __dmul_rn(z, z));
        }
      }
      *reinterpret_cast<uint4*>(&sm.outz[il * kExpZPitch + kg * 16]) = make_uint4(h[0], h[1], h[2], h[3]);
      if (part == 0 && QK != 0) {
        if constexpr (QK == 2) {
          // flag t to bit 4t + 1: e2m1 1.0 = 0b0010 in nibble t
          uint32_t x = (flags & 0xFu) | ((flags & 0xF0u) << 12);  // SNPs 4..7 to the upper half-word
          x = (x & 0x00030003u) | ((x & 0x000C000Cu) << 6);        // pairs to separate bytes
          x = (x & 0x01010101u) | ((x & 0x02020202u) << 3);        // each flag to the base of its nibble
          *reinterpret_cast<uint32_t*>(&sm.outq[il * kExpQ4Pitch + kg * 4]) = x << 1;
        } else if constexpr (QK == 3) {
          sm.outq[il * kExpMPitch + kg] = (uint8_t)flags;
        } else {  // e4m3 1.0 = 0x38, one byte per SNP
          const uint32_t ql = (((flags & 0xFu) * 0x00204081u) & 0x01010101u) * 0x38u;
          const uint32_t qh = (((flags >> 4) * 0x00204081u) & 0x01010101u) * 0x38u;
          *reinterpret_cast<uint2*>(&sm.outq[il * kExpQPitch + kg * 8]) = make_uint2(ql, qh);
        }
      }
    }
    if (part == 0) sm.dpart[kh][il] = dsum;
    __syncthreads();
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

}  // namespace dgrm
