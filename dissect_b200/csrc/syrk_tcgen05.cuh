// grm_syrk — lower-triangle symmetric rank-k update on the 5th-gen tensor cores.
//
// Replaces pdsyrk_('L','T', n, k=m, 1.0, G, 0.0, C) (matrix.cpp:2677-2683), which the reference
// reaches twice from Kernel::Kernel (kernel.cpp:94 for S = Z^T Z, kernel.cpp:100 for N = M^T M).
//
// One launch contracts every resident SNP chunk for a range of output tiles, tile-major and chunk-inner:
// a CTA (pair) keeps its tile while it walks the chunks, accumulating each chunk in fp32 in TMEM and
// folding it into the fp64 (int32) tile, which therefore stays in L2 between chunks.  The operand is
// the K-major matrix X[n_pad][chunks * Kc] written by bed_expand (row = individual, 128-byte-swizzle
// friendly row pitch).  Output tiles are 128 rows x 256 columns, stored column-major inside a tile
// (what DISSECT's Matrix uses locally, matrix.cpp:1748-1758), only for column panels that intersect
// the lower triangle.
//
// Structure (persistent, warp specialised):
//   warp 0   TMA producer  : cp.async.bulk.tensor 2-D boxes of 128 rows x 128 B -> swizzled smem ring
//   warp 1   MMA issuer    : tcgen05.mma, fp32 (exact for the 0/1 count) accumulators in TMEM,
//                            two accumulator stages of 256 columns (all 512 TMEM columns)
//   warp 2   TMEM allocator
//   warps 4-7 epilogue     : tcgen05.ld -> registers -> read-modify-write of the fp64 / int32 tile
//                            (cross-chunk accumulation), overlapped with the next tile's mainloop
// CG = 2 pairs two SMs on one 256x256 tile (cta_group::2, UMMA M = 256): each CTA stages its own
// 128 A rows and half of the B panel, the leader issues the MMAs and multicasts the commits.
#pragma once
#include "ptx_sm100.cuh"

namespace dgrm {

constexpr int kMaxChunks = 32;  // SNP chunks resident at once (one fp32 TMEM accumulation each)

struct SyrkParams {
  const int4* work;   // {first A row, first B row, storage index (rank 0), storage index (rank 1)}
  int work_begin;     // this launch covers work[work_begin, work_end)
  int work_end;
  int nchunks;                  // resident chunks contracted per tile, in order (tile-major, chunk-inner)
  int chunk_col[kMaxChunks];    // element column of chunk j in the operand
  int chunk_kpad[kMaxChunks];   // SNP columns of chunk j, padded to 128
  int npass;          // 1, or 3 for the hi/lo split
  int a_off[3];       // per-pass element column offset of the A operand
  int b_off[3];       // per-pass element column offset of the B operand
  void* out;          // tile storage, [storage index][256 cols][128 rows]
  int accumulate;     // chunk 0 of this launch: 0 = overwrite, 1 = add to what is there (later chunks always add)
  uint32_t idesc;     // tcgen05 instruction descriptor
  int l2_policy;      // 0: no preference; 1: output tiles evict_last when revisited in-launch; 2: 1 + operands evict_first;
                      // 3: 1 + A panels (reused by the next waves of the raster row group) evict_last
};

constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kTileElems = kTileM * kTileN;

template <int CG>
struct SyrkCfg {
#ifndef DGRM_CG2_STAGES
#define DGRM_CG2_STAGES 4
#endif
  static constexpr int kStages = (CG == 1) ? 4 : DGRM_CG2_STAGES;
  static constexpr int kABytes = 128 * 128;                    // 128 rows x 128 B
  static constexpr int kBRows = (CG == 1) ? 256 : 128;         // CG 2: each CTA stages half of B
  static constexpr int kBBytes = kBRows * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + alignment slack
  static constexpr int kThreads = 256;
};

template <typename OutT>
__device__ __forceinline__ OutT acc_convert(uint32_t bits);
template <>
__device__ __forceinline__ double acc_convert<double>(uint32_t bits) {
  return (double)__uint_as_float(bits);
}
template <>
__device__ __forceinline__ int acc_convert<int>(uint32_t bits) {
  return __float2int_rn(__uint_as_float(bits));
}

// KIND: 0 = kind::f16 (fp16/bf16 operands), 1 = kind::f8f6f4 (e4m3 0/1 operands), 2 = kind::mxf4 block-scaled (e2m1 0/1
// operands, all scale factors 1.0); kinds 1 and 2 address the operand in bytes (128 per k-block).
template <int CG, int KIND, typename OutT>
__global__ void __launch_bounds__(256, 1)
grm_syrk_kernel(const __grid_constant__ CUtensorMap tmap, const SyrkParams p) {
  using Cfg = SyrkCfg<CG>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kKElems = KIND == 0 ? 64 : 128;  // operand-map elements per 128-byte k-block
  constexpr int kAccStages = KIND == 2 ? 1 : 2;   // mxf4 keeps TMEM columns for its scale factors: one accumulator
  constexpr uint32_t kSfaCol = 256, kSfbCol = 320;  // 64 columns each, every byte = UE8M0 1.0

  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms are 1024 B: align the ring by hand (same offset in both CTAs of a pair).
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const int unit = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int nunits = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;

  if constexpr (CG == 2) ptx::cluster_sync();

  if (warp == 0 && lane == 0) ptx::prefetch_tensormap(&tmap);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], CG);   // CG 2: leader's expect_tx arrive + peer's remote arrive
      ptx::mbar_init(&empty_bar[s], 1);   // one tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);          // one tcgen05.commit
      ptx::mbar_init(&tempty_bar[s], CG * 128);  // every epilogue thread of the pair
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<CG>(tmem_slot, 512);
  ptx::tcgen05_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if constexpr (KIND == 2) {
    // scale factors: 1.0 (0x7F) in every byte of the SFA / SFB column ranges, whatever sub-layout the MMA reads
    if (warp >= 4) {
      const uint32_t lane_base = tmem_base + ((uint32_t)((warp - 4) * 32) << 16);
#pragma unroll
      for (int c = 0; c < 128; c += 32) ptx::tmem_st_32x32_splat(lane_base + kSfaCol + c, 0x7F7F7F7Fu);
      ptx::tmem_st_wait();
    }
    ptx::tcgen05_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
    ptx::tcgen05_fence_after();
  }

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // operand panels are shared by the CTAs of a wave and then dead: let them go first so that the fp64
      // tiles (evict_last below) survive in L2 from one chunk to the next
      const uint64_t pol = p.l2_policy == 2 ? ptx::policy_evict_first() : ptx::policy_evict_normal();
      const uint64_t pol_a = p.l2_policy == 3 ? ptx::policy_evict_last() : pol;
      for (int w = p.work_begin + unit; w < p.work_end; w += nunits) {
        const int4 t = p.work[w];
        const int a_row = t.x + (CG == 2 ? (int)rank * 128 : 0);
        const int b_row = t.y + (CG == 2 ? (int)rank * 128 : 0);
        for (int j = 0; j < p.nchunks; ++j) {
        const int num_kblocks = p.chunk_kpad[j] / kKElems;
        for (int pass = 0; pass < p.npass; ++pass) {
          const int a_col0 = p.a_off[pass] + p.chunk_col[j];
          const int b_col0 = p.b_off[pass] + p.chunk_col[j];
          for (int kb = 0; kb < num_kblocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            if constexpr (CG == 1) {
              ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              ptx::tma_load_2d(sa, &tmap, &full_bar[stage], a_col0 + kb * kKElems, a_row, pol_a);
              ptx::tma_load_2d(sb, &tmap, &full_bar[stage], b_col0 + kb * kKElems, b_row, pol);
              ptx::tma_load_2d(sb + 128 * 128, &tmap, &full_bar[stage], b_col0 + kb * kKElems, b_row + 128, pol);
            } else {
              ptx::tma_load_2d_2sm(sa, &tmap, &full_bar[stage], a_col0 + kb * kKElems, a_row, pol_a);
              ptx::tma_load_2d_2sm(sb, &tmap, &full_bar[stage], b_col0 + kb * kKElems, b_row, pol);
              if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
              else ptx::mbar_arrive_cluster(&full_bar[stage], 0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = p.work_begin + unit; w < p.work_end; w += nunits) {
      for (int j = 0; j < p.nchunks; ++j, ++it) {
        const int kb_total = p.npass * (p.chunk_kpad[j] / kKElems);
        const int as = it % kAccStages;
        const uint32_t aphase = (uint32_t)(it / kAccStages) & 1u;
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1u);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kTileN);
        for (int kb = 0; kb < kb_total; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          if (ptx::elect_one()) {
            const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
            const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 4 x (UMMA_K * element size = 32 B) per 128-byte k-block
              const uint64_t adesc = ptx::make_smem_desc_kmajor_sw128(a_addr + k * 32);
              const uint64_t bdesc = ptx::make_smem_desc_kmajor_sw128(b_addr + k * 32);
              if constexpr (KIND == 2)
                ptx::umma_mxf4<CG>(d_tmem, adesc, bdesc, p.idesc, tmem_base + kSfaCol, tmem_base + kSfbCol,
                                   (kb > 0 || k > 0) ? 1u : 0u);
              else
                ptx::umma<CG, KIND == 1>(d_tmem, adesc, bdesc, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit<CG>(&empty_bar[stage]);                      // smem slot reusable
            if (kb == kb_total - 1) ptx::umma_commit<CG>(&tfull_bar[as]);  // accumulator complete
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      }
      if constexpr (CG == 2) {
        // keep the leader's barriers alive until the peer's last tmem-empty arrivals landed
        if (it > 0) {
          const int last = it - 1;
          ptx::mbar_wait(&tempty_bar[last % kAccStages], (uint32_t)(last / kAccStages) & 1u);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------- epilogue (fp64 / int32 RMW)
    const int q = warp - 4;  // == warp % 4: the TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    OutT* out = reinterpret_cast<OutT*>(p.out);
    // several resident chunks: the tile is revisited one chunk later by the same threads -> keep it in L2;
    // a single chunk per launch: the tile is dead until the next launch -> stream it
    const uint64_t pol_rmw = p.l2_policy == 0 ? ptx::policy_evict_normal()
                             : p.nchunks > 1 ? ptx::policy_evict_last() : ptx::policy_evict_first();
    int it = 0;
    for (int w = p.work_begin + unit; w < p.work_end; w += nunits) {
      const int4 t = p.work[w];
      const int sidx = (rank == 0) ? t.z : t.w;
      OutT* tile = out + (size_t)sidx * kTileElems + row;
    for (int j = 0; j < p.nchunks; ++j, ++it) {
      const int as = it % kAccStages;
      const uint32_t aphase = (uint32_t)(it / kAccStages) & 1u;
      const bool accumulate = p.accumulate || j > 0;
      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kTileN);
#pragma unroll 1
      for (int cc = 0; cc < kTileN / 32; ++cc) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + cc * 32, v);
        ptx::tmem_ld_wait();
        if (cc == kTileN / 32 - 1) {
          // accumulator stage fully read into registers: hand it back to the MMA warp
          ptx::tcgen05_fence_before();
          if constexpr (CG == 1) ptx::mbar_arrive(&tempty_bar[as]);
          else ptx::mbar_arrive_cluster(&tempty_bar[as], 0);
        }
        OutT* po = tile + (size_t)(cc * 32) * kTileM;
        const uint64_t pol_st = (p.l2_policy != 0 && j == p.nchunks - 1) ? ptx::policy_evict_first() : pol_rmw;  // last visit
        if (accumulate) {
          OutT old[32];
#pragma unroll
          for (int q2 = 0; q2 < 32; ++q2) old[q2] = ptx::ld_hint<OutT>(po + q2 * kTileM, pol_rmw);
#pragma unroll
          for (int q2 = 0; q2 < 32; ++q2) ptx::st_hint(po + q2 * kTileM, old[q2] + acc_convert<OutT>(v[q2]), pol_st);
        } else {
#pragma unroll
          for (int q2 = 0; q2 < 32; ++q2) ptx::st_hint(po + q2 * kTileM, acc_convert<OutT>(v[q2]), pol_st);
        }
      }
    }
    }
  }

  ptx::tcgen05_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<CG>(tmem_base, 512);
}

// Block-scaled descriptor for kind::mxf4: e2m1 operands (format 1), UE8M0 scales, K = 64, scale-factor ids 0.
inline uint32_t make_idesc_mxf4(int cg) {
  uint32_t d = 0;
  d |= 1u << 7;                               // A format: E2M1
  d |= 1u << 10;                              // B format: E2M1
  d |= (uint32_t)(256 >> 3) << 17;            // N >> 3
  d |= 1u << 23;                              // scale format: UE8M0
  d |= (uint32_t)((128 * cg) >> 4) << 24;     // M >> 4
  return d;
}

// tcgen05 instruction descriptor: fp32 accumulate, both operands K-major, N = 256, M = 128 * CG.
// a/b format: kind::f16 -> 0 = fp16, 1 = bf16; kind::f8f6f4 -> 0 = e4m3.
inline uint32_t make_idesc(int cg, int ab_format) {
  uint32_t d = 0;
  d |= 1u << 4;                               // accumulator format: F32
  d |= (uint32_t)ab_format << 7;              // A format
  d |= (uint32_t)ab_format << 10;             // B format
  d |= (uint32_t)(kTileN >> 3) << 17;         // N >> 3
  d |= (uint32_t)((128 * cg) >> 4) << 24;     // M >> 4
  return d;
}

}  // namespace dgrm
