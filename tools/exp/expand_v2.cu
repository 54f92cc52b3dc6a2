// Stand-alone experiment for the next round: a register-resident rewrite of bed_expand_kernel.
// NOT part of libdgrm.so and not built by __graft_entry__.build(); written after the round's GPU budget was spent, so it
// has been compiled (nvcc -arch=sm_100a) but never run.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/expand_v2 tools/exp/expand_v2.cu
//   /tmp/expand_v2 [n=20000] [kc=16384]
//
// It builds a random packed chunk and per-SNP tables, runs the shipped kernel (bed_expand_kernel<false,false,2>, included
// from dissect_b200/csrc/bed_kernels.cuh) and the v2 kernel below on the same inputs, requires Z and Q4 to be bit
// identical and the fp64 diagonal partials to agree to 1e-12 relative, and prints both times against the HBM roofline.
//
// Why: the shipped kernel is shared-memory bound: per element one LDS.U8 (packed byte), one LDS.U16 (table) and one
// LDS.64 (square), ~3.3 LDS per element, 61 M wavefronts per 16 384-SNP chunk at n = 20 000, 0.29-0.38 ms against a
// 0.135 ms HBM roofline.  v2: a thread owns 8 consecutive SNPs x 8 consecutive individuals.  Its 8 packed half-words,
// its 8 four-entry fp16 tables (8 x 64 bit) and its 8 x 3 fp64 squares sit in registers (~100), every element is then
// pure ALU (BFE + PRMT for the operand, predicated selects + DADD for the exact diagonal), the 8 fp16 of one individual
// leave as one 16-byte global store (16 lanes = 256 contiguous bytes of a K-major operand row) and the e2m1 nibbles as
// one 4-byte store: no output staging in shared memory at all.  Shared memory only holds the packed tile and the
// tables: 40 LDS per thread per 64 elements.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../../dissect_b200/csrc/bed_kernels.cuh"

namespace v2 {

constexpr int kK = 128;       // SNPs per block
constexpr int kI = 128;       // individuals per block
constexpr int kPkPitch = 36;  // bytes per packed row in smem (32 + 4: rows 8 apart fall into different banks)

struct Smem {
  uint8_t pk[kK][kPkPitch];
  uint2 lut_hi[kK];           // four 16-bit operand values, pattern p at bits 16p
  double lut_sq[kK][4];
};

template <bool BF16>
__global__ void __launch_bounds__(256) expand_v2_kernel(const dgrm::ExpandArgs a) {
  __shared__ __align__(16) Smem sm;
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * kI;
  const int k0 = blockIdx.y * kK;

  // packed tile, same boundary convention as the shipped kernel: rows past the chunk and individuals past n = 01
  {
    const int r = tid >> 1, h = tid & 1;
    uint4 v = make_uint4(0x55555555u, 0x55555555u, 0x55555555u, 0x55555555u);
    if (k0 + r < a.kc) {
      const int64_t off = (int64_t)(k0 + r) * a.pitch + (i0 >> 2) + h * 16;
      if ((i0 >> 2) + h * 16 < a.pitch) v = __ldg(reinterpret_cast<const uint4*>(a.chunk + off));
    }
    if (i0 + kI > a.n) {
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int valid = a.n - (i0 + (h * 4 + q) * 16);
        valid = valid < 0 ? 0 : (valid > 16 ? 16 : valid);
        const uint32_t keep = valid == 16 ? 0xFFFFFFFFu : ((1u << (2 * valid)) - 1u);
        w[q] = (w[q] & keep) | (0x55555555u & ~keep);
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.pk[r][h * 16]);   // 36-byte pitch: 4-byte aligned only
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  if (tid < kK) {
    double z[4] = {0., 0., 0., 0.};
    if (k0 + tid < a.kc) {
      const double* zp = a.zt + (int64_t)(k0 + tid) * 4;
      z[0] = zp[0]; z[2] = zp[2]; z[3] = zp[3];
    }
    uint32_t h[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      h[c] = dgrm::round16<BF16>(z[c]);
      sm.lut_sq[tid][c] = __dmul_rn(z[c], z[c]);
    }
    sm.lut_hi[tid] = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
  }
  __syncthreads();

  const int sg = tid & 15;   // SNPs  k0 + 8 sg .. + 7
  const int ig = tid >> 4;   // individuals i0 + 8 ig .. + 7
  uint32_t w[8];             // 16 bits each: the 8 individuals' patterns for SNP t
  uint2 hi[8];
  double s0[8], s2[8], s3[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int k = sg * 8 + t;
    w[t] = *reinterpret_cast<const uint16_t*>(&sm.pk[k][ig * 2]);
    hi[t] = sm.lut_hi[k];
    s0[t] = sm.lut_sq[k][0];
    s2[t] = sm.lut_sq[k][2];
    s3[t] = sm.lut_sq[k][3];
  }

  uint16_t* zrow = a.Z + (int64_t)(i0 + ig * 8) * a.z_pitch + k0 + sg * 8;
  uint8_t* qrow = a.Q4 ? a.Q4 + (int64_t)(i0 + ig * 8) * a.q4_pitch + ((k0 + sg * 8) >> 1) : nullptr;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint32_t out[4];
    uint32_t present = 0;
    double d = 0.;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const uint32_t code = (w[t] >> (2 * q)) & 3u;
      // operand: 16 bits at bit 16*code of the 64-bit table -> byte permute with selector (2c+1, 2c)
      const uint32_t sel = code * 0x22u + 0x10u;            // nibbles: low byte <- 2c, next byte <- 2c + 1
      const uint32_t v = __byte_perm(hi[t].x, hi[t].y, sel) & 0xFFFFu;
      if (t & 1) out[t >> 1] |= v << 16; else out[t >> 1] = v;
      // exact diagonal: z[code]^2, missing (01) contributes 0
      const double sq = (code & 2u) ? ((code & 1u) ? s3[t] : s2[t]) : ((code & 1u) ? 0. : s0[t]);
      d = __dadd_rn(d, sq);
      present |= (uint32_t)(code != 1u) << (4 * t + 1);      // e2m1 1.0 = 0b0010 in nibble t
    }
    *reinterpret_cast<uint4*>(zrow + (int64_t)q * a.z_pitch) = make_uint4(out[0], out[1], out[2], out[3]);
    if (qrow) *reinterpret_cast<uint32_t*>(qrow + (int64_t)q * a.q4_pitch) = present;
    // sum over the 16 threads (128 SNPs) of this individual: lanes sg = 0..15 of a half-warp
#pragma unroll
    for (int s = 8; s > 0; s >>= 1) d = __dadd_rn(d, __shfl_xor_sync(0xffffffffu, d, s));
    if (sg == 0) a.diag_part[(int64_t)blockIdx.y * a.n_pad + i0 + ig * 8 + q] = d;
  }
}

}  // namespace v2

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 20000;
  const int kc = argc > 2 ? atoi(argv[2]) : 16384;
  const int n_pad = (n + 255) / 256 * 256;
  const int kc_pad = (kc + 255) / 256 * 256;
  const int64_t pitch = ((n + 3) / 4 + 127) / 128 * 128;
  std::vector<uint8_t> chunk((size_t)kc * pitch);
  std::vector<double> zt((size_t)kc * 4);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (auto& b : chunk) b = (uint8_t)rnd();
  for (int k = 0; k < kc; ++k) {
    const double p = 0.05 + 0.45 * (double)(rnd() % 10000) / 10000.0, sd = sqrt(2 * p * (1 - p));
    zt[(size_t)k * 4 + 0] = (1 - (1 + 2 * p)) / sd;
    zt[(size_t)k * 4 + 1] = 0.;
    zt[(size_t)k * 4 + 2] = (2 - (1 + 2 * p)) / sd;
    zt[(size_t)k * 4 + 3] = (3 - (1 + 2 * p)) / sd;
  }
  uint8_t *d_chunk, *d_q4[2];
  double *d_zt, *d_dp[2];
  uint16_t* d_z[2];
  const int64_t z_pitch = kc_pad, q4_pitch = kc_pad / 2;
  CK(cudaMalloc(&d_chunk, chunk.size()));
  CK(cudaMalloc(&d_zt, zt.size() * 8));
  CK(cudaMemcpy(d_chunk, chunk.data(), chunk.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_zt, zt.data(), zt.size() * 8, cudaMemcpyHostToDevice));
  for (int v = 0; v < 2; ++v) {
    CK(cudaMalloc(&d_z[v], (size_t)n_pad * z_pitch * 2));
    CK(cudaMalloc(&d_q4[v], (size_t)n_pad * q4_pitch));
    CK(cudaMalloc(&d_dp[v], (size_t)(kc_pad / 128) * n_pad * 8));
    CK(cudaMemset(d_z[v], 0xEE, (size_t)n_pad * z_pitch * 2));
    CK(cudaMemset(d_q4[v], 0xEE, (size_t)n_pad * q4_pitch));
  }
  const dim3 grid(n_pad / 128, kc_pad / 128);
  auto args = [&](int v) {
    dgrm::ExpandArgs a{};
    a.chunk = d_chunk; a.pitch = pitch; a.kc = kc; a.n = n; a.zt = d_zt;
    a.Z = d_z[v]; a.z_pitch = z_pitch; a.lo_off = 0;
    a.Q = nullptr; a.q_pitch = 0; a.Q4 = d_q4[v]; a.q4_pitch = q4_pitch; a.MB = nullptr; a.mb_pitch = 0;
    a.diag_part = d_dp[v]; a.n_pad = n_pad;
    return a;
  };
  CK(cudaFuncSetAttribute(dgrm::bed_expand_kernel<false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)sizeof(dgrm::ExpandSmem)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms[2] = {0, 0};
  for (int v = 0; v < 2; ++v) {
    for (int rep = 0; rep < 23; ++rep) {
      if (rep == 3) CK(cudaEventRecord(e0));
      if (v == 0) dgrm::bed_expand_kernel<false, false, 2><<<grid, 256, sizeof(dgrm::ExpandSmem)>>>(args(0));
      else v2::expand_v2_kernel<false><<<grid, 256>>>(args(1));
    }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms[v], e0, e1));
    ms[v] /= 20;
  }
  // compare
  std::vector<uint16_t> z0((size_t)n_pad * z_pitch), z1(z0.size());
  std::vector<uint8_t> q0((size_t)n_pad * q4_pitch), q1(q0.size());
  std::vector<double> p0((size_t)(kc_pad / 128) * n_pad), p1(p0.size());
  CK(cudaMemcpy(z0.data(), d_z[0], z0.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(z1.data(), d_z[1], z1.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(q0.data(), d_q4[0], q0.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(q1.data(), d_q4[1], q1.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(p0.data(), d_dp[0], p0.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(p1.data(), d_dp[1], p1.size() * 8, cudaMemcpyDeviceToHost));
  size_t bad_z = 0, bad_q = 0;
  for (size_t i = 0; i < z0.size(); ++i) bad_z += z0[i] != z1[i];
  for (size_t i = 0; i < q0.size(); ++i) bad_q += q0[i] != q1[i];
  double worst = 0;
  for (size_t i = 0; i < p0.size(); ++i) worst = fmax(worst, fabs(p0[i] - p1[i]) / fmax(1.0, fabs(p0[i])));
  const double bytes = 0.25 * (double)n * kc + 2.5 * (double)n_pad * kc_pad;
  printf("n=%d kc=%d: shipped %.3f ms (%.0f GB/s), v2 %.3f ms (%.0f GB/s); HBM roofline at 6550 GB/s: %.3f ms\n", n, kc,
         ms[0], bytes / ms[0] / 1e6, ms[1], bytes / ms[1] / 1e6, bytes / 6550e6);
  printf("mismatches: Z %zu, Q4 %zu; diag partial max rel diff %.3g -> %s\n", bad_z, bad_q, worst,
         (bad_z == 0 && bad_q == 0 && worst < 1e-12) ? "PASS" : "FAIL");
  return (bad_z == 0 && bad_q == 0 && worst < 1e-12) ? 0 : 1;
}
