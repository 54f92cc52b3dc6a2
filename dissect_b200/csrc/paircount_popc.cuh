// pair_count (popcount form): N_ij += popc(MB_i & MB_j) over the chunk's bit-packed non-missing
// masks.  Exact-integer replacement for the reference's second fp64 pdsyrk_ on the 0/1 `missings`
// matrix (kernel.cpp:100 -> matrix.cpp:2682).  Bound by the integer popc pipe (16 lanes/clk/SM),
// not by HBM or the tensor cores; the tcgen05 e4m3 path (syrk_tcgen05.cuh, KIND8) is the fast one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "syrk_tcgen05.cuh"

namespace dgrm {

constexpr int kPopcWords = 16;  // 32-bit words (512 SNPs) staged per step

// One block per 128x256 output tile (same work list and tile store as grm_syrk, CG 1 form).
__global__ void __launch_bounds__(256)
pair_count_popc_kernel(const uint32_t* __restrict__ MB, int64_t mb_pitch, int nwords, const int4* __restrict__ work,
                       int num_work, int32_t* __restrict__ out, int accumulate) {
  __shared__ uint32_t As[kPopcWords][kTileM + 1];
  __shared__ uint32_t Bs[kPopcWords][kTileN + 1];
  const int tid = threadIdx.x;
  const int ty = tid >> 4;  // rows ty*8 .. ty*8+7
  const int tx = tid & 15;  // cols tx + 16*j
  for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
    const int4 t = work[w];
    const int a_row = t.x, b_row = t.y;
    int acc[8][16];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[i][j] = 0;
    for (int w0 = 0; w0 < nwords; w0 += kPopcWords) {
      __syncthreads();
      for (int e = tid; e < kTileM * kPopcWords; e += 256) {
        const int r = e / kPopcWords, ww = e % kPopcWords;
        As[ww][r] = (w0 + ww < nwords) ? MB[(int64_t)(a_row + r) * mb_pitch + w0 + ww] : 0u;
      }
      for (int e = tid; e < kTileN * kPopcWords; e += 256) {
        const int r = e / kPopcWords, ww = e % kPopcWords;
        Bs[ww][r] = (w0 + ww < nwords) ? MB[(int64_t)(b_row + r) * mb_pitch + w0 + ww] : 0u;
      }
      __syncthreads();
#pragma unroll 2
      for (int ww = 0; ww < kPopcWords; ++ww) {
        uint32_t av[8], bv[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = As[ww][ty * 8 + i];
#pragma unroll
        for (int j = 0; j < 16; ++j) bv[j] = Bs[ww][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[i][j] += __popc(av[i] & bv[j]);
      }
    }
    int32_t* tile = out + (size_t)t.z * kTileElems;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int32_t* po = tile + (size_t)(tx + 16 * j) * kTileM + ty * 8;
      int4 lo = make_int4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      int4 hi = make_int4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
      if (accumulate) {
        const int4 ol = *reinterpret_cast<const int4*>(po);
        const int4 oh = *reinterpret_cast<const int4*>(po + 4);
        lo.x += ol.x; lo.y += ol.y; lo.z += ol.z; lo.w += ol.w;
        hi.x += oh.x; hi.y += oh.y; hi.z += oh.z; hi.w += oh.w;
      }
      *reinterpret_cast<int4*>(po) = lo;
      *reinterpret_cast<int4*>(po + 4) = hi;
    }
  }
}

}  // namespace dgrm
