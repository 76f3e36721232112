// conv_ffma.cu — the same tap-shifted GEMM-shaped convolution as conv_tc.cu, on CUDA cores in
// exact fp32 (FFMA).  It serves two purposes:
//   * HG_PREC_FP32_FFMA: an on-device, bit-faithful-to-fp32 cross-check of the tensor-core path at
//     sizes the CPU oracle cannot reach, and
//   * layers too narrow for a dense tensor-core contraction (C_in < 32: the V2-style 16/8-channel
//     stages), in every precision mode.
// Follows SURVEY.md A.1 / A.3 (reference hifi/models.py:19-81,161-171).
//
// Tile: 64 GEMM rows x 64 columns per 256-thread block, 4x4 outputs per thread; K is walked in
// chunks of 16 input channels; the activation slab (tile rows + dilation halo) is staged in shared
// memory once per chunk and re-read for every tap.
#include "common.cuh"

namespace hg {

constexpr int FT_M = 64, FT_N = 64, FT_K = 16;

__global__ void __launch_bounds__(256) conv_ffma_kernel(const FfmaConvParams p, int tiles_per_item, int min_off,
                                                       int span) {
  extern __shared__ float fsm[];
  const int slab_rows = FT_M + span;
  float* slab = fsm;                             // [slab_rows][FT_K + 1]
  float* wsm = fsm + ((slab_rows * (FT_K + 1) + 3) & ~3);  // [FT_K][FT_N], 16B aligned
  int b, tile;
  decode_tile(p.rag, tiles_per_item, static_cast<int>(blockIdx.x), b, tile);
  const int m0 = tile * FT_M;
  const int n0 = blockIdx.y * FT_N;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < p.cin; c0 += FT_K) {
    // stage the slab chunk: rows m0+min_off .. , channels c0..c0+15 (zero outside the sequence)
    for (int e = threadIdx.x; e < slab_rows * FT_K; e += 256) {
      const int r = e / FT_K, kk = e - r * FT_K;
      const int row = m0 + min_off + r, c = c0 + kk;
      float v = 0.f;
      if (row >= 0 && row < p.L_in && c < p.cin)
        v = load_operand(p.a0, p.a1, p.a_fmt, (static_cast<long long>(b) * p.L_in + row) * p.a_pitch + c);
      slab[r * (FT_K + 1) + kk] = v;
    }
    for (int t = 0; t < p.ntaps; ++t) {
      for (int e = threadIdx.x; e < FT_K * FT_N; e += 256) {
        const int kk = e / FT_N, nn = e - kk * FT_N;
        const int c = c0 + kk, n = n0 + nn;
        wsm[e] = (c < p.cin && n < p.n_total) ? p.w[(static_cast<long long>(t) * p.cin + c) * p.n_total + n] : 0.f;
      }
      __syncthreads();
      const int roff = p.tap_off[t] - min_off;
#pragma unroll
      for (int kk = 0; kk < FT_K; ++kk) {
        float a[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = slab[(ty * 4 + i + roff) * (FT_K + 1) + kk];
        {
          const float4 wv = *reinterpret_cast<const float4*>(&wsm[kk * FT_N + tx * 4]);
          w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // float4 epilogue only when every flat offset it produces is a multiple of 4
  const bool vec = ((static_cast<long long>(p.n_total) | p.epi.out_offset | p.epi.out_extent) & 3) == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long q = static_cast<long long>(m0) + ty * 4 + i;
    if (q >= p.rows) continue;
    const int n = n0 + tx * 4;
    if (vec) {
      if (n < p.n_total) epilogue_vec4(p.epi, b, q, n, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < p.n_total) epilogue_scalar(p.epi, b, q, n + j, acc[i][j]);
    }
  }
}

cudaError_t launch_conv_ffma(FfmaConvParams p, cudaStream_t st, const RaggedItems* items) {
  int min_off = p.tap_off[0], max_off = p.tap_off[0];
  for (int t = 1; t < p.ntaps; ++t) {
    min_off = p.tap_off[t] < min_off ? p.tap_off[t] : min_off;
    max_off = p.tap_off[t] > max_off ? p.tap_off[t] : max_off;
  }
  const int span = max_off - min_off;
  const int tiles = (p.rows + FT_M - 1) / FT_M;
  const size_t smem = (((static_cast<size_t>(FT_M + span) * (FT_K + 1) + 3) & ~size_t(3)) + FT_K * FT_N) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(conv_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  const int total = ragged_fill(&p.rag, items, p.B, p.rows, FT_M);
  dim3 grid(static_cast<unsigned>(total), static_cast<unsigned>((p.n_total + FT_N - 1) / FT_N));
  conv_ffma_kernel<<<grid, 256, smem, st>>>(p, tiles, min_off, span);
  return cudaGetLastError();
}

}  // namespace hg
