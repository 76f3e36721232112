// conv_narrow.cu — same-length dilated Conv1d for very narrow layers (C_in = C_out = C in {8, 16}):
// the 16- and 8-channel late stages of the V2-style config (BASELINE cfg-4).
//
// At these widths a layer moves 4-12 bytes per output element for only 2*C*k FLOPs, and C is below
// the K = 16 granularity of the tensor-core path's smallest tile, so this is a CUDA-core kernel built
// around the memory system: one thread owns one time step and all C output channels; the operand
// slab (tile rows + dilation halo, converted to fp32 once) and the layer's whole weight tensor live
// in shared memory; every weight read is a warp-wide broadcast LDS.128 (one wavefront), every
// activation read a conflict-free LDS.128 of the thread's own padded row; the fused epilogue is the
// shared common.cuh one (bias, residual, MRF accumulate, divide, fp32 x and operand copy).
// Follows SURVEY.md A.1 (reference hifi/models.py:19-81, ResBlock convs).
#include "common.cuh"

namespace hg {

constexpr int kNarrowThreads = 256;
constexpr int kNarrowRPT = 2;                          // time steps per thread: every weight LDS feeds RPT rows
constexpr int kNarrowRows = kNarrowThreads * kNarrowRPT;  // time steps per block

template <int C>
__global__ void __launch_bounds__(kNarrowThreads) conv_narrow_kernel(const NarrowConvParams p) {
  constexpr int PITCH = C + 4;  // floats; keeps a quarter-warp's LDS.128 of 8 consecutive rows conflict-free
  extern __shared__ __align__(16) float nsm[];
  const int span = (p.k - 1) * p.dil;
  const int pad = span >> 1;
  const int slab_rows = kNarrowRows + span;
  float* slab = nsm;                         // [slab_rows][PITCH]
  float* ws = nsm + slab_rows * PITCH;       // [k][C][C]
  int b, tile;
  decode_tile(p.rag, p.tiles_per_item, static_cast<int>(blockIdx.x), b, tile);
  const int t0 = tile * kNarrowRows;

  for (int e = threadIdx.x; e < p.k * C * C / 4; e += kNarrowThreads)
    reinterpret_cast<float4*>(ws)[e] = reinterpret_cast<const float4*>(p.w)[e];

  // stage the slab: units of 8 channels (16 B of bf16 / 32 B of fp32), several loads in flight
  constexpr int UPR = C / 8;  // units per row
  const int units = slab_rows * UPR;
#pragma unroll 4
  for (int e = threadIdx.x; e < units; e += kNarrowThreads) {
    const int r = e / UPR, c0 = (e - r * UPR) * 8;
    const int row = t0 - pad + r;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (row >= 0 && row < p.L) {
      const long long idx = (static_cast<long long>(b) * p.L + row) * C + c0;
      if (p.a_fmt == A_F32) {
        const float4 lo = *reinterpret_cast<const float4*>(static_cast<const float*>(p.a0) + idx);
        const float4 hi = *reinterpret_cast<const float4*>(static_cast<const float*>(p.a0) + idx + 4);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
      } else {
        const uint4 h = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.a0) + idx);
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(hp[i]); v[2 * i + 1] = __high2float(hp[i]); }
        if (p.a_fmt == A_BF16_SPLIT) {
          const uint4 l = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.a1) + idx);
          const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
          for (int i = 0; i < 4; ++i) { v[2 * i] += __low2float(lp[i]); v[2 * i + 1] += __high2float(lp[i]); }
        }
      }
    }
    float* dst = slab + r * PITCH + c0;
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();

  // thread t owns rows t and t + 256 of the tile
  float acc[kNarrowRPT][C];
#pragma unroll
  for (int r = 0; r < kNarrowRPT; ++r)
#pragma unroll
    for (int i = 0; i < C; ++i) acc[r][i] = 0.f;
  for (int j = 0; j < p.k; ++j) {
    float x[kNarrowRPT][C];
#pragma unroll
    for (int r = 0; r < kNarrowRPT; ++r) {
      const float* xr = slab + (threadIdx.x + r * kNarrowThreads + j * p.dil) * PITCH;
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 t = *reinterpret_cast<const float4*>(xr + 4 * c4);
        x[r][4 * c4] = t.x; x[r][4 * c4 + 1] = t.y; x[r][4 * c4 + 2] = t.z; x[r][4 * c4 + 3] = t.w;
      }
    }
    const float* wj = ws + j * C * C;
#pragma unroll
    for (int ci = 0; ci < C; ++ci) {
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wj + ci * C + 4 * c4);  // warp-wide broadcast
#pragma unroll
        for (int r = 0; r < kNarrowRPT; ++r) {
          acc[r][4 * c4] = fmaf(x[r][ci], w4.x, acc[r][4 * c4]);
          acc[r][4 * c4 + 1] = fmaf(x[r][ci], w4.y, acc[r][4 * c4 + 1]);
          acc[r][4 * c4 + 2] = fmaf(x[r][ci], w4.z, acc[r][4 * c4 + 2]);
          acc[r][4 * c4 + 3] = fmaf(x[r][ci], w4.w, acc[r][4 * c4 + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kNarrowRPT; ++r) {
    const long long q = static_cast<long long>(t0) + threadIdx.x + r * kNarrowThreads;
    if (q >= p.L) continue;
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4)
      epilogue_vec4(p.epi, b, q, 4 * c4, acc[r][4 * c4], acc[r][4 * c4 + 1], acc[r][4 * c4 + 2], acc[r][4 * c4 + 3]);
  }
}

template <int C>
static cudaError_t launch_narrow_c(const NarrowConvParams& p, int total_tiles, cudaStream_t st) {
  const int span = (p.k - 1) * p.dil;
  const size_t smem = (static_cast<size_t>(kNarrowRows + span) * (C + 4) + static_cast<size_t>(p.k) * C * C) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(conv_narrow_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  conv_narrow_kernel<C><<<static_cast<unsigned>(total_tiles), kNarrowThreads, smem, st>>>(p);
  return cudaGetLastError();
}

// C must be 8 or 16; k odd.
cudaError_t launch_conv_narrow(int c, NarrowConvParams p, cudaStream_t st, const RaggedItems* items) {
  p.tiles_per_item = (p.L + kNarrowRows - 1) / kNarrowRows;
  const int total = ragged_fill(&p.rag, items, p.B, p.L, kNarrowRows);
  if (c == 16) return launch_narrow_c<16>(p, total, st);
  if (c == 8) return launch_narrow_c<8>(p, total, st);
  return cudaErrorInvalidValue;
}

}  // namespace hg
