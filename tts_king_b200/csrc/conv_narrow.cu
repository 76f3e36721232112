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
// ------------------------------------------------------------------------------------------------
// ConvTranspose1d 16 -> 16 channels, stride s, kernel 2s (two taps per output phase), bf16 operands: the
// upsampler in front of a stage that runs zero-padded to 16 channels (V2-style: 16 -> 8, ups.3).  A
// [L_in*s rows] x [16] output from a [L_in] x [16] input is 32 MACs per output element — nothing for a
// tensor core to do; it is a streaming kernel: one thread per output row reads its two input rows (32 B
// each), contracts them with the phase's [2][16][16] weight slice from shared memory and writes the fp32 row
// (64 B) and the bf16 operand row (32 B).  Polyphase form of SURVEY.md A.3 (reference hifi/models.py:161-171).
constexpr int kConvtRows = 256;

__global__ void __launch_bounds__(kConvtRows) convt_narrow16_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ w,
                                                                    int L_in, int stride, int pad, int tiles_per_item,
                                                                    const RaggedPrefix rag, const EpiParams e) {
  extern __shared__ __align__(16) float cws[];  // [2 taps][16 c_in][stride*16]
  const int n_total = stride * 16;
  for (int i = threadIdx.x; i < 2 * 16 * n_total; i += kConvtRows) cws[i] = w[i];
  __syncthreads();
  int b, tile;
  decode_tile(rag, tiles_per_item, static_cast<int>(blockIdx.x), b, tile);
  const int L_out = L_in * stride;
  const int n = tile * kConvtRows + threadIdx.x;  // output row
  if (n >= L_out) return;
  const int q = (n + pad) / stride, r = (n + pad) - q * stride;  // y[n] = sum_m W[., ., r + s*m] x[q - m]
  float acc[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) acc[o] = e.bias[o];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int row = q - m;
    if (row < 0 || row >= L_in) continue;
    const uint4* ap = reinterpret_cast<const uint4*>(a + (static_cast<long long>(b) * L_in + row) * 16);
    const uint4 h0 = ap[0], h1 = ap[1];
    const __nv_bfloat162* p0 = reinterpret_cast<const __nv_bfloat162*>(&h0);
    const __nv_bfloat162* p1 = reinterpret_cast<const __nv_bfloat162*>(&h1);
    float x[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[2 * i] = __low2float(p0[i]); x[2 * i + 1] = __high2float(p0[i]);
      x[8 + 2 * i] = __low2float(p1[i]); x[8 + 2 * i + 1] = __high2float(p1[i]);
    }
    const float* wm = cws + (m * 16) * n_total + r * 16;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4* wr = reinterpret_cast<const float4*>(wm + c * n_total);
#pragma unroll
      for (int o4 = 0; o4 < 4; ++o4) {
        const float4 ww = wr[o4];
        acc[4 * o4] = fmaf(x[c], ww.x, acc[4 * o4]); acc[4 * o4 + 1] = fmaf(x[c], ww.y, acc[4 * o4 + 1]);
        acc[4 * o4 + 2] = fmaf(x[c], ww.z, acc[4 * o4 + 2]); acc[4 * o4 + 3] = fmaf(x[c], ww.w, acc[4 * o4 + 3]);
      }
    }
  }
  const long long idx = (static_cast<long long>(b) * L_out + n) * 16;
  if (e.out_x) {
    float4* xp = reinterpret_cast<float4*>(e.out_x + idx);
#pragma unroll
    for (int o4 = 0; o4 < 4; ++o4) xp[o4] = make_float4(acc[4 * o4], acc[4 * o4 + 1], acc[4 * o4 + 2], acc[4 * o4 + 3]);
  }
  if (e.out_a0) {
    uint2* hp = reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(e.out_a0) + idx);
#pragma unroll
    for (int o4 = 0; o4 < 4; ++o4)
      hp[o4] = pack_bf16x4(lrelu_fast(acc[4 * o4], e.slope), lrelu_fast(acc[4 * o4 + 1], e.slope),
                           lrelu_fast(acc[4 * o4 + 2], e.slope), lrelu_fast(acc[4 * o4 + 3], e.slope));
  }
}

// w: the layer's CUDA-core weight image [2 taps][16][stride*16] (plan.h w_ffma); epi.bias: [stride*16], first 16 used
cudaError_t launch_convt_narrow16(const void* a, const float* w, int B, int L_in, int stride, int pad, EpiParams epi,
                                  cudaStream_t st, const RaggedItems* items) {
  const int L_out = L_in * stride;
  RaggedItems out_items;
  const RaggedItems* it = nullptr;
  if (items && items->n) {  // valid GEMM rows of the input -> valid output rows
    out_items.n = items->n;
    for (int b = 0; b < items->n; ++b) {
      const long long v = static_cast<long long>(items->valid_rows[b]) * stride;
      out_items.valid_rows[b] = static_cast<int>(v < L_out ? v : L_out);
    }
    it = &out_items;
  }
  RaggedPrefix rag;
  const int total = ragged_fill(&rag, it, B, L_out, kConvtRows);
  const size_t smem = static_cast<size_t>(2) * 16 * stride * 16 * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  convt_narrow16_kernel<<<static_cast<unsigned>(total), kConvtRows, smem, st>>>(static_cast<const __nv_bfloat16*>(a), w, L_in, stride,
                                                                                pad, (L_out + kConvtRows - 1) / kConvtRows, rag, epi);
  return cudaGetLastError();
}

cudaError_t launch_conv_narrow(int c, NarrowConvParams p, cudaStream_t st, const RaggedItems* items) {
  p.tiles_per_item = (p.L + kNarrowRows - 1) / kNarrowRows;
  const int total = ragged_fill(&p.rag, items, p.B, p.L, kNarrowRows);
  if (c == 16) return launch_narrow_c<16>(p, total, st);
  if (c == 8) return launch_narrow_c<8>(p, total, st);
  return cudaErrorInvalidValue;
}

}  // namespace hg
