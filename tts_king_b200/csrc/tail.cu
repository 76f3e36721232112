// tail.cu — the memory-bound ends of the generator.
//
//   conv_post_kernel   leaky_relu(0.01) -> Conv1d(C,1,7,padding=3) -> tanh  (hifi/models.py:197-199)
//                      optionally followed by HIFIapi.generate's  * MAX_WAV_VALUE -> int16 cast
//                      (hifiapi.py:50-51) fused into the store.  Pure bandwidth: reads C fp32 per
//                      sample, writes one sample (AI 3.4 FLOP/B).
//   mel_to_operand     strided fp32 mel [B,80,T] (possibly the transpose of a time-major tensor,
//                      tts_king.py:48) -> channels-last operand planes [B][T][C_pad] for conv_pre.
//   f32_to_operand     fp32 channels-last -> operand planes with leaky_relu (op-level entry points).
#include <math.h>

#include "common.cuh"

namespace hg {

constexpr int kPostTile = 256;  // output samples per block
constexpr int kPostK = 7;

// Where the waveform goes: sample t of item b lands at out[b * item_stride + (t - skip)] when
// skip <= t < skip + keep, and nowhere otherwise.  The dense forward uses {L, 0, L}; a time chunk
// computed with halo frames drops its halo samples here and writes its owned samples straight into
// their place in the final buffer — which may live on another GPU (peer-mapped: the stores travel
// over NVLink, hg_forward_window).
struct PostWindow {
  long long item_stride;
  int skip, keep;
};

// x: fp32 [B][L][C] raw stage output.  Block = kPostTile consecutive samples of one item.  The
// (tile + 6) x C window is staged in shared memory with coalesced float4 loads (leaky_relu applied
// once per element), rows padded by 4 floats so that a quarter-warp's LDS.128 hit distinct banks.
template <bool VEC4>
__global__ void __launch_bounds__(kPostTile) conv_post_kernel(const float* __restrict__ x, int L, int C,
                                                              const float* __restrict__ w,  // [7][C] tap-major
                                                              float bias, int tiles_per_item, float* __restrict__ out_f32,
                                                              int16_t* __restrict__ out_i16, float out_scale,
                                                              const RaggedPrefix rag, const PostWindow win) {
  extern __shared__ float psm[];
  const int pitch = VEC4 ? C + 4 : C + 1;
  float* xs = psm;                                   // [kPostTile + 6][pitch]
  float* ws = psm + (kPostTile + kPostK - 1) * pitch;  // [7][C]
  int b, tile;
  decode_tile(rag, tiles_per_item, static_cast<int>(blockIdx.x), b, tile);
  const int t0 = tile * kPostTile;
  if (t0 + kPostTile <= win.skip || t0 >= win.skip + win.keep) return;  // whole tile outside the window
  const float* xb = x + static_cast<long long>(b) * L * C;
  for (int e = threadIdx.x; e < kPostK * C; e += kPostTile) ws[e] = w[e];
  if (VEC4) {
    const int c4 = C >> 2;
    for (int e = threadIdx.x; e < (kPostTile + kPostK - 1) * c4; e += kPostTile) {
      const int r = e / c4, cc = (e - r * c4) << 2;
      const int row = t0 - 3 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row >= 0 && row < L) {
        v = *reinterpret_cast<const float4*>(xb + static_cast<long long>(row) * C + cc);
        v.x = lrelu(v.x, 0.01f); v.y = lrelu(v.y, 0.01f); v.z = lrelu(v.z, 0.01f); v.w = lrelu(v.w, 0.01f);
      }
      *reinterpret_cast<float4*>(xs + r * pitch + cc) = v;
    }
  } else {
    for (int e = threadIdx.x; e < (kPostTile + kPostK - 1) * C; e += kPostTile) {
      const int r = e / C, cc = e - r * C;
      const int row = t0 - 3 + r;
      xs[r * pitch + cc] = (row >= 0 && row < L) ? lrelu(xb[static_cast<long long>(row) * C + cc], 0.01f) : 0.f;
    }
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= L || t < win.skip || t >= win.skip + win.keep) return;
  float acc = bias;
#pragma unroll
  for (int j = 0; j < kPostK; ++j) {
    const float* xr = xs + (threadIdx.x + j) * pitch;
    const float* wr = ws + j * C;
    if (VEC4) {
      for (int c = 0; c < C; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(xr + c);
        const float4 ww = *reinterpret_cast<const float4*>(wr + c);
        acc = fmaf(a.x, ww.x, acc); acc = fmaf(a.y, ww.y, acc);
        acc = fmaf(a.z, ww.z, acc); acc = fmaf(a.w, ww.w, acc);
      }
    } else {
      for (int c = 0; c < C; ++c) acc = fmaf(xr[c], wr[c], acc);
    }
  }
  const float y = tanhf(acc);
  const long long o = static_cast<long long>(b) * win.item_stride + (t - win.skip);
  if (out_f32) out_f32[o] = y;
  if (out_i16) {
    // numpy float32 -> int16: truncate toward zero to int32, keep the low 16 bits (+1.0 -> -32768)
    const int v = __float2int_rz(y * out_scale);
    out_i16[o] = static_cast<int16_t>(static_cast<uint16_t>(static_cast<uint32_t>(v) & 0xFFFFu));
  }
}

// V1 fast path (C = 32).  The generic kernel above re-reads every staged row seven times (once per
// tap) and is bound by shared-memory bandwidth (ncu: 1.9 TB/s of HBM).  Here each staged row is read
// ONCE: its owner thread forms the seven per-tap partial dot products against weights that live in
// the kernel-parameter constant bank (no loads at all), parks them in a small padded tile, and each
// output sample is then the sum of seven partials from neighbouring rows.
struct PostW32 {
  float w[kPostK][32];
};

__global__ void __launch_bounds__(kPostTile) conv_post32_kernel(const float* __restrict__ x, int L, const PostW32 W,
                                                                float bias, int tiles_per_item, float* __restrict__ out_f32,
                                                                int16_t* __restrict__ out_i16, float out_scale,
                                                                const RaggedPrefix rag, const PostWindow win) {
  constexpr int C = 32, ROWS = kPostTile + kPostK - 1, PITCH = C + 4, PP = 9;
  __shared__ __align__(16) float xs[ROWS * PITCH];
  __shared__ float ps[ROWS * PP];
  int b, tile;
  decode_tile(rag, tiles_per_item, static_cast<int>(blockIdx.x), b, tile);
  const int t0 = tile * kPostTile;
  if (t0 + kPostTile <= win.skip || t0 >= win.skip + win.keep) return;  // whole tile outside the window
  const float* xb = x + static_cast<long long>(b) * L * C;
  // stage the window: all of a thread's loads are issued before the first is consumed (the rolled
  // loop had one 16-byte load in flight per thread and ran at 2 TB/s, latency-bound)
  constexpr int NV = ROWS * (C / 4);            // float4 elements in the window (2096)
  constexpr int FULL = NV / kPostTile;          // 8 unrolled rounds ...
  float4 v[FULL + 1];
#pragma unroll
  for (int i = 0; i <= FULL; ++i) {
    const int e = threadIdx.x + i * kPostTile;  // ... plus a 48-element tail round
    const int r = e >> 3, cc = (e & 7) << 2;
    const int row = t0 - 3 + r;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < NV && row >= 0 && row < L) v[i] = *reinterpret_cast<const float4*>(xb + static_cast<long long>(row) * C + cc);
  }
#pragma unroll
  for (int i = 0; i <= FULL; ++i) {
    const int e = threadIdx.x + i * kPostTile;
    if (e < NV) {
      const int r = e >> 3, cc = (e & 7) << 2;
      float4 t = v[i];
      t.x = lrelu(t.x, 0.01f); t.y = lrelu(t.y, 0.01f); t.z = lrelu(t.z, 0.01f); t.w = lrelu(t.w, 0.01f);
      *reinterpret_cast<float4*>(xs + r * PITCH + cc) = t;
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < ROWS; r += kPostTile) {
    float p[kPostK];
#pragma unroll
    for (int j = 0; j < kPostK; ++j) p[j] = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 a = *reinterpret_cast<const float4*>(xs + r * PITCH + 4 * c4);
#pragma unroll
      for (int j = 0; j < kPostK; ++j) {
        p[j] = fmaf(a.x, W.w[j][4 * c4], p[j]);
        p[j] = fmaf(a.y, W.w[j][4 * c4 + 1], p[j]);
        p[j] = fmaf(a.z, W.w[j][4 * c4 + 2], p[j]);
        p[j] = fmaf(a.w, W.w[j][4 * c4 + 3], p[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kPostK; ++j) ps[r * PP + j] = p[j];
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= L || t < win.skip || t >= win.skip + win.keep) return;
  float acc = bias;
#pragma unroll
  for (int j = 0; j < kPostK; ++j) acc += ps[(threadIdx.x + j) * PP + j];
  const float y = tanhf(acc);
  const long long o = static_cast<long long>(b) * win.item_stride + (t - win.skip);
  if (out_f32) out_f32[o] = y;
  if (out_i16) {
    const int v = __float2int_rz(y * out_scale);
    out_i16[o] = static_cast<int16_t>(static_cast<uint16_t>(static_cast<uint32_t>(v) & 0xFFFFu));
  }
}

cudaError_t launch_conv_post(const float* x, int B, int L, int C, const float* w_tapmajor, float bias,
                             float* out_f32, int16_t* out_i16, float out_scale, cudaStream_t st,
                             const float* w_host_tapmajor, const RaggedItems* items, long long item_stride, int skip,
                             int keep) {
  const int tiles = (L + kPostTile - 1) / kPostTile;
  PostWindow win;
  win.item_stride = item_stride > 0 ? item_stride : L;
  win.skip = skip;
  win.keep = keep > 0 ? keep : L - skip;
  RaggedPrefix rag;
  const int total = ragged_fill(&rag, items, B, L, kPostTile);
  if (C == 32 && w_host_tapmajor) {
    PostW32 W;
    for (int j = 0; j < kPostK; ++j)
      for (int c = 0; c < 32; ++c) W.w[j][c] = w_host_tapmajor[j * 32 + c];
    conv_post32_kernel<<<static_cast<unsigned>(total), kPostTile, 0, st>>>(x, L, W, bias, tiles, out_f32, out_i16, out_scale, rag, win);
    return cudaGetLastError();
  }
  const bool vec = (C & 3) == 0;
  const int pitch = vec ? C + 4 : C + 1;
  const size_t smem = (static_cast<size_t>(kPostTile + kPostK - 1) * pitch + kPostK * C) * sizeof(float);
  dim3 grid(static_cast<unsigned>(total));
  if (vec) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(conv_post_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    conv_post_kernel<true><<<grid, kPostTile, smem, st>>>(x, L, C, w_tapmajor, bias, tiles, out_f32, out_i16, out_scale, rag, win);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(conv_post_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    conv_post_kernel<false><<<grid, kPostTile, smem, st>>>(x, L, C, w_tapmajor, bias, tiles, out_f32, out_i16, out_scale, rag, win);
  }
  return cudaGetLastError();
}

// mel [B, C, T] with element strides (sB, sC, sT)  ->  operand planes [B][T][c_pad], zero padded.
// act: 0 none, 1 leaky_relu(slope), 2 tanh — the activation between two layers of a conv stack
// (hg_stack_forward; the generator's own leaky_relus are fused into its conv epilogues instead).
__global__ void mel_to_operand_kernel(const float* __restrict__ mel, long long sB, long long sC, long long sT, int B,
                                      int C, int T, int c_pad, int a_fmt, void* __restrict__ a0, void* __restrict__ a1,
                                      int act, float slope) {
  const long long total = static_cast<long long>(B) * T * c_pad;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(e % c_pad);
    const long long bt = e / c_pad;
    const int t = static_cast<int>(bt % T);
    const int b = static_cast<int>(bt / T);
    float v = c < C ? mel[b * sB + c * sC + t * sT] : 0.f;
    if (act == 1) v = lrelu(v, slope);
    else if (act == 2) v = tanhf(v);
    if (a_fmt == A_F32) {
      static_cast<float*>(a0)[e] = v;
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      static_cast<__nv_bfloat16*>(a0)[e] = h;
      if (a_fmt == A_BF16_SPLIT) static_cast<__nv_bfloat16*>(a1)[e] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

cudaError_t launch_mel_to_operand(const float* mel, long long sB, long long sC, long long sT, int B, int C, int T,
                                  int c_pad, int a_fmt, void* a0, void* a1, cudaStream_t st, int act, float slope) {
  const long long total = static_cast<long long>(B) * T * c_pad;
  const int blocks = static_cast<int>((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  mel_to_operand_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(mel, sB, sC, sT, B, C, T, c_pad, a_fmt, a0, a1, act, slope);
  return cudaGetLastError();
}

// x fp32 [n] -> operand planes of leaky_relu(x, slope)  (slope 1.0 = identity)
__global__ void f32_to_operand_kernel(const float* __restrict__ x, long long n, float slope, int a_fmt,
                                      void* __restrict__ a0, void* __restrict__ a1) {
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = lrelu(x[e], slope);
    if (a_fmt == A_F32) {
      static_cast<float*>(a0)[e] = v;
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      static_cast<__nv_bfloat16*>(a0)[e] = h;
      if (a_fmt == A_BF16_SPLIT) static_cast<__nv_bfloat16*>(a1)[e] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

cudaError_t launch_f32_to_operand(const float* x, long long n, float slope, int a_fmt, void* a0, void* a1,
                                  cudaStream_t st) {
  const int blocks = static_cast<int>((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
  f32_to_operand_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(x, n, slope, a_fmt, a0, a1);
  return cudaGetLastError();
}

}  // namespace hg
