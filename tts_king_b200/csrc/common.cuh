// common.cuh — structures shared by the host plan and the kernels.
//
// Data layout in HBM (DESIGN.md "Layout"): every activation is channels-last (time-major)
// [B][L][C].  Each stage tensor exists in up to two forms:
//   * the fp32 residual stream x (what hifi/models.py:94 adds back), and
//   * the "operand" copy a = leaky_relu(x) that the next convolution contracts over
//     (hifi/models.py:90,92,188), stored in the format the arithmetic mode consumes:
//       A_BF16        one bf16 plane                        (HG_PREC_BF16)
//       A_BF16_SPLIT  two bf16 planes hi + lo, a ~= hi+lo    (HG_PREC_FP32, bf16x3 products)
//       A_F32         one fp32 plane                        (HG_PREC_FP32_FFMA)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hg {

constexpr int kMaxTaps = 16;

// Ragged batches: item b only needs its first valid_rows[b] GEMM rows, so the tile
// index space is compacted — tile t belongs to the item b with prefix[b] <= t < prefix[b+1] and is
// that item's (t - prefix[b])-th tile.  The table travels in the kernel parameters (constant bank),
// hence the bound on the batch size; n == 0 means a dense batch (every item has tiles_per_item tiles).
constexpr int kMaxRaggedItems = 64;
struct RaggedPrefix {
  int n;
  int rev_total;  // > 0: the launch walks its tiles backwards (tile index t -> rev_total - 1 - t), see hg_reverse_tiles
  int prefix[kMaxRaggedItems + 1];
};
#ifdef __CUDACC__
__device__ __forceinline__ void decode_tile(const RaggedPrefix& r, int tiles_per_item, int t, int& b, int& tile) {
  if (r.rev_total) t = r.rev_total - 1 - t;
  if (r.n == 0) {
    b = t / tiles_per_item;
    tile = t - b * tiles_per_item;
    return;
  }
  int lo = 0, hi = r.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (r.prefix[mid] <= t) lo = mid; else hi = mid;
  }
  b = lo;
  tile = t - r.prefix[lo];
}
#endif
// host side: valid GEMM rows per item of one launch (n == 0: dense) -> prefix table; returns the tile count
struct RaggedItems {
  int n;
  int valid_rows[kMaxRaggedItems];
};
// Tile order: consecutive launches of a forward walk their tiles in OPPOSITE directions.  Each layer streams
// hundreds of MB through the 126 MB L2; what is left there when a launch ends is the END of the tensors it wrote,
// which a forward-walking consumer would only reach after having evicted it.  Walking backwards, the next launch
// starts on those lines (profiles/r2_tile_order_l2.md).  The flag flips with every table built (= every launch) and
// is reset at the start of a forward, so the order is a pure function of the layer sequence.
extern thread_local int hg_reverse_tiles;  // api.cu; -1 = alternation switched off (HG_TILE_ORDER=0)
inline int ragged_finish(RaggedPrefix* r, int total) {
  r->rev_total = hg_reverse_tiles > 0 ? total : 0;
  if (hg_reverse_tiles >= 0) hg_reverse_tiles ^= 1;
  return total;
}
inline int ragged_fill(RaggedPrefix* r, const RaggedItems* items, int B, int rows, int tile_rows) {
  const int dense = (rows + tile_rows - 1) / tile_rows;
  r->n = 0;
  if (!items || items->n == 0) return ragged_finish(r, B * dense);
  r->n = B;
  r->prefix[0] = 0;
  for (int b = 0; b < B; ++b) {
    int v = items->valid_rows[b] < rows ? items->valid_rows[b] : rows;
    if (v < 1) v = 1;
    r->prefix[b + 1] = r->prefix[b] + (v + tile_rows - 1) / tile_rows;
  }
  return ragged_finish(r, r->prefix[B]);
}

enum OperandFmt : int { A_BF16 = 0, A_BF16_SPLIT = 1, A_F32 = 2 };

// Epilogue shared by every GEMM-shaped layer.  A GEMM element (item b, row q, column n) lands at
// flat offset f = q*out_row_stride + n + out_offset inside item b's [L_out][C_out] tensor and is
// dropped unless 0 <= f < out_extent.  For a same-length Conv1d: out_row_stride = C_out,
// out_offset = 0.  For the polyphase ConvTranspose1d the GEMM row holds all
// `stride` phases of one input position: out_row_stride = stride*C_out, out_offset = -pad*C_out.
struct EpiParams {
  const float* bias;    // [n_total]
  const float* res;     // fp32 residual, same indexing as the output (nullable)   :94  x = xt + x
  const float* acc_in;  // fp32 running MRF sum (nullable)                          :195 xs += ...
  float* out_x;         // fp32 result (nullable)
  void* out_a0;         // operand copy of leaky_relu(result): bf16 hi plane or fp32 (nullable)
  void* out_a1;         // bf16 lo plane (A_BF16_SPLIT only)
  int a_fmt;            // OperandFmt of out_a*
  float slope;          // leaky_relu slope baked into the operand copy (0.1, hifi/models.py:9)
  float post_div;       // > 0: result = (acc_in + result) / post_div                :196 xs / num_kernels
  int n_valid;          // > 0: GEMM columns >= n_valid are computed on zero weights and not stored (an output width
                        // padded up to the MMA's N granularity: the 80-mel layers of the FastSpeech2 tail run as N = 96)
  long long out_batch_stride;  // elements between items (= L_out*C_out)
  long long out_offset;
  long long out_extent;  // L_out*C_out
  int out_row_stride;
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// Four consecutive columns n..n+3 of GEMM row q, item b (all offsets are multiples of 4 by
// construction, so the float4 / 8-byte accesses are aligned).  `v` holds accumulator + nothing
// yet: bias is added here.
__device__ __forceinline__ void epilogue_vec4(const EpiParams& e, int b, long long q, int n, float v0, float v1,
                                              float v2, float v3) {
  const long long f = q * e.out_row_stride + n + e.out_offset;
  if (f < 0 || f + 4 > e.out_extent || (e.n_valid > 0 && n + 4 > e.n_valid)) return;
  const long long idx = static_cast<long long>(b) * e.out_batch_stride + f;
  const float4 bb = *reinterpret_cast<const float4*>(e.bias + n);
  v0 += bb.x; v1 += bb.y; v2 += bb.z; v3 += bb.w;
  if (e.res) {
    const float4 r = *reinterpret_cast<const float4*>(e.res + idx);
    v0 += r.x; v1 += r.y; v2 += r.z; v3 += r.w;
  }
  if (e.acc_in) {
    const float4 r = *reinterpret_cast<const float4*>(e.acc_in + idx);
    v0 = r.x + v0; v1 = r.y + v1; v2 = r.z + v2; v3 = r.w + v3;
  }
  if (e.post_div > 0.f) {
    v0 = __fdiv_rn(v0, e.post_div); v1 = __fdiv_rn(v1, e.post_div);
    v2 = __fdiv_rn(v2, e.post_div); v3 = __fdiv_rn(v3, e.post_div);
  }
  if (e.out_x) *reinterpret_cast<float4*>(e.out_x + idx) = make_float4(v0, v1, v2, v3);
  if (e.out_a0) {
    const float a0 = lrelu(v0, e.slope), a1 = lrelu(v1, e.slope), a2 = lrelu(v2, e.slope), a3 = lrelu(v3, e.slope);
    if (e.a_fmt == A_F32) {
      *reinterpret_cast<float4*>(static_cast<float*>(e.out_a0) + idx) = make_float4(a0, a1, a2, a3);
    } else {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
      const __nv_bfloat16 h2 = __float2bfloat16_rn(a2), h3 = __float2bfloat16_rn(a3);
      __nv_bfloat162 p0(h0, h1), p1(h2, h3);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(e.out_a0) + idx) = pk;
      if (e.a_fmt == A_BF16_SPLIT) {
        __nv_bfloat162 q0(__float2bfloat16_rn(a0 - __bfloat162float(h0)), __float2bfloat16_rn(a1 - __bfloat162float(h1)));
        __nv_bfloat162 q1(__float2bfloat16_rn(a2 - __bfloat162float(h2)), __float2bfloat16_rn(a3 - __bfloat162float(h3)));
        uint2 pl;
        pl.x = *reinterpret_cast<uint32_t*>(&q0);
        pl.y = *reinterpret_cast<uint32_t*>(&q1);
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(e.out_a1) + idx) = pl;
      }
    }
  }
}

// NR rows (q0, q0+row_step, ...) x four columns per thread with every global load issued before the
// first use — the memory-level parallelism the HBM-bound layers live on.  Same arithmetic and
// ordering as epilogue_vec4; index math is one 64-bit base plus a constant step.
__device__ __forceinline__ uint2 pack_bf16x4(float a0, float a1, float a2, float a3) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a0, a1), p1 = __floats2bfloat162_rn(a2, a3);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&p0);
  pk.y = *reinterpret_cast<uint32_t*>(&p1);
  return pk;
}
// leaky_relu for 0 <= slope <= 1 in two instructions
__device__ __forceinline__ float lrelu_fast(float v, float slope) { return fmaxf(v, v * slope); }

template <int NR, bool WITH_RES = true>
__device__ __forceinline__ void epilogue_rows(const EpiParams& e, int b, long long q0, int row_step, int n,
                                              float (&v)[NR][4], long long q_limit = 0x7fffffffffffffffLL,
                                              long long q_min = -0x7fffffffffffffffLL) {
  const long long f0 = q0 * e.out_row_stride + n + e.out_offset;
  const long long fstep = static_cast<long long>(row_step) * e.out_row_stride;
  const long long base = static_cast<long long>(b) * e.out_batch_stride;
  if (e.n_valid > 0 && n + 4 > e.n_valid) return;  // a padded output column group
  bool ok[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const long long f = f0 + i * fstep, q = q0 + static_cast<long long>(i) * row_step;
    ok[i] = f >= 0 && f + 4 <= e.out_extent && q < q_limit && q >= q_min;
  }
  const float4 bb = *reinterpret_cast<const float4*>(e.bias + n);
  if (WITH_RES && e.res) {
    const float* rp = e.res + base + f0;
    float4 r[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) r[i] = ok[i] ? *reinterpret_cast<const float4*>(rp + i * fstep) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      v[i][0] = (v[i][0] + bb.x) + r[i].x; v[i][1] = (v[i][1] + bb.y) + r[i].y;
      v[i][2] = (v[i][2] + bb.z) + r[i].z; v[i][3] = (v[i][3] + bb.w) + r[i].w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NR; ++i) { v[i][0] += bb.x; v[i][1] += bb.y; v[i][2] += bb.z; v[i][3] += bb.w; }
  }
  if (e.acc_in) {  // MRF accumulate: 8 of 72 layers — loads in two batches to keep registers down
    const float* ap = e.acc_in + base + f0;
#pragma unroll
    for (int h = 0; h < NR; h += 4) {
      float4 r[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        r[i] = (h + i < NR && ok[h + i]) ? *reinterpret_cast<const float4*>(ap + (h + i) * fstep) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (h + i >= NR) continue;
        v[h + i][0] = r[i].x + v[h + i][0]; v[h + i][1] = r[i].y + v[h + i][1];
        v[h + i][2] = r[i].z + v[h + i][2]; v[h + i][3] = r[i].w + v[h + i][3];
      }
    }
  }
  if (e.post_div > 0.f) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      v[i][0] = __fdiv_rn(v[i][0], e.post_div); v[i][1] = __fdiv_rn(v[i][1], e.post_div);
      v[i][2] = __fdiv_rn(v[i][2], e.post_div); v[i][3] = __fdiv_rn(v[i][3], e.post_div);
    }
  }
  if (e.out_x) {
    float* xp = e.out_x + base + f0;
#pragma unroll
    for (int i = 0; i < NR; ++i)
      if (ok[i]) *reinterpret_cast<float4*>(xp + i * fstep) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
  }
  if (e.out_a0) {
    const float sl = e.slope;
    if (e.a_fmt == A_BF16) {
      __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(e.out_a0) + base + f0;
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if (ok[i])
          *reinterpret_cast<uint2*>(hp + i * fstep) = pack_bf16x4(lrelu_fast(v[i][0], sl), lrelu_fast(v[i][1], sl),
                                                                  lrelu_fast(v[i][2], sl), lrelu_fast(v[i][3], sl));
    } else if (e.a_fmt == A_BF16_SPLIT) {
      __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(e.out_a0) + base + f0;
      __nv_bfloat16* lp = static_cast<__nv_bfloat16*>(e.out_a1) + base + f0;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        if (!ok[i]) continue;
        const float a0 = lrelu_fast(v[i][0], sl), a1 = lrelu_fast(v[i][1], sl), a2 = lrelu_fast(v[i][2], sl), a3 = lrelu_fast(v[i][3], sl);
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(a0, a1), h23 = __floats2bfloat162_rn(a2, a3);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&h01);
        pk.y = *reinterpret_cast<const uint32_t*>(&h23);
        *reinterpret_cast<uint2*>(hp + i * fstep) = pk;
        *reinterpret_cast<uint2*>(lp + i * fstep) =
            pack_bf16x4(a0 - __low2float(h01), a1 - __high2float(h01), a2 - __low2float(h23), a3 - __high2float(h23));
      }
    } else {
      float* fp = static_cast<float*>(e.out_a0) + base + f0;
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if (ok[i])
          *reinterpret_cast<float4*>(fp + i * fstep) = make_float4(lrelu_fast(v[i][0], sl), lrelu_fast(v[i][1], sl),
                                                                   lrelu_fast(v[i][2], sl), lrelu_fast(v[i][3], sl));
    }
  }
}

// Packed fp32 pairs (FADD2 / FMUL2 on sm_100): one issue slot for two IEEE round-to-nearest operations — the same
// bits as the scalar forms, half the instructions in the issue-bound pair epilogues.
__device__ __forceinline__ float2 add_f32x2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 mul_f32x2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {  // a + b, operand order kept
  const float2 lo = add_f32x2(make_float2(a.x, a.y), make_float2(b.x, b.y));
  const float2 hi = add_f32x2(make_float2(a.z, a.w), make_float2(b.z, b.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// bf16x4 of leaky_relu(v) for 0 <= slope <= 1 (lrelu_fast on four values)
__device__ __forceinline__ uint2 lrelu_pack4(float4 v, float slope) {
  const float2 sl = make_float2(slope, slope);
  const float2 lo = mul_f32x2(make_float2(v.x, v.y), sl), hi = mul_f32x2(make_float2(v.z, v.w), sl);
  return pack_bf16x4(fmaxf(v.x, lo.x), fmaxf(v.y, lo.y), fmaxf(v.z, hi.x), fmaxf(v.w, hi.y));
}

// x / 3 in three packed instructions instead of __fdiv_rn's ~10 per element (hifi/models.py:196 divides the MRF sum
// by num_kernels = 3): q0 = x * RN(1/3); rem = fma(-3, q0, x) (exact); q = fma(rem, RN(1/3), q0) is the correctly
// rounded quotient (Markstein); the sign is copied from x so that -0 stays -0.  Checked against x / 3.0f for ALL
// 2^32 bit patterns (tests/test_oracle.py::test_div3_sequence_matches_ieee_division runs the C restatement over a
// sample plus the edge cases): the only differences are +-inf (rem = inf - inf), which take the __fdiv_rn path.
__device__ __forceinline__ float2 fma_f32x2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float copy_sign_bit(float mag, float sgn) {
  return __uint_as_float((__float_as_uint(mag) & 0x7fffffffu) | (__float_as_uint(sgn) & 0x80000000u));
}
__device__ __forceinline__ float4 div3_rn(float4 v) {
  const float big = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
  if (!(big < __int_as_float(0x7f800000)))  // an infinity or a NaN among the four
    return make_float4(__fdiv_rn(v.x, 3.f), __fdiv_rn(v.y, 3.f), __fdiv_rn(v.z, 3.f), __fdiv_rn(v.w, 3.f));
  const float2 r = make_float2(0x1.555556p-2f, 0x1.555556p-2f), m3 = make_float2(-3.f, -3.f);
  const float2 lo = make_float2(v.x, v.y), hi = make_float2(v.z, v.w);
  const float2 ql = mul_f32x2(lo, r), qh = mul_f32x2(hi, r);
  const float2 l2 = fma_f32x2(fma_f32x2(m3, ql, lo), r, ql), h2 = fma_f32x2(fma_f32x2(m3, qh, hi), r, qh);
  return make_float4(copy_sign_bit(l2.x, v.x), copy_sign_bit(l2.y, v.y), copy_sign_bit(h2.x, v.z), copy_sign_bit(h2.y, v.w));
}

// The tail of a fused pair's E2 (conv_pair_tc.cu, conv_pair_fold.cu), written for INSTRUCTION COUNT — those kernels'
// epilogue warps are issue-bound (profiles/r2_pair_epilogue_issue_bound.md).  The thread holds eight float4: rows
// 0, 4, ..., 28 of its warp item at four consecutive columns, residual already added.  `off` is the element offset of
// row 0 (item, row and column folded into one 64-bit value by the caller), rows step by the compile-time STEP
// elements, and row ii is stored iff 4*ii < nv (one 32-bit count instead of eight 64-bit range checks).  Same
// arithmetic, in the same order, as epilogue_rows<8, false>.
template <int STEP>
__device__ __forceinline__ void epilogue_tail8(const EpiParams& e, long long off, int nv, const float4 bias, float4 (&v)[8]) {
#pragma unroll
  for (int ii = 0; ii < 8; ++ii) v[ii] = add4(v[ii], bias);
  if (e.acc_in) {  // MRF accumulate (the last pair of a ResBlock)
    const float* ap = e.acc_in + off;
    float4 a[8];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
      a[ii] = 4 * ii < nv ? *reinterpret_cast<const float4*>(ap + ii * STEP) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) v[ii] = add4(a[ii], v[ii]);
  }
  if (e.post_div == 3.f) {  // V1's three resblock kernels
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) v[ii] = div3_rn(v[ii]);
  } else if (e.post_div > 0.f) {
    const float d = e.post_div;
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      v[ii].x = __fdiv_rn(v[ii].x, d); v[ii].y = __fdiv_rn(v[ii].y, d);
      v[ii].z = __fdiv_rn(v[ii].z, d); v[ii].w = __fdiv_rn(v[ii].w, d);
    }
  }
  if (e.out_x) {
    float* xp = e.out_x + off;
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
      if (4 * ii < nv) *reinterpret_cast<float4*>(xp + ii * STEP) = v[ii];
  }
  if (e.out_a0) {  // bf16 operand copy of leaky_relu(result) (the pair kernels run in bf16 mode only)
    __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(e.out_a0) + off;
    const float sl = e.slope;
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
      if (4 * ii < nv)
        *reinterpret_cast<uint2*>(hp + ii * STEP) = lrelu_pack4(v[ii], sl);
  }
}

// epilogue_rows for a caller that still has to fetch the residual itself, with the summation order of the
// pair kernels' staged epilogue: (accumulator + residual) + bias.  All residual loads are issued first.
template <int NR>
__device__ __forceinline__ void epilogue_rows_res_first(EpiParams e, int b, long long q0, int row_step, int n,
                                                        float (&v)[NR][4], long long q_limit) {
  const long long f0 = q0 * e.out_row_stride + n + e.out_offset;
  const long long fstep = static_cast<long long>(row_step) * e.out_row_stride;
  const float* rp = e.res + static_cast<long long>(b) * e.out_batch_stride + f0;
  float4 r[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const long long f = f0 + i * fstep;
    const bool ok = f >= 0 && f + 4 <= e.out_extent && q0 + static_cast<long long>(i) * row_step < q_limit;
    r[i] = ok ? *reinterpret_cast<const float4*>(rp + i * fstep) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) { v[i][0] += r[i].x; v[i][1] += r[i].y; v[i][2] += r[i].z; v[i][3] += r[i].w; }
  e.res = nullptr;
  epilogue_rows<NR, false>(e, b, q0, row_step, n, v, q_limit);
}

// Split form of epilogue_rows for software-pipelined epilogues: epi_prefetch issues the residual
// loads of NR rows (and computes their validity) early; epi_finish consumes them later.
template <int NR>
__device__ __forceinline__ void epi_prefetch(const EpiParams& e, int b, long long q0, int row_step, int n,
                                             long long q_limit, float4 (&r)[NR], unsigned& okmask) {
  const long long f0 = q0 * e.out_row_stride + n + e.out_offset;
  const long long fstep = static_cast<long long>(row_step) * e.out_row_stride;
  const long long base = static_cast<long long>(b) * e.out_batch_stride;
  okmask = 0;
  const bool col_ok = !(e.n_valid > 0 && n + 4 > e.n_valid);
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const long long f = f0 + i * fstep;
    if (col_ok && f >= 0 && f + 4 <= e.out_extent && q0 + static_cast<long long>(i) * row_step < q_limit) okmask |= 1u << i;
  }
  if (e.res) {
    const float* rp = e.res + base + f0;
#pragma unroll
    for (int i = 0; i < NR; ++i)
      r[i] = (okmask >> i) & 1u ? *reinterpret_cast<const float4*>(rp + i * fstep) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
#pragma unroll
    for (int i = 0; i < NR; ++i) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NR>
__device__ __forceinline__ void epi_finish(const EpiParams& e, int b, long long q0, int row_step, int n,
                                           float (&v)[NR][4], const float4 (&r)[NR], unsigned okmask) {
  const long long f0 = q0 * e.out_row_stride + n + e.out_offset;
  const long long fstep = static_cast<long long>(row_step) * e.out_row_stride;
  const long long base = static_cast<long long>(b) * e.out_batch_stride;
  const float4 bb = *reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    v[i][0] = (v[i][0] + bb.x) + r[i].x; v[i][1] = (v[i][1] + bb.y) + r[i].y;
    v[i][2] = (v[i][2] + bb.z) + r[i].z; v[i][3] = (v[i][3] + bb.w) + r[i].w;
  }
  if (e.acc_in) {
    const float* ap = e.acc_in + base + f0;
    float4 a[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i)
      a[i] = (okmask >> i) & 1u ? *reinterpret_cast<const float4*>(ap + i * fstep) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      v[i][0] = a[i].x + v[i][0]; v[i][1] = a[i].y + v[i][1]; v[i][2] = a[i].z + v[i][2]; v[i][3] = a[i].w + v[i][3];
    }
  }
  if (e.post_div > 0.f) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      v[i][0] = __fdiv_rn(v[i][0], e.post_div); v[i][1] = __fdiv_rn(v[i][1], e.post_div);
      v[i][2] = __fdiv_rn(v[i][2], e.post_div); v[i][3] = __fdiv_rn(v[i][3], e.post_div);
    }
  }
  if (e.out_x) {
    float* xp = e.out_x + base + f0;
#pragma unroll
    for (int i = 0; i < NR; ++i)
      if ((okmask >> i) & 1u) *reinterpret_cast<float4*>(xp + i * fstep) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
  }
  if (e.out_a0) {
    const float sl = e.slope;
    if (e.a_fmt == A_BF16) {
      __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(e.out_a0) + base + f0;
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if ((okmask >> i) & 1u)
          *reinterpret_cast<uint2*>(hp + i * fstep) = pack_bf16x4(lrelu_fast(v[i][0], sl), lrelu_fast(v[i][1], sl),
                                                                  lrelu_fast(v[i][2], sl), lrelu_fast(v[i][3], sl));
    } else if (e.a_fmt == A_BF16_SPLIT) {
      __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(e.out_a0) + base + f0;
      __nv_bfloat16* lp = static_cast<__nv_bfloat16*>(e.out_a1) + base + f0;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        if (!((okmask >> i) & 1u)) continue;
        const float a0 = lrelu_fast(v[i][0], sl), a1 = lrelu_fast(v[i][1], sl), a2 = lrelu_fast(v[i][2], sl), a3 = lrelu_fast(v[i][3], sl);
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(a0, a1), h23 = __floats2bfloat162_rn(a2, a3);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&h01);
        pk.y = *reinterpret_cast<const uint32_t*>(&h23);
        *reinterpret_cast<uint2*>(hp + i * fstep) = pk;
        *reinterpret_cast<uint2*>(lp + i * fstep) =
            pack_bf16x4(a0 - __low2float(h01), a1 - __high2float(h01), a2 - __low2float(h23), a3 - __high2float(h23));
      }
    } else {
      float* fp = static_cast<float*>(e.out_a0) + base + f0;
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if ((okmask >> i) & 1u)
          *reinterpret_cast<float4*>(fp + i * fstep) = make_float4(lrelu_fast(v[i][0], sl), lrelu_fast(v[i][1], sl),
                                                                   lrelu_fast(v[i][2], sl), lrelu_fast(v[i][3], sl));
    }
  }
}

// Scalar variant for layers whose width is not a multiple of 4 (tiny / narrow configs on the
// CUDA-core path).
__device__ __forceinline__ void epilogue_scalar(const EpiParams& e, int b, long long q, int n, float v) {
  const long long f = q * e.out_row_stride + n + e.out_offset;
  if (f < 0 || f >= e.out_extent || (e.n_valid > 0 && n >= e.n_valid)) return;
  const long long idx = static_cast<long long>(b) * e.out_batch_stride + f;
  v += e.bias[n];
  if (e.res) v += e.res[idx];
  if (e.acc_in) v = e.acc_in[idx] + v;
  if (e.post_div > 0.f) v = __fdiv_rn(v, e.post_div);
  if (e.out_x) e.out_x[idx] = v;
  if (e.out_a0) {
    const float a = lrelu(v, e.slope);
    if (e.a_fmt == A_F32) {
      static_cast<float*>(e.out_a0)[idx] = a;
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(a);
      static_cast<__nv_bfloat16*>(e.out_a0)[idx] = h;
      if (e.a_fmt == A_BF16_SPLIT)
        static_cast<__nv_bfloat16*>(e.out_a1)[idx] = __float2bfloat16_rn(a - __bfloat162float(h));
    }
  }
}

// Read one operand element in any format (CUDA-core path).
__device__ __forceinline__ float load_operand(const void* a0, const void* a1, int fmt, long long idx) {
  if (fmt == A_F32) return static_cast<const float*>(a0)[idx];
  float v = __bfloat162float(static_cast<const __nv_bfloat16*>(a0)[idx]);
  if (fmt == A_BF16_SPLIT) v += __bfloat162float(static_cast<const __nv_bfloat16*>(a1)[idx]);
  return v;
}

// ------------------------------------------------------------------ tcgen05 implicit-GEMM conv
struct TcConvParams {
  int B;
  int rows;            // GEMM rows per item (L_in for conv, L_in + 1 for the polyphase convT)
  int tiles_per_item;  // ceil(rows / (MS*128))
  int nc;              // K chunks of KC channels
  int ntaps;
  int tap_row[kMaxTaps];  // slab-relative first row of each tap (tap offset - min offset) >= 0
  int min_off;            // slab row 0 = tile row 0 + min_off (<= 0 normally)
  int slab_rows;          // rows per slab buffer = nboxes*box_rows
  int box_rows, nboxes;
  int nbuf;     // slab chunk ring depth
  int stages;   // weight ring depth (== nc*ntaps*planes when w_resident)
  int w_resident;  // 1: every weight tile stays in shared memory for the CTA's lifetime
  int n_blocks;    // N tiles (grid-strided together with the M tiles)
  int total_work;  // B * tiles_per_item * n_blocks
  int epi_tma;     // 1: TMA epilogue (residual tiles loaded, x / operand tiles stored by TMA); 0: generic
  int has_res, has_x, has_a;  // what the TMA epilogue reads / writes (maps are kernel arguments)
  int has_acc;                // conv_tc2 only: the MRF running sum is a second TMA-loaded input tile
  int in_bufs;                // conv_tc2 only: input (residual / running sum) tiles in flight per epilogue warp, 1 or 2
  int epi_slot_bytes;         // shared memory per epilogue warp (TMA: res | x | a_hi | a_lo tiles; generic: 2 KB)
  // TMA epilogue of a polyphase ConvTranspose1d: GEMM column n = phase * out_cmod + channel, GEMM row q lands on output
  // row q * out_rstride + phase + out_roff.  The output maps are 4-D views [item][L_out / stride][stride][C], in which the
  // 32 rows of an item (one phase) are a plain box.  out_cmod = 0: same-length conv, 3-D maps.
  int out_cmod, out_rstride, out_roff;
  int desc_mode;  // how a tap's row shift enters the UMMA descriptor (see conv_tc.cu)
  long long* dbg;       // optional [grid][8] cycle counters (HG_TC_DEBUG_TIMING): MMA-warp wait breakdown
  RaggedPrefix rag;     // ragged batch: compacted tile index space (n == 0: dense)
  const uint8_t* w_hi;  // packed swizzled weight tiles [n_blk][chunk][tap][N_T rows][KC]
  const uint8_t* w_lo;
  EpiParams epi;
};

// ------------------------------------------------------------------ fused ResBlock pair (tcgen05)
// xt = lrelu(c1(a) + b1) stays in shared memory as the bf16 operand of c2; y = c2(xt) + b2 (+res ...)
// goes through the usual epilogue (hifi/models.py:90-94).  C_in = C_out = C in {32, 64}.
struct TcPairParams {
  int B;
  int L;               // sequence length (rows per item)
  int r_out;           // output rows per tile = MS*128 - (k-1)
  int tiles_per_item;  // ceil(L / r_out)
  int total_work;      // B * tiles_per_item
  int k;               // taps of both convs
  int d1;              // dilation of c1 (c2 has dilation 1)
  int slab_rows, box_rows, nboxes;  // input slab = MS*128 + d1*(k-1) rows (rounded)
  int t_rows;          // rows of the intermediate tile buffer (MS*128 + k - 1, rounded to 16)
  int t_bufs;          // 1 or 2 intermediate buffers
  int stages;          // weight ring depth (== 2*k when w_resident)
  int w_resident;
  const uint8_t* w1;   // packed swizzled tiles [tap][C rows][C]  (conv 1)
  const uint8_t* w2;   //                                         (conv 2)
  const float* bias1;  // [C]
  float slope;         // leaky_relu slope applied to xt (0.1)
  long long* dbg;      // optional [grid][16] cycle counters (HG_TC_DEBUG_TIMING): wait breakdown
  RaggedPrefix rag;    // ragged batch: compacted tile index space (n == 0: dense)
  EpiParams epi;       // epilogue of c2
};

// ------------------------------------------------------------------ fused ResBlock (tcgen05)
// conv_chain_tc.cu: all NP (dilated conv, conv) pairs of a ResBlock1 in one kernel; the residual stream and the
// operand stay in shared memory between the pairs.  C in {32, 64}, small k (halo = sum of the reaches).
constexpr int kChainMaxPairs = 3;
struct TcChainParams {
  int B;
  int L;               // sequence length (rows per item)
  int r_out;           // output rows per tile = MS*128 - 2*halo
  int tiles_per_item;  // ceil(L / r_out)
  int total_work;
  int k;               // taps of every conv
  int np;              // pairs
  int d1[kChainMaxPairs];  // dilation of c1 of each pair (c2 has dilation 1)
  int halo;            // rows lost on each side of a tile: sum over the pairs of (d1 + 1) * (k - 1) / 2
  int buf_rows;        // rows of every on-chip operand buffer: guard + MS*128 + guard (rounded to 1024 bytes)
  int box_rows, nboxes;  // TMA boxes of the input slab
  int stages;          // weight ring depth (== 2*np*k when w_resident)
  int w_resident;
  const uint8_t* w1[kChainMaxPairs];  // packed swizzled tiles [tap][C rows][C] of c1 / c2 of each pair
  const uint8_t* w2[kChainMaxPairs];
  const float* bias1[kChainMaxPairs];  // [C]
  const float* bias2[kChainMaxPairs];  // [C]; the last pair's is applied by the fused epilogue (epi.bias)
  float slope;
  RaggedPrefix rag;
  EpiParams epi;       // epilogue of the last c2
};

// ------------------------------------------------------------------ fused ResBlock pair, time-folded (tcgen05)
// conv_pair_fold.cu: the same pair as TcPairParams, with F = 128 / C consecutive (dilation-strided) time
// rows folded into the N dimension of every MMA, so that one A-operand fetch feeds N = 128 columns.
// One MMA group ("op") = one A start (phase slab + row shift) against a run of consecutive taps stacked
// along N.  The kernel derives its groups at compile time (template on the tap count); FoldOp is the
// host-side description of the same schedule (api.cu::fold_schedule -> hg_fold_info, replayed by
// tests/test_fold_schedule.py).
constexpr int kFoldMaxOps = 24;  // k + F - 1 <= 15 + 8
struct FoldOp {
  int a_off16;  // A start: (bytes >> 4) from the operand buffer base (phase slab + row shift)
  int b_blk;    // first tap (weight block) of B
  int nblk;     // taps stacked along N: N = nblk * C
  int d_col;    // first accumulator column
  int rel;      // streamed weights: the oldest held weight block has had its last use after this op
};
struct TcFoldParams {
  int B;
  int L;               // sequence length (rows per item), a multiple of F
  int r_out;           // output rows per tile (a multiple of F * d1)
  int tiles_per_item;  // ceil(L / r_out)
  int total_work;
  int d1;              // dilation of conv 1 (conv 2 has dilation 1; the tap count is a template parameter)
  int delta;           // tile origin: xt row 0 of tile t is global row t * r_out - delta
  int fdiv;            // F * d1: rows per block group of the de-interleaved input slab
  int blk_off;         // first slab block = origin / fdiv + blk_off (<= 0)
  int nblk_item;       // ceil(L / fdiv): extent of the block dimension of the input map
  int nb_slab;         // blocks per slab box
  int slab_phase_bytes, xt_phase_bytes;  // one phase slab of the input / of xt (multiples of 1024)
  int a1_row0;         // conv 1: slab row of block-group shift 0 (= -blk_off * d1); shift s adds s * d1 rows
  int a2_row0;         // conv 2: xt phase row of shift 0 for M row 0 (= delta / F); shift s adds s rows
  int t_bufs;          // 1 or 2 xt buffers
  int stages;          // streamed weights: ring depth (resident kernels hold all 2k blocks)
  int dbg;             // EXPERIMENT: bit switches that drop parts of the kernel's work (timing only)
  const uint8_t* w1;   // packed swizzled tiles [tap][C rows][C]  (the Layer's w_hi)
  const uint8_t* w2;
  const float* bias1;  // [C]
  float slope;
  RaggedPrefix rag;
  EpiParams epi;       // epilogue of c2 over the FOLDED view [L/F][128] (bias replicated F times)
};

// ------------------------------------------------------------------ CUDA-core (FFMA) conv
struct FfmaConvParams {
  int B, L_in, rows, cin, n_total;
  int ntaps;
  int tap_off[kMaxTaps];  // input row = q + tap_off[j]
  int a_pitch;            // channels per operand row (>= cin when the producer padded)
  const void* a0;         // operand planes, [B][L_in][a_pitch]
  const void* a1;
  int a_fmt;
  const float* w;  // [tap][cin][n_total]
  RaggedPrefix rag;
  EpiParams epi;
};

// ------------------------------------------------------------------ narrow CUDA-core conv (C = 8 / 16)
struct NarrowConvParams {
  int B, L, k, dil, tiles_per_item;
  const void* a0;   // operand planes [B][L][C]
  const void* a1;
  int a_fmt;
  const float* w;   // [k][C_in][C_out]  (the CUDA-core weight layout of plan.h)
  RaggedPrefix rag;
  EpiParams epi;
};

}  // namespace hg
