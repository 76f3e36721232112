// conv_chain_tc.cu — a whole ResBlock1 (all its (dilated conv, conv) pairs) as ONE tcgen05 kernel
// (bf16 mode, C = 32 or 64, small kernels — k = 3 in HiFi-GAN V1):
//
//     for m in 0..NP-1:   xt = leaky_relu(c1_m(a) + b1_m)        hifi/models.py:90-92
//                         x  = c2_m(xt) + b2_m + x               hifi/models.py:93-94
//                         a  = leaky_relu(x)
//     (+ MRF accumulate / divide / next operand copy in the last epilogue, as in conv_pair_tc.cu)
//
// Why: at k = 3 a fused pair is HBM-bound — 12 bytes per element per pair (operand in, fp32 residual in,
// fp32 residual out, operand out) for 2*2*3*C MACs — and a ResBlock is three such launches.  Here the
// residual stream and the operand never leave the SM between the pairs: the block reads `a` (2 B) and `x`
// (4 B) once and writes the final `x` / `a` once, 36 -> 12 bytes per element.
//
// How: conv_pair_tc.cu's structure with the hand-off kept on chip.  A tile is MT = MS*128 rows; every conv is
// evaluated on all MT rows (tap shifts are descriptor row shifts into buffers that carry G guard rows on
// either side), and the range of rows that are actually correct shrinks by d*(k-1)/2 per conv — after the
// NP pairs the middle R = MT - 2*H rows (H = sum of the reaches, 12 for k = 3 and dilations 1/3/5) are kept,
// tiles advance by R.  Per pair: G1 (c1) accumulates in TMEM; E1 applies bias + leaky_relu and writes xt as
// the bf16 UMMA operand of c2; G2 (c2) accumulates; E2 adds the accumulator and b2 into the fp32 residual
// tile — which is TMA-prefetched into the epilogue warps' staging slots once per tile and stays there — and,
// for all but the last pair, writes leaky_relu(x) as the bf16 operand of the next c1 over the input slab; the
// last pair's E2 is the pair kernel's (transposed read of the slot, fused epilogue to HBM).  Rows outside
// [0, L) are forced to zero in xt and in the operand, exactly as the zero padding of the separate convs sees
// them.  Every operation happens in the same order as in the three separate launches, so the result is
// bit-identical to them (tests/test_gpu_ops.py::test_fused_resblock_parity).
//
// The pairs of one tile are sequential (each GEMM waits for the previous epilogue), so tensor pipe and
// epilogue warps alternate inside a tile; consecutive tiles overlap at the seam (the next tile's slab and
// residual loads, and its first GEMM, run under the last epilogue).  Warp roles as in conv_pair_tc.cu:
// 0 weight producer | 1 MMA issuer + TMEM owner | 2 slab producer (TMA) | 3..18 epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

constexpr int kChainEpiWarps = 16;
constexpr int kChainThreads = (3 + kChainEpiWarps) * 32;
constexpr int kChainStageFloats = 32 * 32;  // per-warp residual tile: 32 rows x 32 fp32 columns
constexpr int kChainGuard = 8;              // guard rows around every on-chip operand buffer (>= max d*(k-1)/2)

template <int C, int MS>
__global__ void __launch_bounds__(kChainThreads, 1)
conv_chain_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_res,
                     const __grid_constant__ TcChainParams p) {
  constexpr int KC = C, N_T = C;
  constexpr int ROWB = KC * 2;
  constexpr int STAGE_BYTES = N_T * ROWB;
  constexpr int KSTEPS = KC / 16;
  constexpr int G = kChainGuard;
  constexpr uint32_t ACC_COLS = MS * N_T;  // 128
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr uint32_t LAYOUT = (KC == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
  static_assert(ACC_COLS == 128, "one accumulator = 128 TMEM columns");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int buf_bytes = p.buf_rows * ROWB;    // multiple of 1024 (host rounds the rows)
  uint8_t* slab = smem;                       // [2][buf_bytes]  operand of c1: input slab, then leaky_relu(x) of pairs 0..NP-2
  uint8_t* tbuf = slab + 2 * buf_bytes;       // [buf_bytes]     xt, operand of c2
  float* staging = reinterpret_cast<float*>(tbuf + buf_bytes);  // [16][4 KB] the fp32 residual tile, one item per warp
  uint8_t* wst = reinterpret_cast<uint8_t*>(staging + kChainEpiWarps * kChainStageFloats);  // [stages][STAGE_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + p.stages * STAGE_BYTES);
  uint64_t* slab_full = bars;        // [2]
  uint64_t* slab_empty = bars + 2;   // [2]
  uint64_t* d1_full = bars + 4;      // G1 -> E1
  uint64_t* d1_empty = bars + 5;     // E1 -> G1
  uint64_t* t_full = bars + 6;       // E1 -> G2
  uint64_t* d2_full = bars + 7;      // G2 -> E2
  uint64_t* d2_empty = bars + 8;     // E2 -> G2
  uint64_t* a_full = bars + 9;       // E2 (not last) -> next G1: the operand of the next pair is in the slab
  uint64_t* res_bar = bars + 10;     // [16] residual tile landed in a warp's staging slot (TMA)
  uint64_t* w_full = bars + 26;      // [stages]
  uint64_t* w_empty = w_full + p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_empty + p.stages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h2 = (p.k - 1) >> 1;
  const int n_my = p.total_work > static_cast<int>(blockIdx.x)
                       ? (p.total_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&map_in); prefetch_tensormap(&map_res); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) { mbar_init(&slab_full[i], 1); mbar_init(&slab_empty[i], 1); }
      mbar_init(d1_full, 1); mbar_init(d1_empty, kChainEpiWarps);
      mbar_init(t_full, kChainEpiWarps);
      mbar_init(d2_full, 1); mbar_init(d2_empty, kChainEpiWarps);
      mbar_init(a_full, kChainEpiWarps);
      for (int s = 0; s < p.stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
      for (int w = 0; w < kChainEpiWarps; ++w) mbar_init(&res_bar[w], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: setup and the first weight stages overlap the previous layer's tail
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------ weight producer: conv order c1_0 c2_0 c1_1 c2_1 ...
    if (lane == 0 && n_my > 0) {
      int stage = 0; uint32_t phase = 0;
      auto load_conv = [&](const uint8_t* w) {
        for (int t = 0; t < p.k; ++t) {
          mbar_wait(&w_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&w_full[stage], STAGE_BYTES);
          bulk_load_1d(wst + stage * STAGE_BYTES, w + static_cast<size_t>(t) * STAGE_BYTES, STAGE_BYTES, &w_full[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      };
      const int reps = p.w_resident ? 1 : n_my;
      for (int i = 0; i < reps; ++i)
        for (int m = 0; m < p.np; ++m) { load_conv(p.w1[m]); load_conv(p.w2[m]); }
    }
  } else if (warp == 2) {
    // ------------------------------------------------ input slab producer (TMA): the operand of pair 0
    if (lane == 0) {
      for (int i = 0; i < n_my; ++i) {
        const int work = blockIdx.x + i * gridDim.x;
        int b, tile;
        decode_tile(p.rag, p.tiles_per_item, work, b, tile);
        const int r0 = tile * p.r_out - p.halo;  // global row of tile row 0
        const int buf = i & 1;
        mbar_wait(&slab_empty[buf], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&slab_full[buf], buf_bytes);
        uint8_t* dst = slab + buf * buf_bytes;
        for (int bx = 0; bx < p.nboxes; ++bx)
          tma_load_3d(dst + bx * p.box_rows * ROWB, &map_in, &slab_full[buf], 0, r0 - G + bx * p.box_rows, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, N_T);
    constexpr uint32_t desc_hi = ((SBO >> 4) & 0x3FFFu) | (1u << 14) | (LAYOUT << 29);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t slab_lo = (smem_u32(slab) & 0x3FFFFu) >> 4;
    const uint32_t t_lo = (smem_u32(tbuf) & 0x3FFFFu) >> 4;
    const uint32_t wst_lo = (smem_u32(wst) & 0x3FFFFu) >> 4;
    int stage = 0; uint32_t wphase = 0;
    bool w_seen = false;  // resident mode: every stage has been waited for once
    uint32_t n_d1 = 0, n_t = 0, n_d2 = 0, n_a = 0;  // uses of the per-CTA barriers so far (their parity)

    // one GEMM: k taps, A = base + tap*row_step rows, accumulate into `acc`
    auto gemm = [&](uint32_t a_base_lo, uint32_t tap_step_lo, uint32_t acc, int conv) {
      for (int t = 0; t < p.k; ++t) {
        const int st = p.w_resident ? conv * p.k + t : stage;
        if (!p.w_resident || !w_seen) {
          mbar_wait(&w_full[st], p.w_resident ? 0u : wphase);
          tc_fence_after();
        }
        const uint32_t b_lo = wst_lo + static_cast<uint32_t>(st) * (STAGE_BYTES >> 4);
        const uint32_t a_lo0 = a_base_lo + static_cast<uint32_t>(t) * tap_step_lo;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int ms = 0; ms < MS; ++ms)
              umma_bf16_lohi(acc + ms * N_T, a_lo0 + static_cast<uint32_t>(ms * ((128 * ROWB) >> 4) + ks * 2), b_lo + ks * 2,
                             desc_hi, idesc, (ks == 0 && t == 0) ? 0u : 1u);
          }
          if (!p.w_resident) umma_commit(&w_empty[stage]);
        }
        __syncwarp();
        if (!p.w_resident && ++stage == p.stages) { stage = 0; wphase ^= 1; }
      }
    };
    for (int i = 0; i < n_my; ++i) {
      const int buf = i & 1;
      const uint32_t sl = slab_lo + static_cast<uint32_t>(buf) * (static_cast<uint32_t>(buf_bytes) >> 4);
      for (int m = 0; m < p.np; ++m) {
        const int d = p.d1[m];
        // G1: operand = slab rows G + row + (t - h2) * d
        if (m == 0) mbar_wait(&slab_full[buf], (i >> 1) & 1);
        else { mbar_wait(a_full, n_a & 1); ++n_a; }
        mbar_wait(d1_empty, (n_d1 & 1) ^ 1);
        tc_fence_after();
        gemm(sl + (static_cast<uint32_t>(G - h2 * d) * ROWB >> 4), (static_cast<uint32_t>(d) * ROWB) >> 4, tmem_u, 2 * m);
        if (elect_one()) {
          if (m == p.np - 1) umma_commit(&slab_empty[buf]);  // last read of this slab buffer
          umma_commit(d1_full);
        }
        __syncwarp();
        ++n_d1;
        // G2: operand = xt rows G + row + (t - h2)
        mbar_wait(t_full, n_t & 1); ++n_t;
        mbar_wait(d2_empty, (n_d2 & 1) ^ 1);
        tc_fence_after();
        gemm(t_lo + (static_cast<uint32_t>(G - h2) * ROWB >> 4), ROWB >> 4, tmem_u + ACC_COLS, 2 * m + 1);
        if (elect_one()) umma_commit(d2_full);
        __syncwarp();
        ++n_d2;
      }
      w_seen = true;
    }
  } else {
    // ------------------------------------------------ epilogue warps (all 16 do E1 and E2 of every pair)
    const int e = warp - 3;
    const int quarter = warp & 3;
    const int sub = e >> 2;  // 0..3: which of the four warps sharing this TMEM lane quarter
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    float* stg = staging + e * kChainStageFloats;
    const int c4 = lane & 7, rsub = lane >> 3;
    // the one (sub-tile, 32-column) item of every tile this warp owns in E2
    constexpr int E2_CH = N_T / 32;
    const int ms2 = sub / E2_CH, c02 = (sub - ms2 * E2_CH) * 32;
    const int n2 = c02 + c4 * 4;
    uint32_t n_d1 = 0, n_d2 = 0;

    auto coords = [&](int i, int& b, int& r0) {
      const int work = blockIdx.x + i * gridDim.x;
      int tile;
      decode_tile(p.rag, p.tiles_per_item, work, b, tile);
      r0 = tile * p.r_out - p.halo;
    };
    auto prefetch_res = [&](int i) {  // lane 0 only: this warp's item of the tile's fp32 residual
      int b, r0;
      coords(i, b, r0);
      mbar_arrive_expect_tx(&res_bar[e], kChainStageFloats * 4);
      tma_load_3d(stg, &map_res, &res_bar[e], c02, r0 + ms2 * 128 + quarter * 32, b);
    };
    // E1: D1 -> (+b1, leaky_relu, bf16) -> xt in UMMA layout; two (sub-tile, 16-column) items per warp
    auto e1 = [&](int r0, const float* bias1) {
      mbar_wait(d1_full, n_d1 & 1); ++n_d1;
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + lane_base;
      constexpr int ITEMS = MS * (N_T / 16);
#pragma unroll 1
      for (int j = sub; j < ITEMS; j += 4) {
        const int ms = j / (N_T / 16), c0 = (j - ms * (N_T / 16)) * 16;
        uint32_t r[16];
        tmem_ld_32x16(tmem_acc + ms * N_T + c0, r);
        tmem_ld_wait();
        if (j + 4 >= ITEMS) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d1_empty);
        }
        const int row = ms * 128 + quarter * 32 + lane;  // tile row
        const int grow = r0 + row;                       // global time index of that row
        const bool inside = grow >= 0 && grow < p.L;
        uint32_t pk[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bb = *reinterpret_cast<const float4*>(bias1 + c0 + 4 * q);
          const float v0 = inside ? lrelu_fast(__uint_as_float(r[4 * q]) + bb.x, p.slope) : 0.f;
          const float v1 = inside ? lrelu_fast(__uint_as_float(r[4 * q + 1]) + bb.y, p.slope) : 0.f;
          const float v2 = inside ? lrelu_fast(__uint_as_float(r[4 * q + 2]) + bb.z, p.slope) : 0.f;
          const float v3 = inside ? lrelu_fast(__uint_as_float(r[4 * q + 3]) + bb.w, p.slope) : 0.f;
          const uint2 u = pack_bf16x4(v0, v1, v2, v3);
          pk[2 * q] = u.x; pk[2 * q + 1] = u.y;
        }
        const int brow = G + row;  // buffer row
        const uint32_t swz = (KC == 64) ? (brow & 7) : ((brow >> 1) & 3);
        const int ch = c0 >> 3;    // first 16-byte chunk of this item within the row
        uint8_t* rp = tbuf + brow * ROWB;
        *reinterpret_cast<uint4*>(rp + (((ch) ^ swz) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rp + (((ch + 1) ^ swz) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async();  // generic-proxy writes of xt -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(t_full);
    };
    // E2 of every pair but the last: x += D2 + b2 in the staging slot, leaky_relu(x) -> bf16 -> the slab
    auto e2_mid = [&](int i, int r0, int buf, const float* bias2, bool first_pair) {
      mbar_wait(d2_full, n_d2 & 1); ++n_d2;
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ACC_COLS + lane_base + ms2 * N_T + c02, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty);
      if (first_pair) mbar_wait(&res_bar[e], i & 1);
      const int row = ms2 * 128 + quarter * 32 + lane;
      const int grow = r0 + row;
      const bool inside = grow >= 0 && grow < p.L;
      uint32_t pk[16];
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        float4* sp = reinterpret_cast<float4*>(stg + lane * 32 + ((k4 ^ (lane & 7)) << 2));
        const float4 bb = *reinterpret_cast<const float4*>(bias2 + c02 + 4 * k4);
        float4 t = *sp;
        // same order of operations as the separate launch: (residual + accumulator) + bias
        t.x = (t.x + __uint_as_float(r[4 * k4])) + bb.x; t.y = (t.y + __uint_as_float(r[4 * k4 + 1])) + bb.y;
        t.z = (t.z + __uint_as_float(r[4 * k4 + 2])) + bb.z; t.w = (t.w + __uint_as_float(r[4 * k4 + 3])) + bb.w;
        *sp = t;
        const uint2 u = inside ? pack_bf16x4(lrelu_fast(t.x, p.slope), lrelu_fast(t.y, p.slope), lrelu_fast(t.z, p.slope),
                                             lrelu_fast(t.w, p.slope))
                               : make_uint2(0u, 0u);
        pk[2 * k4] = u.x; pk[2 * k4 + 1] = u.y;
      }
      const int brow = G + row;
      const uint32_t swz = (KC == 64) ? (brow & 7) : ((brow >> 1) & 3);
      const int ch = c02 >> 3;
      uint8_t* rp = slab + buf * buf_bytes + brow * ROWB;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(rp + (((ch + q) ^ swz) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      fence_proxy_async();  // generic-proxy writes of the operand -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
    };
    // E2 of the last pair: conv_pair_tc.cu's — accumulator into the slot, transposed read, fused epilogue to HBM
    auto e2_last = [&](int i, int b, int r0, bool first_pair) {
      mbar_wait(d2_full, n_d2 & 1); ++n_d2;
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ACC_COLS + lane_base + ms2 * N_T + c02, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty);
        if (first_pair) mbar_wait(&res_bar[e], i & 1);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          float4* sp = reinterpret_cast<float4*>(stg + lane * 32 + ((k4 ^ (lane & 7)) << 2));
          float4 t = *sp;
          t.x += __uint_as_float(r[4 * k4]); t.y += __uint_as_float(r[4 * k4 + 1]);
          t.z += __uint_as_float(r[4 * k4 + 2]); t.w += __uint_as_float(r[4 * k4 + 3]);
          *sp = t;
        }
      }
      __syncwarp();
      float v[8][4];
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int row = ii * 4 + rsub;
        const float4 t4 = *reinterpret_cast<const float4*>(stg + row * 32 + ((c4 ^ (row & 7)) << 2));
        v[ii][0] = t4.x; v[ii][1] = t4.y; v[ii][2] = t4.z; v[ii][3] = t4.w;
      }
      fence_proxy_async();  // our generic reads of the slot happen-before the next TMA write into it
      __syncwarp();
      if (lane == 0 && i + 1 < n_my) prefetch_res(i + 1);
      const long long m0 = static_cast<long long>(r0) + p.halo;  // first row this tile keeps
      epilogue_rows<8, false>(p.epi, b, static_cast<long long>(r0) + ms2 * 128 + quarter * 32 + rsub, 4, n2, v, m0 + p.r_out, m0);
    };
    if (n_my > 0 && lane == 0) prefetch_res(0);
    for (int i = 0; i < n_my; ++i) {
      int b, r0;
      coords(i, b, r0);
      const int buf = i & 1;
      for (int m = 0; m < p.np; ++m) {
        e1(r0, p.bias1[m]);
        if (m + 1 < p.np) e2_mid(i, r0, buf, p.bias2[m], m == 0);
        else e2_last(i, b, r0, m == 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
size_t conv_chain_smem_bytes(int c, int buf_rows, int stages) {
  const int rowb = c * 2;
  return 1024 + 3 * static_cast<size_t>(buf_rows) * rowb + static_cast<size_t>(stages) * c * rowb +
         kChainEpiWarps * kChainStageFloats * 4 + (26 + 2 * stages) * 8 + 16;
}
int conv_chain_guard_rows() { return kChainGuard; }

template <int C, int MS>
static cudaError_t launch_chain(const CUtensorMap& m, const CUtensorMap& mr, const TcChainParams& p, size_t smem, int grid,
                                cudaStream_t st) {
  auto kern = conv_chain_tc_kernel<C, MS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, m, mr, p);
}

cudaError_t launch_conv_chain_tc(int c, const CUtensorMap& m, const CUtensorMap& mr, const TcChainParams& p, size_t smem,
                                 int grid, cudaStream_t st) {
  if (c == 64) return launch_chain<64, 2>(m, mr, p, smem, grid, st);
  if (c == 32) return launch_chain<32, 4>(m, mr, p, smem, grid, st);
  return cudaErrorInvalidValue;
}

}  // namespace hg
